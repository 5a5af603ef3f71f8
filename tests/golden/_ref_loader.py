"""Moved to oracle/ref_loader.py (also used by bench.py --impl reference); kept as an import alias for make_golden.py."""
from oracle.ref_loader import *  # noqa: F401,F403
from oracle.ref_loader import CfgNode, reference, REF  # noqa: F401
