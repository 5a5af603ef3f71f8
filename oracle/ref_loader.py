"""Shim loader that imports the UNMODIFIED reference on a CPU-only box.

TEST / BASELINE INFRASTRUCTURE, like everything under ``oracle/``: used by ``tests/golden/make_golden.py`` (build
container, reference at /root/reference) and by ``bench.py --impl reference`` (the copy ``oracle/install_reference.py``
put under ``baseline/_ref``, which travels to the GPU box).  Nothing in the product package imports it.
The root is ``$DML_REFERENCE_ROOT``, else /root/reference, else <repo>/baseline/_ref.
"""
from __future__ import annotations

import collections
import collections.abc
import contextlib
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    for cand in (os.environ.get("DML_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "..", "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "anomaly")):
            return os.path.abspath(cand)
    return None


REF = reference_root() or "/root/reference"


class _Stub(types.ModuleType):
    """Module stub: any attribute is a no-op callable/namespace (dunder lookups fail)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class CfgNode(dict):
    """Minimal stand-in for yacs.config.CfgNode (attribute dict)."""

    def __init__(self, init=None, **kw):
        super().__init__()
        for k, v in dict(init or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, *a, **k):
        pass

    def merge_from_list(self, *a, **k):
        pass

    def freeze(self):
        pass

    def clone(self):
        import copy
        return copy.deepcopy(self)


def _install_common_stubs():
    import torch
    collections.Sequence = collections.abc.Sequence
    collections.Mapping = collections.abc.Mapping
    if not torch.cuda.is_available() or os.environ.get("DML_REF_FORCE_CPU") == "1":
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.set_device = lambda *a, **k: None
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn", "statsmodels", "statsmodels.distributions",
                 "statsmodels.distributions.empirical_distribution", "visdom"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if "yacs" not in sys.modules:
        try:
            importlib.import_module("yacs.config")
        except Exception:
            yacs = types.ModuleType("yacs")
            ycfg = types.ModuleType("yacs.config")
            ycfg.CfgNode = CfgNode
            yacs.config = ycfg
            sys.modules["yacs"] = yacs
            sys.modules["yacs.config"] = ycfg
    try:
        importlib.import_module("distutils.version")
    except Exception:
        du = types.ModuleType("distutils")
        dv = types.ModuleType("distutils.version")

        class LooseVersion(str):
            pass
        dv.LooseVersion = LooseVersion
        du.version = dv
        sys.modules["distutils"] = du
        sys.modules["distutils.version"] = dv
    # torchvision.models.utils was removed upstream (DeepLab backbone/resnet.py:3)
    try:
        importlib.import_module("torchvision.models.utils")
    except Exception:
        import torchvision
        tvu = types.ModuleType("torchvision.models.utils")
        tvu.load_state_dict_from_url = lambda *a, **k: {}
        sys.modules["torchvision.models.utils"] = tvu
        torchvision.models.utils = tvu


_TOP_LEVEL_REF_MODULES = ("models", "utils", "lib", "config", "dataset", "anom_utils", "network", "metrics",
                          "datasets", "eval_ood_traditional", "test_embedding", "test_self_distillation")


@contextlib.contextmanager
def reference(subproject: str):
    """Context manager: cwd + sys.path[0] set to the sub-project, its top-level module
    names purged afterwards so both sub-projects can be loaded in one process."""
    _install_common_stubs()
    root = os.path.join(reference_root() or REF, subproject)
    old_cwd = os.getcwd()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split(".")[0] in _TOP_LEVEL_REF_MODULES}
    sys.path.insert(0, root)
    os.chdir(root)
    try:
        yield root
    finally:
        os.chdir(old_cwd)
        sys.path.remove(root)
        for k in list(sys.modules):
            if k.split(".")[0] in _TOP_LEVEL_REF_MODULES:
                sys.modules.pop(k)
        sys.modules.update(saved)
