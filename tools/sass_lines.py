#!/usr/bin/env python
"""Static instruction budget of one kernel: SASS instructions per source line (needs -lineinfo, which build.py passes)
and the opcode histogram.  Works without a GPU (cuobjdump / nvdisasm on the in-tree object files).

    python tools/sass_lines.py open-world-semantic-segmentation_b200/build/ood_sort.o onesweep_kernelIjE [--min 8]

`kernel` is a substring of the mangled name.  Counts are STATIC (both sides of a branch, unrolled loops counted once per
copy); divide a loop body's count by the items it handles to get instructions per item."""
import argparse
import collections
import os
import re
import subprocess
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--min", type=int, default=8, help="hide source lines with fewer instructions")
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=td, check=True, capture_output=True)
        cubins = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
        text = "".join(subprocess.run(["nvdisasm", "-g", c], capture_output=True, text=True).stdout for c in cubins)
    lines = text.split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and a.kernel in l]
    if not starts:
        raise SystemExit(f"no kernel matching {a.kernel!r}")
    start = starts[0]
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//---")), len(lines))
    print("#", lines[start].rstrip(":"))
    cur = ("?", 0)
    per_line, ops = collections.Counter(), collections.Counter()
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            per_line[cur] += 1
            ops[m.group(2).split(".")[0]] += 1
    total = sum(per_line.values())
    print(f"# {total} SASS instructions")
    print("# opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
    for (f, ln), n in sorted(per_line.items()):
        if n >= a.min:
            print(f"{f}:{ln:<5d} {n:5d}")


if __name__ == "__main__":
    main()
