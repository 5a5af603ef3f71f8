#!/usr/bin/env python
"""Per-source-line executed warp instructions / stall samples of one kernel from an ncu report captured with
`--set full --import-source on` (the kernels are compiled with -lineinfo):
    python tools/ncu_lines.py gpurun_out/X.ncu-rep rank_kernel [--top 40]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if not his:
        raise SystemExit("no source page for that kernel")
    hi = his[0]
    h = rows[hi]
    ii, st = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data, tot = [], 0
    end = his[1] if len(his) > 1 else len(rows)
    for r in rows[hi + 1:end]:
        if len(r) <= ii or not r[0]:
            continue
        try:
            n, s = int(r[ii]), int(r[st])
        except ValueError:
            continue
        tot += n
        data.append((n, s, r[0], r[1].strip()[:110]))
    print(f"# {kern}: {tot} warp instructions executed (first captured launch)")
    for n, s, l, src in sorted(data, reverse=True)[:top]:
        print(f"{n:11d} {100 * n / max(tot, 1):5.1f}%  stall_samples={s:6d}  L{l}: {src}")


if __name__ == "__main__":
    main()
