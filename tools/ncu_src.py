"""Per-CUDA-line table from an ncu report captured with --import-source on:
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv ; python tools/ncu_src.py X.csv [topN] [by]
Aggregates the CUDA-line rows of every file: stall samples and executed warp instructions; `by` = inst | samp."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    by = sys.argv[3] if len(sys.argv) > 3 else "samp"
    fname = None
    hdr = None
    per = {}
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ci = {n: i for i, n in enumerate(hdr)}
            i_s, i_n = ci["# Samples"], ci["Instructions Executed"]
            continue
        if hdr is None or not r[0]:
            continue
        try:
            s, n = int(r[i_s]), int(r[i_n])
        except ValueError:
            continue
        key = (fname, int(r[0]))
        if key in per:
            per[key][0] += s; per[key][1] += n
        else:
            per[key] = [s, n, r[1].strip()]
    S = sum(v[0] for v in per.values()) or 1
    I = sum(v[1] for v in per.values()) or 1
    print(f"samples {S}, warp instructions {I/1e9:.3f} G")
    byfile = defaultdict(lambda: [0, 0])
    for (f, _), v in per.items():
        byfile[f][0] += v[0]; byfile[f][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
        print(f"  {f:28s} {100*v[0]/S:5.1f}% samp {100*v[1]/I:5.1f}% inst")
    idx = 1 if by == "inst" else 0
    for (f, ln), v in sorted(per.items(), key=lambda kv: -kv[1][idx])[:topn]:
        print(f"{f[:16]:16s} L{ln:5d} {100*v[0]/S:5.1f}% samp {100*v[1]/I:5.1f}% inst | {v[2][:100]}")


if __name__ == "__main__":
    main()
