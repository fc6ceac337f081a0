"""Phase buckets (by marker comments) of an ncu source-page CSV for one kernel.
usage: python tools/ncu_buckets.py src.csv listing.sass mangled_prefix first_line marker1 marker2 ..."""
import csv
import sys
from collections import Counter

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines  # noqa: E402

SRC = "/root/repo/strutopy_b200/csrc/estep_kernel.cuh"


def main():
    src_csv, listing, mangled, first = sys.argv[1:5]
    first = int(first)
    markers = sys.argv[5:]
    src = open(SRC).read().split("\n")
    marks = [("setup", first)]
    for m in markers:
        for i, l in enumerate(src):
            if m in l and i + 1 >= first:
                marks.append((m[:28], i + 1))
                break
    marks.append(("end", 10 ** 9))
    lst = ncu_lines.load_listing(listing, mangled)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    exe = Counter(); smp = Counter(); stat = Counter()
    for (off, ln, op), r in zip(lst, data):
        name = "helpers"
        for (nm, st), (_, en) in zip(marks, marks[1:]):
            if st <= ln < en:
                name = nm
        exe[name] += int(r[ci["Instructions Executed"]] or 0)
        smp[name] += int(r[ci["# Samples"]] or 0)
        stat[name] += 1
    I = sum(exe.values()); S = sum(smp.values())
    for k, v in exe.most_common():
        print(f"{k:30s} inst {100*v/I:5.1f}% ({v/1e6:9.1f}M)  samp {100*smp[k]/S:5.1f}%  static {stat[k]:5d}")


if __name__ == "__main__":
    main()
