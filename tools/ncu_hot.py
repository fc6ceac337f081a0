"""Per-source-line table of an ncu source-page CSV for one kernel: samples, executed instructions, top stall
reasons.  usage: python tools/ncu_hot.py src.csv listing.sass mangled_prefix [topN]"""
import csv
import sys
from collections import Counter, defaultdict

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines  # noqa: E402

SRC = "/root/repo/strutopy_b200/csrc/estep_kernel.cuh"


def main():
    src_csv, listing, mangled = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lst = ncu_lines.load_listing(listing, mangled)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    assert len(data) == len(lst), (len(data), len(lst))
    sc = [c for c in hdr if c.startswith("stall_") and "Not" not in c]
    smp = Counter(); exe = Counter(); st = defaultdict(Counter); tot = Counter()
    for (off, ln, op), r in zip(lst, data):
        smp[ln] += int(r[ci["# Samples"]] or 0)
        exe[ln] += int(r[ci["Instructions Executed"]] or 0)
        for c in sc:
            v = int(r[ci[c]] or 0)
            st[ln][c[6:]] += v
            tot[c[6:]] += v
    S = sum(smp.values()) or 1
    I = sum(exe.values()) or 1
    T = sum(tot.values()) or 1
    print("stalls:", {k: round(100 * v / T, 1) for k, v in tot.most_common(8)})
    print(f"warp instructions {I/1e9:.2f} G, static {len(lst)} ({len(lst)*16/1024:.0f} KB)")
    src = open(SRC).read().split("\n")
    for ln, n in smp.most_common(topn):
        top = ", ".join(f"{k}:{100*v/max(1,sum(st[ln].values())):.0f}%" for k, v in st[ln].most_common(3))
        print(f"L{ln:5d} {100*n/S:5.1f}% samp {100*exe[ln]/I:5.1f}% inst | {top:45s} | {src[ln-1].strip()[:80]}")


if __name__ == "__main__":
    main()
