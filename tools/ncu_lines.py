"""Map an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) to CUDA source lines using
an nvdisasm -g -c listing of the same cubin; aggregate stall samples / executed instructions per line.

usage: python tools/ncu_lines.py src.csv kpl2.sass '_ZN3stm12estep_kernelILi2ELi5E' [topN]"""
import csv
import re
import sys
from collections import defaultdict


def load_listing(path, mangled):
    lines = open(path).read().split("\n")
    start = None
    for i, l in enumerate(lines):
        if l.startswith(".text.") and mangled in l:
            start = i
            break
    assert start is not None, "function not found in listing"
    cur = None
    pending = []
    out = []  # (offset, line, opcode)
    for l in lines[start + 1:]:
        if l.startswith("\t.section") or (l.startswith(".text.") and mangled not in l):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            pending.append(m.groups())
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            if pending:
                # outermost location: walk the inline chain; keep the last estep_kernel.cuh kernel-body line
                cand = None
                for f, ln, f2, ln2 in pending:
                    for ff, ll in ((f, ln), (f2, ln2)):
                        if ff and ff.endswith("estep_kernel.cuh") and int(ll) >= 240:
                            cand = int(ll)
                if cand is None:
                    f, ln, f2, ln2 = pending[-1]
                    cand = int(ln2) if ln2 else int(ln)
                cur = cand
                pending = []
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    src_csv, listing, mangled = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lst = load_listing(listing, mangled)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    off2line = {o: (ln, op) for o, ln, op in lst}
    per = defaultdict(lambda: [0, 0, defaultdict(int)])
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot_s = tot_i = 0
    missing = 0
    for r in data:
        off = int(r[0], 16) - base
        ln, op = off2line.get(off, (None, None))
        if ln is None:
            missing += 1
        s = int(r[ci["# Samples"]])
        n = int(r[ci["Instructions Executed"]])
        per[ln][0] += s
        per[ln][1] += n
        for c in stall_cols:
            v = int(r[ci[c]])
            if v:
                per[ln][2][c] += v
        tot_s += s
        tot_i += n
    print(f"total samples {tot_s}, warp instructions {tot_i/1e6:.1f}M, unmapped rows {missing}/{len(data)}")
    src = open("/root/repo/strutopy_b200/csrc/estep_kernel.cuh").read().split("\n")
    for ln, (s, n, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
        top = ", ".join(f"{k[6:]}:{v*100//max(s,1)}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        text = src[ln - 1].strip()[:90] if ln and ln <= len(src) else "?"
        print(f"L{ln!s:>5} {s/tot_s*100:5.1f}% samp {n/tot_i*100:5.1f}% inst | {top:45s} | {text}")


if __name__ == "__main__":
    main()
