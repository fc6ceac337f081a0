"""Per-phase breakdown of an ncu source-page CSV of stm::estep_kernel (samples, instructions, no_inst
share, static code size), phases located by marker comments in estep_kernel.cuh.
usage: python tools/ncu_phases.py src.csv listing.sass mangled_prefix"""
import csv
import sys
from collections import defaultdict

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import ncu_lines  # noqa: E402

SRC = "/root/repo/strutopy_b200/csrc/estep_kernel.cuh"
MARKERS = [
    ("0 misc helpers", None),
    ("0a dcstep", "static __device__ __noinline__ void dcstep("),
    ("0b cubic/quadmin", "static __device__ __noinline__ double cubicmin("),
    ("0c log", "static __device__ STM_NOINLINE double log_noinline("),
    ("0d exp", "static __device__ STM_NOINLINE double exp_noinline("),
    ("0e logprod inl", "struct LogProd {"),
    ("0f tma/misc", "---- TMA bulk copy + mbarrier (PTX)"),
    ("0g chol_factor", "static __device__ __noinline__ int chol_factor("),
    ("0h inverse_and_nu", "static __device__ __noinline__ void inverse_and_nu("),
    ("a setup+gather", "template <int KPL, int J>"),
    ("b a_k precompute", "a_k = sum_v beta_kv c_v / colsum_v"),
    ("c0 bfgs init", "BFGS (scipy/optimize/_optimize.py:1345-1526) as a warp-uniform"),
    ("c1 eval head", "evaluate f, g at x + alpha p"),
    ("c2 exp+partials", "red[]: cnt, ssum, quad"),
    ("c3 contraction", "data term: sum_v c_v (m + log"),
    ("c4 logprod", "for (int j = 0; j < J; ++j) {\n                            const int v = w0 + lane + 32 * j;"),
    ("c5 reduce+lse+grad", "red[3] = logprod_value(lp);"),
    ("d1 consume INIT/W1", "consume the evaluation"),
    ("d2 consume W2/zoom", "} else if (ls == LS_W2) {"),
    ("d3 start w2/zoom/next", "if (start_w2) {"),
    ("e accept+Hupdate", "if (accept) {"),
    ("f new_iter", "if (new_iter && !done) {"),
    ("g post theta", "post-optimisation: theta"),
    ("g2 colsum+loglik", "colsum_v = sum_k e_k beta_kv and"),
    ("h hessian blocks+phi", "Hessian data term  sum_v b_v b_v'"),
    ("i assemble", "assemble H = data"),
    ("j chol+repair", "PD test + repairs"),
    ("k bound", "// bound (stm.py"),
    ("l nu", "// nu = H^-1 = L^-T L^-1 (stm.py:1052-1066), accumulated"),
]


def main():
    src_csv, listing, mangled = sys.argv[1:4]
    text = open(SRC).read()
    starts = []
    for name, mk in MARKERS:
        if mk is None:
            starts.append((1, name))
            continue
        pos = text.rfind(mk) if name.startswith("l nu") else text.find(mk)
        assert pos >= 0, mk
        starts.append((text.count("\n", 0, pos) + 1, name))
    starts.sort()

    def bucket(line):
        if line is None:
            return "none"
        b = starts[0][1]
        for ln, name in starts:
            if line >= ln:
                b = name
        return b

    lst = ncu_lines.load_listing(listing, mangled)
    off2line = {o: l for o, l, op in lst}
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot_st = defaultdict(int)
    for r in data:
        a = agg[bucket(off2line.get(int(r[0], 16) - base))]
        a[0] += int(r[ci["# Samples"]])
        a[1] += int(r[ci["Instructions Executed"]])
        a[2] += int(r[ci["stall_no_inst"]])
        a[3] += 1
        a[4] += int(r[ci["stall_wait"]])
        for c in stall_cols:
            tot_st[c] += int(r[ci[c]])
    TS = sum(a[0] for a in agg.values())
    TI = sum(a[1] for a in agg.values())
    S = sum(tot_st.values())
    print("stalls:", {k[6:]: round(v / S * 100, 1) for k, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:7]})
    print(f"warp instructions {TI/1e9:.2f} G, static {len(data)} instrs ({len(data)*16/1024:.0f} KB)")
    for k, a in sorted(agg.items()):
        print(f"{k:32s} samp {a[0]/TS*100:5.1f}%  inst {a[1]/TI*100:5.1f}%  no_inst {a[2]/max(a[0],1)*100:3.0f}%  "
              f"wait {a[4]/max(a[0],1)*100:3.0f}%  static {a[3]:5d} ({a[3]*16/1024:4.1f} KB)")


if __name__ == "__main__":
    main()
