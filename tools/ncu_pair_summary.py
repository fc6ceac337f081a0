"""Text summary of an `ncu --set full --import-source on` report of the E-step kernel pair (what profiles/*_ncu_summary.txt
hold): raw metrics, an excerpt of the details page, stall reasons, hottest CUDA lines.  usage:
    python tools/ncu_pair_summary.py gpurun_out/prof_X.ncu-rep > profiles/X_estep_pair_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__icc_request_hit_rate.pct",
       "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
       "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
       "sm__inst_executed.sum.per_cycle_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
       "sm__warps_active.avg.per_cycle_active", "launch__shared_mem_per_block_dynamic", "lts__t_sectors.sum",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
DETAILS = ["DRAM Throughput", "Duration", "Compute (SM) Throughput", "Executed Ipc Active", "L1/TEX Hit Rate", "L2 Hit Rate",
           "Issued Warp Per Scheduler", "No Eligible", "Eligible Warps Per Scheduler", "Warp Cycles Per Issued Instruction",
           "Registers Per Thread", "Theoretical Active Warps per SM", "Achieved Active Warps Per SM", "FP64", "fused"]


def run(*a):
    return subprocess.run(["ncu", "-i", *a], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    print("## raw metrics")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d["Kernel Name"])
        for k in RAW:
            if k in d:
                print(f"    {k:80s} {d[k]} {units[hdr.index(k)]}")
        st = sorted(((float(d[k].replace(',', '')), k) for k in hdr
                     if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and d[k]),
                    reverse=True)
        print("    stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, k in st[:8]))
    print("\n## details page (excerpt)")
    for line in run(rep, "--page", "details").split("\n"):
        if "Context 1" in line or any(k in line for k in DETAILS):
            print(line.rstrip()[:118])
    print("\n## hottest CUDA source lines by stall samples (tools/ncu_src.py)")
    src = run(rep, "--page", "source", "--print-source", "cuda,sass", "--csv")
    open("/root/repo/gpurun_out/_src_tmp.csv", "w").write(src)
    print(subprocess.run([sys.executable, __file__.rsplit("/", 1)[0] + "/ncu_src.py", "/root/repo/gpurun_out/_src_tmp.csv", "45"],
                         capture_output=True, text=True).stdout)


if __name__ == "__main__":
    main()
