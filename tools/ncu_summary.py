"""Markdown summary of ncu reports (run here, no GPU needed):
    python tools/ncu_summary.py "title" report1.ncu-rep [report2.ncu-rep ...] > profiles/rNN_ncu_summary.md
Per kernel launch: duration, DRAM traffic, occupancy, pipe utilisation, instruction count and
the top warp-stall reasons, read from `ncu -i <rep> --page raw --csv`.
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    title, reports = sys.argv[1], sys.argv[2:]
    print(f"# {title}\n")
    print("`ncu --set full --clock-control none`; per-launch values; ncu times are cold-cache and serialised, "
          "so compare shares, not absolutes.\n")
    for rep in reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"Source: `{rep}` (scratch; the numbers below are what is kept)\n")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            print(f"## `{name[:150]}`\n")
            print("| metric | value |\n|---|---|")
            for m in METRICS:
                if m in ix:
                    v = r[ix[m]]
                    try:
                        v = f"{float(v.replace(',', '')):.6g}"
                    except ValueError:
                        pass
                    print(f"| `{m}` | {v} {units[ix[m]]} |")
            stalls = []
            for h, i in ix.items():
                if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("| top stall reasons (warps per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + " |")
            print()


if __name__ == "__main__":
    main()
