"""Condense an .ncu-rep into the handful of numbers DESIGN.md / profiles/ cite.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--source N]"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "lts__t_bytes.sum.per_second", "lts__t_sectors_srcunit_tex_op_read.sum"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    hdr, units, rows = raw(path)
    for r in rows:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "?")[:80])
        for h, u in zip(hdr, units):
            if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[h].replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in h and v < 0.05:
                    continue
                print(f"  {h:88s} {v:16.3f} {u}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[0]
        print("source columns:", h[:12])


if __name__ == "__main__":
    main()
