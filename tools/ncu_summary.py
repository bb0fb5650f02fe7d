"""Key metrics of an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "sm__inst_executed_pipe_xu",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor", "sm__pipe_xu_cycles_active",
        "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled", "tensor"]


def main(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"=== {r[ki][:100]}")
        for j, h in enumerate(hdr):
            if any(k in h for k in list(KEYS) + list(extra)):
                print(f"   {h:95s} {r[j]:>18s} {units[j]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
