"""Summarise .ncu-rep captures: one block of counters per distinct kernel (first launch of each).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [...] > profiles/rNx_ncu_summary.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print("# %s: empty" % rep)
            continue
        hdr, units = rows[0], rows[1]
        kcol = hdr.index("Kernel Name")
        seen = set()
        print("# %s" % rep)
        for r in rows[2:]:
            name = r[kcol]
            if name in seen:
                continue
            seen.add(name)
            print("----- %s" % name)
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print("  %-86s %s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    main()
