"""Print the key roofline metrics per kernel from an `ncu --page raw --csv` dump."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("----", r[idx["Kernel Name"]][:60])
        for w in WANT:
            if w in idx:
                print(f"  {w} = {r[idx[w]]} {units[idx[w]]}")


if __name__ == "__main__":
    main(sys.argv[1])
