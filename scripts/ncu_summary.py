"""Summarises an ncu report (`ncu --set full ... -o X`) into a small JSON for profiles/: the metrics the roofline
discussion uses.  Usage: python scripts/ncu_summary.py X.ncu-rep out.json "what was captured" ["reading"]"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "l1tex__t_bytes.sum"]


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines()) if r]
    hdr = next(r for r in rows if r[0] == "ID")
    units = rows[rows.index(hdr) + 1]
    vals = rows[-1]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            d[h] = "%s %s" % (v, u) if u else v
        if h == "Kernel Name":
            d["kernel"] = v
    d["_what"] = what
    if len(sys.argv) > 4:
        d["_reading"] = sys.argv[4]
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
