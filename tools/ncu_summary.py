"""Summarise an ncu report (--page raw --csv) into the handful of numbers the profiles/ notes quote."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("## %s" % d["Kernel Name"].split("(")[0])
        for k in KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], u.get(k, "")))
        st = {k: float(v) for k, v in d.items() if "issue_stalled" in k and k.endswith("ratio") and v not in ("", "n/a")}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print("  top stalls: " + ", ".join("%s=%.2f" % (k.split("issue_stalled_")[1].split("_per")[0], v) for k, v in top))


if __name__ == "__main__":
    main(sys.argv[1])
