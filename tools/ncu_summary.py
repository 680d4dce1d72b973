"""Summarise ncu reports for profiles/: `python tools/ncu_summary.py key=report.ncu-rep [...]` prints a markdown table and updates
profiles/r2_ncu_spmv.json (read by bench.py for roofline.traffic).  `key` names the operator, e.g. n200_periodic_x1."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}


def read(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = float(r[i]) * SCALE.get(units[i], 1.0)
        res.append(d)
    return res


def main():
    path = os.path.join(ROOT, "profiles", "r2_ncu_spmv.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    print("| key | kernel | launches | duration (us) | DRAM read+write per launch (GB) | regs | warps active % | issue active % | instructions |")
    print("|---|---|---|---|---|---|---|---|---|")
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        rows = read(rep)
        n = len(rows)
        avg = lambda k: sum(r.get(k, 0.0) for r in rows) / n
        traffic = avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum")
        db[key] = {"kernel": rows[0]["kernel"].split("(")[0].strip(), "launches": n, "duration_us": avg("gpu__time_duration.sum") * 1e6,
                   "dram_bytes_per_launch": traffic, "registers": avg("launch__registers_per_thread"),
                   "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                   "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   "instructions": avg("smsp__inst_executed.sum"), "report": os.path.basename(rep), "source": "ncu --set full --clock-control none"}
        e = db[key]
        print("| %s | `%s` | %d | %.1f | %.4f | %d | %.1f | %.1f | %.3g |" % (key, e["kernel"], n, e["duration_us"], traffic / 1e9, e["registers"],
                                                                           e["warps_active_pct"], e["issue_active_pct"], e["instructions"]))
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
