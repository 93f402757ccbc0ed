"""profiles/traffic.json from ncu reports: per workload key (bench.py::workload_key) the DRAM bytes of ONE launch of the
transport kernel, the packets of that launch, and the utilisation figures of the same capture.

    python scripts/make_traffic.py key=report.ncu-rep:packets[:summary.json] ...

Every report is also summarised to profiles/<summary.json> (scripts/ncu_summary.py's metric list) when a name is given."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    """`rep`: an .ncu-rep, or the csv that `ncu -i <rep> --page raw --csv` printed (what travels back from the GPU box)"""
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(txt.splitlines()) if len(r) > 10]
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def num(d, k):
    unit, v = d[k]
    v = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(unit, 1.0)
    return v * scale


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    for arg in sys.argv[1:]:
        key, rest = arg.split("=", 1)
        parts = rest.split(":")
        rep, packets = parts[0], int(float(parts[1]))
        d = raw(rep)
        b = num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
        out[key] = {
            "bytes": b, "packets": packets, "dram_bytes_per_packet": b / packets,
            "kernel": d["Kernel Name"][1], "kernel_ms": num(d, "gpu__time_duration.sum") / (1e6 if d["gpu__time_duration.sum"][0] in ("ns", "nsecond") else 1.0),
            "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_instruction": num(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "occupancy_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": num(d, "launch__registers_per_thread"),
            "l2_hit_pct": num(d, "lts__t_sector_hit_rate.pct"),
            "dram_pct_of_peak": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "source": "profiles/" + (parts[2] if len(parts) > 2 else os.path.basename(rep)) + " (dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full)",
        }
        print(key, json.dumps(out[key]))
    json.dump(out, open(path, "w"), indent=1)


main()
