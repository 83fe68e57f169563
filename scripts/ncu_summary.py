"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export: one line per profiled launch (duration, DRAM bytes read + written,
tensor-pipe and SM throughput, registers) and the per-launch average DRAM traffic that bench.py reports as `roofline.traffic`.
Usage: python scripts/ncu_summary.py raw.csv [--json out.json] [--kernel regex]"""
import argparse
import csv
import json
import re

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json", default=None)
    ap.add_argument("--kernel", default=".")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=True):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        return v * UNIT.get(units[i], 1.0) if scale else v

    out = []
    for r in data:
        name = r[col["Kernel Name"]]
        if not re.search(a.kernel, name):
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        out.append({
            "kernel": name.split("(")[0], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
            "us": val(r, "gpu__time_duration.sum"), "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_bytes": (rd or 0.0) + (wr or 0.0),
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
            "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
            "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "registers": val(r, "launch__registers_per_thread", False),
        })
    for o in out:
        print(json.dumps(o))
    if out:
        summ = {"launches": len(out), "avg_us": sum(o["us"] for o in out) / len(out),
                "avg_dram_bytes_per_launch": sum(o["dram_bytes"] for o in out) / len(out), "rows": out}
        print(json.dumps({k: v for k, v in summ.items() if k != "rows"}))
        if a.json:
            json.dump(summ, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
