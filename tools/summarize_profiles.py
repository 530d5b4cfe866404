"""Digest the `<name>_raw.csv` files of tools/capture_profiles.sh (one ncu --set
full capture each, `--page raw --csv`) into one JSON: duration, DRAM bytes,
issue-slot / pipe utilisation, registers, executed instructions.
usage: python tools/summarize_profiles.py <dir with *_raw.csv> > summary.json"""
import csv
import glob
import json
import os
import sys

UNITS = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0,
         "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}
FIELDS = {
    "duration_ms": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
}


def number(text, unit):
    try:
        value = float(text.replace(",", ""))
    except ValueError:
        return None
    return value * UNITS.get(unit, 1.0)


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(sys.argv[1], "*_raw.csv"))):
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(hdr)}
        entry = {"kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else None}
        for key, metric in FIELDS.items():
            i = col.get(metric)
            entry[key] = number(vals[i], units[i]) if i is not None else None
        if entry["dram_read_bytes"] is not None and entry["dram_write_bytes"] is not None:
            entry["dram_bytes"] = entry["dram_read_bytes"] + entry["dram_write_bytes"]
        out[os.path.basename(path)[:-len("_raw.csv")]] = entry
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
