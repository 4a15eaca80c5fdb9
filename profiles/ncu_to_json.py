#!/usr/bin/env python
"""Fold `ncu --set full` captures (read here, no GPU needed) into profiles/ncu_metrics.json, the tracked file
bench.py reads its ncu-only numbers from (DRAM traffic per launch, FP64 pipe utilisation): never constants in
bench.py, always stamped with the commit the captured library was built from.

  python profiles/ncu_to_json.py --commit $(git rev-parse --short HEAD) --command "<what was profiled>" \
         gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]

One entry per kernel (keyed by its name without template and argument lists); when a kernel appears in several
launches the last one wins, so pass the captures of the steady state.
"""
import argparse
import csv
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PICK = {
    "duration_ns": "gpu__time_duration.sum",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "registers": "launch__registers_per_thread",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9,
              "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="+")
    ap.add_argument("--commit", required=True)
    ap.add_argument("--command", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ncu_metrics.json"))
    a = ap.parse_args()
    kernels = {}
    for rep in a.reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = re.sub(r"<.*", "", r[hdr.index("Kernel Name")].replace("void ", "")).split("(")[0].strip()
            full = r[hdr.index("Kernel Name")]
            e = {"kernel_name": full, "capture": os.path.basename(rep)}
            for key, metric in PICK.items():
                if metric in hdr:
                    i = hdr.index(metric)
                    try:
                        e[key] = float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
                    except ValueError:
                        pass
            if "dram_read" in e and "dram_write" in e:
                e["dram_bytes"] = e["dram_read"] + e["dram_write"]
            kernels[name] = e
    with open(a.out, "w") as f:
        json.dump({"commit": a.commit, "command": a.command, "kernels": kernels}, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps({k: {x: v.get(x) for x in ("duration_ns", "fp64_pipe_pct", "dram_bytes")} for k, v in kernels.items()}, indent=1))


if __name__ == "__main__":
    main()
