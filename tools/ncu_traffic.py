#!/usr/bin/env python
"""Add one kernel's per-launch traffic / instruction counters from an ncu --set full capture to profiles/traffic.json (read by bench.py
for roofline.traffic, roofline.l2_traffic and the issue-rate model).

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep KEY profiles/summary.md STEPS_PER_LAUNCH
KEY = <kernel>_<N>cubed_<rays>rays_rec<bytes>, e.g. divq_indep_128cubed_100rays_rec8
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, key, src = sys.argv[1], sys.argv[2], sys.argv[3]
    steps = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))

    def val(k):
        v = float(d[k])
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0}.get(u.get(k, ""), 1.0)
        return v * scale
    e = {"dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum")}
    e["dram_bytes_per_launch"] = e["dram_read"] + e["dram_write"]
    e["l2_bytes_per_launch"] = 32.0 * float(d["lts__t_sectors_srcunit_tex_op_read.sum"])
    e["warp_instructions_per_launch"] = float(d["smsp__inst_executed.sum"])
    e["kernel_ms_under_ncu"] = val("gpu__time_duration.sum") if u.get("gpu__time_duration.sum") == "ms" else float(d["gpu__time_duration.sum"])
    if steps:
        e["steps_per_launch"] = steps
    e["source"] = f"{src} (ncu --set full --clock-control none, one launch)"
    p = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(p)) if os.path.exists(p) else {}
    t[key] = e
    json.dump(t, open(p, "w"), indent=1)
    print(key, json.dumps(e))


if __name__ == "__main__":
    main()
