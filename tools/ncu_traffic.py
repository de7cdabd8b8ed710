"""profiles/traffic.json from an ncu capture of all k_extend launches of one 1-spp pass.

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum -k regex:k_extend \
      -c <max path length> --csv --log-file gpurun_out/x.csv python bench.py --spp 1 --steps 1 --warmup 0 ... > gpurun_out/x.json
  python tools/ncu_traffic.py gpurun_out/x.csv gpurun_out/x.json

The bench line (x.json) carries roofline.stat_rays_1spp = closest-hit rays of one 1-spp pass; the capture covers exactly
the k_extend launches of such a pass, so bytes / rays is the measured DRAM traffic per ray.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "software-raytracing_b200"))
from pyraylib import device_source_hash  # noqa: E402


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale[unit]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    line = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    h = rows[0]
    ni, ui, vi, ki = h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
    dram = l2 = ms = 0.0
    launches = set()
    for r in rows[1:]:
        launches.add(r[ki])
        if r[ni] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram += to_bytes(r[vi], r[ui])
        elif r[ni] == "lts__t_bytes.sum":
            l2 += to_bytes(r[vi], r[ui])
        elif r[ni] == "gpu__time_duration.sum":
            ms += float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
    rays = line["roofline"]["stat_rays_1spp"]
    name = line["config"]["workload"]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    out[name] = {"dram_bytes_per_ray": dram / rays, "l2_bytes_per_ray": l2 / rays, "rays": rays, "launches": len(launches),
                 "extend_ms_under_ncu": ms, "device_source_hash": device_source_hash(), "source": "ncu %s (dram__bytes_read.sum + dram__bytes_write.sum over the %d k_extend launches of one 1-spp pass)"
                 % (os.path.basename(sys.argv[1]), len(launches))}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out[name]))


if __name__ == "__main__":
    main()
