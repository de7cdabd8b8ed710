"""Condenses one .ncu-rep (ncu --set full --import-source on) into a text summary for profiles/.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt

Sections: headline metrics (details page), warp-stall breakdown and DRAM/L2 bytes (raw page), and the hottest
CUDA source lines with lanes-per-instruction (source page).  Needs the `ncu` CLI (present in the build image).
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["Duration", "SM Frequency", "Compute (SM) Throughput", "Memory Throughput", "DRAM Throughput", "L1/TEX Cache Throughput",
        "L2 Cache Throughput", "Executed Ipc Active", "Issue Slots Busy", "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Busy", "Max Bandwidth",
        "No Eligible", "Eligible Warps Per Scheduler", "Active Warps Per Scheduler", "Warp Cycles Per Issued Instruction",
        "Avg. Active Threads Per Warp", "Avg. Not Predicated Off Threads Per Warp", "Registers Per Thread", "Grid Size", "Block Size",
        "Theoretical Occupancy", "Achieved Occupancy", "Block Limit Registers", "Local Memory Spilling Requests", "Branch Efficiency"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = page(rep, "details")
    h = rows[0]
    print("# %s" % rep.split("/")[-1])
    kernel = None
    for r in rows[1:]:
        d = dict(zip(h, r))
        if kernel is None:
            kernel = d.get("Kernel Name")
            print("kernel: %s" % kernel)
        if d.get("Metric Name") in KEEP:
            print("  %-45s %14s %s" % (d["Metric Name"], d["Metric Value"], d["Metric Unit"]))
    raw = page(rep, "raw")
    if len(raw) >= 3:
        names, units, vals = raw[0], raw[1], raw[2]
        m = dict(zip(names, zip(units, vals)))
        print("\n## warp stalls (cycles per issued instruction, smsp__average_warps_issue_stalled_*_per_issue_active)")
        stalls = []
        for k, (u, v) in m.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, k in sorted(stalls, reverse=True)[:8]:
            print("  %-28s %8.3f" % (k, v))
        print("\n## memory")
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
                  "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "smsp__inst_executed.sum",
                  "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "gpu__time_duration.sum"):
            if k in m:
                print("  %-55s %16s %s" % (k, m[k][1], m[k][0]))
    src = page(rep, "source", ("--print-source", "sass,cuda"))
    agg = collections.OrderedDict()
    cur_file, hdr = None, None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("Function Name",) or r[0] == "":
            continue
        try:
            ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            d = agg.setdefault((cur_file, r[0], r[1].strip()[:100]), [0, 0, 0])
            d[0] += int(r[ie]); d[1] += int(r[te]); d[2] += int(r[sm])
        except (ValueError, IndexError):
            pass
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[2] for v in agg.values()) or 1
    print("\n## hottest source lines: %% of warp instructions, %% of stall samples, lanes per instruction")
    print("  total warp instructions %d, mean lanes/instruction %.2f" % (ti, sum(v[1] for v in agg.values()) / ti))
    for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] / ti + kv[1][2] / ts))[:top]:
        print("  %5.1f%% inst %5.1f%% smp %5.1f lanes  %s:%s  %s" % (100.0 * v[0] / ti, 100.0 * v[2] / ts, (v[1] / v[0]) if v[0] else 0, k[0], k[1], k[2]))


if __name__ == "__main__":
    main()
