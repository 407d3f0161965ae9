#!/usr/bin/env python
"""Turns the scratch captures of profiles/capture.sh (gpurun_out/<round>_*) into the committed summaries:

  profiles/<round>_launches.csv      per-kernel launch count / total / mean / share of the bench command (ncu launch list;
                                     per-launch times are cold-cache and serialised: use the SHARE, not the absolute)
  profiles/<round>_kernels.csv       key raw metrics of every --set full capture (one row per captured launch)
  profiles/dram_traffic.json         dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel (bench.py reads it
                                     for roofline.traffic)
Usage (in the build container, ncu is on PATH):  python profiles/summarize.py r1
"""
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SCRATCH = os.path.join(ROOT, "gpurun_out")

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
KEEP = [  # (column prefix in the raw page, short name)
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read_B"),
    ("dram__bytes_write.sum", "dram_write_B"),
    ("lts__t_bytes.sum", "l2_bytes_B"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_static", "smem_static_B"),
    ("launch__shared_mem_per_block_dynamic", "smem_dynamic_B"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__warps_active.avg.per_cycle_active", "warps_active_per_sm"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem_throughput_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"<[^()]*>(?=\()", "", name)  # template arguments of the kernel itself
    m = re.search(r"(?:vb::)?(\w+?)(?:_kernel)?\(", name)
    return m.group(1) if m else name


def launches(round_):
    src = os.path.join(SCRATCH, f"{round_}_bench_launches.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, rows = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows:
        k = short(r[ik])
        us = float(r[iv].replace(",", "")) * UNITS.get(r[iu], 1.0)
        a = agg.setdefault(k, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = min(a[2], us)
        a[3] = max(a[3], us)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"{round_}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c3\n")
        f.write("kernel,launches,total_us,mean_us,min_us,max_us,share_of_gpu_time\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / a[0]:.2f},{a[2]:.2f},{a[3]:.2f},{a[1] / tot:.4f}\n")
    print("launch list:", len(rows), "launches,", len(agg), "kernels")


def full_captures(round_):
    reps = sorted(glob.glob(os.path.join(SCRATCH, f"{round_}_*.raw.csv")) or glob.glob(os.path.join(SCRATCH, f"{round_}_*.ncu-rep")))
    out_rows, traffic = [], {}
    tpath = os.path.join(OUT, "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for rep in reps:
        txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = {"capture": os.path.basename(rep), "kernel": short(r[hdr.index("Kernel Name")] + "(")}
            for col, name in KEEP:
                if col in hdr:
                    i = hdr.index(col)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    if name.endswith("_B") or name.endswith("_us"):
                        v *= UNITS.get(units[i], 1.0)
                    rec[name] = v
            out_rows.append(rec)
            if "_batch_" in rec["capture"]:   # one launch serves 64 sequences: kept apart from the per-sequence traffic table
                rec["kernel"] += "@64"
            if "dram_read_B" in rec:
                t = traffic.setdefault(rec["kernel"], {"samples": 0, "dram_bytes_per_launch": 0.0, "round": round_})
                if t.get("round") != round_:
                    t.update(samples=0, dram_bytes_per_launch=0.0, round=round_)
                tot = t["dram_bytes_per_launch"] * t["samples"] + rec["dram_read_B"] + rec.get("dram_write_B", 0.0)
                t["samples"] += 1
                t["dram_bytes_per_launch"] = tot / t["samples"]
    if not out_rows:
        return
    cols = ["capture", "kernel"] + [n for _, n in KEEP]
    with open(os.path.join(OUT, f"{round_}_kernels.csv"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one row per captured launch (profiles/capture.sh)\n")
        w = csv.DictWriter(f, cols, extrasaction="ignore")
        w.writeheader()
        for rec in out_rows:
            w.writerow({k: (f"{v:.4g}" if isinstance(v, float) else v) for k, v in rec.items()})
    # gftt_tail in bench.py = candidates + sort + select; clahe = lut + apply
    groups = {"gftt_tail": ["gftt_candidates", "sort_keys_desc", "gftt_select"], "clahe": ["clahe_lut", "clahe_apply"]}
    for g, members in groups.items():
        if all(m in traffic for m in members):
            traffic[g] = {"samples": min(traffic[m]["samples"] for m in members), "round": round_,
                          "dram_bytes_per_launch": sum(traffic[m]["dram_bytes_per_launch"] for m in members) / len(members)}
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print("full captures:", len(out_rows), "launches from", len(reps), "reports")


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
    launches(rnd)
    full_captures(rnd)
    for f in (f"{rnd}_clocks.csv", f"{rnd}_micro_lat.txt", f"{rnd}_micro_chol_bench.txt", f"{rnd}_micro_eig_bench.txt"):
        src = os.path.join(SCRATCH, f)
        if os.path.exists(src):
            open(os.path.join(OUT, f), "w").write(open(src).read())
