#!/usr/bin/env python
"""Turns the raw ncu output of a round into the tracked summaries under profiles/ (run here, on the CPU box, with ncu).

    python profiles/summarize_ncu.py r02 gpurun_out/r02_launches_cfgB.csv gpurun_out/r02_render.ncu-rep

  <tag>_launches_cfgB.csv / .txt   launch list of `bench.py --steps 2 --warmup 3` (ncu --metrics gpu__time_duration.sum
                                   --clock-control none): per-kernel mean time and share of the step
  <tag>_prof_render_summary.csv    ncu --set full raw metrics of surfel_render_fwd / surfel_render_bwd (selected rows)
  <tag>_prof_bwd_sass_hotloop.txt  SASS of surfel_render_bwd with per-instruction execution counts, stall samples and
                                   shared-memory wavefronts (instructions executed >= 0.5 M times)
  traffic.json                     per-launch DRAM bytes / warp instructions that bench.py copies into roofline.traffic
                                   and roofline.side_bound
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    shutil.copy(path, os.path.join(HERE, f"{tag}_launches_cfgB.csv"))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4], []).append(float(r[14]))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(HERE, f"{tag}_launches_cfgB.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 "
                "--no-train --no-extras --no-cpu-baseline (cold-cache, serialised: shares, not absolutes, compare with bench.py)\n")
        for k, v in agg.items():
            f.write(f"{k[:70]:70s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:9.1f} us  share={sum(v) / total * 100:5.1f}%\n")


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def render(tag, rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    traffic = {"_comment": f"per-launch figures from the ncu --set full capture of `python tests/gpu_profile.py steps` (cfg-B: 2M surfels "
                           f"seed 0, 1600x1060), profiles/{tag}_prof_render_summary.csv: dram_bytes = dram__bytes_read.sum + "
                           "dram__bytes_write.sum, inst = smsp__inst_executed.sum (warp instructions; deterministic for this workload up to "
                           "mbarrier spin counts), issue_active_pct = smsp__issue_active.avg.pct_of_peak_sustained_active. bench.py copies "
                           "dram_bytes into roofline.traffic and uses inst for the issue-rate side bound."}
    with open(os.path.join(HERE, f"{tag}_prof_render_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")][:60] for r in rows[2:]])
        stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
        for k in KEEP + stall:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in rows[2:]])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = "render_bwd" if "bwd" in d["Kernel Name"] else "render_fwd"
        unit = dict(zip(hdr, units))["dram__bytes_read.sum"]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[unit]
        traffic[name] = {"dram_bytes": int((float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * scale),
                         "inst": int(float(d["smsp__inst_executed.sum"])),
                         "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                         "lsu_wavefront_pct": float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
                         "source": f"profiles/{tag}_prof_render_summary.csv"}
    with open(os.path.join(HERE, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    # SASS hot loop of the backward
    src = ncu_csv(rep, "source", ["--print-source", "sass"])
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    b = [b for b in blocks if "bwd" in b["name"]][0]
    h = b["rows"][0]
    c = {k: i for i, k in enumerate(h)}
    with open(os.path.join(HERE, f"{tag}_prof_bwd_sass_hotloop.txt"), "w") as f:
        f.write(f"# {b['name']}\n# ncu --set full --import-source on, --page source --print-source sass; instructions executed >= 0.5 M times\n"
                "# addr | executed (M warp-instructions) | avg active threads | stall samples | smem wavefronts (M) | SASS\n")
        for r in b["rows"][1:]:
            if len(r) < len(h):
                continue
            ex = float(r[c["Instructions Executed"]] or 0) / 1e6
            if ex >= 0.5:
                f.write(f"{r[c['Address']][-5:]} {ex:8.2f} {r[c['Avg. Threads Executed']]:>3s} {r[c['# Samples']]:>6s} "
                        f"{float(r[c['L1 Wavefronts Shared']] or 0) / 1e6:7.2f}  {r[c['Source']]}\n")


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    launches(tag, lcsv)
    render(tag, rep)
    print("written", sorted(p for p in os.listdir(HERE) if p.startswith(tag) or p == "traffic.json"))
