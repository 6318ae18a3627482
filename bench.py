#!/usr/bin/env python
"""bench.py -- rasterizer fwd+bwd Gaussians/s on the BASELINE workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric config, SURVEY.md section 8(d) cfg-B): one synthetic scene of
2,000,000 surfels rendered at 1600x1060 with precomputed colours per rank (N ranks = N
independent VastGaussian-style tiles with different seeds -> weak scaling, no data-path
collective).  A "step" = GaussianRasterizer forward + autograd backward with fixed upstream
gradients through the drop-in Python API (-> C ABI -> sm_100a kernels).

One JSON line on rank 0:
  value  : whole-job Gaussians/s, inputs resident in HBM (CUDA events, max over ranks)
  e2e    : same metric with the step's inputs copied from pinned host memory and the rendered
           colour image copied back inside the timed region
  step_ms: median / p10 / p90 of the K per-step event times (BASELINE.md's reporting protocol)
  evals  : K_eval (per tile 256 x list entries walked before the block exits, reference definition,
           counted by the CPU port on the full frame) and FP32 pair evaluations/s = K_eval / step time
  extras : the other hot-path workloads (cfg-A surfel, 3DGS 1M, plane cfg-4, visible_filter 2M, distCUDA2 1M, SSIM loss, TSDF fusion, mesh extraction), same
           timing for both arms (tests/bench_extras.py)
  train  : BASELINE's "train iters/s" on config 2, with the fused image-space ops and rasterizer-only
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".
--impl reference times the UNMODIFIED reference CUDA rasterizer (oracle/_ref/libref_surfel.so,
compiled from /root/reference by oracle/build_ref.sh) on the same workload; when that library
is absent it falls back to the CPU oracle port on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gs-sr_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

P_DEFAULT, W_DEFAULT, H_DEFAULT = 2_000_000, 1600, 1060
METRIC = "rasterizer fwd+bwd Gaussians/s @ 2M surfels x 1600x1060"
PROF_NAMES = ["preprocess_fwd", "scan", "duplicate", "sort", "build_records", "render_fwd", "render_bwd",
              "preprocess_bwd"]
OWN_KERNELS_PER_STEP = 7  # preprocess_fwd, tile_scan, scatter_keys, sort_build_records, render_fwd, render_bwd, preprocess_bwd


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel, key="dram_bytes"):
    """dram__bytes_read.sum + dram__bytes_write.sum (or another figure) per launch from the committed
    ncu --set full capture of the same workload (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel][key]
    except Exception:
        return None


def issue_side_bound(kernel, kernel_ms, sm_mhz, default_workload):
    """The render kernels are instruction-issue bound, not HBM bound: warp
    instructions per launch (ncu smsp__inst_executed.sum of the committed capture of the default
    workload) / live kernel time, against 148 SMs x 4 schedulers x 1 warp-instruction per clock."""
    inst = ncu_traffic(kernel, "inst")
    if inst is None or not default_workload or not sm_mhz:
        return None
    peak = 148 * 4 * sm_mhz * 1e6
    ach = inst / (kernel_ms * 1e-3)
    return {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": ach / peak,
            "inst_per_launch": inst, "source": "profiles/traffic.json (ncu smsp__inst_executed.sum) / live cudaEvent time"}


def algorithmic_bytes(P, V, R, N):
    """SURVEY 8(d) per-unit figures restated for this repo's layouts (DESIGN.md 'Measurement')."""
    return {
        "preprocess_fwd": P * 40 + V * (64 + 16) + P * 8,
        "render_fwd": R * 96 + N * 76,
        "render_bwd": R * 96 + N * 76 + V * 80,
        "whole_step": P * (40 + 24 + 112) + V * (64 + 12 + 76 + 148 + 44) + R * (12 + 24 * 6 + 8 + 80 + 80)
        + N * (76 + 76),
    }


def pipelined_e2e(step_fn, host, template, out_host, steps, barrier):
    """End-to-end loop: every step's inputs travel host(pinned)->device and its rendered image
    device->host(pinned) INSIDE the timed region.  Copies run on side streams, double buffered,
    so step i+1's H2D and step i-1's D2H overlap step i's kernels (what a prefetching data
    loader does); the first H2D and the last D2H are exposed and counted.  Returns milliseconds."""
    import torch
    cur = torch.cuda.current_stream()
    h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
    # one packed device buffer per stage (views per tensor): the step's inputs travel as ONE cudaMemcpyAsync from the
    # packed pinned host buffer instead of one copy per tensor
    packed_host = host["_packed"]
    staged_packed = [torch.empty(packed_host.shape, dtype=packed_host.dtype, device=cur.device) for _ in range(2)]
    staged = [unpack_views(sp, host["_layout"]) for sp in staged_packed]
    outs = [torch.empty(out_host.shape, dtype=out_host.dtype, device=cur.device) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    rendered = [torch.cuda.Event() for _ in range(2)]
    fetched = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(h2d):
            if i >= 2:
                h2d.wait_event(consumed[b])
            staged_packed[b].copy_(packed_host, non_blocking=True)
            ready[b].record(h2d)

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    prefetch(0)
    for i in range(steps):
        b = i & 1
        if i + 1 < steps:
            prefetch(i + 1)
        cur.wait_event(ready[b])
        if i >= 2:
            cur.wait_event(fetched[b])          # outs[b] has been copied out
        color = step_fn(staged[b])
        outs[b].copy_(color.detach())
        consumed[b].record(cur)
        rendered[b].record(cur)
        with torch.cuda.stream(d2h):
            d2h.wait_event(rendered[b])
            out_host.copy_(outs[b], non_blocking=True)
            fetched[b].record(d2h)
    cur.wait_stream(d2h)
    cur.wait_stream(h2d)
    e1.record(cur)
    barrier()
    return e0.elapsed_time(e1)


def unpack_views(packed, layout):
    """{name: view} of one packed float32 buffer; layout = [(name, offset, shape)]."""
    out = {}
    for name, off, shape in layout:
        n = int(np.prod(shape))
        out[name] = packed[off: off + n].view(shape)
    return out


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated: pages
    are placed on the allocating thread's node (first touch), so every rank's per-step upload then reads local DRAM
    instead of crossing the socket interconnect.  Under torchrun the ranks are otherwise unbound.  Returns the node or None."""
    import torch
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "none reported (single-node host)"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def h2d_copy_only_gbs(host, device, barrier, reps=20):
    """Bandwidth of the packed upload alone (all ranks at once): what the host link / host DRAM can feed this rank."""
    import torch
    dst = torch.empty(host["_packed"].shape, dtype=host["_packed"].dtype, device=device)
    for _ in range(3):
        dst.copy_(host["_packed"], non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(host["_packed"], non_blocking=True)
    e1.record()
    barrier()
    return host["_packed"].numel() * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def build_inputs(P, W, H, seed, device):
    import torch
    import synth
    sc = synth.make_scene(P, W, H, seed=seed)
    gc, go = synth.make_upstream_grads(W, H, seed=seed + 1)
    names = ("means3D", "scales", "rotations", "opacities", "colors")
    arrays = {k: np.ascontiguousarray(getattr(sc, k), dtype=np.float32) for k in names}
    layout, off = [], 0
    for k in names:
        layout.append((k, off, tuple(arrays[k].shape)))
        off += (arrays[k].size + 63) // 64 * 64                     # keep every tensor 256-byte aligned
    packed = torch.empty(off, dtype=torch.float32).pin_memory()
    host = unpack_views(packed, layout)
    for k in names:
        host[k].copy_(torch.from_numpy(arrays[k]))
    dev = {k: v.to(device) for k, v in host.items()}
    host["_packed"], host["_layout"] = packed, layout
    cam = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in
           dict(bg=sc.cam.bg, view=sc.cam.viewmatrix, proj=sc.cam.projmatrix, campos=sc.cam.campos).items()}
    g = (torch.from_numpy(gc).to(device), torch.from_numpy(go).to(device))
    return sc, host, dev, cam, g


def run_ours(args, rank, world, device):
    import torch
    import gsr_b200
    from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from gsr_b200 import shard
    P, W, H = args.P, args.W, args.H
    sc, host, dev, cam, (gct, got) = build_inputs(P, W, H, seed=rank * 17, device=device)
    rs = GaussianRasterizationSettings(H, W, sc.cam.tanfovx, sc.cam.tanfovy, cam["bg"], 1.0, cam["view"], cam["proj"],
                                       0, cam["campos"], False, False)
    rast = GaussianRasterizer(rs)
    state = {}

    def step(inputs):
        leaves = {k: v.detach().requires_grad_(True) for k, v in inputs.items()}
        m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
        color, radii, others = rast(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                    colors_precomp=leaves["colors"], scales=leaves["scales"],
                                    rotations=leaves["rotations"])
        torch.autograd.backward([color, others], [gct, got])
        state["radii"], state["color"], state["grad"] = radii, color, leaves["means3D"].grad
        return color

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts BEFORE the warm-up so that nvidia-smi's own start-up (driver
    # attach, ~0.5 s) is over when the timed region begins; it keeps sampling through both
    # timed regions.
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)
    for _ in range(args.warmup):
        step(dev)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(dev)
        ev[i + 1].record()
    barrier()
    ms_resident = shard.max_over_ranks(ev[0].elapsed_time(ev[-1]), device)
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]

    # end-to-end: H2D of the step's inputs from pinned memory + D2H of the rendered image
    out_host = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
    pipelined_e2e(step, host, dev, out_host, 2, barrier)  # warm the side streams / allocator
    ms_e2e = shard.max_over_ranks(pipelined_e2e(step, host, dev, out_host, args.steps, barrier), device)
    clocks = sampler.stop() if rank == 0 else None
    h2d = host["_packed"].numel() * 4
    d2h = out_host.numel() * out_host.element_size()
    copy_gbs = h2d_copy_only_gbs(host, device, barrier)
    copy_all = [round(m["gbs"], 2) for m in shard.gather_metrics({"gbs": copy_gbs})] if world > 1 else [round(copy_gbs, 2)]

    # per-kernel durations (cudaEvents inside the library, outside the timed regions)
    L = gsr_b200.lib()
    import ctypes
    acc = np.zeros(16)
    nprof = 3
    for _ in range(nprof):
        L.gsr_profile_enable(1)
        step(dev)
        buf = (ctypes.c_float * 16)()
        L.gsr_profile_read(buf)
        acc += np.array(list(buf))
    L.gsr_profile_enable(0)
    kernel_ms = {n: float(acc[i] / nprof) for i, n in enumerate(PROF_NAMES)}
    V = int((state["radii"] > 0).sum().item())
    return dict(ms_resident=ms_resident, ms_e2e=ms_e2e, clocks=clocks, h2d=h2d, d2h=d2h, kernel_ms=kernel_ms, V=V,
                sc=sc, checksum=float(state["color"].double().sum().item()), step_ms=step_ms, copy_gbs=copy_gbs, copy_all=copy_all)


def run_reference_cuda(args, rank, world, device):
    import torch
    from gsr_b200 import shard
    from oracle.refcuda import RefSurfel
    P, W, H = args.P, args.W, args.H
    sc, host, dev, cam, (gct, got) = build_inputs(P, W, H, seed=rank * 17, device=device)
    r = RefSurfel()
    state = {}

    def step(inp):
        color, radii, others, R = r.forward(cam["bg"], cam["view"], cam["proj"], cam["campos"], W, H, sc.cam.tanfovx,
                                            sc.cam.tanfovy, inp["means3D"], inp["opacities"], inp["scales"],
                                            inp["rotations"], colors=inp["colors"])
        r.backward(gct, got)
        state["R"], state["radii"] = R, radii
        return color

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step(dev)
        ev[i + 1].record()
    barrier()
    ms_resident = shard.max_over_ranks(ev[0].elapsed_time(ev[-1]), device)
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    out_host = torch.empty((3, H, W), dtype=torch.float32).pin_memory()
    pipelined_e2e(step, host, dev, out_host, 2, barrier)
    ms_e2e = shard.max_over_ranks(pipelined_e2e(step, host, dev, out_host, args.steps, barrier), device)
    h2d = host["_packed"].numel() * 4
    return dict(ms_resident=ms_resident, ms_e2e=ms_e2e, h2d=h2d, d2h=out_host.numel() * 4, R=int(state["R"]),
                V=int((state["radii"] > 0).sum().item()), step_ms=step_ms, copy_gbs=h2d_copy_only_gbs(host, device, barrier))


def step_stats(step_ms):
    return {"median": float(np.median(step_ms)), "p10": float(np.percentile(step_ms, 10)), "p90": float(np.percentile(step_ms, 90)),
            "n": len(step_ms)}


def cpu_oracle_full(P, W, H, stride=4):
    """CPU port (oracle/liborc.so, OpenMP on all host threads) on the SAME workload: the full frame once (preprocess +
    binning + render fwd+bwd of every tile) and, for comparison with round 1, the 1/stride^2 tile sample with its
    extrapolation rule.  Also returns K_eval of the full frame (reference definition)."""
    import synth
    from oracle.oracle import SurfelOracle
    sc = synth.make_scene(P, W, H, seed=0)
    gc, go = synth.make_upstream_grads(W, H, seed=1)
    o = SurfelOracle()
    t0 = time.perf_counter()
    o.forward(sc.cam, sc.means3D, sc.opacities, sc.scales, sc.rotations, colors=sc.colors, tile_stride=1 << 20)
    t_bin = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.forward(sc.cam, sc.means3D, sc.opacities, sc.scales, sc.rotations, colors=sc.colors, tile_stride=stride)
    o.backward(gc, go, tile_stride=stride)
    t_sample = max(time.perf_counter() - t0 - t_bin, 1e-9)
    t0 = time.perf_counter()
    f = o.forward(sc.cam, sc.means3D, sc.opacities, sc.scales, sc.rotations, colors=sc.colors)
    o.backward(gc, go)
    t_full = time.perf_counter() - t0
    est = t_bin + stride * stride * t_sample
    return dict(value=P / t_full, cores=os.cpu_count(), t_full=t_full, k_eval=int(f["k_eval"]), R=int(f["num_rendered"]),
                sample=f"oracle/liborc.so (OpenMP, {os.cpu_count()} threads) on the full workload once: forward+backward of "
                       f"the {P}-surfel scene, every tile, {t_full:.1f}s (preprocess+binning {t_bin:.1f}s of it); the every-"
                       f"{stride}th-tile sample of round 1 extrapolates to {est:.1f}s ({P / est:.0f} Gaussians/s)")


def mesh_gather_check(rank, world, device):
    """Config 5's one collective (extract_mesh_split.py:58-119): every rank fuses ITS tile's views into the same bounded
    TSDF lattice, the partial volumes are summed onto rank 0 with one NCCL reduce (gsr_b200.tsdf.BoundedTSDFVolume), and
    rank 0 checks the result against fusing all ranks' views itself, then meshes it on its GPU (gsr_b200.mesh: marching
    cubes + cluster filter).  Returns the record on rank 0, None elsewhere."""
    import torch
    from gsr_b200.tsdf import BoundedTSDFVolume
    from tsdf_synth import build_tsdf_case
    c = build_tsdf_case("world_rgb", n=8)
    nv = len(c["projs"])
    grid = dict(origin=(-0.8, -0.8, -0.8), voxel_size=1.6 / 127, dims=(128, 128, 128), sdf_trunc=0.06, depth_trunc=5.0)
    pick = lambda idx: ([torch.from_numpy(c["projs"][i]) for i in idx], [torch.from_numpy(c["depthmaps"][i]) for i in idx],  # noqa: E731
                        [torch.from_numpy(c["rgbmaps"][i]) for i in idx])
    mine = [(2 * rank) % nv, (2 * rank + 1) % nv]                      # two views per tile
    from gsr_b200.mesh import post_process_mesh
    # the first pass pays for what is not the gather: NCCL's lazy channel set-up, cudaMalloc behind the caching allocator
    # (several ms each, and they synchronise the device), module loads; the second pass is timed
    for timed in (False, True):
        vol = BoundedTSDFVolume(with_rgb=True, device=device, **grid).integrate(*pick(mine))
        torch.distributed.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vol.reduce_to(0)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            # ... and meshes the gathered volume where it lies (extract_mesh_split.py:119-128)
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            mesh = post_process_mesh(vol.extract_triangle_mesh(), cluster_to_keep=50)
            m1.record(); torch.cuda.synchronize()
    if rank != 0:
        return None
    every = [v for r in range(world) for v in ((2 * r) % nv, (2 * r + 1) % nv)]
    full = BoundedTSDFVolume(with_rgb=True, device=device, **grid).integrate(*pick(every))
    ref_mesh = post_process_mesh(full.extract_triangle_mesh(), cluster_to_keep=50)
    return {"voxels": 128 ** 3, "views": len(every), "reduce_ms": e0.elapsed_time(e1), "bytes_per_rank": 128 ** 3 * 20,
            "max_abs_err_tsdf": float((vol.tsdf - full.tsdf).abs().max()), "weights_equal": bool(torch.equal(vol.weight, full.weight)),
            "max_abs_err_rgb": float((vol.rgb - full.rgb).abs().max()),
            "mesh_ms": m0.elapsed_time(m1), "mesh_triangles": int(mesh.triangles.shape[0]),
            "mesh_triangles_single_gpu_volume": int(ref_mesh.triangles.shape[0])}


def train_iters(impl):
    """Second half of BASELINE's metric ("train iters/s") on BASELINE config 2 (cfg-A: 2DGS, 100k surfels, 800x800, SH 3):
    the GS-SR-style iteration of tests/train_harness.py (render + post-processing + L1/SSIM/normal/dist losses + backward +
    densification statistics + Adam, loss.item() every step).  `value` = our arm with the drop-in rasterizer AND the fused
    SSIM / post-processing ops; `rasterizer_only` = the same iteration with the reference's torch ops around the drop-in
    rasterizer, which isolates the extension.  The reference arm runs the unmodified reference kernels + torch ops."""
    from train_harness import measure_iters_per_s
    ours = impl == "ours"
    arm = "ours" if ours else "reference"
    v, _ = measure_iters_per_s(arm, 100_000, 800, 800, iters=40, warmup=5, fused_ssim=ours, fused_post=ours)
    out = {"metric": "train iters/s", "value": v, "unit": "iters/s",
           "workload": "2DGS iteration, P=100000 synthetic surfels, 800x800, SH degree 3 (BASELINE config 2), Adam, "
                       "loss.item() per step; tests/train_harness.py",
           "ops": "drop-in rasterizer + fused SSIM + fused post-processing" if ours else
                  "reference CUDA rasterizer + reference torch ops"}
    if ours:
        v2, _ = measure_iters_per_s("ours", 100_000, 800, 800, iters=40, warmup=5, fused_ssim=False, fused_post=False)
        out["rasterizer_only"] = {"value": v2, "unit": "iters/s", "ops": "drop-in rasterizer + the reference's torch ops"}
        try:    # the whole iteration replayed from ONE CUDA graph (the drop-in forward never blocks the stream; the reference's does)
            from train_harness import measure_graph_iters_per_s
            v3, _, overflow = measure_graph_iters_per_s(100_000, 800, 800, iters=200, warmup=5)
            out["cuda_graph"] = {"value": v3, "unit": "iters/s", "capture_overflow": overflow,
                                 "ops": "same iteration, loss kept on the device, recorded once with torch.cuda.graph and replayed "
                                        "(render + losses + backward + statistics + capturable Adam = 1 graph launch per iteration)"}
        except Exception as e:  # noqa: BLE001
            out["cuda_graph"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    else:
        out["cuda_graph"] = {"unavailable": "the reference forward blocks on a cudaMemcpy of num_rendered "
                                            "(S/cuda_rasterizer/rasterizer_impl.cu:282): it cannot be captured"}
    for name, kw in (("config3_scaffold2dgs", dict(P=400_000, W=1600, H=1060, scaffold=True)),
                     ("config4_pgsr", dict(P=1_000_000, W=1600, H=900, pgsr=True)),
                     # the octree scene classes of BASELINE configs 4 / 5: anchors x 5 neural Gaussians behind the
                     # level-of-detail mask and the octree prefilter (P = anchors)
                     ("config4_octree_pgsr", dict(P=200_000, W=1600, H=900, pgsr=True, octree=True)),
                     ("config5_octree_2dgs", dict(P=400_000, W=1600, H=1060, octree=True))):
        try:
            vv, _ = measure_iters_per_s(arm, iters=15, warmup=3, fused_ssim=ours, fused_post=ours, **kw)
            out[name] = {"value": vv, "unit": "iters/s"}
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--P", type=int, default=P_DEFAULT)
    ap.add_argument("--W", type=int, default=W_DEFAULT)
    ap.add_argument("--H", type=int, default=H_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    workload = f"2DGS surfel rasterizer fwd+bwd, P={args.P} synthetic surfels (SURVEY 8(d) cfg-B), {args.W}x{args.H}, " \
               "precomputed colours, one independent scene per rank"
    base = {"metric": METRIC, "unit": "Gaussians/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs + record stream (>850 MB/step) exceed the 126 MB L2; no flush",
                       "parallelism": f"independent tiles x{world}, no data-path collective"}}

    if args.impl == "reference":
        from oracle import refcuda
        if refcuda.available("surfel"):
            r = run_reference_cuda(args, rank, world, device)
            if rank == 0:
                v = args.P * world * args.steps / (r["ms_resident"] * 1e-3)
                ve = args.P * world * args.steps / (r["ms_e2e"] * 1e-3)
                base.update({"impl": "reference", "value": v, "ms_per_step": r["ms_resident"] / args.steps,
                             "e2e": {"value": ve, "unit": "Gaussians/s", "h2d_bytes_per_step": r["h2d"],
                                     "d2h_bytes_per_step": r["d2h"], "numa_node": numa_node,
                                     "h2d_copy_only_gbs_rank0": r["copy_gbs"]},
                             "cpu_baseline": {"value": v, "unit": "Gaussians/s", "cores": 0, "kind": "reference",
                                              "sample": "the reference has no CPU path: its own UNMODIFIED CUDA "
                                                        "kernels (diff-surfel-rasterization compiled for sm_100a "
                                                        "by oracle/build_ref.sh) on the full workload on the GPU"},
                             "stats": {"num_rendered": r["R"], "visible": r["V"]}, "gpu_launches": 0,
                             "step_ms": step_stats(r["step_ms"])})
                if world == 1 and not args.no_train:
                    base["train"] = train_iters("reference")
                if world == 1 and not args.no_extras:
                    import bench_extras
                    base["extras"] = bench_extras.run_all("reference")
                print(json.dumps(base))
        else:
            if rank == 0:
                c = cpu_oracle_full(args.P, args.W, args.H)
                base.update({"impl": "reference", "value": c["value"], "ms_per_step": args.P / c["value"] * 1e3,
                             "n_gpus": 1, "e2e": {"value": c["value"], "unit": "Gaussians/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0},
                             "cpu_baseline": {"value": c["value"], "unit": "Gaussians/s", "cores": c["cores"],
                                              "kind": "port", "sample": c["sample"]}, "gpu_launches": 0})
                print(json.dumps(base))
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    r = run_ours(args, rank, world, device)
    gather = mesh_gather_check(rank, world, device) if world > 1 else None
    if rank == 0:
        v = args.P * world * args.steps / (r["ms_resident"] * 1e-3)
        ve = args.P * world * args.steps / (r["ms_e2e"] * 1e-3)
        peak, peak_src = measured_peak()
        N = args.W * args.H
        # R is read back from the oracle-free path: num_rendered of rank 0's scene
        from diff_surfel_rasterization import last_num_rendered
        R = last_num_rendered()
        ab = algorithmic_bytes(args.P, r["V"], R, N)
        dom = max(("render_fwd", "render_bwd"), key=lambda k: r["kernel_ms"][k])
        achieved = ab[dom] / (r["kernel_ms"][dom] * 1e-3) / 1e9
        base.update({
            "value": v, "ms_per_step": r["ms_resident"] / args.steps,
            "e2e": {"value": ve, "unit": "Gaussians/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "numa_node": numa_node, "h2d_copy_only_gbs_rank0": r["copy_gbs"],
                    "h2d_copy_only_gbs_per_rank": r["copy_all"],   # all ranks uploading at once, nothing else running: the host-link ceiling
                    "h2d_needed_gbs_per_rank": r["h2d"] / (r["ms_resident"] / args.steps * 1e-3) / 1e9,
                    "upload": "one packed pinned buffer per step (1 cudaMemcpyAsync), process bound to the GPU's NUMA node"},
            "gpu_launches": OWN_KERNELS_PER_STEP * args.steps,
            "clocks": r["clocks"],
            "roofline": {"bound": "hbm", "kernel": f"gsr::surfel_{dom}", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(dom), "peak_source": peak_src,
                         "algorithmic_bytes": ab[dom], "kernel_ms": r["kernel_ms"][dom],
                         "whole_step_frac": ab["whole_step"] / (r["ms_resident"] / args.steps * 1e-3) / 1e9 / peak,
                         "side_bound": issue_side_bound(dom, r["kernel_ms"][dom], (r["clocks"] or {}).get("sm_mhz"),
                                                        (args.P, args.W, args.H) == (P_DEFAULT, W_DEFAULT, H_DEFAULT))},
            "kernel_ms": r["kernel_ms"], "step_ms": step_stats(r["step_ms"]), "mesh_gather": gather,
            "stats": {"num_rendered": R, "visible": r["V"], "pixels": N, "checksum": r["checksum"]},
        })
        if world == 1 and not args.no_train:
            base["train"] = train_iters("ours")
        if world == 1 and not args.no_extras:
            import bench_extras
            base["extras"] = bench_extras.run_all("ours")
        if world == 1 and not args.no_cpu_baseline:
            c = cpu_oracle_full(args.P, args.W, args.H)
            base["cpu_baseline"] = {"value": c["value"], "unit": "Gaussians/s", "cores": c["cores"], "kind": "port",
                                    "sample": c["sample"]}
            step_s = r["ms_resident"] / args.steps * 1e-3
            base["evals"] = {"k_eval": c["k_eval"], "evals_per_s": c["k_eval"] / step_s, "unit": "(pixel, splat) evaluations/s",
                             "definition": "K_eval = sum over tiles of 256 x list entries walked before the whole block is done "
                                           "(what the reference kernel evaluates, counted by the CPU port on rank 0's scene); "
                                           "ours skips the culled share of them"}
        print(json.dumps(base))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
