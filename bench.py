#!/usr/bin/env python
"""bench.py -- the hot path's benchmark of record (one JSON line on stdout, rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config 5|4|3|2|1]

Workload (default): BASELINE.json configs[4] -- 1 000 000 synthetic Gaussians (seed 43, sigma = 10^U(-2.6,-2.0)) at
4096x4096, 256 reference tiles per axis, mode-8 semantics (A&S erf, round-to-nearest, alpha from colour.w), lists =
"reference predicate AND 6-sigma bound" (VRT_CUDA_LIST_REFERENCE_BOUND).  The metric is pixel-Gaussian evaluations
(inner terms of src/vrt/rt.h:107-124) per second; a step is one frame: K1 tile/cull + K2/K3 render (+ the NCCL gather of
the row bands to rank 0 when N > 1).  N GPUs split the SAME frame into work-balanced row bands => "scaling": "strong".

  value     terms actually executed by all ranks / max-over-ranks device time of the K timed steps; scene resident in HBM.
  e2e       the same frame through the C ABI with HOST buffers: pinned-host scene -> device (40 MB), tile, render, image ->
            pinned host (67 MB) inside the timed region, every step.
  roofline  FP32-pipe roofline of the render kernel (15 flops per term); the kernel is not HBM- or tensor-
            bound, so "bound" is "fp32": peak = SMs x 128 lanes x 2 x max SM clock, and the FFMA throughput
            measured live on the same GPU is reported beside it.
  cpu_baseline  the reference's own production entry vrt::simd_render_image (mode 8, all host threads) from oracle/_ref,
            timed on a bounded sample of this workload's tiles with the lists the workload defines.

--impl reference times only that CPU path (oracle/_ref; the C restatement if the compiled reference cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION on some boxes) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

CONFIGS = {
    # name: (scene builder args, width, tiles, description)
    5: dict(kind="synthetic", n=1_000_000, seed=43, lo=-2.6, hi=-2.0, width=4096, tiles=256, name="config5: 1M synthetic Gaussians @4096^2, 256 tiles"),
    4: dict(kind="synthetic", n=100_000, seed=42, lo=-2.1, hi=-1.5, width=4096, tiles=256, name="config4: 100k synthetic Gaussians @4096^2, 256 tiles"),
    3: dict(kind="grid", dim=64, width=2048, tiles=16, name="config3: 64x64 grid @2048^2, 16 tiles"),
    1: dict(kind="grid", dim=4, width=256, tiles=16, name="config1: 4x4 grid @256^2, 16 tiles"),
}
FLOPS_PER_TERM = 15.0  # 6 FMA + 2 MUL + 1 ADD after hoisting, A&S variant
BOUND_SIGMAS = 6.0


def build_scene(pkg, cfg):
    if cfg["kind"] == "synthetic":
        return pkg.scenes.synthetic(cfg["n"], cfg["seed"], cfg["lo"], cfg["hi"])
    return pkg.scenes.grid(cfg["dim"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8] or [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = []
        for i, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows), "reasons": reasons}


def reference_sample(pkg, cfg, scene, n_tiles, threads):
    """(lists, plane, origin, tile_px, terms) for a power-of-two sample of reference tiles spread over the frame."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_lists

    W, tiles = cfg["width"], cfg["tiles"]
    tile_px = W // tiles
    cam, origin = pkg.vrt.camera_t.app(W, W)
    # a square block of tiles starting a quarter into the frame: mixes central and oblique rays
    side = int(np.sqrt(n_tiles))
    assert side * side == n_tiles and side <= tiles
    t0 = max(0, min(tiles - side, tiles // 4))
    ids = [(t0 + j) * tiles + (t0 + i) for j in range(side) for i in range(side)]
    lists = cpu_lists.sample_tile_lists(scene, cam.view_matrix, origin, W, W, tiles, ids, BOUND_SIGMAS, use_reference=True)
    plane = np.zeros((3, tile_px, n_tiles * tile_px), np.float32)
    for k, t in enumerate(ids):
        tx, ty = t % tiles, t // tiles
        pts = cpu_lists.plane_points(cam.view_matrix, W, W, np.arange(ty * tile_px, (ty + 1) * tile_px), np.arange(tx * tile_px, (tx + 1) * tile_px))
        plane[:, :, k * tile_px : (k + 1) * tile_px] = pts
    terms = float(sum(5.0 * tile_px * tile_px * len(l) * len(l) for l in lists))
    return [scene[l] for l in lists], plane, origin, tile_px, terms


def time_reference(pkg, cfg, scene, n_tiles, threads, reps):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Ref

    lists, plane, origin, tile_px, terms = reference_sample(pkg, cfg, scene, n_tiles, threads)
    times = []
    if Ref.available():
        kind, simd = "reference", Ref.simd_floats()
        for _ in range(reps):
            ms, _ = Ref.render_tile_strip(lists, tile_px, tile_px, plane, origin, threads, -1)
            times.append(ms * 1e-3)
        what = f"vrt::simd_render_image (mode 8, oracle/_ref, SIMD_FLOATS={simd})"
    else:
        # the compiled reference cannot run on this CPU: time the plain-C restatement (scalar A&S path, all threads)
        kind = "port"
        pix = plane.reshape(3, -1).T
        for _ in range(reps):
            t = time.time()
            for k, l in enumerate(lists):
                sel = pix[[r * plane.shape[2] + c for r in range(tile_px) for c in range(k * tile_px, (k + 1) * tile_px)]]
                d = sel - origin[:3]
                d = np.concatenate([d / np.linalg.norm(d, axis=1, keepdims=True), np.zeros((len(d), 1), np.float32)], 1).astype(np.float32)
                Oracle.radiance(l, origin, d, 1)
            times.append(time.time() - t)
        what = "oracle/vrt_oracle.c scalar radiance (A&S)"
    return dict(kind=kind, what=what, terms=terms, times=times, n_tiles=n_tiles, tile_px=tile_px, mean_list=float(np.mean([len(l) for l in lists])))


def run_reference(args, pkg, cfg, rank):
    if rank != 0:
        return
    scene = build_scene(pkg, cfg)
    threads = os.cpu_count() or 1
    n_tiles = 1024 if cfg["tiles"] >= 32 else (256 if cfg["tiles"] >= 16 else cfg["tiles"] ** 2)
    r = time_reference(pkg, cfg, scene, n_tiles, threads, args.warmup + args.steps)
    timed = r["times"][args.warmup :]
    total = sum(timed)
    value = r["terms"] * len(timed) / total
    sample = f"{r['n_tiles']} reference tiles of {r['tile_px']}x{r['tile_px']} px (mean list {r['mean_list']:.0f}), {r['terms']:.3e} evaluations per step; {r['what']}"
    line = {
        "impl": "reference", "metric": "pixel-Gaussian evaluations/s", "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["name"], "lists": "reference predicate AND 6-sigma bound", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cuda(args, pkg, cfg, rank, world):
    import torch

    V = pkg.vrt
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, tiles = cfg["width"], cfg["tiles"]
    flags = (V.MODE8 & ~V.LIST_MASK) | V.LIST_REFERENCE_BOUND
    cam, origin = V.camera_t.app(W, W)
    r = V.Renderer(local)
    stream = torch.cuda.ExternalStream(r.stream, device=torch.device("cuda", local))

    from vrt_b200 import bands as B  # registered by load_package()

    # ---- scene: rank 0 builds it in pinned host memory; the device copy is broadcast once over NCCL ----
    n = cfg["n"] if cfg["kind"] == "synthetic" else cfg["dim"] ** 2
    host_scene = torch.empty((n, 10), dtype=torch.float32, pin_memory=True)
    if rank == 0:
        host_scene.numpy()[:] = build_scene(pkg, cfg)
    with torch.cuda.stream(stream):
        dev_scene = torch.empty((n, 10), dtype=torch.float32, device="cuda")
        if rank == 0:
            dev_scene.copy_(host_scene, non_blocking=True)
        if world > 1:
            dist.broadcast(dev_scene, 0)
        r.set_gaussians_device(dev_scene.data_ptr(), n)
        image = torch.zeros((W, W), dtype=torch.int32, device="cuda")
        host_image = torch.empty((W, W), dtype=torch.int32, pin_memory=True)
    r.sync()

    # ---- row bands: balanced by the per-row cost K1 reports for the full frame (computed once per camera) ----
    full = r.frame(cam.view_matrix, origin, W, W, flags, (tiles, tiles), BOUND_SIGMAS)
    if world > 1:
        r.tile(full)
        bounds = B.balanced_bands(r, world, W, align=W // tiles)
    else:
        bounds = [0, W]
    rows = (bounds[rank], bounds[rank + 1])
    frame = r.frame(cam.view_matrix, origin, W, W, flags, (tiles, tiles), BOUND_SIGMAS, rows=rows if world > 1 else (0, 0))

    def step(e2e):
        """one frame; returns this rank's stats"""
        with torch.cuda.stream(stream):
            if e2e:
                if rank == 0:
                    dev_scene.copy_(host_scene, non_blocking=True)
                if world > 1:
                    dist.broadcast(dev_scene, 0)
                r.set_gaussians_device(dev_scene.data_ptr(), n)
            r.tile(frame)
            st = r.render_device(frame, image.data_ptr(), 0, want_stats=True)
            if world > 1:
                B.gather_bands(image, bounds, rank, world, dist)
            if e2e and rank == 0:
                host_image.copy_(image, non_blocking=True)
        return st

    def timed(e2e, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        r.sync()
        t0 = time.time()
        ev0.record(stream)
        stats = [step(e2e) for _ in range(steps)]
        ev1.record(stream)
        r.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t1 = time.time()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        terms = torch.tensor([sum(s["terms_executed"] for s in stats), sum(s["terms_listed"] for s in stats)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(terms, op=dist.ReduceOp.SUM)
        return float(ms.item()), terms.tolist(), stats, (t0, t1)

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    ms, (terms_exec, terms_listed), stats, (t0, t1) = timed(False, args.steps)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    step(True)
    ms_e2e, (terms_exec_e2e, _), _, _ = timed(True, args.steps)

    # the same frame in depth-window mode (VRT_CUDA_DEPTH_WINDOW), reported beside the headline, not as it: most listed terms
    # are resolved by the saturation shortcut there, so its strict evals/s (fully evaluated terms only) is lower while the
    # frame is several times faster
    plain_frame = frame
    frame = r.frame(cam.view_matrix, origin, W, W, flags | V.DEPTH_WINDOW, (tiles, tiles), BOUND_SIGMAS, rows=rows if world > 1 else (0, 0))
    step(False)
    ms_win, _, stats_win, _ = timed(False, args.steps)
    win_terms = torch.tensor([sum(s["terms_executed"] for s in stats_win), sum(s["terms_saturated"] for s in stats_win)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(win_terms, op=dist.ReduceOp.SUM)
    win_exec, win_sat = (v / args.steps for v in win_terms.tolist())
    frame = plain_frame

    k2_ms = float(np.mean([s["ms_render"] for s in stats]))
    k1_ms = float(np.mean([s["ms_tile"] for s in stats]))
    k2 = torch.tensor([k2_ms, k1_ms, stats[0]["terms_executed"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(k2, op=dist.ReduceOp.MAX)
    def cleanup():
        # ordered teardown: tensors that live on the context's stream must be released before the stream is destroyed
        # (the pinned-host allocator records an event on every stream a block was used on when the block is freed)
        nonlocal dev_scene, image, host_image, host_scene, k2
        import gc

        r.sync()
        torch.cuda.synchronize()
        dev_scene = image = host_image = host_scene = k2 = None
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        if world > 1:
            dist.destroy_process_group()
        r.close()

    if rank != 0:
        cleanup()
        return

    value = terms_exec / (ms * 1e-3)
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz", (clocks or {}).get("sm_max_mhz") or 1965.0))
    peak_tflops = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    # dominant kernel = k2_render on this rank: its own executed terms / its own CUDA-event duration
    k2_terms = float(stats[0]["terms_executed"])
    achieved = k2_terms * FLOPS_PER_TERM / (float(np.mean([s["ms_render"] for s in stats])) * 1e-3) / 1e12
    ffma, ffma2 = r.fp32_peak(False), r.fp32_peak(True)
    roofline = {
        "bound": "fp32", "kernel": "k2_render", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
        "peak_source": f"nominal: {sm_count} SMs x 128 lanes x 2 x {sm_max_mhz:.0f} MHz (max SM clock, MEASURED_PEAKS.json); MEASURED_PEAKS has no fp32 figure",
        "measured_ffma_tflops": ffma, "measured_ffma2_tflops": ffma2, "frac_of_measured_ffma": achieved / max(ffma, ffma2),
        "flops_per_term": FLOPS_PER_TERM, "terms_per_launch": k2_terms, "ms_per_launch": float(np.mean([s["ms_render"] for s in stats])),
        # dram__bytes_read.sum + dram__bytes_write.sum of one k2_render launch of this command under `ncu --set full`
        # (profiles/r01_k2_render_final_ncu_raw.csv: 358.1 MB + 63.8 MB); algorithmic: 48 MB records + 172 MB lists + 67 MB image
        "traffic": 421867776 if args.config == 5 and world == 1 else None, "traffic_unit": "bytes per launch (ncu, round-1 capture)",
    }
    line = {
        "metric": "pixel-Gaussian evaluations/s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "lists": "reference predicate AND 6-sigma bound per 8x4-pixel cell", "semantics": "mode 8 (A&S erf, nearest, alpha from w)",
                   "l2": "inputs (scene 40 MB + lists) are rebuilt every step by K0/K1; the working set exceeds L2", "bands": bounds,
                   "terms_listed_per_frame": terms_listed / args.steps, "terms_executed_per_frame": terms_exec / args.steps,
                   "ms_tile_max_rank": float(k2[1].item()), "ms_render_max_rank": float(k2[0].item()),
                   "depth_window_mode": {"ms_per_step": ms_win / args.steps, "terms_evaluated_per_frame": win_exec, "terms_saturated_per_frame": win_sat,
                                         "strict_evals_per_s": win_exec / (ms_win / args.steps * 1e-3),
                                         "resolved_evals_per_s": (win_exec + win_sat) / (ms_win / args.steps * 1e-3),
                                         "note": "opt-in flag VRT_CUDA_DEPTH_WINDOW; same image (fp32 sums reordered); not the headline"}},
        "e2e": {"value": terms_exec_e2e / (ms_e2e * 1e-3), "unit": "evals/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(n * 40), "d2h_bytes_per_step": int(W * W * 4)},
        "gpu_launches": int(sum(s["n_launches"] for s in stats)),
        "clocks": clocks, "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_tiles = 1024 if cfg["tiles"] >= 32 else (256 if cfg["tiles"] >= 16 else cfg["tiles"] ** 2)
        cb = time_reference(pkg, cfg, build_scene(pkg, cfg), n_tiles, threads, 2)
        best = min(cb["times"])
        line["cpu_baseline"] = {"value": cb["terms"] / best, "unit": "evals/s", "cores": threads, "kind": cb["kind"],
                                "sample": f"{cb['n_tiles']} reference tiles of {cb['tile_px']}x{cb['tile_px']} px (mean list {cb['mean_list']:.0f}), {cb['terms']:.3e} evaluations, best of 2; {cb['what']}"}
    print(json.dumps(line), flush=True)
    cleanup()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import __graft_entry__ as ge

    pkg = ge.load_package()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, pkg, cfg, rank)
    else:
        run_cuda(args, pkg, cfg, rank, world)


if __name__ == "__main__":
    main()
