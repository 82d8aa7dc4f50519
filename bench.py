#!/usr/bin/env python
"""bench.py — emosaic hot path on B200: match + compose of BASELINE config 4 (C4).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference's path

Workload (config.workload): C4 = `--mode 1 -s 8`, 100 000 synthetic 8x8 tiles, 4096x4096 synthetic source
-> 32768x32768x3 output, source block rows sharded over the N ranks (strong scaling, no collective in the
loop; the library + source are replicated once by the product's own NCCL broadcast, emo_comm_set_library_dev /
emo_comm_broadcast_dev of libemosaic_cuda.so — torch.distributed only provides the barrier and the max-over-ranks reduction).  A step = one pass (match every
source pixel against the whole library, then composite the output stripe).  The library side is prepared once
per rank before the loop, like the replicated library of the north star: analysis, the (tile, mirror) pixel
store and the 1to1 search index (the GPU stand-in for build_kiddo, tileset.rs:178-190; its build time is
reported as extra.index_build_ms).  The same step with the brute-force scan kernel instead of the index is
measured right after the timed region and reported as extra.match_scan.

metric = matched source px/s for the whole job.  `value` has inputs resident in HBM; `e2e` goes through
the host-pointer C-ABI call emo_mosaic() with pinned host buffers (H2D of the source stripe and D2H of
the composited stripe inside the timed region).  PyTorch is used for device memory, NCCL and barriers only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C4 = dict(T=100_000, ts=8, W=4096, H=4096, N=1)
WORKLOAD = "C4: 1to1 (--mode 1 -s 8), 100k synthetic 8x8 tiles, 4096x4096 source -> 32768x32768x3, row-sharded"
METRIC = "matched source px/s (1to1, match+compose, whole job)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel: str, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` on this workload at N = 1, from
    profiles/traffic.json — written by `tools/summarise_ncu.py traffic` from an `ncu --set full` capture together with the
    commit it was taken at.  None when there is no capture (or at N > 1, where the stripe is a different launch)."""
    if world != 1:
        return None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["kernels"].get(kernel)
        return int(rec["dram_bytes"]) if rec else None
    except Exception:  # noqa: BLE001
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "25", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(pw)))
        return out


def gpu_index_for(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def synth_c4(cfg):
    tiles = np.random.default_rng(1234).integers(0, 256, (cfg["T"], cfg["ts"], cfg["ts"], 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (cfg["H"], cfg["W"], 3), dtype=np.uint8)
    return tiles, src


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm class (bucketed KD-tree nearest_one + row-copy render), OpenMP
# ------------------------------------------------------------------------------------------------
def cpu_reference_setup(cfg):
    import oracle
    tiles, src = synth_c4(cfg)
    colors = oracle.analyse_tiles(tiles, cfg["N"])
    t0 = time.perf_counter()
    kd = oracle.KdTree(colors)  # build_kiddo equivalent: outside the timed region like emo_set_library
    build_s = time.perf_counter() - t0
    return oracle, tiles, src, colors, kd, build_s


def cpu_step(oracle, kd, tiles, src, r0, rows):
    s = src[r0:r0 + rows]
    t0 = time.perf_counter()
    item, dist = kd.match(s)
    out = oracle.render(tiles, item)
    dt = time.perf_counter() - t0
    return dt, rows * src.shape[1], out.nbytes


def cpu_baseline(cfg, target_s=12.0):
    oracle, tiles, src, colors, kd, build_s = cpu_reference_setup(cfg)
    dt, px, _ = cpu_step(oracle, kd, tiles, src, 0, 8)  # calibration (also warms the threads)
    rows = int(max(8, min(512, target_s / max(dt / 8, 1e-6))))
    dt, px, _ = cpu_step(oracle, kd, tiles, src, 8, rows)
    return {"value": px / dt, "unit": "px/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{rows} of {cfg['H']} source rows ({px} px): bucketed KD-tree (leaf 640) nearest_one L1 + row-copy "
                      f"render, OpenMP over block rows; tree build ({build_s:.2f} s) excluded",
            "seconds": dt}


def cpu_baseline_spread(cfg, rows=64):
    """Context for cpu_baseline: the same CPU code on a KD-tree-friendly library (tile colours uniform in the cube
    instead of clustered at 127 +- 9, which is what averaging uniform-random tile pixels produces)."""
    import oracle
    rng = np.random.default_rng(4321)
    colors = rng.integers(0, 256, (cfg["T"], 1, 3), dtype=np.uint8)
    src = np.random.default_rng(5678).integers(0, 256, (rows, cfg["W"], 3), dtype=np.uint8)
    kd = oracle.KdTree(colors)
    kd.match(src[:8])
    t0 = time.perf_counter()
    kd.match(src)
    dt = time.perf_counter() - t0
    return {"value": rows * cfg["W"] / dt, "unit": "px/s", "cores": oracle.num_threads(),
            "sample": f"{rows} source rows, match only, library colours uniform in the colour cube (KD-tree best case)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU arm is meant to use all host threads
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    cfg = C4
    oracle, tiles, src, colors, kd, build_s = cpu_reference_setup(cfg)
    dt, px, _ = cpu_step(oracle, kd, tiles, src, 0, 8)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    rows = int(max(8, min(cfg["H"] // 2, min(budget, 12.0) / max(dt / 8, 1e-6))))
    for w in range(args.warmup):
        cpu_step(oracle, kd, tiles, src, (w * rows) % (cfg["H"] - rows), rows)
    tot, tot_px, tot_out = 0.0, 0, 0
    for k in range(args.steps):
        dt, px, ob = cpu_step(oracle, kd, tiles, src, ((k + args.warmup) * rows) % (cfg["H"] - rows), rows)
        tot += dt; tot_px += px; tot_out += ob
    v = tot_px / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "px/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_rows_per_step": rows, "tiles": cfg["T"], "tile_size": cfg["ts"]},
        "cpu_baseline": {"value": v, "unit": "px/s", "cores": oracle.num_threads(), "kind": "port",
                         "sample": f"{rows} source rows per step ({rows * cfg['W']} px) of the C4 workload; CPU restatement of "
                                   "the reference's algorithm (bucketed KD-tree nearest_one L1 + render), OpenMP, all host "
                                   "threads; the Rust reference cannot be built here (no cargo)"},
        "e2e": {"value": v, "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "composed_output_gbs": tot_out / tot / 1e9,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import emosaic_b200 as emo
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = C4
    T, ts, W, H, N = cfg["T"], cfg["ts"], cfg["W"], cfg["H"], cfg["N"]
    ctx = emo.Context(local_rank)
    info = ctx.device_info()
    comm = None
    if world > 1:   # the product's own communicator: the 128-byte NCCL id travels through torchrun's store
        box = [emo.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init_rank(box[0], rank, world)
        comm = ctx.comm_info()

    # ---- inputs: rank 0 synthesises and analyses; libemosaic_cuda.so replicates library + source over NCCL ----
    src_d = torch.empty(H * W * 3, dtype=torch.uint8, device=dev)
    if rank == 0:
        tiles_h, src_h = synth_c4(cfg)
        tiles_d = torch.from_numpy(tiles_h).reshape(-1).to(dev)
        src_d.copy_(torch.from_numpy(src_h).reshape(-1))
        colors_d = torch.empty(T * N * 3, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.analyse_dev(tiles_d.data_ptr(), T, ts, 1, colors_d.data_ptr())   # library analysis on the GPU
        ctx.comm_set_library_dev(colors_d.data_ptr(), tiles_d.data_ptr(), T, N, ts, root=0)
    else:
        ctx.comm_set_library_dev(0, 0, 0, 0, 0, root=0)
    ctx.comm_broadcast_dev(src_d.data_ptr(), H * W * 3, root=0)
    ctx.sync()
    assert (ctx.T, ctx.N, ctx.ts) == (T, N, ts), "library did not arrive"

    a, b = emo.stripe_bounds(H, world, rank)           # emo_stripe_bounds; dim == 1: block rows == source rows
    Hs = b - a
    src_ptr = src_d.data_ptr() + a * W * 3
    item_d = torch.empty(Hs * W, dtype=torch.int32, device=dev)
    dist_d = torch.empty(Hs * W, dtype=torch.int32, device=dev)
    out_d = torch.empty(Hs * ts * W * ts * 3, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    # the 1to1 search index: built once per library, like the KD-tree of rendering.rs:136 (time reported, not hidden)
    idx_ms = []
    for _ in range(5):
        ctx.timer_start(); ctx.build_index(); idx_ms.append(ctx.timer_stop())
    index_build_ms = float(np.median(idx_ms))

    def step(k=None, base=0, marks=(True, True, True)):
        """match + compose of the stripe.  k = None: one emo_mosaic_dev call (the two launches back to back).  k = slot of the
        CUDA-event marks around the two launches: [mark] emo_match_dev, mark, emo_compose_dev [mark] — the per-kernel durations
        of the roofline; `marks` says which of the three records are issued."""
        if k is None:
            ctx.mosaic_dev(src_ptr, W, Hs, 3, 0, item_d.data_ptr(), dist_d.data_ptr(), out_d.data_ptr())
            return
        if marks[0]: ctx.mark(base + 3 * k)
        ctx.match_dev(src_ptr, W, Hs, item_d.data_ptr(), dist_d.data_ptr())
        if marks[1]: ctx.mark(base + 3 * k + 1)
        ctx.compose_dev(item_d.data_ptr(), 0, W, Hs, 3, 0, out_d.data_ptr())
        if marks[2]: ctx.mark(base + 3 * k + 2)

    # Per-kernel CUDA events inside the timed region, at every N.  An event record between two kernels serialises them (no
    # programmatic overlap) and costs about 6 us of stream time — 15 % of the 80 us step at N = 8 when every launch of a step is
    # bracketed (tools/bench_stripes.py with and without marks).  So only two records sit between kernels: the first timed step is
    # [mark] lookup [mark] compose — its first record follows the region's start event directly — and gives the lookup's duration;
    # the last one is lookup [mark] compose [mark] — its last record is followed directly by the region's stop event — and gives the
    # compose kernel's.  Every other launch boundary of the region is the product's own (emo_mosaic_dev).
    first, last = 0, args.steps - 1
    one = first == last

    def timed_step(k):
        if k == first:
            step(k, 0, (True, True, one))
        elif k == last:
            step(k, 0, (False, True, True))
        else:
            step()

    for _ in range(args.warmup):
        step()
    ctx.sync()
    barrier()
    sampler = ClockSampler(gpu_index_for(local_rank)) if rank == 0 else None
    launches0 = ctx.launch_count()
    ctx.timer_start()
    for k in range(args.steps):
        timed_step(k)
    ms = ctx.timer_stop()
    ctx.sync()
    launches = ctx.launch_count() - launches0
    barrier()
    ms = max_over_ranks(ms)
    match_ms = float(ctx.mark_elapsed(3 * first, 3 * first + 1))
    comp_ms = float(ctx.mark_elapsed(3 * last + 1, 3 * last + 2))
    match_ms_max, comp_ms_max = max_over_ranks(match_ms), max_over_ranks(comp_ms)
    Q_total = H * W
    value = Q_total * args.steps / (ms * 1e-3)

    # ---- the same step with the brute-force scan kernel (north star's match kernel), outside the timed region.
    # The clock sampler keeps running: the timed region above is shorter than one nvidia-smi period.
    ctx.set_match_mode("scan")
    step(); ctx.sync()
    scan_steps = max(3, int(np.ceil(0.35 * world / 0.12)))
    base = 3 * args.steps
    for k in range(scan_steps):
        step(k, base)
    ctx.sync()
    scan_ms = float(np.mean([ctx.mark_elapsed(base + 3 * k, base + 3 * k + 1) for k in range(scan_steps)]))
    scan_ms_max = max_over_ranks(scan_ms)
    ctx.set_match_mode("auto")
    barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = (f"the {args.steps} timed steps ({ms:.1f} ms) plus the {scan_steps} scan-kernel steps that follow "
                            f"({scan_steps * scan_ms:.0f} ms), sampled every 25 ms")

    # ---- e2e: host-pointer C ABI (emo_mosaic), pinned host buffers, copies inside the timed region ---
    src_pin = ctx.host_alloc(Hs * W * 3)
    out_pin = ctx.host_alloc(Hs * ts * W * ts * 3)
    stripe_t = torch.empty(Hs * W * 3, dtype=torch.uint8)
    stripe_t.copy_(src_d[a * W * 3:b * W * 3])
    src_pin[:] = stripe_t.numpy()
    src_img = src_pin.reshape(Hs, W, 3)
    out_img = out_pin.reshape(Hs * ts, W * ts, 3)
    e2e_steps = max(1, min(args.steps, 3))
    ctx.mosaic(src_img, 3, 0, out=out_img, want_maps=False)  # warm-up (allocates the staging buffers)
    barrier()
    ctx.timer_start()
    for _ in range(e2e_steps):
        ctx.mosaic(src_img, 3, 0, out=out_img, want_maps=False)
    e2e_ms = ctx.timer_stop()
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = Q_total * e2e_steps / (e2e_ms * 1e-3)
    # spot-check the e2e output against the resident path (same bytes)
    chk = torch.from_numpy(out_pin[:1 << 20].copy()).to(dev)
    assert bool((chk == out_d[:1 << 20]).all()), "e2e output differs from the device-resident path"
    ctx.host_free(src_pin)
    ctx.host_free(out_pin)

    # ---- the same step on a photo-like source (smooth gradient + noise): neighbouring pixels fall into neighbouring
    # cells of the colour-cube table, so the gather shares cache lines; the uniform-random source of the headline is its worst case
    photo = None
    if not args.no_extras:
        yy, xx = np.mgrid[a:b, 0:W]
        ph = np.stack([xx * 255 // (W - 1), yy * 255 // (H - 1), (xx + yy) * 255 // (W + H - 2)], -1)
        ph = np.clip(ph + np.random.default_rng(99 + rank).integers(-6, 7, ph.shape), 0, 255).astype(np.uint8)
        ph_d = torch.from_numpy(ph.reshape(-1)).to(dev)
        torch.cuda.synchronize()
        pbase = 20000
        for _ in range(3):
            ctx.mosaic_dev(ph_d.data_ptr(), W, Hs, 3, 0, item_d.data_ptr(), dist_d.data_ptr(), out_d.data_ptr())
        ctx.sync()
        barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            ctx.mosaic_dev(ph_d.data_ptr(), W, Hs, 3, 0, item_d.data_ptr(), dist_d.data_ptr(), out_d.data_ptr())
        p_ms = max_over_ranks(ctx.timer_stop())
        for k in range(3):
            ctx.mark(pbase + 2 * k)
            ctx.match_dev(ph_d.data_ptr(), W, Hs, item_d.data_ptr(), dist_d.data_ptr())
            ctx.mark(pbase + 2 * k + 1)
        ctx.sync()
        p_match = max_over_ranks(float(np.mean([ctx.mark_elapsed(pbase + 2 * k, pbase + 2 * k + 1) for k in range(3)])))
        photo = {"value": Q_total * args.steps / (p_ms * 1e-3), "unit": "px/s", "ms_per_step": p_ms / args.steps, "match_ms": p_match,
                 "source": "smooth RGB gradient + uniform noise of +-6 per channel, same library"}
        del ph_d
        barrier()

    # ---- what the host side can absorb: every rank drains a buffer of its stripe's size to pinned memory at the same
    # time (one cudaMemcpyAsync per 64 MB piece, like emo_mosaic), and rank 0 once alone.  e2e is bound by these figures.
    host = None
    try:        # measurement-only library (tools/libemosaic_probe.so); without it the line simply carries no ceiling / pipe rates
        from tools import probe
        probe.load()
        have_probe = 1.0
    except Exception:  # noqa: BLE001
        probe, have_probe = None, 0.0
    have_probe = -max_over_ranks(-have_probe) == 1.0      # every rank or none: the section below has barriers
    if not args.no_extras and have_probe:
        stripe_out = Hs * ts * W * ts * 3
        reps = 3
        # every rank starts its timed copies at the same wall-clock instant (the ranks share the box's clock): allocation and
        # the warm-up pass of 8 x 400 MB of pinned memory take different times per process
        barrier()
        t_all, lead = None, 4.0
        for attempt in range(2):
            box = [(time.time_ns() + int(lead * 1e9) if world > 1 else 0) if rank == 0 else None]   # one rank: nothing to wait for
            if world > 1:
                dist.broadcast_object_list(box, src=0)
            ok = 1.0
            try:
                _, per = probe.host_copy([local_rank], stripe_out, 64 << 20, "d2h", reps, start_unix_ns=box[0])
                mine = stripe_out * reps / (per[0] * 1e9)
            except RuntimeError:
                ok, mine = 0.0, 0.0
            barrier()
            if -max_over_ranks(-ok) == 1.0:          # min over ranks: everyone started on time
                t_all = max_over_ranks(mine)           # slowest rank's time for its passes
                break
            lead *= 3
        solo = None
        if rank == 0:
            try:
                _, per0 = probe.host_copy([local_rank], stripe_out, 64 << 20, "d2h", reps)
                solo = per0[0]
            except RuntimeError:
                solo = None
        barrier()
        host = {"d2h_all_ranks_gbs": (H * ts * W * ts * 3 * reps / t_all / 1e9) if t_all else None, "d2h_one_rank_alone_gbs": solo,
                "bytes_per_rank": stripe_out, "piece": 64 << 20,
                "note": "pinned cudaHostAlloc buffers, every rank copying a buffer of its stripe's size device->host, all ranks starting "
                        "at the same wall-clock instant (tools/probe/hostcopy.cu); aggregate = whole-image bytes / slowest rank's time"}

    # ---- context for the strong-scaling number: the same step with a whole 4096 x 4096 source per rank (weak scaling) ----
    weak = None
    if world > 1 and not args.no_extras:
        item_w = torch.empty(H * W, dtype=torch.int32, device=dev)
        dist_w = torch.empty(H * W, dtype=torch.int32, device=dev)
        out_w = torch.empty(H * ts * W * ts * 3, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()

        def wstep():
            ctx.match_dev(src_d.data_ptr(), W, H, item_w.data_ptr(), dist_w.data_ptr())
            ctx.compose_dev(item_w.data_ptr(), 0, W, H, 3, 0, out_w.data_ptr())

        for _ in range(args.warmup):
            wstep()
        ctx.sync()
        barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            wstep()
        wms = max_over_ranks(ctx.timer_stop())
        barrier()
        weak = {"value": world * Q_total * args.steps / (wms * 1e-3), "unit": "px/s", "ms_per_step": wms / args.steps,
                "note": "every rank matches + composes its own full 4096x4096 source (per-GPU work fixed); not the headline"}
        del item_w, dist_w, out_w
        torch.cuda.empty_cache()

    hbm_peak, peak_src = peaks()
    c3_sharded = c3_across_ranks(ctx, emo, torch, dev, world, rank, max_over_ranks, barrier) if world > 1 and not args.no_extras else None

    # ---- one process driving all N GPUs (emo_group_mosaic: stripes copied straight into ONE pinned host image) next to
    # the N-process e2e above; rank 0 runs it while the other ranks wait at the barrier.
    group_e2e = None
    if world > 1 and not args.no_extras:
        barrier()
        if rank == 0:
            try:
                group_e2e = group_e2e_record(emo, ctx, world, cfg, tiles_h, colors_d.cpu().numpy().reshape(T, N, 3), src_h)
            except Exception as e:  # noqa: BLE001
                group_e2e = {"error": str(e)}
        barrier()

    # ---- pipe rates measured in this run (tools/libemosaic_probe.so; rank 0) ---------------------
    probes = None
    if rank == 0 and have_probe:
        try:
            probes = {"imad": probe.probe_int_pipe(local_rank, 0), "vabsdiff4": probe.probe_int_pipe(local_rank, 1),
                      "vimnmx3": probe.probe_int_pipe(local_rank, 2)}
        except RuntimeError:
            probes = None
    # ---- C2 (BASELINE configs[1], 4to1): first-class record at every N, row-sharded like the headline ----
    c2 = None
    if not args.no_extras:
        try:
            c2 = c2_record(ctx, emo, torch, dev, world, rank, args, max_over_ranks, barrier, probes["vabsdiff4"] if probes else 0.0)
        except Exception as e:  # noqa: BLE001
            c2 = {"error": str(e)}
    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel of the step: compose_tile_kernel<8> (HBM writes) -----------
        out_bytes = Hs * ts * W * ts * 3
        comp_bytes = out_bytes + Hs * W * 4 + T * ts * ts * 3  # §8(d): output + item map + library once
        roofline = {
            "kernel": "compose_tile_kernel<8>", "bound": "hbm",
            "achieved": comp_bytes / (comp_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": comp_bytes / (comp_ms * 1e-3) / 1e9 / hbm_peak,
            "traffic": measured_traffic("compose_tile_kernel", world),
            "peak_source": peak_src, "ms_per_launch": comp_ms, "share_of_step": comp_ms / (comp_ms + match_ms),
            "note": "achieved = (output stripe + item map + tile library once) bytes / CUDA-event time of the compose "
                    "launch (the last timed step, rank 0); the peak is the measured read+write copy figure, which a "
                    "write-mostly stream can exceed by a few per cent; traffic = dram bytes of one launch from the ncu capture "
                    "recorded in profiles/traffic.json (null when no capture of this build exists)",
        }
        # the index lookup: 3 B of source in, 8 B of item/dist out per block, plus one gather from the L2-resident table
        look_bytes = Hs * W * 11
        sad = probes["vabsdiff4"] if probes else None
        D = 3 * N
        L = T if N == 1 else 2 * T
        pairs_per_launch = (Hs * W) * L
        pairs_per_s = pairs_per_launch / (scan_ms * 1e-3)
        step_ms = ms / args.steps
        extra = {
            "match_ms": match_ms_max, "compose_ms": comp_ms_max, "index_build_ms": index_build_ms,
            # rendering.rs:136 builds the KD-tree inside render_nto1, i.e. once per render: the per-render figure pays the
            # index build (the GPU stand-in for that tree) in every step
            "value_per_render": Q_total / ((step_ms + index_build_ms) * 1e-3),
            "composed_output_gbs": (H * ts * W * ts * 3) / (comp_ms_max * 1e-3) / 1e9,
            "roofline_match_index": {"kernel": "match_index16_kernel<2>", "bound": "hbm", "achieved": look_bytes / (match_ms * 1e-3) / 1e9,
                                     "peak": hbm_peak, "unit": "GB/s", "frac": look_bytes / (match_ms * 1e-3) / 1e9 / hbm_peak,
                                     "traffic": measured_traffic("match_index16_kernel", world), "ms_per_launch": match_ms,
                                     "note": "algorithmic bytes = 11 B per block (3 B of source in, 8 B of item + dist out); the 32 MiB "
                                             "compact table is gathered from L2; the duration is of an event-bracketed launch (the first of "
                                             "the timed steps), which includes ~5 us of launch latency"},
            "match_scan": {
                "kernel": "match_kernel<1,8,256>", "bound": "int32-pipe (VABSDIFF4)", "ms_per_launch": scan_ms_max, "steps": scan_steps,
                "matched_px_per_s": Q_total / ((scan_ms_max + comp_ms_max) * 1e-3),
                "achieved": pairs_per_s / 1e12, "peak": sad / 1e12 if sad else None, "unit": "T pairs/s",
                "frac": pairs_per_s / sad if sad else None,
                "traffic": measured_traffic("match_kernel_c4", world),
                "note": "the same step with EMO_MATCH_SCAN (the north star's brute-force kernel): achieved = (query, candidate) "
                        "pairs per second of the scan launch; peak = VABSDIFF4 issue rate measured in this run "
                        "(tools/probe/probe.cu): one VABSDIFF4.U8.ACC per pair is the floor for a 3-byte vector; the ncu "
                        "capture in profiles/ shows the ALU pipe 97.8 % active (VABSDIFF4 + VIMNMX3 share it)",
                "algorithmic_int_ops_per_s": 2 * D * pairs_per_s,
                "probe_thread_inst_per_s": probes,
            },
        }
        if photo is not None:
            extra["c4_photo_like_source"] = photo
        if comm is not None:
            extra["comm"] = dict(comm, library_replication="emo_comm_set_library_dev (ncclBroadcast inside libemosaic_cuda.so)")
        if c2 is not None:
            extra["c2_4to1"] = c2
        if world == 1 and not args.no_extras:
            extra.update(extras(ctx, torch, dev, hbm_peak, peak_src))
        if c3_sharded is not None:
            extra["c3_analysis_sharded"] = c3_sharded
        if weak is not None:
            extra["weak_scaling"] = weak
        if group_e2e is not None:
            extra["e2e_single_process_group"] = group_e2e
        cpu = cpu_baseline(cfg) if world == 1 and not args.no_cpu else None
        if cpu is not None:
            extra["cpu_baseline_spread_library"] = cpu_baseline_spread(cfg)
        d2h_bytes = H * ts * W * ts * 3
        e2e = {"value": e2e_value, "unit": "px/s", "h2d_bytes_per_step": H * W * 3, "d2h_bytes_per_step": d2h_bytes,
               "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "api": "emo_mosaic (host pointers, pinned), one call per rank on its stripe",
               "d2h_gbs": d2h_bytes / (e2e_ms / e2e_steps * 1e-3) / 1e9}
        if host is not None:
            e2e["host_d2h_ceiling_gbs"] = host["d2h_all_ranks_gbs"]
            e2e["frac_of_host_ceiling"] = e2e["d2h_gbs"] / host["d2h_all_ranks_gbs"] if host["d2h_all_ranks_gbs"] else None
            e2e["host_ceiling"] = host
        line = {
            "metric": METRIC, "value": value, "unit": "px/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tiles": T, "tile_size": ts, "source": [H, W], "mode": "1to1",
                       "match": "colour-cube index lookup (exact; built once per library, extra.index_build_ms; per-render figure "
                                "in extra.value_per_render); brute-force scan of the same step in extra.match_scan",
                       "parallelism": f"row-stripes x{world}", "rows_per_rank": Hs,
                       "l2": "no explicit flush: every step writes a 3.2 GB/N output stripe and re-reads 50 MB/N of source, "
                             "far more than the 126 MB L2", "gpu": info["name"]},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": e2e,
            "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)
    ctx.close()


def group_e2e_record(emo, ctx, world, cfg, tiles_h, colors_h, src_h):
    """The whole C4 image through ONE process: emo_group_mosaic over `world` GPUs, host buffers pinned, every GPU copying its
    stripe to its row offset of one output image.  Timed with CUDA events on every member's stream (max over members)."""
    T, ts, W, H = cfg["T"], cfg["ts"], cfg["W"], cfg["H"]
    g = emo.Group(list(range(world)))
    try:
        g.set_library(colors_h, tiles_h)
        src_pin = ctx.host_alloc(H * W * 3)
        out_pin = ctx.host_alloc(H * ts * W * ts * 3)
        src_pin[:] = src_h.reshape(-1)
        src_img, out_img = src_pin.reshape(H, W, 3), out_pin.reshape(H * ts, W * ts, 3)
        g.mosaic(src_img, 3, 0, out=out_img, want_maps=False)      # warm-up: staging buffers, index build on every GPU
        steps = 3
        for m in g.members:
            m.timer_start()
        for _ in range(steps):
            g.mosaic(src_img, 3, 0, out=out_img, want_maps=False)
        ms = max(m.timer_stop() for m in g.members)
        # same bytes as one GPU: the first and the last 8 block rows against rank 0's own ctx (same library resident)
        top, _, _ = ctx.mosaic(src_img[:8], 3, 0, want_maps=False)
        bot, _, _ = ctx.mosaic(src_img[H - 8:], 3, 0, want_maps=False)
        ok = bool((out_img[:8 * ts] == top).all() and (out_img[(H - 8) * ts:] == bot).all())
        rec = {"value": H * W * steps / (ms * 1e-3), "unit": "px/s", "ms_per_step": ms / steps, "gpus": world,
               "d2h_gbs": H * ts * W * ts * 3 / (ms / steps * 1e-3) / 1e9, "api": "emo_group_mosaic (one process, one worker thread per GPU)",
               "same_bytes_as_one_gpu": ok}
        ctx.host_free(src_pin)
        ctx.host_free(out_pin)
        return rec
    finally:
        g.close()


def c2_record(ctx, emo, torch, dev, world, rank, args, max_over_ranks, barrier, sad_rate):
    """C2 = BASELINE configs[1]: 4to1 (--mode 2 -s 16), 10k synthetic tiles, 1024x1024 source -> 8192x8192x3; block rows
    sharded over the ranks like the headline, library replicated from rank 0 by emo_comm_set_library.  `value` resident,
    `e2e` through emo_mosaic with pinned host buffers; the scan kernel against the VABSDIFF4 issue rate."""
    T2, ts2, S2, dim = 10_000, 16, 1024, 2
    if rank == 0:
        tiles = np.random.default_rng(1234).integers(0, 256, (T2, ts2, ts2, 3), dtype=np.uint8)
        ctx.comm_set_library(ctx.analyse_tiles(tiles, dim), tiles, root=0)
    else:
        ctx.comm_set_library(None, None, root=0)
    src = np.random.default_rng(5678).integers(0, 256, (S2, S2, 3), dtype=np.uint8)
    bh = bw = S2 // dim
    a, b = emo.stripe_bounds(bh, world, rank)
    rows = b - a
    stripe = np.ascontiguousarray(src[a * dim:b * dim])
    src_d = torch.from_numpy(stripe.reshape(-1)).to(dev)
    item_d = torch.empty(max(rows * bw, 1), dtype=torch.int32, device=dev)
    dist_d = torch.empty(max(rows * bw, 1), dtype=torch.int32, device=dev)
    out_d = torch.empty(max(rows * ts2 * bw * ts2 * 3, 1), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    base = 40000

    def step(k=None):
        if rows == 0:
            return
        if k is None:
            ctx.mosaic_dev(src_d.data_ptr(), S2, rows * dim, 3, 0, item_d.data_ptr(), dist_d.data_ptr(), out_d.data_ptr())
            return
        ctx.mark(base + 3 * k)
        ctx.match_dev(src_d.data_ptr(), S2, rows * dim, item_d.data_ptr(), dist_d.data_ptr())
        ctx.mark(base + 3 * k + 1)
        ctx.compose_dev(item_d.data_ptr(), 0, S2, rows * dim, 3, 0, out_d.data_ptr())
        ctx.mark(base + 3 * k + 2)

    marked = sorted({0, args.steps - 1})      # per-kernel events on two of the timed steps, like the headline
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.sync()
    barrier()
    ctx.timer_start()
    for k in range(args.steps):
        step(k if k in marked else None)
    ms = ctx.timer_stop()
    barrier()
    ms = max_over_ranks(ms)
    m_ms = float(np.mean([ctx.mark_elapsed(base + 3 * k, base + 3 * k + 1) for k in marked])) if rows else 0.0
    c_ms = float(np.mean([ctx.mark_elapsed(base + 3 * k + 1, base + 3 * k + 2) for k in marked])) if rows else 0.0
    m_ms_max, c_ms_max = max_over_ranks(m_ms), max_over_ranks(c_ms)
    # e2e: host stripe in, host stripe out
    e2e_steps = 3
    src_pin = ctx.host_alloc(max(stripe.nbytes, 1))
    out_pin = ctx.host_alloc(max(rows * ts2 * bw * ts2 * 3, 1))
    src_pin[:stripe.nbytes] = stripe.reshape(-1)
    if rows:
        s_img = src_pin[:stripe.nbytes].reshape(rows * dim, S2, 3)
        o_img = out_pin.reshape(rows * ts2, bw * ts2, 3)
        ctx.mosaic(s_img, 3, 0, out=o_img, want_maps=False)
    barrier()
    ctx.timer_start()
    for _ in range(e2e_steps):
        if rows:
            ctx.mosaic(s_img, 3, 0, out=o_img, want_maps=False)
    e_ms = ctx.timer_stop()
    barrier()
    e_ms = max_over_ranks(e_ms)
    same = True
    if rows:
        same = bool((torch.from_numpy(out_pin[:1 << 20].copy()).to(dev) == out_d[:1 << 20]).all())
    ctx.host_free(src_pin)
    ctx.host_free(out_pin)
    pairs = rows * bw * 2 * T2                               # this rank's (block, candidate) pairs per launch
    rec = {
        "workload": "C2: 4to1 (--mode 2 -s 16), 10k synthetic 16x16 tiles, 1024x1024 source -> 8192x8192x3, row-sharded",
        "metric": "matched source px/s (4to1, match+compose, whole job)", "value": S2 * S2 * args.steps / (ms * 1e-3), "unit": "px/s",
        "n_gpus": world, "ms_per_step": ms / args.steps, "match_ms": m_ms_max, "compose_ms": c_ms_max,
        "composed_output_gbs": bh * ts2 * bw * ts2 * 3 / (c_ms_max * 1e-3) / 1e9 if c_ms_max else None,
        "e2e": {"value": S2 * S2 * e2e_steps / (e_ms * 1e-3), "unit": "px/s", "ms_per_step": e_ms / e2e_steps,
                "h2d_bytes_per_step": S2 * S2 * 3, "d2h_bytes_per_step": bh * ts2 * bw * ts2 * 3, "api": "emo_mosaic (host pointers, pinned)",
                "same_bytes_as_resident_path": same},
    }
    if sad_rate and m_ms:
        rec["roofline"] = {"kernel": "match_kernel<3,...> (scan, D = 12)", "bound": "int32-pipe (VABSDIFF4)", "unit": "T abs-diff-words/s",
                           "achieved": 3 * pairs / (m_ms * 1e-3) / 1e12, "peak": sad_rate / 1e12, "frac": 3 * pairs / (m_ms * 1e-3) / sad_rate,
                           "traffic": measured_traffic("match_kernel_c2", world), "ms_per_launch": m_ms,
                           "note": "3 VABSDIFF4 per (block, candidate) pair (12 bytes) on rank 0's stripe; peak = VABSDIFF4 issue rate "
                                   "measured in this run"}
    return rec


def c3_across_ranks(ctx, emo, torch, dev, world, rank, max_over_ranks, barrier):
    """C3 (analysis cache build, 1M tiles of 64x64, fused 1to1+4to1) sharded over the ranks: every rank analyses
    its contiguous range of tiles, emo_comm_allgather_analysis_dev (ncclAllGather inside libemosaic_cuda.so) assembles [T,3]
    and [T,12] on every rank.  Kernel and collectives run on the ctx stream, timed with its CUDA events, max over ranks.
    Outside the timed region of the headline number."""
    T3, ts3 = 1_000_000, 64
    a, b = emo.stripe_bounds(T3, world, rank)
    n = b - a
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    tiles = torch.randint(0, 256, (n * ts3 * ts3 * 3,), dtype=torch.uint8, device=dev, generator=g)
    o1 = torch.empty(n * 3, dtype=torch.uint8, device=dev)
    o4 = torch.empty(n * 12, dtype=torch.uint8, device=dev)
    f1 = torch.empty(T3 * 3, dtype=torch.uint8, device=dev)
    f4 = torch.empty(T3 * 12, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    times, k_times = [], []
    for rep in range(4):
        barrier()
        ctx.mark(50000)
        ctx.analyse_fused_dev(tiles.data_ptr(), n, ts3, o1.data_ptr(), o4.data_ptr())
        ctx.mark(50001)
        ctx.comm_allgather_analysis_dev(o1.data_ptr(), T3, 3, f1.data_ptr())
        ctx.comm_allgather_analysis_dev(o4.data_ptr(), T3, 12, f4.data_ptr())
        ctx.mark(50002)
        ctx.sync()
        if rep:                                   # first pass warms NCCL up
            times.append(max_over_ranks(ctx.mark_elapsed(50000, 50002)))
            k_times.append(max_over_ranks(ctx.mark_elapsed(50000, 50001)))
    same = bool((f1[a * 3:b * 3] == o1).all()) and bool((f4[a * 12:b * 12] == o4).all())
    ms, kms = float(np.median(times)), float(np.median(k_times))
    del tiles, o1, o4, f1, f4
    torch.cuda.empty_cache()
    return {"tiles": T3, "ranks": world, "ms": ms, "kernel_ms": kms, "allgather_ms": ms - kms, "tiles_per_s": T3 / (ms * 1e-3),
            "input_gbs_whole_job": T3 * 12303 / (ms * 1e-3) / 1e9, "own_range_intact_after_gather": same,
            "collective": "emo_comm_allgather_analysis_dev (ncclAllGather, libemosaic_cuda.so)"}


def extras(ctx, torch, dev, hbm_peak, peak_src):
    """Other BASELINE configs, device-resident, outside the timed region of the headline number (N=1 only)."""
    ex = {}

    def timeit(fn, reps=5):
        fn(); ctx.sync()
        ts_ = []
        for _ in range(reps):
            ctx.timer_start(); fn(); ts_.append(ctx.timer_stop())
        return float(np.median(ts_))

    # C3: analysis cache build, 1M tiles of 64x64 (12.29 GB read), fused 1to1+4to1
    T3 = 1_000_000
    try:
        g = torch.Generator(device=dev); g.manual_seed(1234)
        tiles = torch.randint(0, 256, (T3 * 64 * 64 * 3,), dtype=torch.uint8, device=dev, generator=g)
        o1 = torch.empty(T3 * 3, dtype=torch.uint8, device=dev)
        o4 = torch.empty(T3 * 12, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ms = timeit(lambda: ctx.analyse_fused_dev(tiles.data_ptr(), T3, 64, o1.data_ptr(), o4.data_ptr()))
        gbs = T3 * 12303 / (ms * 1e-3) / 1e9
        ex["c3_analysis"] = {"tiles": T3, "ms": ms, "tiles_per_s": T3 / (ms * 1e-3),
                             "roofline": {"kernel": "analyse_fast_kernel<64,1,1>", "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                                          "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": measured_traffic("analyse_fast_kernel", 1),
                                          "peak_source": peak_src, "bytes_per_tile": 12303}}
        del tiles, o1, o4
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        ex["c3_analysis"] = {"error": str(e)}

    # C5: 1to1 + tint 0.5, ts 32, 1024x1024 source -> 32768x32768x4 RGBA (4.29 GB written)
    try:
        T5, ts5, S5 = 4096, 32, 1024
        rng = np.random.default_rng(1234)
        tiles5 = torch.from_numpy(rng.integers(0, 256, (T5 * ts5 * ts5 * 3,), dtype=np.uint8)).to(dev)
        src5 = torch.from_numpy(np.random.default_rng(5678).integers(0, 256, (S5 * S5 * 3,), dtype=np.uint8)).to(dev)
        col5 = torch.empty(T5 * 3, dtype=torch.uint8, device=dev)
        item5 = torch.empty(S5 * S5, dtype=torch.int32, device=dev)
        dist5 = torch.empty(S5 * S5, dtype=torch.int32, device=dev)
        out5 = torch.empty(S5 * ts5 * S5 * ts5 * 4, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.analyse_dev(tiles5.data_ptr(), T5, ts5, 1, col5.data_ptr())
        ctx.set_library_dev(col5.data_ptr(), tiles5.data_ptr(), T5, 1, ts5)
        m_ms = timeit(lambda: ctx.match_dev(src5.data_ptr(), S5, S5, item5.data_ptr(), dist5.data_ptr()))
        c_ms = timeit(lambda: ctx.compose_dev(item5.data_ptr(), src5.data_ptr(), S5, S5, 4, 127, out5.data_ptr()))
        r_ms = timeit(lambda: ctx.compose_dev(item5.data_ptr(), 0, S5, S5, 3, 0, out5.data_ptr()))
        b5 = S5 * ts5 * S5 * ts5 * 4 + S5 * S5 * 4 + S5 * S5 * 3 + T5 * ts5 * ts5 * 3
        ex["c5_tint"] = {"match_ms": m_ms, "compose_tint_ms": c_ms, "compose_rgb_ms": r_ms,
                         "matched_px_per_s": S5 * S5 / (m_ms * 1e-3),
                         "roofline": {"kernel": "compose_tint_kernel<0>", "bound": "hbm", "achieved": b5 / (c_ms * 1e-3) / 1e9,
                                      "peak": hbm_peak, "unit": "GB/s", "frac": b5 / (c_ms * 1e-3) / 1e9 / hbm_peak,
                                      "traffic": measured_traffic("compose_tint_kernel", 1), "peak_source": peak_src}}
        del tiles5, src5, out5
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        ex["c5_tint"] = {"error": str(e)}

    # Lanczos3 resize (image 0.25.2 imageops::resize, bit-exact): the source step of main.rs:567-595 on a C4-sized image that is
    # not divisible (4098 x 4097 -> 4096 x 4096) and tile preparation (tiles/utils.rs:188-189) of 64 photos of 2048^2 -> 64^2.
    # Bound: the exactness contract forbids FMA and fixes the tap order, so the vertical pass costs 2 FP32 instructions per tap
    # (6 taps per source byte when shrinking) — an FP32-pipe bound, reported as algorithmic bytes/s next to the HBM peak.
    try:
        rs = {}
        for name, n, h, w, nh, nw in (("source_4098x4097_to_4096", 1, 4097, 4098, 4096, 4096), ("tiles_64x2048sq_to_64", 64, 2048, 2048, 64, 64)):
            g = torch.Generator(device=dev); g.manual_seed(77)
            imgs = torch.randint(0, 256, (n * h * w * 3,), dtype=torch.uint8, device=dev, generator=g)
            outr = torch.empty(n * nh * nw * 3, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ms = timeit(lambda: ctx.resize_dev(imgs.data_ptr(), n, w, h, None, nw, nh, outr.data_ptr()))
            byt = n * (h * w * 3 + nh * nw * 3)
            rs[name] = {"ms": ms, "algorithmic_gbs": byt / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": byt / (ms * 1e-3) / 1e9 / hbm_peak,
                        "source_mpx_per_s": n * h * w / (ms * 1e-3) / 1e6}
            del imgs, outr
            torch.cuda.empty_cache()
        ex["resize_lanczos3"] = rs
    except Exception as e:  # noqa: BLE001
        ex["resize_lanczos3"] = {"error": str(e)}
    return ex


def emit(line: dict):
    """Exactly one JSON line on the real stdout (library banners such as NCCL's go to stderr, see main)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    # NCCL / torchrun print banners on stdout; keep stdout for the single JSON line the driver parses
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the C2/C3/C5 side measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
