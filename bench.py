#!/usr/bin/env python
"""bench.py -- stereo frames/s of the ZenSLAM front-end hot path (detect + describe + match + KLT).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): synthetic 752x480 stereo sequence; per stereo frame 2 optical-flow
pyramids, 2 grid FAST detections (cells 16x16, threshold 10) + ORB, 1 stereo Hamming kNN-ratio match and
4 forward+backward KLT pairs (31x31 window, 4 levels, 99 its / eps 1e-3, min-eig 1e-4, FB gate 1 px).
A step = one batch of B consecutive stereo frames through the batched front-end.

  value  whole-job stereo frames/s, inputs resident in HBM (device->device staging + every kernel timed)
  e2e    same metric through the host-buffer C-ABI call: H2D of the frames from pinned memory and D2H of
         every result inside the timed region
  roofline      the dominant kernel (fused forward+backward KLT) against the measured HBM peak
  cpu_baseline  the reference's call pattern driven through cv2 on the host cores (oracle/cv2_ref.py)

--impl reference times that CPU pipeline itself (all host cores, one frame per worker per step).
Multi-GPU: the path shards by sequence -- every rank runs its own sequence on its own GPU, no collective on
the data path ("weak" scaling); timing is barrier + synchronize bracketed, max over ranks.
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
sys.path.insert(0, ROOT)

METRIC, UNIT = "stereo_frames_per_s", "stereo frames/s"
RATIO = 0.8
# name -> frame size, detector cells / FAST threshold / algorithm, KLT window / max_level / FB threshold, default batch
CONFIGS = {
    "C2":       dict(w=752,  h=480,  cell=(16, 16), thr=10, alg="GRID",          win=(31, 31), ml=3, klt_thr=1.0, batch=128),
    "C4":       dict(w=1280, h=1024, cell=(32, 32), thr=10, alg="GRID",          win=(31, 31), ml=3, klt_thr=1.0, batch=64),
    "C5":       dict(w=3840, h=2160, cell=(32, 32), thr=10, alg="GRID",          win=(31, 31), ml=3, klt_thr=1.0, batch=16),
    # the only configuration the reference ships (zenslam_options/options/tumvi.yaml:38-47) at the TUM-VI frame size
    "TUMVI":    dict(w=1024, h=1024, cell=(64, 64), thr=1,  alg="PARALLEL_GRID", win=(63, 63), ml=4, klt_thr=2.0, batch=128),
    "TUMVI752": dict(w=752,  h=480,  cell=(64, 64), thr=1,  alg="PARALLEL_GRID", win=(63, 63), ml=4, klt_thr=2.0, batch=128),
}


def set_config(name):
    global CFG, W, H, CELL, FAST_T, WIN, MAX_LEVEL, KLT_THR, WORKLOAD
    CFG = dict(CONFIGS[name], name=name)
    W, H, CELL, FAST_T = CFG["w"], CFG["h"], CFG["cell"], CFG["thr"]
    WIN, MAX_LEVEL, KLT_THR = CFG["win"], CFG["ml"], CFG["klt_thr"]
    WORKLOAD = ("%s: %dx%d synthetic stereo sequence; per frame 2 pyramids + 2 grid FAST (%dx%d cells, thr %d)%s + ORB "
                "+ 1 stereo Hamming kNN-ratio + 4 fwd/bwd KLT pairs (%dx%d, %d levels, 99 its, eps 1e-3)"
                % (name, W, H, CELL[0], CELL[1], FAST_T, " + cornerSubPix" if CFG["alg"] == "PARALLEL_GRID" else "",
                   WIN[0], WIN[1], fe_levels(W, H)))
    return CFG


def cpu_options():
    from oracle import FrontendOptions
    return FrontendOptions(CELL, FAST_T, WIN, MAX_LEVEL, KLT_THR, RATIO, CFG["alg"] == "PARALLEL_GRID")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_sequence(frames, seed):
    from zenslam_b200 import synthetic as syn
    seq, _ = syn.stereo_sequence(W, H, frames, seed, subpixel=True)
    return seq


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU pipeline (reference arm / cpu_baseline)
# --------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """one process: `frames` stereo frames of the reference call pattern through cv2, 1 OpenCV thread"""
    seed, frames, cfg_name = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import cv2_ref
    set_config(cfg_name)          # spawned workers import this module afresh: the configuration travels with the task
    opts = cpu_options()
    seq = make_sequence(frames + 1, seed)
    # frame 0 primes the "previous frame" state (not timed)
    prev = cv2_ref.stereo_frame(seq[0, 0], seq[0, 1], seq[0, 0], seq[0, 1], np.zeros((0, 2), np.float32),
                                np.zeros((0, 2), np.float32), opts)
    t0 = time.perf_counter()
    for t in range(1, frames + 1):
        prev = cv2_ref.stereo_frame(seq[t - 1, 0], seq[t - 1, 1], seq[t, 0], seq[t, 1], prev["kp_l"], prev["kp_r"], opts)
    return time.perf_counter() - t0, len(prev["kp_l"])


def cpu_pool(cores):
    import multiprocessing as mp
    return mp.get_context("spawn").Pool(cores)


def cpu_step(pool, cores, frames_per_worker, seed0):
    """all workers run concurrently; returns (stereo frames, wall seconds) for the step"""
    t0 = time.perf_counter()
    out = pool.map(_cpu_worker, [(seed0 + i, frames_per_worker, CFG["name"]) for i in range(cores)])
    wall = time.perf_counter() - t0
    # the per-worker timers exclude process spawn, imports and synthetic-data generation
    busy = max(o[0] for o in out)
    return cores * frames_per_worker, busy, wall


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    fpw = 2
    pool = cpu_pool(cores)
    try:
        for i in range(args.warmup):
            cpu_step(pool, cores, 1, 100 + 1000 * i)
        frames = 0; secs = 0.0
        for i in range(args.steps):
            f, busy, _ = cpu_step(pool, cores, fpw, 5000 + 1000 * i)
            frames += f; secs += busy
    finally:
        pool.close(); pool.join()
    value = frames / secs
    import cv2
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": cores * fpw},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d steps x %d processes x %d frames; reference glue restated over real cv2 %s calls "
                                   "(oracle/cv2_ref.py), 1 OpenCV thread per process" % (args.steps, cores, fpw, cv2.__version__)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"].update(cpu_overhead(value, cores))
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from zenslam_b200.runtime import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pin_rank_to_cores(local_rank, world)
    ctx = Context(local_rank)
    if args.extras_only:
        emit(extras(args, ctx))
        return
    line = measure_frontend(args, rank, world, local_rank, ctx, light=False)
    c4 = c4_sharded(args, rank, world, local_rank, ctx) if (world > 1 and not args.no_extra and CFG["name"] == "C2") else None
    ctx.close()
    if rank == 0:
        if world == 1 and not args.no_extra:
            line["extra"] = extras_in_child()
        elif c4 is not None:
            line["extra"] = {"C4_sharded": c4}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def c4_sharded(args, rank, world, local_rank, ctx):
    """BASELINE configs[3] as it is named -- 1280x1024 stereo, 4 096 frames sharded as independent contiguous chunks across
    the GPUs of the box -- inside the multi-GPU run the driver scales (`extra.C4_sharded` of the N > 1 lines).  Every rank owns
    zenslam_b200.sharding.shard_frames(4096, world, rank): its chunk in batches of 64, its first batch preceded by the one
    overlap frame as the carried previous frame; no collective on the data path.  All ranks call this (the timing barrier and
    the max-over-ranks reduction are collectives)."""
    import copy

    from zenslam_b200.sharding import shard_frames
    keep = CFG["name"]
    try:
        set_config("C4")
        a = copy.copy(args)
        chunk = shard_frames(4096, world, rank)
        a.batch, a.warmup, a.batches = CFG["batch"], 3, 2
        a.steps = max(1, -(-chunk.frames // a.batch))
        ln = measure_frontend(a, rank, world, local_rank, ctx, light=True)
        if ln is None:
            return None
        out = {k: ln[k] for k in ("value", "unit", "ms_per_step", "steps", "e2e", "stage_ms_per_step", "rank_ms_per_step") if k in ln}
        out.update({"workload": ln["config"]["workload"], "frames_total": 4096, "frames_per_rank": chunk.frames,
                    "overlap_frames_per_rank": 1, "batch_stereo_frames": a.batch,
                    "sharding": "zenslam_b200.sharding.shard_frames: contiguous chunks, one overlap frame, no collective"})
        return out
    except Exception as e:                          # never take the headline line down
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])} if rank == 0 else None
    finally:
        set_config(keep)


def pin_rank_to_cores(local_rank, world):
    """one rank per GPU and an equal, disjoint share of the host cores per rank: the per-step host work (graph launch,
    staging copies) of N ranks otherwise lands on whatever cores the scheduler picks and jitters the max-over-ranks time"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cores) >= world:
            per = len(cores) // world
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
    except Exception:
        pass


def measure_frontend(args, rank, world, local_rank, ctx, light=False):
    """one configuration (the globals set by set_config) through the batched front-end -> the JSON line (rank 0) or None.
    light: value + e2e + stage times only (the `extra` one-liners): no blocking-call / raw-BGR / sustained / CPU legs."""
    import torch
    import torch.distributed as dist
    from zenslam_b200 import detection_options, slam_options, tracking_options
    from zenslam_b200.frontend import StereoFrontend

    B, K, Wm = args.batch, args.steps, max(3, args.warmup)          # never fewer than three warm-up steps (timing rules)
    opts = slam_options(matcher="KNN", matcher_ratio=RATIO,
                        detection=detection_options(cell_size=CELL, fast_threshold=FAST_T, algorithm=CFG["alg"]),
                        tracking=tracking_options(klt_window_size=WIN, klt_max_level=MAX_LEVEL, klt_threshold=KLT_THR))
    fe = StereoFrontend(ctx, W, H, B, opts)

    # distinct batches cycled through the run: consecutive chunks of one long synthetic sequence per rank
    nb = min(args.batches, K + Wm)
    # every rank runs the SAME synthetic sequence: KLT iteration counts depend on the data (ranks with different sequences
    # differed by 2.8 % in step time at N = 8, profiles/r2_bench_n8_distinct_sequences.json), and weak scaling is defined on
    # identical per-GPU work -- the max-over-ranks time then measures the system, not the luck of the seeds
    seq = make_sequence(nb * B, 20000)                                   # (nb*B, 2, H, W)
    left = torch.from_numpy(np.ascontiguousarray(seq[:, 0])).reshape(nb, B, H, W)
    right = torch.from_numpy(np.ascontiguousarray(seq[:, 1])).reshape(nb, B, H, W)
    left_pin, right_pin = left.pin_memory(), right.pin_memory()
    left_dev, right_dev = left_pin.cuda(non_blocking=True), right_pin.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(v):
        if world == 1:
            return [v]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = v
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    stream = torch.cuda.current_stream()

    # ---- HBM-resident throughput ---------------------------------------------------------------
    for i in range(Wm):
        fe.upload(left_dev[i % nb], right_dev[i % nb]); fe.run()
    barrier()
    # per-stage device times: a separate, untimed pass with CUDA events between the stages (eager launches); the timed
    # region below runs the call exactly as a user gets it (one CUDA-graph replay per batch, no events inside)
    fe.timing_enable(True)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for i in range(K):
        j = (Wm + i) % nb
        fe.upload(left_dev[j], right_dev[j]); fe.run()
    s1.record(stream)
    stage_ms, runs = fe.timing_collect()
    staged_pass_ms = s0.elapsed_time(s1) / K
    fe.timing_enable(False)
    for i in range(2):
        fe.upload(left_dev[i % nb], right_dev[i % nb]); fe.run()          # (re-)capture outside the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        j = (Wm + i) % nb
        fe.upload(left_dev[j], right_dev[j]); fe.run()
    e1.record(stream)
    barrier()
    ms_own = e0.elapsed_time(e1)
    ms = max_over_ranks(ms_own)
    rank_ms = all_ranks(ms_own / K)
    launches = ctx.launches - launches0
    res = fe.download()
    clocks = sampler.stop() if rank == 0 else None
    kp_mean = float(np.mean(np.concatenate([res["n_left"], res["n_right"]])))
    valid = np.arange(res["track_keep"].shape[-1])[None, None, :] < res["track_n"][..., None]      # rows past a job's count are stale
    keep_frac = float((res["track_keep"].astype(bool) & valid).sum() / max(1, res["track_n"].sum()))
    value = world * B * K / (ms / 1000.0)

    # ---- sustained: the same loop for at least --min-seconds (same step count on every rank), clocks sampled throughout
    sustained = None
    if not light and args.min_seconds > 0:
        n_sus = int(np.ceil(args.min_seconds * 1000.0 / (ms / K)))
        samp2 = ClockSampler(local_rank)
        if rank == 0:
            samp2.start()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        u0.record(stream)
        for i in range(n_sus):
            j = i % nb
            fe.upload(left_dev[j], right_dev[j]); fe.run()
            if (i & 15) == 15:
                torch.cuda.synchronize()          # bound the launch queue; 16 steps of work stay queued behind it
        u1.record(stream)
        barrier()
        sus_ms = max_over_ranks(u0.elapsed_time(u1))
        sustained = {"value": world * B * n_sus / (sus_ms / 1000.0), "unit": UNIT, "seconds": sus_ms / 1000.0, "steps": n_sus,
                     "ms_per_step": sus_ms / n_sus, "clocks": samp2.stop() if rank == 0 else None}

    # ---- end to end: host buffers, H2D + D2H of every step inside the timed region ---------------------
    # (a) the pipelined public call: submit/wait keeps two batches in flight, so the PCIe copies of the neighbouring
    #     batches overlap this batch's kernels; the pipeline starts empty and is drained inside the timed region
    # (b) the blocking call process(): H2D -> kernels -> D2H strictly in sequence (what a per-frame caller sees)
    def e2e_pipelined(steps, first):
        fe.submit(left_pin[first % nb], right_pin[first % nb])
        for i in range(1, steps):
            j = (first + i) % nb
            fe.submit(left_pin[j], right_pin[j])
            fe.wait()
        return fe.wait()

    e2e_pipelined(max(2, min(Wm, 3)), 0)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    g0.record(stream)
    out = e2e_pipelined(K, Wm)
    g1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    e2e_ms = max_over_ranks(max(g0.elapsed_time(g1), wall_ms))
    e2e_value = world * B * K / (e2e_ms / 1000.0)
    assert int(out["n_left"].min()) > 0

    e2e_sync_value = None
    if not light:
        Ks = max(2, K // 2)
        for i in range(2):
            fe.process(left_pin[i % nb], right_pin[i % nb])
        barrier()
        t0 = time.perf_counter()
        for i in range(Ks):
            j = (Wm + i) % nb
            out = fe.process(left_pin[j], right_pin[j])
        barrier()
        sync_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
        e2e_sync_value = world * B * Ks / (sync_ms / 1000.0)

    raw_line = None
    if args.raw and not light:
        # optional: raw BGR camera frames + CLAHE + rectification maps in front of the same path (processor::process,
        # SURVEY 8 f1); 3x the H2D bytes, three more streaming kernels per step
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        r2 = ((xx - W / 2) ** 2 + (yy - H / 2) ** 2) / (W * W / 4)
        maps = [((W / 2 + (xx - W / 2) * (1 + kk * r2)).astype(np.float32), (H / 2 + (yy - H / 2) * (1 + kk * r2)).astype(np.float32))
                for kk in (0.05, -0.04)]
        lb = torch.stack([left, left.roll(2, -1), left // 2 + 40], -1).contiguous().pin_memory()
        rb = torch.stack([right, right.roll(2, -1), right // 2 + 40], -1).contiguous().pin_memory()
        fe.set_preprocess(3, True, 4.0, maps)

        def raw_pipelined(steps, first):
            fe.submit(lb[first % nb], rb[first % nb])
            for i in range(1, steps):
                j = (first + i) % nb
                fe.submit(lb[j], rb[j])
                fe.wait()
            return fe.wait()
        raw_pipelined(3, 0)
        barrier()
        t0 = time.perf_counter()
        raw_pipelined(K, Wm)
        barrier()
        raw_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
        raw_line = {"value": world * B * K / (raw_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": fe.h2d_bytes,
                    "input": "BGR frames; BGR2GRAY + CLAHE(4.0) + remap(INTER_LINEAR) on the device before the pyramids"}
        fe.set_preprocess(1, False, 4.0, None)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        P = sum(((W + (1 << l) - 1) >> l) * ((H + (1 << l) - 1) >> l) for l in range(fe_levels(W, H)))
        klt_bytes = B * 8 * (2 * P + 21 * kp_mean)          # 8 KLT calls x (two pyramids + points in/out)
        klt_s = stage_ms["klt"] / 1000.0
        achieved = klt_bytes / klt_s / 1e9 if klt_s > 0 else 0.0
        # per-stage algorithmic bytes of SURVEY 8(d) (per stereo frame: both images) against the stage's device time
        cells = (W // CELL[0]) * (H // CELL[1])
        stage_bytes = {"pyramid": 2 * P, "fast_grid": 2 * (W * H + 16 * cells), "orb": 4 * W * H + 2 * kp_mean * (512 + 32),
                       "match": 80 * kp_mean, "klt": 8 * (2 * P + 21 * kp_mean)}
        stage_roofline = {k: {"algorithmic_bytes_per_step": B * b, "ms": stage_ms[k],
                              "achieved_gbs": (B * b / (stage_ms[k] * 1e-3) / 1e9) if stage_ms[k] > 0 else None,
                              "frac_of_hbm_peak": (B * b / (stage_ms[k] * 1e-3) / 1e9 / peak) if stage_ms[k] > 0 else None}
                          for k, b in stage_bytes.items()}
        if stage_ms["match"] > 0:
            # the match stage is a contraction, not a stream: B x N x N x 256-bit Hamming distances, on the tcgen05 kind::i8 kernel
            # (bits expanded to bytes) when the call is large enough, else on the CUDA cores (ZS_HAMMING_NO_TENSOR forces those)
            stage_roofline["match"]["giga_distances_per_s"] = B * kp_mean * kp_mean / (stage_ms["match"] * 1e-3) / 1e9
            stage_roofline["match"]["tera_ops_per_s"] = 2.0 * 256 * B * kp_mean * kp_mean / (stage_ms["match"] * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_stereo_frames": B, "keypoints_per_image": kp_mean,
                       "fb_keep_fraction": keep_frac, "distinct_batches": nb,
                       "l2": "inputs larger than L2: one batch's pyramids are %.0f MB, %d distinct batches cycled"
                             % (2 * B * 3.2 * (W * H) / (752.0 * 480.0), nb),
                       "parallelism": "one independent sequence per GPU (the same synthetic content on every rank: identical per-GPU work), no collective",
                       "note": "the contract's metric: every cell of every frame is detected again and every keypoint goes through all "
                               "four forward/backward KLT jobs (SURVEY C2) -- more work per frame than the reference's stateful flow "
                               "(occupancy-aware detection, stereo tracks only of what the other camera lacks), which is measured as "
                               "extra.tracker; the per-frame seam calls the reference's slam_thread would make are extra.seams_e2e"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": fe.h2d_bytes, "d2h_bytes_per_step": fe.d2h_bytes,
                    "ms_per_step": e2e_ms / K, "call": "zs_frontend_submit_host/zs_frontend_wait (2 batches in flight)",
                    "blocking_call_value": e2e_sync_value},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms_per_step": stage_ms,
            "stage_roofline": stage_roofline,
            "rank_ms_per_step": {"min": min(rank_ms), "median": float(np.median(rank_ms)), "max": max(rank_ms)},
            "stage_timing": {"ms_per_step": staged_pass_ms, "steps": K,
                             "note": "same K batches run eagerly with CUDA events between the stages, immediately before the "
                                     "timed region; the timed region replays one CUDA graph per batch"},
            "roofline": {"kernel": "k_klt_track (fused forward+backward LK, 4 pairs x B frames per launch)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": klt_traffic(B) if CFG["name"] in ("C2", "TUMVI") else None,
                         "algorithmic_bytes_per_launch": klt_bytes, "avg_launch_ms": stage_ms["klt"],
                         "peak_source": peak_src,
                         "note": "KLT is instruction-issue bound (ncu: 86 % issue-active at C2, 76 % at TUMVI; DRAM 1.3 % / 3 % of peak), "
                                 "not HBM bound (SURVEY 8d); the compulsory-bytes figure is reported as the contract asks, see DESIGN.md"},
        }
        issue = klt_issue(B, stage_ms["klt"], clocks) if CFG["name"] in ("C2", "TUMVI") else None
        if issue:
            line["roofline"]["issue"] = issue
        if sustained:
            line["sustained"] = sustained
        if raw_line:
            line["e2e_raw_bgr"] = raw_line
        if world == 1 and not args.no_cpu_baseline and not light:
            line["cpu_baseline"] = cpu_baseline()
    fe.close()
    return line


def extras_in_child():
    """the `extra` block is measured by a child process (python bench.py --extras-only): a CUDA fault in one of the side
    paths is sticky for its process and must not take the headline line down with it"""
    import torch
    torch.cuda.empty_cache()
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--extras-only"], capture_output=True, text=True, timeout=900)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": "no JSON from the extras child (exit %d): %s" % (r.returncode, r.stderr[-600:])}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def extras(args, ctx):
    """Cheap measurements of the other paths, carried in the default line (`extra`): the per-frame seam calls, the stateful
    device tracker, and one-liners of BASELINE configs 3-5 and the shipped tumvi.yaml configuration.  Rank 0, one GPU."""
    import copy
    out = {}
    keep = CFG["name"]

    def guarded(name, fn):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:                      # an extra must never take the headline line down with it
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        if isinstance(out[name], dict):
            out[name]["bench_seconds"] = round(time.perf_counter() - t0, 2)

    def seams():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_seams
        return bench_seams.measure(ctx, frames=30, warm=6)

    def tracker():
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_tracker
        r = bench_tracker.measure(ctx, [1, 32], frames=18)
        r["call"] = "zs_tracker_track_host / zs_tracker_track / zs_tracker_submit_host: keypoint_tracker::track with its state on the device"
        return r

    def config(name, steps):
        def run():
            a = copy.copy(args)
            set_config(name)
            a.batch, a.steps, a.warmup, a.batches = CFG["batch"], steps, 3, 2
            ln = measure_frontend(a, 0, 1, ctx.device, ctx, light=True)
            return {k: ln[k] for k in ("value", "unit", "ms_per_step", "steps", "e2e", "stage_ms_per_step", "stage_roofline",
                                       "gpu_launches")} | {"workload": ln["config"]["workload"], "batch_stereo_frames": a.batch,
                                                           "keypoints_per_image": ln["config"]["keypoints_per_image"],
                                                           "roofline_klt_frac": ln["roofline"]["frac"]} | (
                    {"roofline_klt_issue": ln["roofline"]["issue"], "roofline_klt_traffic": ln["roofline"]["traffic"]}
                    if "issue" in ln["roofline"] else {})
        return run

    def c3():
        a = copy.copy(args)
        a.batch, a.steps, a.warmup = 64, 5, 3
        ln = measure_c3(a, 0, 1, ctx, light=True)
        return {k: ln[k] for k in ("value", "unit", "ms_per_step", "steps", "e2e", "u8_rows", "roofline", "gpu_launches")}

    guarded("seams_e2e", seams)
    guarded("tracker", tracker)
    guarded("C3", c3)
    guarded("C4", config("C4", 3))
    guarded("C5", config("C5", 3))
    guarded("TUMVI", config("TUMVI", 3))
    set_config(keep)
    return out


# --------------------------------------------------------------------------------------------------
# BASELINE config 3: SIFT-shaped L2 matching on the tensor cores (match stage only; SIFT extraction is input generation)
# --------------------------------------------------------------------------------------------------
C3_N, C3_DIM = 2000, 128
C3_WORKLOAD = ("C3: 1280x720 stereo, 2000 SIFT-shaped descriptors per image (128-d, integer-valued, norm ~512); per stereo "
               "frame L2 kNN-2 + ratio 0.8 AND mutual cross-check (both result sets); match stage only")


def c3_descriptors(pairs, seed):
    rng = np.random.default_rng(seed)
    g = rng.gamma(0.6, 40.0, (2, pairs, C3_N, C3_DIM))
    g = g / np.linalg.norm(g, axis=-1, keepdims=True) * 512
    return np.clip(np.rint(g), 0, 255).astype(np.float32)


def _c3_cpu_worker(args):
    seed, pairs = args
    import cv2
    cv2.setNumThreads(1)
    d = c3_descriptors(pairs, seed)
    t0 = time.perf_counter()
    for k in range(pairs):
        knn = cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(d[0, k], d[1, k], 2)          # matcher.cpp:60-75
        _ = [m for m in knn if len(m) == 2 and m[0].distance < RATIO * m[1].distance]
        _ = cv2.BFMatcher(cv2.NORM_L2, True).match(d[0, k], d[1, k])                    # matcher.cpp:76-80
    return time.perf_counter() - t0


def c3_cpu(cores, pairs_per_worker, seed):
    pool = cpu_pool(cores)
    try:
        busy = max(pool.map(_c3_cpu_worker, [(seed + i, pairs_per_worker) for i in range(cores)]))
    finally:
        pool.close(); pool.join()
    return cores * pairs_per_worker / busy


def run_c3(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from zenslam_b200.runtime import Context
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        c3_cpu(cores, 1, 50)
        vals = [c3_cpu(cores, 1, 900 + 10 * i) for i in range(max(1, args.steps))]
        v = float(np.mean(vals))
        import cv2
        emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": 1000.0 * cores / v, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": C3_WORKLOAD, "frames_per_step": cores},
              "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                               "sample": "%d steps x %d processes x 1 stereo frame, cv2 %s BFMatcher(NORM_L2) knnMatch + "
                                         "cross-check match, 1 OpenCV thread per process" % (args.steps, cores, cv2.__version__)},
              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pin_rank_to_cores(local_rank, world)
    ctx = Context(local_rank)
    line = measure_c3(args, rank, world, ctx, light=False)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_c3(args, rank, world, ctx, light=False):
    import torch
    import torch.distributed as dist
    from zenslam_b200.runtime import match_l2_cross, match_l2_knn2
    local_rank = ctx.device
    B = 64 if args.batch in (0, 128) else args.batch
    K, Wm = args.steps, max(3, args.warmup)
    nb = 3                                                           # 3 x 131 MB of float descriptors > the 126 MB L2
    host = [torch.from_numpy(c3_descriptors(B, 31000 + i)).pin_memory() for i in range(nb)]
    dev = [h.cuda(non_blocking=True) for h in host]
    n = torch.full((B,), C3_N, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step(d):
        r = match_l2_knn2(ctx, d[0], n, d[1], n, RATIO)
        c = match_l2_cross(ctx, d[0], n, d[1], n)
        return r, c

    for i in range(Wm):
        r, c = step(dev[i % nb])          # results kept like in the timed loop, so that torch's allocator has both sets of blocks
    barrier()
    # the kNN call alone, timed with events: its top-2 kernel is the tensor-core kernel the roofline is quoted for
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for i in range(K):
        match_l2_knn2(ctx, dev[i % nb][0], n, dev[i % nb][1], n, RATIO)
    k1.record(stream)
    torch.cuda.synchronize()
    knn_ms = k0.elapsed_time(k1) / K
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(K):
        r, c = step(dev[(Wm + i) % nb])
    e1.record(stream)
    barrier()
    ms = maxr(e0.elapsed_time(e1))
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * K / (ms / 1000.0)
    # end to end: descriptors from pinned host memory, both result sets back on the host, every step
    def e2e_step(h):
        d = h.cuda(non_blocking=True)
        r, c = step(d)
        return [x.cpu() for x in r] + [x.cpu() for x in c]
    for i in range(2):
        e2e_step(host[i % nb])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        out = e2e_step(host[(Wm + i) % nb])
    barrier()
    e2e_ms = maxr((time.perf_counter() - t0) * 1000.0)
    e2e_value = world * B * K / (e2e_ms / 1000.0)
    passed = float(out[2].float().mean())

    # the same end-to-end step, double-buffered the way zs_frontend_submit_host is at C2: the upload of step i+1 (copy stream)
    # and the result copies of step i-1 (second copy stream, pinned destinations) run beside the matching of step i; a slot's
    # buffers are reused only after its results have reached the host.  Every step still moves its own descriptors in and its
    # own results out inside the timed region.
    def e2e_pipelined(hosts):
        cs, ds = torch.cuda.Stream(), torch.cuda.Stream()
        dbuf = [torch.empty_like(hosts[0], device="cuda") for _ in range(2)]
        r0, c0 = step(dbuf[0])
        pin = [[torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in list(r0) + list(c0)] for _ in range(2)]
        up = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        outev = [torch.cuda.Event() for _ in range(2)]
        pending = [None, None]
        torch.cuda.synchronize()

        def run(steps, first):
            for i in range(steps):
                sl = i & 1
                if i >= 2:
                    outev[sl].synchronize()
                with torch.cuda.stream(cs):
                    dbuf[sl].copy_(hosts[(first + i) % nb], non_blocking=True)
                    up[sl].record(cs)
                stream.wait_event(up[sl])
                r, c = step(dbuf[sl])
                done[sl].record(stream)
                with torch.cuda.stream(ds):
                    ds.wait_event(done[sl])
                    for dst, src in zip(pin[sl], list(r) + list(c)):
                        dst.copy_(src, non_blocking=True)
                    outev[sl].record(ds)
                pending[sl] = (r, c)                       # keeps the device results alive until the slot is reused
            for sl in range(2):
                outev[sl].synchronize()

        run(4, 0)
        barrier()
        t = time.perf_counter()
        run(K, Wm)
        barrier()
        ms_ = maxr((time.perf_counter() - t) * 1000.0)
        return ms_, pin

    e2e_pipe_ms, pin_f = e2e_pipelined(host)
    # the pipelined results of the last step equal the blocking call's
    last = (K - 1) & 1
    pipe_same = bool(all(torch.equal(a_, b_) for a_, b_ in zip(pin_f[last], out)))
    # the same descriptors as u8 rows (cv::SIFT with descriptorType CV_8U): a quarter of the bytes over PCIe, no conversion pass
    host8 = [h.to(torch.uint8).pin_memory() for h in host]
    dev8 = [h.cuda(non_blocking=True) for h in host8]
    for i in range(Wm):
        r8, c8 = step(dev8[i % nb])
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record(stream)
    for i in range(K):
        r8, c8 = step(dev8[(Wm + i) % nb])
    u1.record(stream)
    barrier()
    u8_ms = maxr(u0.elapsed_time(u1))
    same_as_f32 = bool(torch.equal(r8[0], step(dev[(Wm + K - 1) % nb])[0][0]))
    for i in range(2):
        e2e_step(host8[i % nb])
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        out8 = e2e_step(host8[(Wm + i) % nb])
    barrier()
    e2e8_ms = maxr((time.perf_counter() - t0) * 1000.0)
    e2e8_pipe_ms, _ = e2e_pipelined(host8)
    ctx.async_error()                                                # raises if any float row was not an integer in 0..255
    if rank == 0:
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
        except Exception:
            peak, peak_src = 1393.0, "fallback (SURVEY 8d)"
        ops = 2.0 * B * C3_N * C3_N * C3_DIM
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32 (tcgen05 kind::i8)",
                "data": "synthetic",
                "config": {"workload": C3_WORKLOAD, "batch_stereo_frames": B, "ratio_pass_fraction": passed, "distinct_batches": nb,
                           "l2": "inputs larger than L2: %d distinct batches of %.0f MB cycled" % (nb, 2 * B * C3_N * C3_DIM * 4 / 1e6),
                           "parallelism": "independent stereo frames per GPU, no collective"},
                "e2e": {"value": world * B * K / (e2e_pipe_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": 2 * B * C3_N * C3_DIM * 4,
                        "d2h_bytes_per_step": int(sum(x.numel() * x.element_size() for x in out)), "ms_per_step": e2e_pipe_ms / K,
                        "blocking_value": e2e_value, "blocking_ms_per_step": e2e_ms / K, "same_results_as_blocking": pipe_same,
                        "call": "match_l2_knn2 + match_l2_cross on host descriptors, double-buffered: upload of step i+1 and result "
                                "copies of step i-1 beside the matching of step i (H2D + D2H of every step inside the timed region); "
                                "blocking_value = the same step with nothing overlapped"},
                "u8_rows": {"value": world * B * K / (u8_ms / 1000.0), "unit": UNIT, "ms_per_step": u8_ms / K,
                            "e2e": {"value": world * B * K / (e2e8_pipe_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": 2 * B * C3_N * C3_DIM,
                                    "d2h_bytes_per_step": int(sum(x.numel() * x.element_size() for x in out8)), "ms_per_step": e2e8_pipe_ms / K,
                                    "blocking_value": world * B * K / (e2e8_ms / 1000.0), "blocking_ms_per_step": e2e8_ms / K},
                            "same_matches_as_float_rows": same_as_f32,
                            "note": "zs_match_l2_knn2_u8 / zs_match_l2_cross_u8: the descriptors as u8 rows (cv::SIFT with descriptorType "
                                    "CV_8U, or narrowed where they are produced) -- the headline value / e2e above use float rows, "
                                    "the reference's default"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"kernel": "k_l2_tc_persist (inside zs_match_l2_knn2; the call also converts f32 -> u8 with the row norms in "
                                       "one pass, merges and applies the ratio test)", "bound": "tensor", "achieved": ops / (knn_ms * 1e-3) / 1e12,
                             "peak": peak, "unit": "TFLOP/s", "frac": ops / (knn_ms * 1e-3) / 1e12 / peak, "traffic": None,
                             "algorithmic_flops_per_launch": ops, "avg_launch_ms": knn_ms, "peak_source": peak_src,
                             "note": "integer ops counted as flops against the dense bf16 peak; avg_launch_ms is the WHOLE kNN call "
                                     "(the tensor kernel alone is 84.5 us of it, profiles/r1_l2_tc_persist_epi2_ncu.md); the drain of "
                                     "the accumulators (exact top-2 per query), not the tensor pipe, bounds the kernel -- DESIGN.md"}}
        if world == 1 and not args.no_cpu_baseline and not light:
            cores = os.cpu_count() or 1
            c3_cpu(cores, 1, 50)
            import cv2
            line["cpu_baseline"] = {"value": c3_cpu(cores, 1, 950), "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": "%d processes x 1 stereo frame, cv2 %s BFMatcher(NORM_L2) knnMatch + cross-check"
                                              % (cores, cv2.__version__)}
        return line
    return None


def klt_capture_file():
    """the committed ncu --set full capture of the KLT launch of the default workload (or of the TUMVI one): newest round first"""
    if CFG["name"] == "TUMVI":
        return os.path.join(ROOT, "profiles", "r2_klt63_traffic.json")
    for name in ("r2_klt_traffic.json", "r1_klt_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return p
    return os.path.join(ROOT, "profiles", "r1_klt_traffic.json")


def klt_traffic(batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one KLT launch from the committed ncu capture (taken at batch
    128; scaled linearly with the batch, which is exact for the compulsory part), or None"""
    try:
        t = json.load(open(klt_capture_file()))
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) * batch / t["batch_stereo_frames"]
    except Exception:
        return None


def klt_issue(batch, klt_ms, clocks):
    """The roofline that actually bounds the KLT kernel: warp instructions issued per second against the chip's issue rate
    (SMs x 4 schedulers x SM clock).  The instruction count of one launch comes from the committed ncu capture of this very
    workload (smsp__inst_executed.sum at batch 128, linear in the batch); the time is this run's."""
    try:
        import torch
        t = json.load(open(klt_capture_file()))
        inst = t["warp_instructions"] * batch / t["batch_stereo_frames"]
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak = sms * 4 * mhz * 1e6
        ach = inst / (klt_ms * 1e-3)
        return {"bound": "issue", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-inst/s", "frac": ach / peak,
                "warp_instructions_per_launch": inst}
    except Exception:
        return None


def fe_levels(w, h):
    lv = 1
    for _ in range(MAX_LEVEL):
        w, h = (w + 1) // 2, (h + 1) // 2
        if w <= WIN[0] or h <= WIN[1]:
            break
        lv += 1
    return lv


def cpu_baseline():
    cores = os.cpu_count() or 1
    fpw = 2
    pool = cpu_pool(cores)
    try:
        cpu_step(pool, cores, 1, 300)                      # warm-up: imports, page-in
        frames, busy, _ = cpu_step(pool, cores, fpw, 7000)
    finally:
        pool.close(); pool.join()
    import cv2
    out = {"value": frames / busy, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d processes x %d frames of the same workload (reference glue over real cv2 %s calls, "
                     "oracle/cv2_ref.py, 1 OpenCV thread per process)" % (cores, fpw, cv2.__version__)}
    out.update(cpu_overhead(out["value"], cores))
    # the reference's own deployment shape -- ONE process, OpenCV's default thread pool
    try:
        from oracle import cv2_ref
        cv2.setNumThreads(-1)
        opts = cpu_options()
        seq = make_sequence(4, 7100)
        prev = cv2_ref.stereo_frame(seq[0, 0], seq[0, 1], seq[0, 0], seq[0, 1], np.zeros((0, 2), np.float32),
                                    np.zeros((0, 2), np.float32), opts)
        t0 = time.perf_counter()
        for t in range(1, 4):
            prev = cv2_ref.stereo_frame(seq[t - 1, 0], seq[t - 1, 1], seq[t, 0], seq[t, 1], prev["kp_l"], prev["kp_r"], opts)
        out["single_process"] = {"value": 3 / (time.perf_counter() - t0), "unit": UNIT, "opencv_threads": cv2.getNumThreads(),
                                 "sample": "3 consecutive frames, one process, OpenCV default threads"}
    except Exception as e:            # the (ii) figure above is the contract's; this one is informational
        out["single_process"] = {"error": str(e)[:200]}
    return out


def cpu_overhead(value, cores):
    """What the Python restatement spends on work the C++ reference does not do (pyramid rebuilds inside the 8 Python LK
    calls, interpreter time of the per-cell loop), measured on this host, and the baseline with it taken out: the GPU/CPU
    ratio against `value_without_overhead` is the conservative one."""
    try:
        import cv2
        from oracle import cv2_ref
        cv2.setNumThreads(1)
        ov = cv2_ref.overhead_estimate(make_sequence(1, 7200)[0, 0], cpu_options())
        per_frame = cores / value                       # seconds one worker spends per stereo frame
        frac = min(0.9, (ov["pyramid_rebuild"] + ov["python_cell_loop"]) / per_frame)
        return {"overhead_not_in_reference": {"pyramid_rebuild_ms_per_frame": 1e3 * ov["pyramid_rebuild"],
                                              "python_cell_loop_ms_per_frame": 1e3 * ov["python_cell_loop"],
                                              "fraction_of_frame": frac},
                "value_without_overhead": value / (1.0 - frac)}
    except Exception as e:
        return {"overhead_not_in_reference": {"error": str(e)[:200]}}


_JSON_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG is set in the environment), so everything except that line is sent to stderr: fd 1 is pointed at fd 2 for the
    whole run and the line goes to a duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="stereo frames per step per GPU (default: the configuration's, 128 at C2)")
    ap.add_argument("--batches", type=int, default=2, help="distinct synthetic batches cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--raw", action="store_true", help="also time the end-to-end path fed with raw BGR frames (device pre-processing)")
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5", "TUMVI", "TUMVI752"],
                    help="C2 (default, the headline): 752x480 cells 16; C3: 2000 x 2000 SIFT-shaped L2 matching per stereo frame "
                         "(match stage only); C4: 1280x1024 cells 32; C5: 3840x2160 cells 32; TUMVI: the reference's shipped "
                         "tumvi.yaml (cells 64, thr 1, PARALLEL_GRID, KLT 63x63 / max_level 4) at 1024x1024; TUMVI752: same at 752x480")
    ap.add_argument("--min-seconds", type=float, default=5.0,
                    help="after the K timed steps, keep stepping for at least this long and report the sustained figure "
                         "(value, clocks) next to the burst one; 0 disables")
    ap.add_argument("--extras-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` block (per-frame seams, stateful tracker, C3/C4/C5 one-liners)")
    args = ap.parse_args()
    set_config("C2" if args.config == "C3" else args.config)
    if args.batch <= 0:
        args.batch = CFG["batch"]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "C3":
        run_c3(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
