#!/usr/bin/env python
"""bench.py — frames/sec of the VINS-Mono hot path (feature tracker + sliding-window BA) on B200.

A "step" is one published (10 Hz) frame of one sequence, fully processed: the two 752x480 camera images that
arrive in that interval go through FeatureTracker::readImage (the second one publishes), the ~20 IMU samples
through processIMU, the feature message through Estimator::processImage (triangulate, 8-iteration dogleg solve,
marginalisation, slide).  Workload = BASELINE.json configs[1]: one synthetic EuRoC-shaped sequence per GPU
(weak scaling: rank r runs sequence seed r).  Inputs are synthetic (harness/synth.py), no dataset is read.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                            the CPU oracle port of the reference path
Under torchrun (N > 1) every rank runs one sequence; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

# Concurrent sequences use two CUDA streams each; the default of 8 hardware work queues would alias them onto each other
# (false serialisation).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from harness import synth, pipeline  # noqa: E402

METRIC = "frames/sec (tracker+BA) at 752x480, 10-KF window, 150 feats; ATE vs ref"
WORKLOAD = ("configs[1]: single 752x480 synthetic sequence per GPU, 10-keyframe window, 150 features, 200 Hz IMU, "
            "fp64 Jacobians")
INIT_PUBS = 12           # published frames consumed before the window is full and seeded (estimator goes NON_LINEAR)
ALGO_BYTES_IMAGE = 1_319_760   # compulsory front-end traffic per input image
ALGO_BYTES_SOLVE = 233_000     # back-end inputs+outputs per solve at C1


def sequence_inputs(seed, n_pub):
    """Rendered frames + IMU of one sequence, enough for n_pub published frames."""
    n_img = 2 * (n_pub + 1) + 2
    seq = synth.Sequence(seed=seed, duration=n_img / 20.0 + 0.5)
    ts, imgs = pipeline.cached_images(seq, n_img)
    t_imu, acc, gyr = seq.imu()
    return seq, ts, np.ascontiguousarray(imgs), (t_imu, acc, gyr)


class BatchedImu(pipeline.ImuFeeder):
    """Same sample selection / interpolation as estimator_node.cpp:98-136, 225-265, delivered in one call."""

    def feed(self, estimator, stamp, td=None):
        rec = _Recorder()
        super().feed(rec, stamp, estimator.states()[1] if td is None else td)
        if rec.dt:
            estimator.processIMU_batch(np.array(rec.dt), np.array(rec.acc), np.array(rec.gyr))


class _Recorder:
    def __init__(self):
        self.dt, self.acc, self.gyr = [], [], []

    def processIMU(self, dt, a, g):
        self.dt.append(dt)
        self.acc.append(np.array(a, float))
        self.gyr.append(np.array(g, float))


class ClockSampler:
    """SM clock and throttle reasons polled through NVML (every ~2 ms) while a pass runs; the reported figures use the
    samples that fall inside the timed region (SoloGate marks its begin / end)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread = index, [], threading.Event(), None
        self.t_begin = self.t_end = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-agnostic lookup is not needed: one process per GPU with LOCAL_RANK == device index
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:  # no NVML: report nothing rather than fail the bench
            return

        def poll():
            while not self.stop_flag.is_set():
                try:
                    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    why = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.rows.append((time.perf_counter(), float(mhz), int(why)))
                except Exception:
                    pass
                time.sleep(0.002)

        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(1.0)
        rows = self.rows
        if self.t_begin is not None and self.t_end is not None:
            inside = [r for r in rows if self.t_begin <= r[0] <= self.t_end]
            rows = inside or rows
        sm = [r[1] for r in rows]
        reasons = sorted({name for r in rows for bit, name in self.REASONS.items() if r[2] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "NVML, polled every 2 ms inside the timed region"}


def make_gpu_pair(device):
    from vins_mono_b200 import FeatureTracker, Estimator
    trk = FeatureTracker(device=device, **synth.tracker_config_dict())
    est = Estimator(tic=synth.TIC, ric=synth.RIC, device=device)
    return trk, est


class TwoStage:
    """Runs a tracker stage (producer thread) and an estimator stage (caller's thread) the way the reference runs its two
    nodes: concurrently, coupled by a bounded queue of feature messages.  The producer stops after `n_before` messages
    until `go()` is called, so that a timed region contains exactly the frames produced after it (pipeline fill and
    drain fall inside the region)."""

    def __init__(self, produce, n_before, n_total, depth=2):
        import queue
        self.q, self.gate, self.err = queue.Queue(maxsize=depth), threading.Event(), None
        self.produce, self.n_before, self.n_total = produce, n_before, n_total
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            for k in range(self.n_total):
                if k == self.n_before:
                    self.gate.wait()
                msg = self.produce()
                self.q.put(msg)
                if msg is None:
                    return
            self.q.put(None)
        except BaseException as e:  # surfaced by get()
            self.err = e
            self.q.put(None)

    def go(self):
        self.gate.set()

    def get(self):
        msg = self.q.get()
        if self.err:
            raise self.err
        return msg


class SoloGate:
    """Brackets the timed region of ONE sequence: device drained, L2 evicted, optional cross-rank barrier, CUDA events."""

    def __init__(self, device, flush=None, sync_cb=None, clocks=None):
        import torch
        self.torch, self.device, self.flush, self.sync_cb, self.clocks = torch, device, flush, sync_cb, clocks
        self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def begin(self):
        if self.flush is not None:
            self.flush.fill_(1.0)
        self.torch.cuda.synchronize(self.device)
        if self.sync_cb:
            self.sync_cb()
        if self.clocks:
            self.clocks.mark_begin()
        self.ev0.record()

    def end(self):
        self.torch.cuda.synchronize(self.device)  # includes the last frame's (asynchronous) marginalisation
        self.ev1.record()
        self.ev1.synchronize()
        if self.clocks:
            self.clocks.mark_end()
        return self.ev0.elapsed_time(self.ev1)


def run_ours_pass(seq, ts, imgs, imu, n_init, warmup, steps, device, host_images, profile=False, flush=None, sync_cb=None,
                  gate=None, d_imgs=None, clocks=None):
    """One pass over the sequence.  The K timed steps form ONE region bracketed by device synchronisation and a pair of
    CUDA events (ms per step = region / K); sync_cb, if given, is the cross-rank barrier placed inside the bracket.
    Tracker and estimator handles are driven from two host threads (TwoStage); with profile=True they run serially."""
    import torch
    trk, est = make_gpu_pair(device)
    if profile:
        trk.set_profile(True)
        est.set_profile(True)
    feeder = BatchedImu(*imu)
    n_pub = n_init + warmup + steps
    est.set_seed(pipeline.gt_seed_rows(seq, ts), seq.ba, seq.bg)
    frame_bytes = imgs.shape[1] * imgs.shape[2]
    if not host_images:
        if d_imgs is None:
            d_imgs = torch.from_numpy(imgs).to(f"cuda:{device}")
        base = d_imgs.data_ptr()
    if gate is None:
        gate = SoloGate(device, flush, sync_cb, clocks)
    launches, h2d, d2h, traj_t, traj_p = 0, 0.0, 0.0, [], []
    n_img = len(ts)
    cursor = [0]

    def one_image(k):
        if host_images:
            return trk.node_image(imgs[k], float(ts[k]))[0]
        return trk.node_image_device(base + k * frame_bytes, imgs.shape[2], float(ts[k]))

    def produce():
        """Images until one publishes -> (stamp, ids, observations, launches, h2d, d2h) or None at end of data."""
        torch.cuda.set_device(device)
        n_l, n_h, n_d, r = 0, 0.0, 0.0, 0
        while r != 2 and cursor[0] < n_img:
            r = one_image(cursor[0])
            cursor[0] += 1
            if r:
                n_l += trk.timing()[1]
                a, b = trk.traffic()
                n_h += a
                n_d += b
        if r != 2:
            return None
        ids, d = pipeline.oracle_feature_message(trk.result())
        return float(ts[cursor[0] - 1]), ids, d, n_l, n_h, n_d

    class Serial:
        def go(self):
            pass

        def get(self):
            return produce()

    stage = Serial() if profile else TwoStage(produce, n_init + warmup, n_pub)
    pubs, first_msg, n_timed, started = 0, True, 0, False
    while pubs < n_pub:
        timed = pubs >= n_init + warmup
        if timed and not started:
            # start of the timed region: both stages idle, device drained, L2 evicted once (every timed step reads images
            # that were uploaded long ago and never touched since, i.e. cold in L2; the rest of a step's data is
            # produced inside it)
            gate.begin()
            started = True
            stage.go()
        msg = stage.get()
        if msg is None:
            break
        stamp, ids, d, step_launch, step_h2d, step_d2h = msg
        if first_msg:  # estimator_node.cpp:167-172 drops the first feature message
            first_msg = False
        else:
            feeder.feed(est, stamp)
            est.processImage(ids, d, stamp)
            step_launch += est.launch_count()
            a, b = est.traffic()
            step_h2d += a
            step_d2h += b
        if timed:
            n_timed += 1
            launches += step_launch
            h2d += step_h2d
            d2h += step_d2h
        if est.info()["solver_flag"] == 1:
            st, _ = est.states()
            traj_t.append(stamp)
            traj_p.append(st[-1, 0:3].copy())
        pubs += 1
    region_ms = gate.end() if started else 0.0
    stage.go()
    times = [region_ms / max(n_timed, 1)] * n_timed if started else []
    out = dict(times=times, launches=launches, h2d=h2d, d2h=d2h, traj_t=traj_t, traj_p=traj_p, info=est.info())
    if profile:
        out["trk_k"], out["est_k"] = trk.kernel_times(), est.kernel_times()
    trk.close()
    est.close()
    return out


def run_replay_pass(seq, ts, imgs, imu, n_init, warmup, steps, device, n_seq, host_images=False, flush=None, sync_cb=None):
    """n_seq replicas of the sequence driven by the native replay driver (include/vinsb200/replay.h: the reference's two
    node loops in C++, one thread pair per sequence), all concurrently on one GPU.  The timed region is ONE vr_advance
    call of `steps` published frames per sequence, bracketed like the single-sequence pass.  n_seq = 1 is the plain
    single-sequence pipeline without any Python in the loop."""
    import torch
    from vins_mono_b200 import ReplaySession
    t_imu, acc, gyr = imu
    pairs = [make_gpu_pair(device) for _ in range(n_seq)]
    seed = pipeline.gt_seed_rows(seq, ts)
    for _, est in pairs:
        est.set_seed(seed, seq.ba, seq.bg)
    d_imgs = None
    if host_images:
        images = dict(images=imgs)
    else:
        d_imgs = torch.from_numpy(np.array(imgs, copy=True)).to(f"cuda:{device}")
        images = dict(images=d_imgs.data_ptr(), shape=imgs.shape)
    seqs = [dict(stamps=ts, imu_t=t_imu, acc=acc, gyr=gyr, **images) for _ in range(n_seq)]
    ses = ReplaySession([p[0] for p in pairs], [p[1] for p in pairs], seqs)
    gate = SoloGate(device, flush, sync_cb)
    ses.advance(n_init + warmup)
    before = [ses.stats(k) for k in range(n_seq)]
    gate.begin()
    frames = ses.advance(steps)
    region_ms = gate.end()
    after = [ses.stats(k) for k in range(n_seq)]
    trajs = [ses.trajectory(k) for k in range(n_seq)]
    infos = [p[1].info() for p in pairs]
    ses.close()
    for trk, est in pairs:
        trk.close()
        est.close()
    del d_imgs
    return dict(frames=frames, region_ms=region_ms, fps=frames / (region_ms / 1e3) if region_ms else 0.0,
                launches=sum(a["launches"] - b["launches"] for a, b in zip(after, before)),
                h2d=sum(a["h2d"] - b["h2d"] for a, b in zip(after, before)), d2h=sum(a["d2h"] - b["d2h"] for a, b in zip(after, before)),
                trajs=trajs, infos=infos, final_cost=[i["final_cost"] for i in infos])


def run_reference_pass(seq, ts, imgs, imu, n_init, warmup, steps):
    """The CPU port of the reference path (oracle tracker + estimator twins) on two host threads, tracker and estimator
    coupled by a queue like the reference's two nodes (each hot loop single-threaded: the tracker node is, and Ceres
    runs with num_threads = 1)."""
    import orc
    trk = orc.OracleTracker(synth.tracker_config_dict())
    est = orc.OracleEstimator(orc.be_config())
    est.set_fast_eigen(True)  # tridiagonal QL (the reference's Eigen solver class) instead of the parity tests' Jacobi
    feeder = pipeline.ImuFeeder(*imu)
    est.set_seed(pipeline.gt_seed_rows(seq, ts), seq.ba, seq.bg)
    n_pub = n_init + warmup + steps
    traj_t, traj_p = [], []
    cursor = [0]

    def produce():
        r = 0
        while r != 2 and cursor[0] < len(ts):
            r, _ = trk.node_image(imgs[cursor[0]], float(ts[cursor[0]]))
            cursor[0] += 1
        if r != 2:
            return None
        ids, d = pipeline.oracle_feature_message(trk.result())
        return float(ts[cursor[0] - 1]), ids, d

    stage = TwoStage(produce, n_init + warmup, n_pub)
    pubs, first_msg, n_timed, t0 = 0, True, 0, None
    while pubs < n_pub:
        if pubs == n_init + warmup:
            t0 = time.perf_counter()
            stage.go()
        msg = stage.get()
        if msg is None:
            break
        stamp, ids, d = msg
        if first_msg:
            first_msg = False
        else:
            feeder.feed(est, stamp)
            est.processImage(ids, d, stamp)
        if t0 is not None:
            n_timed += 1
        if est.info()["solver_flag"] == 1:
            st, _ = est.states()
            traj_t.append(stamp)
            traj_p.append(st[-1, 0:3].copy())
        pubs += 1
    total_ms = (time.perf_counter() - t0) * 1e3 if t0 is not None else 0.0
    stage.go()
    return dict(times=[total_ms / max(n_timed, 1)] * n_timed, traj_t=traj_t, traj_p=traj_p)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=200,
                    help="published frames of the CPU-baseline sample inside the b200 arm (~10 s of CPU work)")
    ap.add_argument("--batch", type=int, default=16,
                    help="supplementary pass: this many concurrent replicas of the sequence per GPU (0 = skip)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    n_pub = INIT_PUBS + warmup + a.steps

    if a.impl == "reference":
        if rank != 0:
            return
        seq, ts, imgs, imu = sequence_inputs(0, n_pub)
        r = run_reference_pass(seq, ts, imgs, imu, INIT_PUBS, warmup, a.steps)
        total_s = sum(r["times"]) / 1e3
        fps = len(r["times"]) / total_s
        sample = f"{len(r['times'])} published frames (2 images + 1 window solve each) of sequence seed 0 after {INIT_PUBS} init + {warmup} warm-up frames"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": len(r["times"]),
            "warmup": warmup, "ms_per_step": 1e3 * total_s / len(r["times"]), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sequences_per_gpu": 1,
                       "arm": "configs[0]: the same sequence on the host CPU, 1 stream (tracker thread + estimator thread)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 2, "kind": "port", "sample": sample, "cpu": cpu_model(),
                             "note": "CPU restatement of the reference path (oracle/), not Ceres/OpenCV; no wall-clock solver cap"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ate_rmse_m": pipeline.ate_rmse(seq, r["traj_t"], r["traj_p"]) if len(r["traj_t"]) > 3 else None,
        }))
        return

    from vins_mono_b200 import shard
    # frames are rendered (fork pool) before this process touches CUDA; the CPU-baseline sample needs a longer sequence
    cpu_frames = max(a.steps, a.cpu_frames) if (not a.no_cpu_baseline and world == 1) else a.steps
    seq, ts, imgs, imu = sequence_inputs(shard.sequences_of_rank(rank, world, 1)[0], INIT_PUBS + warmup + cpu_frames)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the vinsb200 library has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keeps NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=f"cuda:{local}")  # 256 MB > 126 MB L2

    # untimed shake-out pass (first CUDA context / module load, pinned allocations)
    run_ours_pass(seq, ts, imgs, imu, INIT_PUBS, 0, 2, local, host_images=True)

    def barrier():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local)

    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    rank_barrier = (lambda: dist.barrier()) if world > 1 else None
    res_dev = run_ours_pass(seq, ts, imgs, imu, INIT_PUBS, warmup, a.steps, local, host_images=False, flush=flush, sync_cb=rank_barrier,
                            clocks=clocks)
    barrier()
    clk = clocks.stop()
    barrier()
    res_e2e = run_ours_pass(seq, ts, imgs, imu, INIT_PUBS, warmup, a.steps, local, host_images=True, flush=flush, sync_cb=rank_barrier)
    barrier()

    def agg(res):
        rate, sec, _ = shard.aggregate_rate(sum(res["times"]), len(res["times"]), device=f"cuda:{local}")
        return rate, sec

    fps_dev, t_dev = agg(res_dev)
    fps_e2e, t_e2e = agg(res_e2e)
    launches = torch.tensor([float(res_dev["launches"])], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(launches)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, measured live with CUDA events in a profile pass
    prof = run_ours_pass(seq, ts, imgs, imu, INIT_PUBS, 2, 10, local, host_images=False, profile=True)
    kt = {}
    for name, (ms, cnt) in list(prof["trk_k"].items()) + list(prof["est_k"].items()):
        if cnt:
            kt[name] = {"total_ms": ms, "launches": cnt, "avg_us": 1e3 * ms / cnt}
    n_prof_frames = 12
    dominant = max(kt, key=lambda k: kt[k]["total_ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_frame = {k: v["launches"] / n_prof_frames for k, v in kt.items()}
    # algorithmic bytes per launch of the dominant kernel: the per-unit figure of SURVEY §8(d)
    # divided by the launches that unit needs
    if dominant in ("clahe", "pyrdown", "lk_track", "mask_discs", "min_eig", "gftt_tail"):
        algo = 2 * ALGO_BYTES_IMAGE / max(per_frame[dominant], 1e-9)
    else:
        algo = ALGO_BYTES_SOLVE / max(per_frame[dominant], 1e-9)
    avg_s = kt[dominant]["avg_us"] * 1e-6
    achieved = algo / avg_s / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(dominant, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": algo, "avg_launch_us": kt[dominant]["avg_us"],
                "note": "single-sequence step: every kernel is latency/launch bound, not HBM bound (see DESIGN.md)",
                "kernels": kt}

    # ---- supplementary: concurrent sequences on the same GPU (BASELINE.json configs[2] direction; not the headline)
    batch = None
    if a.batch > 1 and world == 1:
        rb = run_replay_pass(seq, ts, imgs, imu, INIT_PUBS, warmup, min(a.steps, 30), local, a.batch, flush=flush)
        spread = max(abs(c - rb["final_cost"][0]) for c in rb["final_cost"]) / max(1.0, abs(rb["final_cost"][0]))
        r1 = run_replay_pass(seq, ts, imgs, imu, INIT_PUBS, warmup, min(a.steps, 30), local, 1, flush=flush)
        batch = {"sequences_per_gpu": a.batch, "value": rb["fps"], "unit": "frames/s", "frames": rb["frames"],
                 "ms_region": rb["region_ms"], "gpu_launches": rb["launches"], "replica_final_cost_rel_spread": spread,
                 "single_sequence_same_driver": {"value": r1["fps"], "unit": "frames/s", "frames": r1["frames"]},
                 "spread_note": "replicas see identical inputs; they differ only through the order of the fp64 atomic adds "
                                "of the Hessian assembly (not bit-reproducible run to run, same as a single sequence)",
                 "note": "replicas of sequence 0 (device-resident frames) driven by the native replay driver "
                         "(vinsb200/replay.h: one tracker thread + one estimator thread per sequence); aggregate frames/s "
                         "over one shared timed region"}

    # ---- CPU baseline on this box's host cores (bounded sample of the same workload)
    cpu_b = None
    ate_ref = ate_same = None
    if not a.no_cpu_baseline and world == 1:
        rr = run_reference_pass(seq, ts, imgs, imu, INIT_PUBS, warmup, cpu_frames)
        tot = sum(rr["times"]) / 1e3
        cpu_b = {"value": len(rr["times"]) / tot, "unit": "frames/s", "cores": 2, "kind": "port", "cpu": cpu_model(),
                 "sample": f"{len(rr['times'])} published frames of the same sequence after the same {INIT_PUBS} init + {warmup} "
                           f"warm-up frames (oracle tracker + estimator twins on two threads, {tot:.1f} s)"}
        nn = min(len(rr["traj_t"]), len(res_dev["traj_t"]))
        if nn > 3:  # the same frames of both trajectories, so that the two ATE figures are comparable
            ate_ref = pipeline.ate_rmse(seq, rr["traj_t"][:nn], rr["traj_p"][:nn])
            ate_same = pipeline.ate_rmse(seq, res_dev["traj_t"][:nn], res_dev["traj_p"][:nn])
    ate = pipeline.ate_rmse(seq, res_dev["traj_t"], res_dev["traj_p"]) if len(res_dev["traj_t"]) > 3 else None

    k = len(res_dev["times"])
    print(json.dumps({
        "impl": "b200", "metric": METRIC, "value": fps_dev, "unit": "frames/s", "n_gpus": world, "steps": k, "warmup": warmup,
        "ms_per_step": 1e3 * t_dev / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sequences_per_gpu": 1,
                   "pipelining": "tracker and estimator handles driven from two host threads coupled by a depth-2 queue (the "
                                 "reference's two nodes); the timed region starts with both idle and ends with both drained",
                   "l2": "L2 flushed (256 MB write) at the start of the timed region; every step reads two images that were never touched since upload",
                   "timing": "one CUDA-event pair around the K steps, device synchronised (and ranks barriered) on both sides; max over ranks"},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": 1e3 * t_e2e / k,
                "h2d_bytes_per_step": res_e2e["h2d"] / k, "d2h_bytes_per_step": res_e2e["d2h"] / k},
        "gpu_launches": int(launches.item()),
        "roofline": roofline, "cpu_baseline": cpu_b, "clocks": clk, "batch": batch,
        "ate_rmse_m": ate, "ate_rmse_m_same_frames": ate_same, "ate_rmse_m_cpu_port": ate_ref,
        "ate_rel_diff": (abs(ate_same - ate_ref) / ate_ref) if ate_ref else None,
        "solver": res_dev["info"],
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
