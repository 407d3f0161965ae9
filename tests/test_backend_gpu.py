"""GPU parity tests of the back end: CUDA estimator (through the ve_* C ABI) vs the CPU oracle twin on identical
feature messages and IMU samples.  States are float64 on both sides; the tolerance covers different summation
orders (atomics, tiled products, Jacobi orderings) through 8 trust-region iterations per frame."""
import numpy as np
import pytest

import orc
from harness import synth, pipeline

pytestmark = pytest.mark.gpu

# m, rad, m/s, m/s^2, rad/s.  SURVEY §8d proposes 1e-3 (1e-4 for bg) "tighten after measuring": measured over repeated
# runs (the fp64 atomics make every run different) the worst window-state deviation is 1e-5 .. 5e-5 in steady state, so
# the gate is 3e-4 (3e-5 for bg), never above the survey's figure even where a test multiplies it.
TOL = dict(p=3e-4, q=3e-4, v=3e-4, ba=3e-4, bg=3e-5)


def make_gpu(**kw):
    from vins_mono_b200 import Estimator
    return Estimator(tic=synth.TIC, ric=synth.RIC, **kw)


def quat_angle(a, b):
    d = np.abs(np.sum(a * b, axis=-1))
    return 2 * np.arccos(np.clip(d, -1, 1))


def run_both(seq, msgs, cfg_kw=None, gpu_kw=None, check_prior=True, cost_rel=1e-3):
    cfg_kw, gpu_kw = cfg_kw or {}, gpu_kw or {}
    cpu, gpu = orc.OracleEstimator(orc.be_config(**cfg_kw)), make_gpu(**gpu_kw)
    t_imu, acc, gyr = seq.imu()
    fa, fb = pipeline.ImuFeeder(t_imu, acc, gyr), pipeline.ImuFeeder(t_imu, acc, gyr)
    seeds = pipeline.gt_seed_rows(seq, [m[0] for m in msgs])
    cpu.set_seed(seeds, seq.ba, seq.bg)
    gpu.set_seed(seeds, seq.ba, seq.bg)
    worst = dict(p=0, q=0, v=0, ba=0, bg=0, prior=0)
    n_nl = 0
    for stamp, ids, d in msgs:
        fa.feed(cpu, stamp)
        fb.feed(gpu, stamp)
        cpu.processImage(ids, d, stamp)
        gpu.processImage(ids, d, stamp)
        ia, ib = cpu.info(), gpu.info()
        for k in ("solver_flag", "frame_count", "marginalization_flag", "landmarks", "visual", "n_reboots"):
            assert ia[k] == ib[k], (stamp, k, ia, ib)
        if ia["solver_flag"] != 1:
            continue
        n_nl += 1
        sa, _ = cpu.states()
        sb, _ = gpu.states()
        worst["p"] = max(worst["p"], np.abs(sa[:, 0:3] - sb[:, 0:3]).max())
        worst["q"] = max(worst["q"], quat_angle(sa[:, 3:7], sb[:, 3:7]).max())
        worst["v"] = max(worst["v"], np.abs(sa[:, 7:10] - sb[:, 7:10]).max())
        worst["ba"] = max(worst["ba"], np.abs(sa[:, 10:13] - sb[:, 10:13]).max())
        worst["bg"] = max(worst["bg"], np.abs(sa[:, 13:16] - sb[:, 13:16]).max())
        # The absolute cost carries the prior's constant |r0|^2 = sum (v_k^T b)^2 / lambda_k over the retained eigen-
        # directions; the smallest retained lambda_k are round-off (the prior is rank deficient in the 4 gauge
        # directions and the reference keeps whatever lands above eps = 1e-8), so only cost *decreases* are comparable.
        da, db_ = ia["initial_cost"] - ia["final_cost"], ib["initial_cost"] - ib["final_cost"]
        assert abs(da - db_) <= 2e-2 + cost_rel * abs(da), (stamp, da, db_, ia, ib)
        if check_prior:
            Aa, ba_, blka = cpu.prior()
            Ab, bb_, blkb = gpu.prior()
            assert sorted(b[:2] + (b[3],) for b in blka) == sorted(b[:2] + (b[3],) for b in blkb)
            perm = []
            for (t, i, off, sz) in blkb:   # gpu order -> oracle columns
                o = next(x for x in blka if x[0] == t and (t >= 2 or x[1] == i))
                perm += list(range(o[2], o[2] + sz))
            if len(ba_) == 0:
                continue
            Aa, ba_ = Aa[np.ix_(perm, perm)], ba_[perm]
            # A' entry-wise against its largest entry; b' = g0 + A dx amplifies the ~1e-7 state differences by |A| ~ 1e8,
            # so the gradients are compared through the implied (regularised) prior mean shift instead.
            worst["prior"] = max(worst["prior"], np.abs(Aa - Ab).max() / np.abs(Aa).max())
            reg = 1e-9 * np.abs(np.diag(Aa)).max() * np.eye(len(ba_))
            ya, yb = np.linalg.solve(Aa + reg, ba_), np.linalg.solve(Ab + reg, bb_)
            worst["prior_mean"] = max(worst.get("prior_mean", 0), np.abs(ya - yb).max())
            worst.setdefault("floor_pairs", []).append(gpu.solver_debug()["floor_pairs"])
    return worst, n_nl, cpu, gpu


def test_estimator_matches_oracle_on_synthetic_tracks():
    seq = synth.Sequence(seed=11, duration=6.0)
    msgs = synth.track_messages(seq, 45)
    worst, n_nl, cpu, gpu = run_both(seq, msgs)
    print("worst deviations", worst, "frames", n_nl)
    assert n_nl >= 30
    for k, tol in TOL.items():
        assert worst[k] <= tol, (k, worst)
    assert worst["prior"] <= 1e-5
    # the eps floor of the prior normally separates only the eigenpairs at the noise floor (prior_floor.h: 3 of 75 in steady
    # state); the full decomposition is the fall-back (more than 16 such pairs: the first, weakly constrained priors)
    fp = worst["floor_pairs"]
    print("explicit eigenpairs per marginalisation", fp)
    assert sum(1 for k in fp if 0 <= k <= 16) >= 0.8 * len(fp), fp


@pytest.mark.parametrize("tr", [0.0, 0.033])
def test_estimator_td_and_extrinsic_blocks(tr):
    """ProjectionTdFactor path (estimate_td) together with a free extrinsic (estimate_extrinsic = 1); tr > 0 adds the
    rolling-shutter row term TR / ROW * row of projection_td_factor.cpp:34-46 (rolling_shutter_tr of the YAML)."""
    seq = synth.Sequence(seed=12, duration=5.0)
    msgs = synth.track_messages(seq, 35)
    kw = dict(estimate_td=1, estimate_extrinsic=1)
    worst, n_nl, cpu, gpu = run_both(seq, msgs, cfg_kw=dict(tr=tr, **kw), gpu_kw=dict(tr=tr, **kw))
    print("worst deviations (td + extrinsic, tr = %g)" % tr, worst, "frames", n_nl)
    for k, tol in TOL.items():
        assert worst[k] <= 2 * tol, (k, worst)
    assert abs(cpu.states()[1] - gpu.states()[1]) < 1e-5


def test_estimator_trajectory_error():
    seq = synth.Sequence(seed=13, duration=8.0)
    msgs = synth.track_messages(seq, 70)
    gpu = make_gpu()
    res = pipeline.run_vio(seq, None, gpu, 0, messages=[(0.0, np.zeros(0, np.int32), np.zeros((0, 7)))] + msgs)
    cpu = orc.OracleEstimator(orc.be_config())
    ref = pipeline.run_vio(seq, None, cpu, 0, messages=[(0.0, np.zeros(0, np.int32), np.zeros((0, 7)))] + msgs)
    ate_g, ate_c = pipeline.ate_rmse(seq, res["t"], res["P"]), pipeline.ate_rmse(seq, ref["t"], ref["P"])
    print("ATE rmse gpu", ate_g, "cpu oracle", ate_c)
    assert ate_g < 0.05 and abs(ate_g - ate_c) <= 0.01 * max(ate_c, 1e-3) + 1e-6    # within 1 % of the CPU path


def test_estimator_margin_second_new_path():
    """30 Hz feature messages: the parallax test (feature_manager.cpp:46-79) rejects most frames as keyframes, so
    optimization() takes the MARGIN_SECOND_NEW branch (estimator.cpp:929-1002: prior-only marginalisation of the
    second-newest pose, slideWindowNew with merged pre-integration) interleaved with MARGIN_OLD."""
    seq = synth.Sequence(seed=14, duration=4.0)
    msgs = synth.track_messages(seq, 60, pub_hz=30.0)
    flags = []
    worst, n_nl, cpu, gpu = run_both(seq, msgs)
    print("worst deviations (mixed marginalisation)", worst, "frames", n_nl)
    assert n_nl >= 40
    for k, tol in TOL.items():
        assert worst[k] <= tol, (k, worst)


def test_estimator_feature_dropouts_and_capacity():
    """Frames with no or very few features in the middle of a sequence (tracking loss: every landmark of the window
    loses its newest observation, estimator.cpp:120-150 still optimises with IMU + prior), and the landmark capacity
    bound (NUM_OF_F, estimator.h:111) turned into an error instead of an overflow."""
    seq = synth.Sequence(seed=15, duration=5.0)
    msgs = synth.track_messages(seq, 36)
    for k in (20, 21):                                   # two consecutive frames without a single feature
        msgs[k] = (msgs[k][0], np.zeros(0, np.int32), np.zeros((0, 7)))
    s, ids, d = msgs[27]
    msgs[27] = (s, ids[:5].copy(), d[:5].copy())         # five features only
    # after the outage the window is re-linearised far from its optimum (costs in the thousands, every step at the trust
    # region boundary): summation-order noise is amplified more than in steady state, hence the wider cost tolerance
    worst, n_nl, cpu, gpu = run_both(seq, msgs, check_prior=False, cost_rel=2e-2)
    print("worst deviations (dropouts)", worst, "frames", n_nl)
    for k, tol in TOL.items():
        assert worst[k] <= 3 * tol, (k, worst)
    small = make_gpu(max_features=40)
    small.set_seed(pipeline.gt_seed_rows(seq, [m[0] for m in msgs]), seq.ba, seq.bg)
    fb = pipeline.ImuFeeder(*seq.imu())
    clean = synth.track_messages(seq, 14)
    with pytest.raises(RuntimeError):
        for stamp, ids, d in clean:
            fb.feed(small, stamp)
            small.processImage(ids, d, stamp)


def test_estimator_window20_300_features_rolling_shutter():
    """BASELINE.json configs[3]: 20-keyframe window, 300 features, ProjectionTdFactor with rolling shutter.  The reduced
    camera system (322 columns) and the prior (136 parameters) no longer fit one CTA's shared memory: the step kernel
    factorises in global memory and the marginalisation keeps its reduced system there."""
    seq = synth.Sequence(seed=16, duration=5.0)
    msgs = synth.track_messages(seq, 30, max_feats=300)
    kw = dict(window_size=20, estimate_td=1, tr=0.033)
    worst, n_nl, cpu, gpu = run_both(seq, msgs, cfg_kw=kw, gpu_kw=kw)
    print("worst deviations (W = 20, 300 features, td + rolling shutter)", worst, "frames", n_nl)
    assert n_nl >= 8
    for k, tol in TOL.items():
        assert worst[k] <= 2 * tol, (k, worst)
    assert abs(cpu.states()[1] - gpu.states()[1]) < 1e-5


def _compare_run(seq, msgs, acc, gyr, t_imu, on_message=None, tol_mult=1.0):
    cpu, gpu = orc.OracleEstimator(orc.be_config()), make_gpu()
    seeds = pipeline.gt_seed_rows(seq, [m[0] for m in msgs])
    cpu.set_seed(seeds, seq.ba, seq.bg)
    gpu.set_seed(seeds, seq.ba, seq.bg)
    fa, fb = pipeline.ImuFeeder(t_imu, acc, gyr), pipeline.ImuFeeder(t_imu, acc, gyr)
    trace, worst, per_frame, diverged = [], dict(p=0.0, q=0.0, v=0.0, ba=0.0, bg=0.0), [], [None]
    for k, (stamp, ids, d) in enumerate(msgs):
        if on_message:
            on_message(k, cpu, gpu)
        fa.feed(cpu, stamp)
        fb.feed(gpu, stamp)
        cpu.processImage(ids, d, stamp)
        gpu.processImage(ids, d, stamp)
        ia, ib = cpu.info(), gpu.info()
        for key in ("solver_flag", "frame_count", "marginalization_flag", "n_reboots", "n_solves"):
            assert ia[key] == ib[key], (k, key, ia, ib)
        trace.append((ia["solver_flag"], ia["frame_count"], ia["n_reboots"]))
        if ia["solver_flag"] == 1 and (ia["successful_steps"], ia["iterations"], ia["termination"]) != (ib["successful_steps"], ib["iterations"], ib["termination"]):
            # The trust-region loop took a different accept / reject decision (rho against 1e-3 / 0.25 / 0.75 is a discontinuity:
            # ~1e-7 state differences can flip it when a window is far from converged, here 8 iterations bring the cost
            # from 68 to 47).  Both paths are valid executions of the algorithm; states are comparable only up to that point.
            print("solver path differs at message", k, {q: ia[q] for q in ("iterations", "successful_steps", "termination", "initial_cost", "final_cost")},
                  {q: ib[q] for q in ("iterations", "successful_steps", "termination", "initial_cost", "final_cost")})
            if diverged[0] is None:
                diverged[0] = k
        if ia["solver_flag"] == 1 and diverged[0] is None:
            sa, sb = cpu.states()[0], gpu.states()[0]
            per_frame.append((k, float(np.abs(sa[:, 0:3] - sb[:, 0:3]).max())))
            worst["p"] = max(worst["p"], np.abs(sa[:, 0:3] - sb[:, 0:3]).max())
            worst["q"] = max(worst["q"], quat_angle(sa[:, 3:7], sb[:, 3:7]).max())
            worst["v"] = max(worst["v"], np.abs(sa[:, 7:10] - sb[:, 7:10]).max())
            worst["ba"] = max(worst["ba"], np.abs(sa[:, 10:13] - sb[:, 10:13]).max())
            worst["bg"] = max(worst["bg"], np.abs(sa[:, 13:16] - sb[:, 13:16]).max())
    print("per-frame position deviation:", " ".join("%d:%.1e" % (k, v) for k, v in per_frame))
    for key in worst:
        assert worst[key] <= tol_mult * TOL[key], (key, worst)
    return trace, diverged[0]


def test_failure_detection_reboots_like_the_reference():
    """Estimator::failureDetection (estimator.cpp:621-667) -> clearState + setParameter (:193-201): a burst of corrupted
    accelerometer samples (+300 m/s^2 for one frame interval) drives the newest position past the 5 m / 1 m limits, both
    estimators reboot at the same message, refill the window, re-initialise and agree again."""
    seq = synth.Sequence(seed=4, duration=5.5)
    msgs = synth.track_messages(seq, 44, max_feats=90)
    t_imu, acc, gyr = seq.imu()
    acc = acc.copy()
    t0 = msgs[20][0]
    # the first sample after the stamp is left clean: it also enters message 20 through the interpolation at the image time
    acc[(t_imu > t0 + 0.0051) & (t_imu <= t0 + 0.1), 0] += 300.0
    trace, diverged = _compare_run(seq, msgs, acc, gyr, t_imu)
    assert diverged is None or diverged >= 36   # at least the first frames after the re-initialisation (message 32) are compared
    flags = [t[0] for t in trace]
    reboots = [t[2] for t in trace]
    assert reboots[20] == 0 and reboots[21] == 1 and reboots[-1] == 1
    assert flags[20] == 1 and flags[21] == 0 and flags[-1] == 1          # rebooted, then NON_LINEAR again
    assert [t[1] for t in trace[21:32]] == list(range(0, 11))             # frame_count restarts at 0 and refills the window


def test_clear_state_mid_run():
    """Estimator::clearState + setParameter called by the node on a restart message (estimator_node.cpp:186-203)."""
    seq = synth.Sequence(seed=6, duration=4.5)
    msgs = synth.track_messages(seq, 34, max_feats=90)
    t_imu, acc, gyr = seq.imu()

    def restart(k, cpu, gpu):
        if k == 17:
            cpu.clearState()
            gpu.clearState()

    trace, diverged = _compare_run(seq, msgs, acc, gyr, t_imu, on_message=restart)
    assert diverged is None or diverged >= 30
    assert trace[16][0] == 1 and trace[17][0] == 0 and trace[17][1] == 1 and trace[-1][0] == 1
