"""GPU parity of the relocalisation path (Estimator::setReloFrame estimator.cpp:1128-1146, the relo_Pose block and its
ProjectionFactors :769-801, the drift bookkeeping of double2vector :598-617): CUDA estimator through the C ABI against the CPU
oracle on identical loop messages, and against the drift that was injected."""
import numpy as np
import pytest

import orc
from harness import pipeline, synth

pytestmark = pytest.mark.gpu

TOL = dict(p=3e-4, q=3e-4, v=3e-4, ba=3e-4, bg=3e-5)


def quat_angle(a, b):
    d = np.abs(np.sum(a * b, axis=-1))
    return 2 * np.arccos(np.clip(d, -1, 1))


def rz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def run(seq, msgs, relo_at, cfg_kw=None, window_index=5, t_loop=1.0, yaw_deg=7.0, t_drift=(0.4, -0.3, 0.1), stamp_shift=0.0):
    """Seeded run of both estimators; before each message k in relo_at a loop message between window frame `window_index` and the
    old key frame at t_loop arrives (pose_graph's pose of the old frame = its true pose moved by the injected drift)."""
    from vins_mono_b200 import Estimator
    cfg_kw = cfg_kw or {}
    cpu, gpu = orc.OracleEstimator(orc.be_config(**cfg_kw)), Estimator(tic=synth.TIC, ric=synth.RIC, **cfg_kw)
    seeds = pipeline.gt_seed_rows(seq, [m[0] for m in msgs])
    for e in (cpu, gpu):
        e.set_seed(seeds, seq.ba, seq.bg)
    t_imu, acc, gyr = seq.imu()
    fa, fb = pipeline.ImuFeeder(t_imu, acc, gyr), pipeline.ImuFeeder(t_imu, acc, gyr)
    D, td_ = rz(np.radians(yaw_deg)), np.asarray(t_drift)
    worst = dict(p=0.0, q=0.0, v=0.0, ba=0.0, bg=0.0)
    relos = []
    for k, (stamp, ids, d) in enumerate(msgs):
        fa.feed(cpu, stamp)
        fb.feed(gpu, stamp)
        if k in relo_at:
            # a frame of the window (exactly messages k-10 .. k-1 when all of them were key frames; a later index otherwise)
            j = k - 10 + window_index
            mp, p_old, R_old = synth.loop_frame_matches(seq, t_loop, msgs[j][1], pixel_sigma=0.3, seed=k)
            args = (msgs[j][0] + stamp_shift, 17, mp, D @ p_old + td_, D @ R_old)
            cpu.setReloFrame(*args)
            found = gpu.setReloFrame(*args)
            assert found == cpu.relo()["pending"] and gpu.relo()["pending"] == found
            if found:
                assert gpu.relo()["local_index"] == cpu.relo()["local_index"] >= window_index
        cpu.processImage(ids, d, stamp)
        gpu.processImage(ids, d, stamp)
        ia, ib = cpu.info(), gpu.info()
        for key in ("solver_flag", "frame_count", "marginalization_flag", "landmarks", "visual", "n_solves"):
            assert ia[key] == ib[key], (k, key, ia, ib)
        if ia["solver_flag"] != 1:
            continue
        assert (ia["iterations"], ia["successful_steps"]) == (ib["iterations"], ib["successful_steps"]), (k, ia, ib)
        sa, sb = cpu.states()[0], gpu.states()[0]
        worst["p"] = max(worst["p"], np.abs(sa[:, 0:3] - sb[:, 0:3]).max())
        worst["q"] = max(worst["q"], quat_angle(sa[:, 3:7], sb[:, 3:7]).max())
        worst["v"] = max(worst["v"], np.abs(sa[:, 7:10] - sb[:, 7:10]).max())
        worst["ba"] = max(worst["ba"], np.abs(sa[:, 10:13] - sb[:, 10:13]).max())
        worst["bg"] = max(worst["bg"], np.abs(sa[:, 13:16] - sb[:, 13:16]).max())
        if k in relo_at:
            relos.append((cpu.relo(), gpu.relo()))
    return worst, relos, cpu, gpu


def check_relo_pair(rc, rg, tol=1e-4):
    assert rc["factors"] == rg["factors"] and rc["solves"] == rg["solves"] and not rc["pending"] and not rg["pending"]
    assert np.abs(rc["drift_correct_r"] - rg["drift_correct_r"]).max() < tol
    assert np.abs(rc["drift_correct_t"] - rg["drift_correct_t"]).max() < tol
    assert np.abs(rc["relo_relative_t"] - rg["relo_relative_t"]).max() < tol
    assert quat_angle(rc["relo_relative_q"], rg["relo_relative_q"]) < tol
    assert abs(rc["relo_relative_yaw"] - rg["relo_relative_yaw"]) < np.degrees(tol)


def test_relocalisation_matches_oracle_and_recovers_the_drift():
    seq = synth.Sequence(seed=0, duration=6.0)
    msgs = synth.track_messages(seq, 45, max_feats=150)
    # pose_graph keeps sending matches for a loop: five consecutive frames
    worst, relos, cpu, gpu = run(seq, msgs, relo_at=(30, 31, 32, 33, 34))
    print("worst state deviations with relocalisation blocks", worst)
    for key, tol in TOL.items():
        assert worst[key] <= 2 * tol, (key, worst)
    assert len(relos) == 5
    for rc, rg in relos:
        check_relo_pair(rc, rg)
        assert rg["factors"] >= 20
    assert relos[-1][1]["solves"] == 5
    # the drift that was injected: 7 degrees of yaw and (0.4, -0.3, 0.1) m; a single solve starts from the window frame's own pose
    # (2 m away from the old frame) and gets most of the way, as in the reference
    r = relos[0][1]
    yaw = np.degrees(np.arctan2(r["drift_correct_r"][1, 0], r["drift_correct_r"][0, 0]))
    assert abs(yaw - 7.0) < 1.5 and np.abs(r["drift_correct_t"] - np.array([0.4, -0.3, 0.1])).max() < 0.08
    assert np.abs(r["drift_correct_r"][2] - np.array([0, 0, 1.0])).max() == 0  # yaw only (ypr2R(yaw, 0, 0))


def test_relocalisation_with_td_and_free_extrinsic():
    """relo_Pose next to the extrinsic and time-offset blocks (ProjectionTdFactor records): (ex, relo) and (td, relo) block pairs."""
    seq = synth.Sequence(seed=12, duration=5.0)
    msgs = synth.track_messages(seq, 38, max_feats=120)
    kw = dict(estimate_td=1, estimate_extrinsic=1, tr=0.02)
    worst, relos, cpu, gpu = run(seq, msgs, relo_at=(28, 30), cfg_kw=kw, window_index=3, t_loop=1.4, yaw_deg=-4.0, t_drift=(0.1, 0.2, -0.05))
    print("worst state deviations (td + extrinsic + relocalisation)", worst)
    for key, tol in TOL.items():
        assert worst[key] <= 3 * tol, (key, worst)
    assert len(relos) == 2
    for rc, rg in relos:
        check_relo_pair(rc, rg, tol=3e-4)
        assert rg["factors"] >= 10
    assert abs(cpu.states()[1] - gpu.states()[1]) < 1e-5


def test_relocalisation_stamp_outside_window_is_ignored():
    """setReloFrame with a stamp that matches no window frame leaves relocalization_info unset (estimator.cpp:1136-1145): the run is
    the run without a loop message."""
    seq = synth.Sequence(seed=3, duration=4.5)
    msgs = synth.track_messages(seq, 32, max_feats=100)
    worst, relos, cpu, gpu = run(seq, msgs, relo_at=(25,), stamp_shift=0.013)
    assert relos and relos[0][1]["solves"] == 0 and relos[0][1]["factors"] == 0
    for key, tol in TOL.items():
        assert worst[key] <= tol, (key, worst)
    from vins_mono_b200 import Estimator
    ref = Estimator(tic=synth.TIC, ric=synth.RIC)
    ref.set_seed(pipeline.gt_seed_rows(seq, [m[0] for m in msgs]), seq.ba, seq.bg)
    feeder = pipeline.ImuFeeder(*seq.imu())
    for stamp, ids, d in msgs:
        feeder.feed(ref, stamp)
        ref.processImage(ids, d, stamp)
    assert np.array_equal(ref.states()[0], gpu.states()[0])  # bit-identical: no trace of the ignored message


def test_relocalisation_inside_a_batch():
    """One member of a batch receives loop messages, its neighbours do not: the frame runs the relocalisation variant of the gather
    for everybody.  Every member must come out bit-identical to a stand-alone estimator fed the same way."""
    from vins_mono_b200 import Estimator, EstimatorBatch
    seq = synth.Sequence(seed=0, duration=5.0)
    msgs = synth.track_messages(seq, 36, max_feats=120)
    seeds = pipeline.gt_seed_rows(seq, [m[0] for m in msgs])
    t_imu, acc, gyr = seq.imu()
    relo_at, D, td_ = (28, 29, 31), rz(np.radians(5.0)), np.array([0.2, 0.1, -0.1])

    def loop_args(k):
        j = k - 10 + 4
        mp, p_old, R_old = synth.loop_frame_matches(seq, 1.2, msgs[j][1], pixel_sigma=0.3, seed=k)
        return msgs[j][0], 9, mp, D @ p_old + td_, D @ R_old

    solo = {}
    for with_relo in (False, True):
        e = Estimator(tic=synth.TIC, ric=synth.RIC)
        e.set_seed(seeds, seq.ba, seq.bg)
        f = pipeline.ImuFeeder(t_imu, acc, gyr)
        for k, (stamp, ids, d) in enumerate(msgs):
            f.feed(e, stamp)
            if with_relo and k in relo_at:
                assert e.setReloFrame(*loop_args(k))
            e.processImage(ids, d, stamp)
        solo[with_relo] = (e.states()[0].copy(), e.relo())
        e.close()
    assert solo[True][1]["solves"] == 3 and solo[False][1]["solves"] == 0
    assert not np.array_equal(solo[True][0], solo[False][0])  # the loop factors do move the window

    eb = EstimatorBatch(3, tic=synth.TIC, ric=synth.RIC)
    feeders = [pipeline.ImuFeeder(t_imu, acc, gyr) for _ in range(3)]
    for m in eb.members:
        m.set_seed(seeds, seq.ba, seq.bg)
    for k, (stamp, ids, d) in enumerate(msgs):
        for m, f in zip(eb.members, feeders):
            f.feed(m, stamp)
        if k in relo_at:
            assert eb.members[1].setReloFrame(*loop_args(k))
        assert not eb.processImage([(ids, d, stamp)] * 3).any()
    for j, with_relo in enumerate((False, True, False)):
        assert np.array_equal(eb.members[j].states()[0], solo[with_relo][0]), j
    r = eb.members[1].relo()
    assert r["solves"] == 3 and np.array_equal(r["drift_correct_t"], solo[True][1]["drift_correct_t"])
    assert eb.members[0].relo()["solves"] == 0
    eb.close()
