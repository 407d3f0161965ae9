"""Pins the back-end oracle (oracle/be_*.cpp).  The reference ships no tests or golden vectors for this path
and cannot be compiled offline (ROS/Ceres/Eigen), so the pins are: the known answers of SURVEY.md Appendix E
(an independent numpy restatement made during the survey), finite differences with the reference's own
check() recipe (projection_factor.cpp:176-224: right-multiplicative deltaQ perturbations), and numpy linear
algebra for the small dense kernels."""
import numpy as np
import pytest

import orc


def qnorm(w, x, y, z):
    q = np.array([w, x, y, z], float)
    return q / np.linalg.norm(q)


def qmul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def qrot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def pose(p, q_wxyz):  # Ceres layout: p, qx qy qz qw
    return np.array([p[0], p[1], p[2], q_wxyz[1], q_wxyz[2], q_wxyz[3], q_wxyz[0]])


E1 = dict(Pi=(0.1, -0.2, 0.3), qi=qnorm(0.98, 0.05, -0.1, 0.15), Pj=(0.5, -0.1, 0.35), qj=qnorm(0.97, 0.02, -0.15, 0.18),
          tic=(-0.02, -0.06, 0.01), qic=(0.5, -0.5, 0.5, -0.5), lam=0.25, pts_i=(0.12, -0.07, 1), pts_j=(0.05, -0.11, 1))


def e1_params(td=0.0):
    return np.concatenate([pose(E1["Pi"], E1["qi"]), pose(E1["Pj"], E1["qj"]), pose(E1["tic"], E1["qic"]), [E1["lam"]], [td]])


def test_projection_factor_known_answer():
    res, J = orc.projection_factor(e1_params(), np.r_[E1["pts_i"], E1["pts_j"]])
    assert np.allclose(res, [47.3616956445, 37.3001252650], rtol=0, atol=1e-8)
    assert np.allclose(J[3].ravel(), [11.2680305646, -23.7366784762], rtol=0, atol=1e-8)
    assert np.all(J[0][:, 6] == 0) and np.all(J[1][:, 6] == 0) and np.all(J[2][:, 6] == 0)


def test_projection_td_factor_known_answer():
    data = np.r_[E1["pts_i"], E1["pts_j"], 0.3, -0.1, 0.25, -0.15, 0.001, 0.002, 200, 310]
    res, J = orc.projection_factor(e1_params(td=0.003), data, use_td=True, TR=0.033, ROW=480.0)
    assert np.allclose(res, [47.8872029413, 37.0120033366], rtol=0, atol=1e-8)


@pytest.mark.parametrize("use_td", [False, True])
def test_projection_jacobians_finite_differences(use_td):
    """The reference's check() recipe, turned into an assertion."""
    data = np.r_[E1["pts_i"], E1["pts_j"], 0.3, -0.1, 0.25, -0.15, 0.001, 0.002, 200, 310]
    p0 = e1_params(td=0.003)
    kw = dict(use_td=use_td, TR=0.033, ROW=480.0)
    r0, J = orc.projection_factor(p0, data, **kw)
    eps = 1e-6
    num = []
    for blk, off in ((0, 0), (1, 7), (2, 14)):
        cols = []
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            p = p0.copy()
            p[off:off + 7] = orc.pose_plus(p0[off:off + 7], d)
            cols.append((orc.projection_factor(p, data, **kw)[0] - r0) / eps)
        num.append(np.array(cols).T)
    for k, n_ in enumerate(num):
        assert np.abs(J[k][:, :6] - n_).max() < 2e-3 * max(1.0, np.abs(n_).max())
    for idx, jk in ((21, 3), (22, 4)):
        if jk == 4 and not use_td:
            continue
        p = p0.copy()
        p[idx] += eps
        n_ = (orc.projection_factor(p, data, **kw)[0] - r0) / eps
        assert np.abs(J[jk].ravel() - n_).max() < 2e-3 * max(1.0, np.abs(n_).max())


def e3_samples():
    t = 0.005 * np.arange(21)
    acc = np.c_[0.3 * np.sin(3 * t) + 0.1, 0.2 * np.cos(2 * t), 9.8 + 0.4 * np.sin(5 * t)]
    gyr = np.c_[0.2 * np.sin(2 * t), -0.1 + 0.15 * np.cos(3 * t), 0.3 * np.sin(t + 0.5)]
    return np.full(21, 0.005), acc, gyr


BA, BG = np.array([0.01, -0.02, 0.03]), np.array([0.002, -0.001, 0.003])


def test_preintegration_known_answer():
    dt, acc, gyr = e3_samples()
    r = orc.preintegrate(orc.be_config(), BA, BG, dt, acc, gyr)
    assert abs(r["sum_dt"] - 0.1) < 1e-15
    # SURVEY E3.  The survey's numpy restatement normalised result_delta_q before rotating acc_1; the reference
    # rotates with the un-normalised product (integration_base.h:65-66), which moves delta_p/delta_v by <= 2.2e-10.
    assert np.allclose(r["delta_p"], [0.0006770624474, 0.0010870152918, 0.0491786337654], rtol=0, atol=5e-10)
    assert np.allclose(r["delta_q"], [0.99996708142916, 8.9760247017093e-4, 2.4401899731221e-3, 7.6860809741231e-3], rtol=0, atol=1e-12)
    assert np.allclose(r["delta_v"], [0.0157756965189, 0.0214284275725, 0.9867639056293], rtol=0, atol=5e-10)
    # independent numpy midpoint integration with Eigen's quaternion semantics (q * v = v + 2w(u x v) + 2u x (u x v))
    def tv(q, v):
        u = q[1:]
        uv = 2 * np.cross(u, v)
        return v + q[0] * uv + np.cross(u, uv)
    dp, dv, dq, a0, g0 = np.zeros(3), np.zeros(3), np.array([1.0, 0, 0, 0]), acc[0], gyr[0]
    for k in range(1, 21):
        d = dt[k]
        un_acc_0, un_gyr = tv(dq, a0 - BA), 0.5 * (g0 + gyr[k]) - BG
        rq = qmul(dq, np.r_[1.0, un_gyr * d / 2])
        un_acc = 0.5 * (un_acc_0 + tv(rq, acc[k] - BA))
        dp, dv = dp + dv * d + 0.5 * un_acc * d * d, dv + un_acc * d
        dq, a0, g0 = rq / np.linalg.norm(rq), acc[k], gyr[k]
    assert np.allclose(r["delta_p"], dp, rtol=0, atol=1e-15) and np.allclose(r["delta_v"], dv, rtol=0, atol=1e-15)
    assert np.allclose(r["delta_q"], dq, rtol=0, atol=1e-15)
    J, Pm = r["jacobian"], r["covariance"]
    assert np.allclose(np.diag(J[0:3, 9:12]), [-0.0049998969316, -0.0049999069595, -0.0049999889948], atol=1e-12)
    assert np.allclose(np.diag(J[3:6, 12:15]), [-0.099996154946, -0.0999964052184, -0.0999995944908], atol=1e-11)
    assert np.allclose(J[6, 12:15], [-0.0002642362375, -0.0494977119171, 0.0010701015862], atol=1e-12)
    for (i, j), v in {(0, 0): 5.331944432336e-9, (0, 6): 8.004877663549e-8, (3, 3): 4.000025972708e-9, (6, 6): 1.601304452229e-6,
                      (0, 4): 6.586152335579e-11, (9, 9): 8e-13, (12, 12): 2e-15}.items():
        assert abs(Pm[i, j] - v) <= 1e-9 * abs(v) + 1e-24
    d = np.diag(r["sqrt_info"])
    assert np.allclose(d, np.repeat([2.742e4, 1.582e4, 7.90e2, 1.118e6, 2.236e7], 3), rtol=2e-3)
    # sqrt_info^T sqrt_info = covariance^-1
    assert np.allclose(r["sqrt_info"].T @ r["sqrt_info"] @ Pm, np.eye(15), atol=1e-6)


def e4_params():
    dt, acc, gyr = e3_samples()
    pre = orc.preintegrate(orc.be_config(), BA, BG, dt, acc, gyr)
    G = np.array([0, 0, 9.81007])
    T = 0.1
    Pi, qi = np.array(E1["Pi"]), E1["qi"]
    Vi = np.array([0.4, -0.1, 0.05])
    Bai, Bgi = BA + [0.003, -0.002, 0.001], BG + [2e-4, -1e-4, 3e-4]
    dq = pre["delta_q"]
    th = np.array([0.002, -0.001, 0.0015])
    qj = qmul(qmul(qi, dq), np.r_[1.0, th / 2])
    qj /= np.linalg.norm(qj)
    Ri = qrot(qi)
    Pj = Pi + Vi * T - 0.5 * G * T * T + Ri @ pre["delta_p"] + [0.002, -0.001, 0.001]
    Vj = Vi - G * T + Ri @ pre["delta_v"] + [0.01, 0.005, -0.004]
    Baj, Bgj = BA + [0.0031, -0.0019, 0.0012], BG + [2.1e-4, -0.9e-4, 3.1e-4]
    return np.concatenate([pose(Pi, qi), Vi, Bai, Bgi, pose(Pj, qj), Vj, Baj, Bgj])


def test_imu_factor_known_answer():
    dt, acc, gyr = e3_samples()
    res, raw, J = orc.imu_factor(orc.be_config(), BA, BG, dt, acc, gyr, e4_params())
    raw_ref = [1.8109449737e-3, -1.5025145652e-3, 7.4478498460e-4, 2.0198619992e-3, -1.0100980168e-3, 1.5300557258e-3,
               1.0222171855e-2, 1.2140511465e-3, -6.2614221868e-3, 1e-4, 1e-4, 2e-4, 1e-5, 1e-5, 1e-5]
    assert np.allclose(raw, raw_ref, rtol=0, atol=2e-12)
    white_ref = [35.4091665398, -43.3215553408, 29.0168050345, 31.9771687361, -16.1713117056, 24.2041320068, 8.0818331730,
                 0.9629815745, -4.9425859754, 111.803398875, 111.803398875, 223.60679775, 223.60679775, 223.60679775, 223.60679775]
    assert np.allclose(res, white_ref, rtol=1e-8, atol=1e-7)


def test_imu_jacobians_finite_differences():
    dt, acc, gyr = e3_samples()
    cfg = orc.be_config()
    p0 = e4_params()
    r0, raw0, J = orc.imu_factor(cfg, BA, BG, dt, acc, gyr, p0)
    si = orc.preintegrate(cfg, BA, BG, dt, acc, gyr)["sqrt_info"]
    eps = 1e-7
    for blk, (off, size, is_pose) in enumerate([(0, 7, True), (7, 9, False), (16, 7, True), (23, 9, False)]):
        local = 6 if is_pose else size
        num = np.zeros((15, local))
        for k in range(local):
            p = p0.copy()
            if is_pose:
                d = np.zeros(6)
                d[k] = eps
                p[off:off + 7] = orc.pose_plus(p0[off:off + 7], d)
            else:
                p[off + k] += eps
            num[:, k] = (orc.imu_factor(cfg, BA, BG, dt, acc, gyr, p)[1] - raw0) / eps
        ana = np.linalg.solve(si, J[blk][:, :local])    # un-whiten
        assert np.abs(ana - num).max() < 5e-6 * max(1.0, np.abs(num).max()), f"block {blk}"


def test_small_dense_kernels():
    rng = np.random.default_rng(0)
    B = rng.normal(size=(40, 40))
    A = B @ B.T + np.diag(10.0 ** rng.uniform(-6, 6, 40))
    for solver in (orc.sym_eigen, orc.sym_eigen_ql):   # cyclic Jacobi (the oracle) and tridiagonal QL (the kernel's template)
        w, V = solver(A)
        assert np.allclose(w, np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-9 * np.abs(A).max())
        assert np.allclose(V @ np.diag(w) @ V.T, A, rtol=1e-10, atol=1e-9 * np.abs(A).max())
        assert np.allclose(V.T @ V, np.eye(40), atol=1e-12)


def test_oracle_relocalisation_recovers_injected_drift():
    """The oracle's setReloFrame / relo_Pose block / drift bookkeeping (estimator.cpp:1128-1146, :769-801, :598-617) against physics: a
    loop message whose old key frame is reported 7 degrees of yaw and (0.4, -0.3, 0.1) m away from where the odometry frame puts it
    yields that drift (a single solve starts relo_Pose at the window frame's own pose, 2 m off, and gets most of the way)."""
    from harness import pipeline, synth
    seq = synth.Sequence(seed=0, duration=5.0)
    msgs = synth.track_messages(seq, 34, max_feats=150)
    est = orc.OracleEstimator(orc.be_config())
    est.set_seed(pipeline.gt_seed_rows(seq, [m[0] for m in msgs]), seq.ba, seq.bg)
    feeder = pipeline.ImuFeeder(*seq.imu())
    a = np.radians(7.0)
    D = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    t_drift = np.array([0.4, -0.3, 0.1])
    before = None
    for k, (stamp, ids, d) in enumerate(msgs):
        feeder.feed(est, stamp)
        if k == 29:  # a stamp that is not in the window: nothing is armed
            mp, p_old, R_old = synth.loop_frame_matches(seq, 1.0, msgs[24][1])
            est.setReloFrame(msgs[24][0] + 0.0123, 3, mp, D @ p_old + t_drift, D @ R_old)
            assert not est.relo()["pending"]
        if k == 30:
            mp, p_old, R_old = synth.loop_frame_matches(seq, 1.0, msgs[25][1], pixel_sigma=0.3)
            est.setReloFrame(msgs[25][0], 3, mp, D @ p_old + t_drift, D @ R_old)
            r = est.relo()
            assert r["pending"] and 0 <= r["local_index"] < 10 and r["solves"] == 0
            before = est.info()["visual"]
        est.processImage(ids, d, stamp)
        if k == 30:
            r = est.relo()
            assert not r["pending"] and r["solves"] == 1 and 20 <= r["factors"] <= len(mp)
            assert abs(est.info()["visual"] - before) < 200     # f_m_cnt counts the window's own factors only
            yaw = np.degrees(np.arctan2(r["drift_correct_r"][1, 0], r["drift_correct_r"][0, 0]))
            assert abs(yaw - 7.0) < 1.5 and np.abs(r["drift_correct_t"] - t_drift).max() < 0.08
            assert np.array_equal(r["drift_correct_r"][2], [0, 0, 1.0]) and abs(np.linalg.norm(r["relo_relative_q"]) - 1) < 1e-12
    assert est.relo()["solves"] == 1 and est.info()["solver_flag"] == 1
