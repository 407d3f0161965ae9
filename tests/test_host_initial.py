"""The product's host C++ of the one-shot initialisation (vins_mono_b200/csrc/initial.cpp) through the handle-free
ve_debug_* entries against the OpenCV / numpy / scipy twin in oracle/initial.py, which computes every stage the way the reference
does (cv::findFundamentalMat + cv::recoverPose, cv::solvePnP, SVD triangulation, the visual-inertial alignment).  CPU only.

Stated tolerances: two-view pose 1e-9, PnP 1e-6 (OpenCV stops at FLT_EPSILON parameter change).  The vision-only bundle stops by
Ceres' rule (relative cost decrease below function_tolerance = 1e-6), i.e. a hair before the minimum; Ceres is not available, so
the twin solves the same problem to its minimum with scipy and the comparison is made twice: with the product's tolerance
tightened to 1e-14 (same minimum: equal cost to 1e-9 relative, poses 2e-4 -- the valley along the weakly constrained depths is
flat --, gyroscope bias 2e-5, scale / velocities 3e-3) and with the default (poses 1e-3 .. 2e-3, scale 2 %).  Measured: poses 3e-7 .. 6e-5,
scale 7e-6 .. 6e-4."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from harness import init_inputs, synth  # noqa: E402
from vins_mono_b200 import estimator as ve  # noqa: E402

cv2 = pytest.importorskip("cv2")
import initial as oi  # noqa: E402


def corres_between(tracks, a, b):
    out = []
    for _i, s, xy in tracks:
        if s <= a and s + len(xy) - 1 >= b:
            out.append(np.r_[xy[a - s], xy[b - s]])
    return np.array(out)


@pytest.mark.parametrize("seed,a", [(0, 0), (1, 2), (3, 5), (5, 0)])
def test_relative_rt_matches_opencv(seed, a):
    seq = synth.Sequence(seed=seed, duration=3.0)
    _, _, tracks = init_inputs.first_window(seq)
    c = corres_between(tracks, a, 10)
    ok_c, R_c, T_c, cnt_c = oi.solve_relative_rt(c)
    ok, R, T, cnt = ve.debug_relative_rt(c)
    assert ok == ok_c and cnt == cnt_c and cnt > 12
    assert np.abs(R - R_c).max() < 1e-9 and np.abs(T - T_c).max() < 1e-9
    assert abs(np.linalg.det(R) - 1) < 1e-12 and abs(np.linalg.norm(T) - 1) < 1e-12


def test_relative_rt_rejects_degenerate_input():
    rng = np.random.default_rng(0)
    few = rng.uniform(-0.3, 0.3, (14, 4))
    assert ve.debug_relative_rt(few)[0] is False and oi.solve_relative_rt(few)[0] is False
    # no motion at all (every point at the same place in both views): OpenCV 4.13 raises inside findFundamentalMat, i.e. the
    # reference would abort; the product reports failure
    same = np.tile(rng.uniform(-0.3, 0.3, (40, 2)), (1, 2))
    with pytest.raises(cv2.error):
        oi.solve_relative_rt(same)
    assert ve.debug_relative_rt(same)[0] is False


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_solve_pnp_matches_opencv(seed):
    rng = np.random.default_rng(seed)
    from scipy.spatial.transform import Rotation
    R_true = Rotation.from_rotvec(rng.normal(0, 0.3, 3)).as_matrix()
    t_true = rng.normal(0, 0.5, 3)
    n = 40 if seed else 12
    Xc = np.c_[rng.uniform(-1, 1, n), rng.uniform(-0.7, 0.7, n), rng.uniform(2, 8, n)]
    Xw = (Xc - t_true) @ R_true  # R_true^T (Xc - t)
    uv = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, 0.3 / 460, (n, 2))
    R0 = Rotation.from_rotvec(rng.normal(0, 0.05, 3)).as_matrix() @ R_true
    t0 = t_true + rng.normal(0, 0.1, 3)
    ok_c, R_c, t_c = oi.solve_pnp(Xw, uv, R0, t0)
    ok, R, t = ve.debug_solve_pnp(Xw, uv, R0, t0)
    assert ok and ok_c
    assert np.abs(R - R_c).max() < 1e-6 and np.abs(t - t_c).max() < 1e-6
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12


@pytest.mark.parametrize("seed", [0, 3])
def test_sfm_construct_matches_twin(seed):
    seq = synth.Sequence(seed=seed, duration=3.0)
    _, _, tracks = init_inputs.first_window(seq)
    ok, rel_R, rel_T, l = oi.relative_pose(tracks, 10)
    assert ok
    ok_c, q_c, T_c, pts_c, cost_c = oi.sfm_construct(11, l, rel_R, rel_T, tracks)
    assert ok_c
    for tol, pose_tol, pt_tol in ((1e-14, 2e-4, 2e-2), (0.0, 1e-3, 5e-2)):
        out = ve.debug_sfm_construct(11, l, rel_R, rel_T, tracks, function_tolerance=tol)
        assert out["ok"] and 1 <= out["iterations"] <= 50
        assert set(out["points"]) == set(pts_c)
        for i in range(11):
            dq = min(np.abs(out["q"][i] - q_c[i]).max(), np.abs(out["q"][i] + q_c[i]).max())
            assert dq < pose_tol and np.abs(out["T"][i] - T_c[i]).max() < pose_tol, (tol, i)
        # the gauge: frame l at the origin, |T_last| = |relative_T| = 1
        assert np.abs(out["T"][l]).max() == 0 and abs(np.linalg.norm(out["T"][10]) - 1) < 1e-12
        worst = max(np.abs(out["points"][i] - pts_c[i]).max() / max(1.0, np.abs(pts_c[i]).max()) for i in pts_c)
        assert worst < pt_tol, (tol, worst)  # poorly constrained depths move most; the poses are what the alignment consumes
        assert out["cost"] <= cost_c * (1 + (1e-9 if tol else 1e-3)) + 1e-15


@pytest.mark.parametrize("seed,drop", [(0, ()), (3, ()), (0, (9,)), (3, (4, 8))])
def test_initial_structure_matches_twin(seed, drop):
    """Whole initialStructure (up to VisualIMUAlignment) incl. frames that left the window as non-keyframes but stay in
    all_image_frame (PnP branch, estimator.cpp:306-356)."""
    seq = synth.Sequence(seed=seed, duration=3.0)
    headers, frames, tracks = init_inputs.first_window(seq, drop=drop)
    oframes = init_inputs.oracle_frames(frames)
    ref = oi.initial_structure(oframes, headers, tracks, synth.RIC, synth.TIC, synth.G_NORM)
    assert ref["code"] == 0 and sum(f.is_key_frame for f in oframes) == 11
    s_c = ref["x"][-1]
    for tol, pose_tol, bg_tol, g_tol, rel in ((1e-14, 2e-4, 2e-5, 1e-3, 3e-3), (0.0, 2e-3, 2e-5, 5e-2, 2e-2)):
        res = ve.debug_initial_structure(headers, frames, tracks, synth.RIC, synth.TIC, synth.G_NORM, function_tolerance=tol)
        assert res["code"] == 0 and res["l"] == ref["l"] and res["key_frames"] == 11
        for k, f in enumerate(oframes):
            assert np.abs(res["R"][k] - f.R).max() < pose_tol and np.abs(res["T"][k] - f.T).max() < pose_tol, (tol, k)
        assert np.abs(res["delta_bg"] - ref["delta_bg"]).max() < bg_tol
        assert np.abs(res["g"] - ref["g"]).max() < g_tol and abs(np.linalg.norm(res["g"]) - synth.G_NORM) < 1e-9
        s = res["x"][-1]
        assert s > 0 and abs(s - s_c) < rel * s_c, (tol, s, s_c)
        assert np.abs(res["x"][:-3] - ref["x"][:-3]).max() < rel * max(1.0, np.abs(ref["x"][:-3]).max())  # body-frame velocities


def test_initialisation_recovers_truth_without_noise():
    """Noise-free measurements: gyroscope bias, gravity direction and metric scale come out at their true values (the scale up to
    the accelerometer bias the linear alignment does not model)."""
    seq = synth.Sequence(seed=0, duration=3.0, imu_noise=False)
    headers, frames, tracks = init_inputs.first_window(seq, pixel_sigma=0.0)
    res = ve.debug_initial_structure(headers, frames, tracks, synth.RIC, synth.TIC, synth.G_NORM)
    assert res["code"] == 0
    assert np.abs(res["delta_bg"] - seq.bg).max() < 2e-5
    l = res["l"]
    pose = [seq.pose(t) for t in headers]
    cam = [p[0] + p[1] @ synth.TIC for p in pose]
    true_scale = np.linalg.norm(cam[10] - cam[l])
    assert abs(res["x"][-1] / true_scale - 1) < 0.08
    # gravity is expressed in the camera frame of window frame l: g_c = (R_wb R_ic)^T g_w
    g_true = (pose[l][1] @ synth.RIC).T @ np.array([0, 0, synth.G_NORM])
    cosang = res["g"] @ g_true / (np.linalg.norm(res["g"]) * synth.G_NORM)
    assert np.degrees(np.arccos(min(1.0, cosang))) < 1.0
    # rotations of the window frames against ground truth (relative to frame l): exact up to the pixel quantisation of float
    for k in range(11):
        R_rel_true = (pose[l][1] @ synth.RIC).T @ pose[k][1]
        assert np.abs(res["R"][k] - R_rel_true).max() < 1e-5


def test_initial_structure_reports_missing_parallax():
    """A hovering camera: relativePose finds no frame pair with 30 px of parallax (estimator.cpp:442-471)."""
    seq = synth.Sequence(seed=0, duration=3.0)
    headers, frames, tracks = init_inputs.first_window(seq)
    still = [(i, s, np.repeat(xy[:1], len(xy), axis=0)) for i, s, xy in tracks]
    res = ve.debug_initial_structure(headers, frames, still, synth.RIC, synth.TIC, synth.G_NORM)
    assert res["code"] == 1


# ---- ESTIMATE_EXTRINSIC == 2: InitialEXRotation (initial/initial_ex_rotation.cpp) ----------------------------------------------
def _ex_rotation_inputs(seq, n_frames, pixel_sigma=0.3):
    _, frames, _ = init_inputs.first_window(seq, n_window=n_frames, pixel_sigma=pixel_sigma)
    of = init_inputs.oracle_frames(frames)
    corres, dqs = [], []
    for k in range(1, len(frames)):
        a = dict(zip(frames[k - 1]["ids"], frames[k - 1]["xy"]))
        corres.append(np.array([np.r_[a[i], xy] for i, xy in zip(frames[k]["ids"], frames[k]["xy"]) if i in a]))
        dqs.append(of[k].pre.dq)
    return corres, dqs


def _angle_deg(Ra, Rb):
    return np.degrees(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def test_ex_rotation_linear_system_matches_twin():
    """The quaternion system of CalibrationExRotation (Huber weights against the running estimate, 4N x 4 stack of L(q_cam) - R(q_imu),
    smallest right singular vector, the observability threshold) with the camera rotations handed in: exact pairs
    q_imu (x) q_ic = q_ic (x) q_cam plus two outliers that the Huber weights must tame."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    q_ic = Rotation.from_rotvec([0.3, -1.1, 0.5])
    dqs, rcs = [], []
    for k in range(14):
        # the first pairs rotate by less than 2.5 degrees: their Huber weight is 1 whatever the (still ambiguous) running estimate
        # is; later pairs are large, their weights depend on an estimate that has converged by then
        rv = rng.normal(0, 0.15, 3) if k >= 4 else 0.04 * rng.normal(0, 1, 3) / 3
        r_cam = Rotation.from_rotvec(rv)
        r_imu = q_ic * r_cam * q_ic.inv()
        if k in (5, 9):  # an outlier pair: 20 degrees off
            r_cam = Rotation.from_rotvec([0.35, 0, 0]) * r_cam
        x, y, z, w = r_imu.as_quat()
        dqs.append([w, x, y, z] if w >= 0 else [-w, -x, -y, -z])
        rcs.append(r_cam.as_matrix())
    corres = [np.zeros((0, 4))] * 14
    ric, ok, cov, rcam = ve.debug_ex_rotation(corres, dqs, 10, rc_given=rcs)
    cal = oi.ExRotation()
    for k in range(14):
        ok_c, ric_c = cal.calibrate(corres[k], dqs[k], 10, rc_given=rcs[k])
        assert np.array_equal(rcam[k], rcs[k])
        assert ok[k] == ok_c and abs(cov[k] - cal.cov1) < 1e-10
        if k >= 2:  # the null direction is unique once rotations about different axes are stacked
            assert np.abs(ric[k] - ric_c).max() < 1e-9, k
    assert ok[-1] and not ok[:9].any()
    assert _angle_deg(ric[-1], q_ic.as_matrix()) < 1.0  # ric = R(q_ic): rotation from the camera frame to the IMU frame


@pytest.mark.parametrize("seed", [0, 1])
def test_ex_rotation_calibrates_under_rotation(seed):
    """Camera rotations out of cv::findFundamentalMat's DEFAULT call (RANSAC threshold 3 on normalised coordinates: every hypothesis
    explains every point, the first 7-point sample's first root wins).  Which root comes first depends on OpenCV's SVD null-space
    basis, so the product's and the twin's per-step rotations differ in detail; the calibration converges to the true camera-IMU
    rotation for both, at (nearly) the same step."""
    seq = synth.Sequence(seed=seed, duration=5.0, rot_gain=5.0)
    corres, dqs = _ex_rotation_inputs(seq, 32)
    ric, ok, cov, rcam = ve.debug_ex_rotation(corres, dqs, 10)
    cal, ok_c, ric_c = oi.ExRotation(), [], []
    for c, dq in zip(corres, dqs):
        o, r = cal.calibrate(c, dq, 10)
        ok_c.append(o)
        ric_c.append(r.copy())
    assert ok.any() and any(ok_c)
    first, first_c = int(np.argmax(ok)), ok_c.index(True)
    assert first >= 9 and first_c >= 9 and abs(first - first_c) <= 2      # never before frame_count >= WINDOW_SIZE
    assert not ok[:9].any()
    assert _angle_deg(ric[first], synth.RIC) < 2.5 and _angle_deg(ric_c[first_c], synth.RIC) < 2.5
    assert _angle_deg(ric[-1], synth.RIC) < 2.0 and cov[-1] > 0.25
    for k in range(len(rcam)):  # every extracted camera rotation is a proper rotation
        assert abs(np.linalg.det(rcam[k]) - 1) < 1e-9 and np.abs(rcam[k] @ rcam[k].T - np.eye(3)).max() < 1e-9


def test_ransac_model_matches_opencv_at_the_initialisation_threshold():
    """The F matrix behind solveRelativeRT (RANSAC winner at 0.3 / 460): identical inlier masks and the same model as OpenCV."""
    import ctypes as C
    from vins_mono_b200 import load_library
    lib = load_library()
    seq = synth.Sequence(seed=0, duration=3.0)
    corres, _ = _ex_rotation_inputs(seq, 12)
    for c in corres:
        ll, rr = np.ascontiguousarray(c[:, :2], np.float32), np.ascontiguousarray(c[:, 2:], np.float32)
        E, m = cv2.findFundamentalMat(ll, rr, cv2.FM_RANSAC, 0.3 / 460, 0.99)
        F, st = np.zeros(9), np.zeros(len(ll), np.uint8)
        assert lib.vt_debug_fundamental_ransac_model(ll.ctypes.data_as(C.c_void_p), rr.ctypes.data_as(C.c_void_p), len(ll), C.c_double(0.3 / 460),
                                                     C.c_double(0.99), st.ctypes.data_as(C.c_void_p), F.ctypes.data_as(C.c_void_p)) == 1
        assert np.array_equal(m.ravel() != 0, st != 0)
        assert np.abs(E - F.reshape(3, 3)).max() < 1e-6 * max(1.0, np.abs(E).max())


# ---- properties on exact data (general 3-D scenes, no noise): the stages recover the geometry they were given ----------------------
def _random_scene(rng, n=80):
    from scipy.spatial.transform import Rotation
    X = np.c_[rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(3, 9, n)]     # a cloud with real depth variation
    R = Rotation.from_rotvec(rng.normal(0, 0.12, 3)).as_matrix()                       # camera 1 expressed in camera 0
    t = rng.normal(0, 0.4, 3) + np.array([0.5, 0, 0])
    X1 = (X - t) @ R                                                                   # R^T (X - t)
    return X, R, t, X[:, :2] / X[:, 2:3], X1[:, :2] / X1[:, 2:3]


@pytest.mark.parametrize("seed", range(4))
def test_relative_rt_recovers_exact_two_view_geometry(seed):
    rng = np.random.default_rng(100 + seed)
    X, R, t, x0, x1 = _random_scene(rng)
    ok, Rot, Tr, cnt = ve.debug_relative_rt(np.c_[x0, x1])
    ok_c, Rot_c, Tr_c, cnt_c = oi.solve_relative_rt(np.c_[x0, x1])
    assert ok and ok_c and cnt == cnt_c == len(X)           # float32 rounding of the inputs stays far below the threshold
    assert np.abs(Rot - Rot_c).max() < 1e-8 and np.abs(Tr - Tr_c).max() < 1e-8
    # the reference's convention: Rotation = pose of camera 1 in camera 0, Translation its (unit) position
    assert np.abs(Rot - R).max() < 1e-4 and np.abs(Tr - t / np.linalg.norm(t)).max() < 1e-3


@pytest.mark.parametrize("seed", range(3))
def test_solve_pnp_recovers_exact_pose(seed):
    rng = np.random.default_rng(200 + seed)
    X, R, t, _, x1 = _random_scene(rng, 30)
    from scipy.spatial.transform import Rotation
    R_wc, t_wc = R.T, -R.T @ t                               # world (= camera 0) -> camera 1
    R0 = Rotation.from_rotvec(rng.normal(0, 0.08, 3)).as_matrix() @ R_wc
    ok, Rp, tp = ve.debug_solve_pnp(X, x1, R0, t_wc + rng.normal(0, 0.2, 3))
    assert ok and np.abs(Rp - R_wc).max() < 2e-6 and np.abs(tp - t_wc).max() < 2e-5   # inputs pass through float32
    assert abs(np.linalg.det(Rp) - 1) < 1e-12


def test_sfm_construct_recovers_a_known_structure():
    """Five cameras on a curved path over a general point cloud, exact observations: poses and points come back in the frame of
    camera l with |T_last| = 1 (the gauge GlobalSFM fixes)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(7)
    n_frames, n_pts = 5, 90
    X = np.c_[rng.uniform(-3, 3, n_pts), rng.uniform(-2, 2, n_pts), rng.uniform(4, 10, n_pts)]
    Rs = [Rotation.from_rotvec([0.02 * k, 0.05 * k, -0.01 * k]).as_matrix() for k in range(n_frames)]   # camera k in world
    Ts = [np.array([0.25 * k, 0.05 * k * k, 0.03 * k]) for k in range(n_frames)]
    tracks = []
    for i, p in enumerate(X):
        xy = []
        for k in range(n_frames):
            pc = Rs[k].T @ (p - Ts[k])
            xy.append(pc[:2] / pc[2])
        tracks.append((i, 0, np.array(xy)))
    l = 0
    rel_R = Rs[l].T @ Rs[-1]
    rel_T = Rs[l].T @ (Ts[-1] - Ts[l])
    scale = np.linalg.norm(rel_T)
    out = ve.debug_sfm_construct(n_frames, l, rel_R, rel_T / scale, tracks, function_tolerance=1e-14)
    assert out["ok"] and len(out["points"]) == n_pts and out["cost"] < 1e-12
    for k in range(n_frames):
        q = out["q"][k]
        Rk = oi.quat_to_R(q)
        assert np.abs(Rk - Rs[l].T @ Rs[k]).max() < 1e-5
        assert np.abs(out["T"][k] - Rs[l].T @ (Ts[k] - Ts[l]) / scale).max() < 1e-5
    worst = max(np.abs(out["points"][i] - Rs[l].T @ (X[i] - Ts[l]) / scale).max() for i in range(n_pts))
    assert worst < 1e-4


@pytest.mark.parametrize("seed", [0, 3])
def test_whole_initialisation_is_metric_on_exact_data(seed):
    """No pixel noise, no IMU noise, no accelerometer bias (the one quantity the linear alignment does not model): scale, body-frame
    velocities, gravity and the gyroscope bias come out at their true values up to the mid-point integration error."""
    seq = synth.Sequence(seed=seed, duration=3.0, imu_noise=False)
    seq.ba = np.zeros(3)
    headers, frames, tracks = init_inputs.first_window(seq, pixel_sigma=0.0)
    res = ve.debug_initial_structure(headers, frames, tracks, synth.RIC, synth.TIC, synth.G_NORM, function_tolerance=1e-14)
    assert res["code"] == 0
    pose = [seq.pose(t) for t in headers]
    cam = [p[0] + p[1] @ synth.TIC for p in pose]
    l = res["l"]
    assert abs(res["x"][-1] / np.linalg.norm(cam[10] - cam[l]) - 1) < 1e-4
    for k in range(11):
        assert np.abs(res["x"][3 * k: 3 * k + 3] - pose[k][1].T @ pose[k][2]).max() < 1e-4      # R_wb^T v_w
    assert np.abs(res["delta_bg"] - seq.bg).max() < 1e-5
    g_true = (pose[l][1] @ synth.RIC).T @ np.array([0, 0, synth.G_NORM])
    assert np.abs(res["g"] - g_true).max() < 1e-3
