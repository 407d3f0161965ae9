"""ctypes bindings of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY: nothing
under vins_mono_b200/ may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

u8p, i32p, f32p, f64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int, C.c_float, C.c_double))


def P(a, t):
    return a.ctypes.data_as(t)


class TrackerConfig(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("max_cnt", C.c_int), ("min_dist", C.c_int),
                ("equalize", C.c_int), ("freq", C.c_int), ("focal_length", C.c_int), ("fisheye", C.c_int),
                ("f_threshold", C.c_double), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double),
                ("cy", C.c_double), ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("camera_model", C.c_int), ("xi", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle"))
                if f.endswith((".cpp", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(so)
        _LIB.orc_tracker_create.restype = C.c_void_p
    return _LIB


def clahe(img, clip=3.0, tiles=(8, 8)):
    img = np.ascontiguousarray(img)
    out = np.empty_like(img)
    lib().orc_clahe(P(img, u8p), img.shape[0], img.shape[1], img.shape[1], C.c_double(clip), tiles[0], tiles[1],
                    P(out, u8p), img.shape[1])
    return out


def pyrdown(img):
    img = np.ascontiguousarray(img)
    out = np.empty(((img.shape[0] + 1) // 2, (img.shape[1] + 1) // 2), np.uint8)
    lib().orc_pyrdown(P(img, u8p), img.shape[0], img.shape[1], img.shape[1], P(out, u8p), out.shape[1])
    return out


def min_eig(img):
    img = np.ascontiguousarray(img)
    out = np.empty(img.shape, np.float32)
    lib().orc_min_eig(P(img, u8p), img.shape[0], img.shape[1], img.shape[1], P(out, f32p))
    return out


def gftt(img, max_corners, quality, min_dist, mask=None):
    img = np.ascontiguousarray(img)
    out = np.zeros((max(max_corners, 1), 2), np.float32)
    nc = C.c_int(0)
    mp = P(np.ascontiguousarray(mask), u8p) if mask is not None else None
    n = lib().orc_gftt(P(img, u8p), img.shape[0], img.shape[1], img.shape[1], mp, img.shape[1], max_corners,
                       C.c_double(quality), C.c_double(min_dist), P(out, f32p), C.byref(nc))
    return out[:n].copy(), nc.value


def lk(prev, nxt, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig_thr=1e-4):
    prev, nxt = np.ascontiguousarray(prev), np.ascontiguousarray(nxt)
    pts = np.ascontiguousarray(pts, np.float32)
    n = len(pts)
    out = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    lib().orc_lk(P(prev, u8p), P(nxt, u8p), prev.shape[0], prev.shape[1], prev.shape[1], P(pts, f32p), n, win,
                 max_level, max_iter, C.c_double(eps), C.c_double(min_eig_thr), P(out, f32p), P(st, u8p))
    return out, st


def circle(mask, cx, cy, r, color=0):
    lib().orc_circle(P(mask, u8p), mask.shape[0], mask.shape[1], mask.shape[1], int(cx), int(cy), int(r), color)


def find_fundamental_ransac(p1, p2, thr=1.0, conf=0.99):
    p1, p2 = np.ascontiguousarray(p1, np.float32), np.ascontiguousarray(p2, np.float32)
    n = len(p1)
    mask = np.zeros(n, np.uint8)
    it = C.c_int(0)
    ok = lib().orc_find_fundamental_ransac(P(p1, f32p), P(p2, f32p), n, C.c_double(thr), C.c_double(conf),
                                           P(mask, u8p), C.byref(it))
    return ok, mask, it.value


def fm_7point(p1, p2):
    p1, p2 = np.ascontiguousarray(p1, np.float32), np.ascontiguousarray(p2, np.float32)
    out = np.zeros(27)
    n = lib().orc_fm_7point(P(p1, f32p), P(p2, f32p), P(out, f64p))
    return out[:9 * max(n, 0)].reshape(-1, 3, 3)


def setmask_sort_perm(track_cnt):
    tc = np.ascontiguousarray(track_cnt, np.int32)
    perm = np.zeros(len(tc), np.int32)
    lib().orc_setmask_sort_perm(P(tc, i32p), len(tc), P(perm, i32p))
    return perm


def make_config(d):
    cfg = TrackerConfig()
    for k, v in d.items():
        setattr(cfg, k, v)
    return cfg


def lift_projective(cfg_dict, px):
    cfg = make_config(cfg_dict)
    px = np.ascontiguousarray(px, np.float64)
    out = np.zeros_like(px)
    lib().orc_lift_projective_pinhole(C.byref(cfg), P(px, f64p), len(px), P(out, f64p))
    return out


class OracleTracker:
    """FeatureTracker twin (oracle/fe_tracker.cpp)."""

    def __init__(self, cfg_dict, fisheye_mask=None):
        self.cfg = make_config(cfg_dict)
        self._mask = np.ascontiguousarray(fisheye_mask, np.uint8) if fisheye_mask is not None else None
        self.h = C.c_void_p(lib().orc_tracker_create(C.byref(self.cfg), P(self._mask, u8p) if self._mask is not None else None))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tracker_destroy(self.h)
            self.h = None

    def read_image(self, img, t, pub):
        img = np.ascontiguousarray(img)
        lib().orc_tracker_read_image(self.h, P(img, u8p), img.shape[1], C.c_double(t), int(pub))
        return self.result()

    def node_image(self, img, stamp):
        img = np.ascontiguousarray(img)
        restart = C.c_int(0)
        r = lib().orc_tracker_node_image(self.h, P(img, u8p), img.shape[1], C.c_double(stamp), C.byref(restart))
        return r, restart.value

    def result(self):
        n = lib().orc_tracker_count(self.h)
        ids, tc = np.zeros(n, np.int32), np.zeros(n, np.int32)
        cur, un, vel = (np.zeros((n, 2), np.float32) for _ in range(3))
        lib().orc_tracker_get(self.h, P(ids, i32p), P(tc, i32p), P(cur, f32p), P(un, f32p), P(vel, f32p))
        return dict(ids=ids, track_cnt=tc, cur_pts=cur, un_pts=un, velocity=vel)

    def stats(self):
        s = np.zeros(5, np.int32)
        lib().orc_tracker_stats(self.h, P(s, i32p))
        return dict(lk_in=int(s[0]), lk_ok=int(s[1]), ransac_in=int(s[2]), ransac_ok=int(s[3]), new=int(s[4]))


# ------------------------------------------------------------------------------------------------
# back end
class BeConfig(C.Structure):
    _fields_ = [("window_size", C.c_int), ("num_iterations", C.c_int), ("estimate_extrinsic", C.c_int),
                ("estimate_td", C.c_int), ("focal_length", C.c_double), ("min_parallax", C.c_double),
                ("acc_n", C.c_double), ("gyr_n", C.c_double), ("acc_w", C.c_double), ("gyr_w", C.c_double),
                ("g_norm", C.c_double), ("init_depth", C.c_double), ("td", C.c_double), ("tr", C.c_double),
                ("row", C.c_double), ("tic", C.c_double * 3), ("ric", C.c_double * 9)]


def be_config(window_size=10, num_iterations=8, estimate_extrinsic=0, estimate_td=0, td=0.0, tr=0.0, row=480.0,
              tic=None, ric=None):
    from harness import synth
    c = BeConfig(window_size=window_size, num_iterations=num_iterations, estimate_extrinsic=estimate_extrinsic,
                 estimate_td=estimate_td, focal_length=460.0, min_parallax=10.0 / 460.0, acc_n=synth.ACC_N,
                 gyr_n=synth.GYR_N, acc_w=synth.ACC_W, gyr_w=synth.GYR_W, g_norm=synth.G_NORM, init_depth=5.0, td=td,
                 tr=tr, row=row)
    c.tic[:] = list(synth.TIC if tic is None else tic)
    c.ric[:] = list(np.asarray(synth.RIC if ric is None else ric).ravel())
    return c


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def preintegrate(cfg, ba, bg, dt, acc, gyr, want_sqrt_info=True):
    dt, acc, gyr = _d(dt), _d(acc), _d(gyr)
    out, jac, cov, si = np.zeros(11), np.zeros((15, 15)), np.zeros((15, 15)), np.zeros((15, 15))
    lib().orc_preintegrate(C.byref(cfg), P(_d(ba), f64p), P(_d(bg), f64p), len(dt), P(dt, f64p), P(acc, f64p), P(gyr, f64p),
                           P(out, f64p), P(jac, f64p), P(cov, f64p), P(si, f64p) if want_sqrt_info else None)
    return dict(sum_dt=out[0], delta_p=out[1:4], delta_q=out[4:8], delta_v=out[8:11], jacobian=jac, covariance=cov,
                sqrt_info=si)


def imu_factor(cfg, ba, bg, dt, acc, gyr, params32):
    dt, acc, gyr, p = _d(dt), _d(acc), _d(gyr), _d(params32)
    res, raw = np.zeros(15), np.zeros(15)
    J = [np.zeros((15, 7)), np.zeros((15, 9)), np.zeros((15, 7)), np.zeros((15, 9))]
    lib().orc_imu_factor(C.byref(cfg), P(_d(ba), f64p), P(_d(bg), f64p), len(dt), P(dt, f64p), P(acc, f64p), P(gyr, f64p),
                         P(p, f64p), P(res, f64p), P(raw, f64p), *[P(j, f64p) for j in J])
    return res, raw, J


def projection_factor(params23, data, use_td=False, focal_length=460.0, TR=0.0, ROW=480.0):
    p, d = _d(params23), _d(data)
    res = np.zeros(2)
    J = [np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 7)), np.zeros((2, 1)), np.zeros((2, 1))]
    lib().orc_projection_factor(int(use_td), C.c_double(focal_length), C.c_double(TR), C.c_double(ROW), P(p, f64p), P(d, f64p),
                                P(res, f64p), *[P(j, f64p) for j in J])
    return res, J


def pose_plus(x, delta):
    out = np.zeros(7)
    lib().orc_pose_plus(P(_d(x), f64p), P(_d(delta), f64p), P(out, f64p))
    return out


def sym_eigen_ql(A):
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    w, V = np.zeros(n), np.zeros((n, n))
    lib().orc_sym_eigen_ql(n, P(A, f64p), P(w, f64p), P(V, f64p))
    return w, V


def sym_eigen(A):
    A = _d(A)
    n = A.shape[0]
    w, V = np.zeros(n), np.zeros((n, n))
    lib().orc_sym_eigen(n, P(A, f64p), P(w, f64p), P(V, f64p))
    return w, V


class OracleEstimator:
    """Estimator twin (oracle/be_estimator.cpp)."""

    def __init__(self, cfg):
        lib().orc_est_create.restype = C.c_void_p
        self.cfg = cfg
        self.W = cfg.window_size
        self.h = C.c_void_p(lib().orc_est_create(C.byref(cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_est_destroy(self.h)
            self.h = None

    # ---- solver cross-check hooks (oracle/be_solver.h ProbeEvaluate)
    PROBE_CB = C.CFUNCTYPE(None, C.c_int, C.c_int)

    def set_probe(self, fn, solve_index=-1):
        """fn(n_columns, n_residuals, evaluate) is called right before Solve of the given solve (inside processImage);
        evaluate(delta) returns the robustified residual vector at x (+) delta."""
        def cb(ncols, nres):
            def evaluate(delta):
                delta = np.ascontiguousarray(delta, np.float64)
                out = np.zeros(nres)
                rc = lib().orc_est_probe_residuals(self.h, P(delta, f64p), P(out, f64p))
                assert rc == 0
                return out
            fn(ncols, nres, evaluate)
        self._probe = self.PROBE_CB(cb) if fn else self.PROBE_CB()
        lib().orc_est_set_probe.argtypes = [C.c_void_p, self.PROBE_CB, C.c_int]
        lib().orc_est_set_probe(self.h, self._probe, solve_index)

    def set_iterations(self, n):
        lib().orc_est_set_iterations(self.h, int(n))

    def set_fast_eigen(self, on=True):
        """Tridiagonal QL instead of cyclic Jacobi in the marginalisation (the timed CPU baseline uses it)."""
        lib().orc_est_set_fast_eigen(self.h, int(bool(on)))

    def profile(self):
        out = (C.c_double * 2)()
        lib().orc_est_profile(self.h, out)
        return dict(solve_s=out[0], marg_s=out[1])

    def set_seed(self, rows, ba, bg):
        rows = _d(rows)
        lib().orc_est_set_seed(self.h, len(rows), P(rows, f64p), P(_d(ba), f64p), P(_d(bg), f64p))

    def clearState(self):
        lib().orc_est_clear_state(self.h)

    def setReloFrame(self, stamp, index, match_points, relo_t, relo_r):
        """Estimator::setReloFrame (estimator.cpp:1128-1146); match_points = n x (x, y, feature id)."""
        mp = _d(match_points).reshape(-1, 3)
        lib().orc_est_set_relo_frame(self.h, C.c_double(stamp), int(index), len(mp), P(mp, f64p), P(_d(relo_t), f64p), P(_d(relo_r).reshape(9), f64p))

    def relo(self):
        o = np.zeros(24)
        lib().orc_est_relo(self.h, P(o, f64p))
        return dict(drift_correct_r=o[0:9].reshape(3, 3).copy(), drift_correct_t=o[9:12].copy(), relo_relative_t=o[12:15].copy(),
                    relo_relative_q=o[15:19].copy(), relo_relative_yaw=float(o[19]), pending=bool(o[20]), local_index=int(o[21]),
                    factors=int(o[22]), solves=int(o[23]))

    def processIMU(self, dt, acc, gyr):
        lib().orc_est_process_imu(self.h, C.c_double(dt), P(_d(acc), f64p), P(_d(gyr), f64p))

    def processImage(self, ids, xyz_uv_vel, stamp):
        ids = np.ascontiguousarray(ids, np.int32)
        d = _d(xyz_uv_vel)
        lib().orc_est_process_image(self.h, len(ids), P(ids, i32p), P(d, f64p), C.c_double(stamp))

    def info(self):
        o, c = np.zeros(10, np.int32), np.zeros(2)
        lib().orc_est_info(self.h, P(o, i32p), P(c, f64p))
        keys = ["solver_flag", "frame_count", "marginalization_flag", "n_solves", "n_reboots", "landmarks", "visual",
                "iterations", "successful_steps", "termination"]
        d = {k: int(v) for k, v in zip(keys, o)}
        d["initial_cost"], d["final_cost"] = float(c[0]), float(c[1])
        return d

    def states(self):
        out, td = np.zeros((self.W + 1, 16)), C.c_double(0)
        lib().orc_est_states(self.h, P(out, f64p), C.byref(td))
        return out, td.value

    def prior(self, cap=256):
        """(A', b', blocks) of the last marginalisation; blocks = [(type, index, offset, size)]."""
        A, b = np.zeros(cap * cap), np.zeros(cap)
        n = lib().orc_est_prior(self.h, cap, P(A, f64p), P(b, f64p))
        blk = np.zeros(4 * 64, np.int32)
        nb = lib().orc_est_prior_blocks(self.h, P(blk, i32p))
        return A[: n * n].reshape(n, n).copy(), b[:n].copy(), [tuple(int(v) for v in blk[4 * k:4 * k + 4]) for k in range(nb)]
