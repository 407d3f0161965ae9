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
                ("cy", C.c_double), ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double)]


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

    def __init__(self, cfg_dict):
        self.cfg = make_config(cfg_dict)
        self.h = C.c_void_p(lib().orc_tracker_create(C.byref(self.cfg), None))

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
