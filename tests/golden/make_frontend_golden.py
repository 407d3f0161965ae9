"""Generates tests/golden/frontend_*.npz with the REAL OpenCV (cv2 wheel) — run in the build container:

    python tests/golden/make_frontend_golden.py

Contents
  * frontend_ops.npz    : per-op known answers from cv2 4.13 (CLAHE, pyrDown, cornerMinEigenVal,
                          goodFeaturesToTrack with a disc mask, calcOpticalFlowPyrLK, circle,
                          findFundamentalMat masks) on seeded synthetic inputs.
  * frontend_track.npz  : per-frame (ids, track_cnt, cur_pts, un_pts, velocity) of a cv2-backed twin
                          of FeatureTracker::readImage + the node's gating/ID logic
                          (feature_tracker/src/feature_tracker.cpp:81-306,
                          feature_tracker_node.cpp:28-111) over a rendered sequence.

cv2 is put in its baseline (non-dispatched) mode with cv2.setUseOptimized(False): the dispatched
AVX2/AVX-512 kernels of the wheel contract multiply-adds into FMAs in the Sobel filters (and not even
uniformly across the image width), which is a property of that build, not of the algorithm; the
baseline path is plain IEEE arithmetic like the reference's SSE2 OpenCV 3.3.1.
OpenCV version recorded in each file.  The setMask tie order (std::sort, unstable) is taken from
libstdc++ via oracle's orc_setmask_sort_perm helper because Python cannot reproduce introsort.
"""
import os
import sys
import hashlib

import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import synth  # noqa: E402
import orc  # noqa: E402


def lift_projective(cfg, x, y):
    """PinholeCamera::liftProjective (PinholeCamera.cc:450-510), float64."""
    mx_d = (1.0 / cfg["fx"]) * x + (-cfg["cx"] / cfg["fx"])
    my_d = (1.0 / cfg["fy"]) * y + (-cfg["cy"] / cfg["fy"])
    k1, k2, p1, p2 = cfg["k1"], cfg["k2"], cfg["p1"], cfg["p2"]

    def dist(ux, uy):
        mx2, my2, mxy = ux * ux, uy * uy, ux * uy
        rho2 = mx2 + my2
        rad = k1 * rho2 + k2 * rho2 * rho2
        return (ux * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2),
                uy * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2))

    dx, dy = dist(mx_d, my_d)
    mx_u, my_u = mx_d - dx, my_d - dy
    for _ in range(1, 8):
        dx, dy = dist(mx_u, my_u)
        mx_u, my_u = mx_d - dx, my_d - dy
    return mx_u, my_u


def cv_round(v):
    return int(np.rint(np.float32(v)))


class Cv2Tracker:
    """cv2-backed twin of FeatureTracker + img_callback gating."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.cur_img = None
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.ids = np.zeros(0, np.int32)
        self.track_cnt = np.zeros(0, np.int32)
        self.prev_un_map = {}
        self.prev_time = 0.0
        self.n_id = 0
        self.first_image_flag, self.init_pub, self.pub_count = True, False, 1
        self.first_image_time = self.last_image_time = 0.0

    def in_border(self, p):
        x, y = cv_round(p[0]), cv_round(p[1])
        return 1 <= x < self.cfg["cols"] - 1 and 1 <= y < self.cfg["rows"] - 1

    def read_image(self, raw, t, pub):
        cfg = self.cfg
        img = cv2.createCLAHE(3.0, (8, 8)).apply(raw) if cfg["equalize"] else raw
        if self.cur_img is None:
            self.cur_img = img
        forw_img = img
        cur_pts, ids, tc = self.cur_pts, self.ids, self.track_cnt
        forw_pts = np.zeros((0, 2), np.float32)
        if len(cur_pts) > 0:
            nxt, st, _ = cv2.calcOpticalFlowPyrLK(self.cur_img, forw_img, cur_pts.reshape(-1, 1, 2), None,
                                                  winSize=(21, 21), maxLevel=3)
            nxt, st = nxt.reshape(-1, 2), st.ravel().astype(bool)
            for i in range(len(nxt)):
                if st[i] and not self.in_border(nxt[i]):
                    st[i] = False
            cur_pts, forw_pts, ids, tc = cur_pts[st], nxt[st], ids[st], tc[st]
        tc = tc + 1
        if pub:
            if len(forw_pts) >= 8:
                def virt(pts):
                    out = np.zeros((len(pts), 2), np.float32)
                    for i, (x, y) in enumerate(pts):
                        X, Y = lift_projective(cfg, float(x), float(y))
                        out[i] = (cfg["focal_length"] * X + cfg["cols"] / 2.0, cfg["focal_length"] * Y + cfg["rows"] / 2.0)
                    return out
                _, m = cv2.findFundamentalMat(virt(cur_pts), virt(forw_pts), cv2.FM_RANSAC, cfg["f_threshold"], 0.99)
                m = m.ravel().astype(bool) if m is not None else np.zeros(len(forw_pts), bool)
                cur_pts, forw_pts, ids, tc = cur_pts[m], forw_pts[m], ids[m], tc[m]
            mask = np.full((cfg["rows"], cfg["cols"]), 255, np.uint8)
            perm = orc.setmask_sort_perm(tc)  # libstdc++ std::sort tie order
            keep = []
            for k in perm:
                x, y = cv_round(forw_pts[k, 0]), cv_round(forw_pts[k, 1])
                if mask[y, x] == 255:
                    keep.append(k)
                    cv2.circle(mask, (x, y), cfg["min_dist"], 0, -1)
            keep = np.array(keep, np.int64)
            forw_pts, ids, tc = forw_pts[keep], ids[keep], tc[keep]
            n_max = cfg["max_cnt"] - len(forw_pts)
            if n_max > 0:
                c = cv2.goodFeaturesToTrack(forw_img, n_max, 0.01, cfg["min_dist"], mask=mask)
                c = c.reshape(-1, 2) if c is not None else np.zeros((0, 2), np.float32)
                forw_pts = np.vstack([forw_pts, c]).astype(np.float32)
                ids = np.concatenate([ids, -np.ones(len(c), np.int32)])
                tc = np.concatenate([tc, np.ones(len(c), np.int32)])
        self.cur_img, self.cur_pts, self.ids, self.track_cnt = forw_img, forw_pts.astype(np.float32), ids.astype(np.int32), tc.astype(np.int32)
        # undistortedPoints
        un = np.zeros((len(self.cur_pts), 2), np.float32)
        cur_map = {}
        for i, (x, y) in enumerate(self.cur_pts):
            X, Y = lift_projective(cfg, float(x), float(y))
            un[i] = (X, Y)
            cur_map.setdefault(int(self.ids[i]), un[i].copy())
        vel = np.zeros_like(un)
        if self.prev_un_map:
            dt = t - self.prev_time
            for i in range(len(un)):
                if self.ids[i] != -1 and int(self.ids[i]) in self.prev_un_map:
                    pu = self.prev_un_map[int(self.ids[i])]
                    # double v_x = (cur_un_pts[i].x - it->second.x) / dt: the difference of two floats is a float
                    vel[i] = (float(np.float32(un[i, 0]) - np.float32(pu[0])) / dt, float(np.float32(un[i, 1]) - np.float32(pu[1])) / dt)
        self.prev_un_map = cur_map
        self.prev_time = t
        self.un_pts, self.velocity = un, vel
        # updateID loop
        for i in range(len(self.ids)):
            if self.ids[i] == -1:
                self.ids[i] = self.n_id
                self.n_id += 1

    def node_image(self, raw, stamp):
        if self.first_image_flag:
            self.first_image_flag = False
            self.first_image_time = self.last_image_time = stamp
            return 0
        if stamp - self.last_image_time > 1.0 or stamp < self.last_image_time:
            self.first_image_flag, self.last_image_time, self.pub_count = True, 0, 1
            return 0
        self.last_image_time = stamp
        if np.floor(1.0 * self.pub_count / (stamp - self.first_image_time) + 0.5) <= self.cfg["freq"]:  # C round()
            pub = True
            if abs(1.0 * self.pub_count / (stamp - self.first_image_time) - self.cfg["freq"]) < 0.01 * self.cfg["freq"]:
                self.first_image_time, self.pub_count = stamp, 0
        else:
            pub = False
        self.read_image(raw, stamp, pub)
        if pub:
            self.pub_count += 1
            if not self.init_pub:
                self.init_pub = True
                return 1
            return 2
        return 1

    def result(self):
        return dict(ids=self.ids.copy(), track_cnt=self.track_cnt.copy(), cur_pts=self.cur_pts.copy(),
                    un_pts=self.un_pts.copy(), velocity=self.velocity.copy())


def two_view(rng, n, out_frac, noise=0.3):
    X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(3, 12, n)]
    K = np.array([[460, 0, 376], [0, 460, 240], [0, 0, 1.0]])
    R, _ = cv2.Rodrigues(rng.normal(0, 0.03, 3))
    t = rng.normal(0, 0.2, 3)
    p1 = (K @ X.T).T
    p1 = p1[:, :2] / p1[:, 2:]
    p2 = (K @ ((R @ X.T).T + t).T).T
    p2 = p2[:, :2] / p2[:, 2:]
    p1 += rng.normal(0, noise, p1.shape)
    p2 += rng.normal(0, noise, p2.shape)
    no = int(out_frac * n)
    if no:
        p2[:no] += rng.uniform(-40, 40, (no, 2))
    return p1.astype(np.float32), p2.astype(np.float32)


def image_hash(imgs):
    return hashlib.sha256(np.ascontiguousarray(imgs).tobytes()).hexdigest()


TRACK_SEED, TRACK_FRAMES, TRACK_FULL = 0, 240, 30   # bench.py's sequence (seed 0); full arrays for the first frames, digests for all


def result_digest(res):
    """Bit-level digest of one frame's public result vectors."""
    h = hashlib.sha256()
    for k in ("ids", "track_cnt", "cur_pts", "un_pts", "velocity"):
        h.update(np.ascontiguousarray(res[k]).tobytes())
    return h.hexdigest()[:24]


def main():
    cv2.setUseOptimized(False)
    cv2.setNumThreads(1)
    ops = dict(opencv_version=cv2.__version__)
    rng = np.random.default_rng(11)
    for k, (rows, cols) in enumerate([(480, 752), (123, 157)]):
        img = synth.value_noise_image(rows, cols, 100 + k)
        ops[f"img{k}_shape"] = np.array([rows, cols])
        eq = cv2.createCLAHE(3.0, (8, 8)).apply(img)
        ops[f"clahe{k}_sha"] = image_hash(eq)
        ops[f"pyr{k}_sha"] = image_hash(cv2.pyrDown(eq))
        ops[f"mineig{k}_sub3"] = cv2.cornerMinEigenVal(eq, 3, ksize=3)[::3, ::3].copy()
        mask = np.full((rows, cols), 255, np.uint8)
        centres = np.c_[rng.integers(-10, cols + 10, 30), rng.integers(-10, rows + 10, 30)]
        for cx, cy in centres:
            cv2.circle(mask, (int(cx), int(cy)), 30, 0, -1)
        ops[f"mask{k}_centres"] = centres
        ops[f"mask{k}_sha"] = image_hash(mask)
        c = cv2.goodFeaturesToTrack(eq, 150, 0.01, 30, mask=mask)
        ops[f"gftt{k}"] = c.reshape(-1, 2)
        M = np.float32([[1, 0, 3.3], [0, 1, -2.2]])
        nxt = cv2.warpAffine(eq, M, (cols, rows), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        pts = np.vstack([c.reshape(-1, 2), np.float32([[1.5, 2.5], [cols - 1.8, rows - 1.1], [0, 0]])]).astype(np.float32)
        nx, st, _ = cv2.calcOpticalFlowPyrLK(eq, nxt, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3)
        ops[f"lk{k}_shift"] = M
        ops[f"lk{k}_pts"], ops[f"lk{k}_next"], ops[f"lk{k}_status"] = pts, nx.reshape(-1, 2), st.ravel()
    fm = []
    for trial in range(12):
        n = int(rng.integers(15, 151))
        p1, p2 = two_view(rng, n, [0, 0.05, 0.13, 0.3][trial % 4])
        _, m = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        ops[f"fm{trial}_p1"], ops[f"fm{trial}_p2"], ops[f"fm{trial}_mask"] = p1, p2, m.ravel()
        fm.append(n)
    ops["fm_count"] = np.array(len(fm))
    np.savez_compressed(os.path.join(HERE, "frontend_ops.npz"), **ops)

    seq = synth.Sequence(seed=TRACK_SEED, duration=TRACK_FRAMES / 20.0 + 0.5)
    ts, imgs = seq.images(TRACK_FRAMES)
    cfg = synth.tracker_config_dict()
    tr = Cv2Tracker(cfg)
    out = dict(opencv_version=cv2.__version__, images_sha=image_hash(imgs), seed=np.array(TRACK_SEED), n_frames=np.array(TRACK_FRAMES),
               n_full=np.array(TRACK_FULL))
    rets, digests, counts = [], [], []
    for i in range(TRACK_FRAMES):
        r = tr.node_image(np.ascontiguousarray(imgs[i]), float(ts[i]))
        rets.append(r)
        digests.append(result_digest(tr.result()) if r else "")
        counts.append(len(tr.ids) if r else 0)
        if i < TRACK_FULL:
            out[f"f{i}_ret"] = np.array(r)
            if r:
                for k, v in tr.result().items():
                    out[f"f{i}_{k}"] = v
    out["rets"], out["digests"], out["counts"], out["final_n_id"] = np.array(rets), np.array(digests), np.array(counts), np.array(tr.n_id)
    np.savez_compressed(os.path.join(HERE, "frontend_track.npz"), **out)
    print("wrote golden; last frame tracks:", len(tr.ids), "n_id:", tr.n_id)


if __name__ == "__main__":
    main()
