"""CPU oracle of the one-shot initialisation (TEST INFRASTRUCTURE: only tests/ may import this).

Restates Estimator::initialStructure / visualInitialAlign (vins_estimator/src/estimator.cpp:218-440) and its helpers the way
the reference itself computes them: through OpenCV for the two-view geometry and PnP (the reference calls
cv::findFundamentalMat, cv::recoverPose, cv::solvePnP, cv::Rodrigues; cv2 4.13 is in the image) and numpy for the rest.
    solve_relative_rt     initial/solve_5pts.cpp:193-227
    GlobalSFM.construct   initial/initial_sfm.cpp:117-312   (the Ceres bundle is solved to its minimum with scipy's
                                                             least_squares; Ceres itself is not available: parity unpinned for
                                                             the iteration path, the minimum is what is compared)
    solve_gyroscope_bias / linear_alignment / refine_gravity   initial/initial_aligment.cpp:3-207
    IntegrationBase (mid-point deltas and d(delta_q)/d(bg))    factor/integration_base.h:54-186
The product's C++ (vins_mono_b200/csrc/initial.cpp) shares no code with this file.
"""
from __future__ import annotations

import numpy as np


# ---- rotations ---------------------------------------------------------------------------------------------------------------
def quat_to_R(q):  # q = (w, x, y, z)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat(R):
    from scipy.spatial.transform import Rotation
    x, y, z, w = Rotation.from_matrix(R).as_quat()
    return np.array([w, x, y, z]) * (1.0 if w >= 0 else -1.0)


def qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


# ---- IntegrationBase subset ------------------------------------------------------------------------------------------------
class Preint:
    def __init__(self, acc0, gyr0, ba=np.zeros(3), bg=np.zeros(3)):
        self.lin_acc, self.lin_gyr = np.array(acc0, float), np.array(gyr0, float)
        self.samples = []
        self._reset(ba, bg)

    def _reset(self, ba, bg):
        self.ba, self.bg = np.array(ba, float), np.array(bg, float)
        self.acc_0, self.gyr_0 = self.lin_acc.copy(), self.lin_gyr.copy()
        self.sum_dt = 0.0
        self.dp, self.dv = np.zeros(3), np.zeros(3)
        self.dq = np.array([1.0, 0, 0, 0])
        self.J_R_bg = np.zeros((3, 3))

    def _propagate(self, dt, a1, g1):
        Rq = quat_to_R(self.dq)
        un_acc_0 = Rq @ (self.acc_0 - self.ba)
        un_gyr = 0.5 * (self.gyr_0 + g1) - self.bg
        rq = qmul(self.dq, np.array([1.0, un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2]))
        un_acc_1 = quat_to_R(rq) @ (a1 - self.ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        self.dp = self.dp + self.dv * dt + 0.5 * un_acc * dt * dt
        self.dv = self.dv + un_acc * dt
        self.J_R_bg = (np.eye(3) - skew(un_gyr) * dt) @ self.J_R_bg - np.eye(3) * dt
        self.dq = rq / np.linalg.norm(rq)
        self.sum_dt += dt
        self.acc_0, self.gyr_0 = np.array(a1, float), np.array(g1, float)

    def push_back(self, dt, a, g):
        self.samples.append((float(dt), np.array(a, float), np.array(g, float)))
        self._propagate(*self.samples[-1])

    def repropagate(self, ba, bg):
        self._reset(ba, bg)
        for s in self.samples:
            self._propagate(*s)


class ImageFrame:
    def __init__(self, t, ids, xy, pre):
        self.t, self.ids, self.xy, self.pre = float(t), list(ids), np.asarray(xy, float).reshape(-1, 2), pre
        self.R, self.T, self.is_key_frame = np.eye(3), np.zeros(3), False


# ---- two-view geometry through OpenCV ----------------------------------------------------------------------------------------
def solve_relative_rt(corres):
    """corres: n x 4 (x0 y0 x1 y1).  Returns (ok, Rotation, Translation, inlier_cnt)."""
    import cv2
    corres = np.asarray(corres, float).reshape(-1, 4)
    if len(corres) < 15:
        return False, np.eye(3), np.zeros(3), 0
    ll = corres[:, :2].astype(np.float32)
    rr = corres[:, 2:].astype(np.float32)
    E, mask = cv2.findFundamentalMat(ll, rr, cv2.FM_RANSAC, 0.3 / 460, 0.99)
    if E is None or E.shape != (3, 3):
        return False, np.eye(3), np.zeros(3), 0
    cnt, rot, trans, mask = cv2.recoverPose(E, ll, rr, np.eye(3), mask=mask)
    R, T = np.array(rot), np.array(trans).reshape(3)
    return cnt > 12, R.T, -R.T @ T, int(cnt)


def solve_pnp(pts3, pts2, R_initial, P_initial):
    import cv2
    rvec, _ = cv2.Rodrigues(np.asarray(R_initial, float))
    t = np.asarray(P_initial, float).reshape(3, 1).copy()
    ok, rvec, t = cv2.solvePnP(np.asarray(pts3, np.float32).reshape(-1, 1, 3), np.asarray(pts2, np.float32).reshape(-1, 1, 2),
                               np.eye(3), None, rvec, t, True)
    r, _ = cv2.Rodrigues(rvec)
    return bool(ok), np.array(r), np.array(t).reshape(3)


def triangulate_point(P0, P1, x0, x1):
    A = np.stack([x0[0] * P0[2] - P0[0], x0[1] * P0[2] - P0[1], x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1]])
    v = np.linalg.svd(A)[2][3]
    return v[:3] / v[3]


# ---- GlobalSFM ---------------------------------------------------------------------------------------------------------------
def sfm_construct(frame_num, l, relative_R, relative_T, tracks):
    """tracks: list of (id, start_frame, xy[nobs, 2]).  Returns (ok, q[frame_num] wxyz cam-to-world, T, {id: point})."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    feats = [dict(id=i, obs=[(s + k, np.array(p, float)) for k, p in enumerate(np.asarray(xy).reshape(-1, 2))], state=False, pos=None)
             for i, s, xy in tracks]
    cR, cT, Pose = [None] * frame_num, [None] * frame_num, [None] * frame_num

    def set_pose(i, R, t):
        cR[i], cT[i] = np.array(R), np.array(t)
        Pose[i] = np.hstack([cR[i], cT[i].reshape(3, 1)])

    last = frame_num - 1
    set_pose(l, np.eye(3), np.zeros(3))
    R_last = np.asarray(relative_R).T  # (q[l] * relative_R)^-1
    set_pose(last, R_last, -R_last @ np.asarray(relative_T))

    def tri_two(f0, f1):
        for ft in feats:
            if ft["state"]:
                continue
            d = dict(ft["obs"])
            if f0 in d and f1 in d:
                ft["pos"], ft["state"] = triangulate_point(Pose[f0], Pose[f1], d[f0], d[f1]), True

    def pnp(i, R0, t0):
        p2, p3 = [], []
        for ft in feats:
            if not ft["state"]:
                continue
            for fr, xy in ft["obs"]:
                if fr == i:
                    p2.append(xy)
                    p3.append(ft["pos"])
                    break
        if len(p2) < 10:
            return False, None, None
        return solve_pnp(np.array(p3), np.array(p2), R0, t0)

    for i in range(l, last):
        if i > l:
            ok, R, t = pnp(i, cR[i - 1], cT[i - 1])
            if not ok:
                return False, None, None, None, None
            set_pose(i, R, t)
        tri_two(i, last)
    for i in range(l + 1, last):
        tri_two(l, i)
    for i in range(l - 1, -1, -1):
        ok, R, t = pnp(i, cR[i + 1], cT[i + 1])
        if not ok:
            return False, None, None, None, None
        set_pose(i, R, t)
        tri_two(i, l)
    for ft in feats:
        if not ft["state"] and len(ft["obs"]) >= 2:
            (f0, x0), (f1, x1) = ft["obs"][0], ft["obs"][-1]
            ft["pos"], ft["state"] = triangulate_point(Pose[f0], Pose[f1], x0, x1), True

    # full BA: rotations of all frames but l, translations of all but l and the last, every triangulated point
    pts = [ft for ft in feats if ft["state"]]
    rot_free = [i for i in range(frame_num) if i != l]
    tr_free = [i for i in range(frame_num) if i != l and i != last]
    cam = np.array([fr for ft in pts for fr, _ in ft["obs"]])
    pid = np.array([k for k, ft in enumerate(pts) for _ in ft["obs"]])
    uv = np.array([xy for ft in pts for _, xy in ft["obs"]])
    R0 = np.stack(cR)
    T0 = np.stack(cT)
    X0 = np.stack([ft["pos"] for ft in pts])
    nr, nt = len(rot_free), len(tr_free)

    def unpack(z):
        R, T = R0.copy(), T0.copy()
        w = z[:3 * nr].reshape(-1, 3)
        for k, i in enumerate(rot_free):
            R[i] = Rotation.from_rotvec(w[k]).as_matrix() @ R0[i]
        for k, i in enumerate(tr_free):
            T[i] = T0[i] + z[3 * nr + 3 * k: 3 * nr + 3 * k + 3]
        return R, T, X0 + z[3 * (nr + nt):].reshape(-1, 3)

    def fun(z):
        R, T, X = unpack(z)
        p = np.einsum("nij,nj->ni", R[cam], X[pid]) + T[cam]
        return (p[:, :2] / p[:, 2:3] - uv).ravel()

    # sparsity of the Jacobian (each residual pair touches one camera and one point) keeps the finite differences cheap
    from scipy.sparse import lil_matrix
    nz = 3 * (nr + nt) + 3 * len(pts)
    sp = lil_matrix((2 * len(cam), nz), dtype=int)
    rpos = {i: k for k, i in enumerate(rot_free)}
    tpos = {i: k for k, i in enumerate(tr_free)}
    for k in range(len(cam)):
        cols = list(range(3 * (nr + nt) + 3 * pid[k], 3 * (nr + nt) + 3 * pid[k] + 3))
        if cam[k] in rpos:
            cols += list(range(3 * rpos[cam[k]], 3 * rpos[cam[k]] + 3))
        if cam[k] in tpos:
            cols += list(range(3 * nr + 3 * tpos[cam[k]], 3 * nr + 3 * tpos[cam[k]] + 3))
        sp[2 * k, cols] = 1
        sp[2 * k + 1, cols] = 1
    sol = least_squares(fun, np.zeros(nz), method="trf", jac="3-point", jac_sparsity=sp, x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-15,
                        max_nfev=200)
    R, T, X = unpack(sol.x)
    cost = 0.5 * float(np.sum(sol.fun ** 2))
    q = [R_to_quat(R[i].T) for i in range(frame_num)]
    Tw = [-(R[i].T @ T[i]) for i in range(frame_num)]
    return True, q, Tw, {ft["id"]: X[k] for k, ft in enumerate(pts)}, cost


# ---- alignment ---------------------------------------------------------------------------------------------------------------
def solve_gyroscope_bias(frames, Bgs):
    A, b = np.zeros((3, 3)), np.zeros(3)
    for fi, fj in zip(frames[:-1], frames[1:]):
        q_ij = R_to_quat(fi.R.T @ fj.R)
        J = fj.pre.J_R_bg
        dqi = fj.pre.dq * np.array([1, -1, -1, -1])
        tb = 2 * qmul(dqi, q_ij)[1:]
        A += J.T @ J
        b += J.T @ tb
    delta = np.linalg.solve(A, b)
    for i in range(len(Bgs)):
        Bgs[i] = Bgs[i] + delta
    for fj in frames[1:]:
        fj.pre.repropagate(np.zeros(3), Bgs[0])
    return delta


def tangent_basis(g0):
    a = g0 / np.linalg.norm(g0)
    tmp = np.array([0, 0, 1.0])
    if np.array_equal(a, tmp):
        tmp = np.array([1.0, 0, 0])
    b = tmp - a * (a @ tmp)
    b /= np.linalg.norm(b)
    return np.stack([b, np.cross(a, b)], axis=1)


def _accumulate(A, b, i, tail, tA, tb):
    n = len(b)
    rA, rb = tA.T @ tA, tA.T @ tb
    idx = np.r_[np.arange(i * 3, i * 3 + 6), np.arange(n - tail, n)]
    A[np.ix_(idx, idx)] += rA
    b[idx] += rb


def refine_gravity(frames, tic, g_norm, g):
    g0 = g / np.linalg.norm(g) * g_norm
    n = len(frames)
    n_state = n * 3 + 3
    x = None
    for _ in range(4):
        lxly = tangent_basis(g0)
        A, b = np.zeros((n_state, n_state)), np.zeros(n_state)
        for i, (fi, fj) in enumerate(zip(frames[:-1], frames[1:])):
            dt = fj.pre.sum_dt
            tA, tb = np.zeros((6, 9)), np.zeros(6)
            tA[0:3, 0:3] = -dt * np.eye(3)
            tA[0:3, 6:8] = fi.R.T @ lxly * dt * dt / 2
            tA[0:3, 8] = fi.R.T @ (fj.T - fi.T) / 100.0
            tb[0:3] = fj.pre.dp + fi.R.T @ fj.R @ tic - tic - fi.R.T @ g0 * dt * dt / 2
            tA[3:6, 0:3] = -np.eye(3)
            tA[3:6, 3:6] = fi.R.T @ fj.R
            tA[3:6, 6:8] = fi.R.T @ lxly * dt
            tb[3:6] = fj.pre.dv - fi.R.T @ g0 * dt
            _accumulate(A, b, i, 3, tA, tb)
        x = np.linalg.solve(A * 1000.0, b * 1000.0)
        g0 = g0 + lxly @ x[n_state - 3: n_state - 1]
        g0 = g0 / np.linalg.norm(g0) * g_norm
    return g0, x


def linear_alignment(frames, tic, g_norm):
    n = len(frames)
    n_state = n * 3 + 4
    A, b = np.zeros((n_state, n_state)), np.zeros(n_state)
    for i, (fi, fj) in enumerate(zip(frames[:-1], frames[1:])):
        dt = fj.pre.sum_dt
        tA, tb = np.zeros((6, 10)), np.zeros(6)
        tA[0:3, 0:3] = -dt * np.eye(3)
        tA[0:3, 6:9] = fi.R.T * dt * dt / 2
        tA[0:3, 9] = fi.R.T @ (fj.T - fi.T) / 100.0
        tb[0:3] = fj.pre.dp + fi.R.T @ fj.R @ tic - tic
        tA[3:6, 0:3] = -np.eye(3)
        tA[3:6, 3:6] = fi.R.T @ fj.R
        tA[3:6, 6:9] = fi.R.T * dt
        tb[3:6] = fj.pre.dv
        _accumulate(A, b, i, 4, tA, tb)
    x = np.linalg.solve(A * 1000.0, b * 1000.0)
    s = x[-1] / 100.0
    g = x[n_state - 4: n_state - 1]
    if abs(np.linalg.norm(g) - g_norm) > 1.0 or s < 0:
        return False, g, x
    g, x = refine_gravity(frames, tic, g_norm, g)
    x = x.copy()
    x[-1] = x[-1] / 100.0
    return x[-1] >= 0.0, g, x


def visual_imu_alignment(frames, Bgs, tic, g_norm):
    delta = solve_gyroscope_bias(frames, Bgs)
    ok, g, x = linear_alignment(frames, np.asarray(tic, float), g_norm)
    return ok, g, x, delta


# ---- Estimator::relativePose / initialStructure ------------------------------------------------------------------------------
def relative_pose(tracks, W):
    for i in range(W):
        corres = []
        for _id, s, xy in tracks:
            xy = np.asarray(xy).reshape(-1, 2)
            if s <= i and s + len(xy) - 1 >= W:
                corres.append(np.r_[xy[i - s], xy[W - s]])
        if len(corres) > 20:
            c = np.array(corres)
            average_parallax = np.mean(np.linalg.norm(c[:, :2] - c[:, 2:], axis=1))
            if average_parallax * 460 > 30:
                ok, R, T, _ = solve_relative_rt(c)
                if ok:
                    return True, R, T, i
    return False, None, None, -1


def initial_structure(frames, headers, tracks, ric, tic, g_norm):
    """frames: list of ImageFrame (all images, ascending stamps), headers: stamps of the W + 1 window frames.
    Returns dict(code, l, g, x, delta_bg, Bgs) and fills R / T / is_key_frame of the frames."""
    F = len(headers)
    ok, rel_R, rel_T, l = relative_pose(tracks, F - 1)
    if not ok:
        return dict(code=1)
    out = sfm_construct(F, l, rel_R, rel_T, tracks)
    if not out[0]:
        return dict(code=2)
    _, Q, T, pts, cost = out
    ric = np.asarray(ric, float)
    i = 0
    for fr in frames:
        if fr.t == headers[i]:
            fr.is_key_frame = True
            fr.R = quat_to_R(Q[i]) @ ric.T
            fr.T = T[i]
            i += 1
            continue
        if fr.t > headers[i]:
            i += 1
        R_initial = quat_to_R(Q[i]).T
        P_initial = -R_initial @ T[i]
        fr.is_key_frame = False
        p3 = [pts[j] for j in fr.ids if j in pts]
        p2 = [fr.xy[k] for k, j in enumerate(fr.ids) if j in pts]
        if len(p3) < 6:
            return dict(code=3)
        ok, r, t = solve_pnp(np.array(p3), np.array(p2), R_initial, P_initial)
        if not ok:
            return dict(code=3)
        R_pnp = r.T
        fr.R = R_pnp @ ric.T
        fr.T = R_pnp @ (-t)
    Bgs = [np.zeros(3) for _ in range(F)]
    ok, g, x, delta = visual_imu_alignment(frames, Bgs, tic, g_norm)
    if not ok:
        return dict(code=4)
    return dict(code=0, l=l, g=g, x=x, delta_bg=delta, Bgs=Bgs, sfm_cost=cost)


# ---- visualInitialAlign tail (estimator.cpp:372-440): the window states the optimisation starts from -------------------------
def g2R(g):
    from scipy.spatial.transform import Rotation
    a = np.asarray(g, float) / np.linalg.norm(g)
    b = np.array([0, 0, 1.0])
    axis = np.cross(a, b)
    s = np.sqrt((1 + a @ b) * 2)
    R0 = quat_to_R(np.r_[s / 2, axis / s])  # Eigen::Quaterniond::FromTwoVectors(a, b)
    yaw = R2ypr(R0)[0]
    return Rotation.from_euler("z", -yaw, degrees=True).as_matrix() @ R0


def R2ypr(R):
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = np.arctan2(n[1], n[0])
    p = np.arctan2(-n[2], n[0] * np.cos(y) + n[1] * np.sin(y))
    r = np.arctan2(a[0] * np.sin(y) - a[1] * np.cos(y), -o[0] * np.sin(y) + o[1] * np.cos(y))
    return np.degrees(np.array([y, p, r]))


def window_after_align(frames, headers, x, g, tic):
    """Ps, Rs, Vs of the window frames and the rotated gravity after the 'change state' part of visualInitialAlign (depths are
    not reproduced here: the reference triangulates them with TIC = 0 and scales them by s)."""
    from scipy.spatial.transform import Rotation
    by_t = {f.t: f for f in frames}
    F = len(headers)
    Ps = [by_t[t].T.copy() for t in headers]
    Rs = [by_t[t].R.copy() for t in headers]
    for t in headers:
        by_t[t].is_key_frame = True
    s = x[-1]
    tic = np.asarray(tic, float)
    for i in range(F - 1, -1, -1):
        Ps[i] = s * Ps[i] - Rs[i] @ tic - (s * Ps[0] - Rs[0] @ tic)
    Vs = [np.zeros(3) for _ in range(F)]
    kv = -1
    for f in frames:
        if f.is_key_frame:
            kv += 1
            Vs[kv] = f.R @ x[3 * kv: 3 * kv + 3]  # indexed with the key-frame counter, as the reference does
    R0 = g2R(g)
    yaw = R2ypr(R0 @ Rs[0])[0]
    R0 = Rotation.from_euler("z", -yaw, degrees=True).as_matrix() @ R0
    return [R0 @ p for p in Ps], [R0 @ r for r in Rs], [R0 @ v for v in Vs], R0 @ g


# ---- InitialEXRotation (initial/initial_ex_rotation.cpp): ESTIMATE_EXTRINSIC == 2 ---------------------------------------------
class ExRotation:
    """Camera-IMU rotation from (essential-matrix rotation, gyroscope rotation) pairs; the two-view part through OpenCV as in the
    reference (cv::findFundamentalMat with its defaults, cv::SVD, cv::triangulatePoints)."""

    def __init__(self):
        self.frame_count = 0
        self.Rc, self.Rimu, self.Rc_g = [np.eye(3)], [np.eye(3)], [np.eye(3)]
        self.ric = np.eye(3)
        self.cov1 = 0.0

    @staticmethod
    def _test_triangulation(l, r, R, t):
        import cv2
        P = np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32)
        P1 = np.hstack([R, t.reshape(3, 1)]).astype(np.float32)
        X = cv2.triangulatePoints(P, P1, l.T.copy(), r.T.copy())
        front = 0
        for i in range(X.shape[1]):
            x = X[:, i] / X[3, i]
            if (P.astype(float) @ x)[2] > 0 and (P1.astype(float) @ x)[2] > 0:
                front += 1
        return front / X.shape[1]

    def solve_relative_r(self, corres):
        import cv2
        corres = np.asarray(corres, float).reshape(-1, 4)
        if len(corres) < 9:
            return np.eye(3)
        ll, rr = corres[:, :2].astype(np.float32), corres[:, 2:].astype(np.float32)
        E, _ = cv2.findFundamentalMat(ll, rr)

        def decompose(E):
            _, u, vt = cv2.SVDecomp(E)
            W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
            return u @ W @ vt, u @ W.T @ vt, u[:, 2].copy(), -u[:, 2].copy()

        R1, R2, t1, t2 = decompose(E)
        if np.linalg.det(R1) + 1.0 < 1e-09:
            R1, R2, t1, t2 = decompose(-E)
        ratio1 = max(self._test_triangulation(ll, rr, R1, t1), self._test_triangulation(ll, rr, R1, t2))
        ratio2 = max(self._test_triangulation(ll, rr, R2, t1), self._test_triangulation(ll, rr, R2, t2))
        return (R1 if ratio1 > ratio2 else R2).T

    def calibrate(self, corres, delta_q_imu, window_size=10, rc_given=None):
        """delta_q_imu wxyz.  Returns (ok, ric)."""
        self.frame_count += 1
        self.Rc.append(self.solve_relative_r(corres) if rc_given is None else np.asarray(rc_given, float))
        Rq = quat_to_R(np.asarray(delta_q_imu, float))
        self.Rimu.append(Rq)
        self.Rc_g.append(self.ric.T @ Rq @ self.ric)
        A = np.zeros((4 * self.frame_count, 4))
        for i in range(1, self.frame_count + 1):
            r1, r2 = R_to_quat(self.Rc[i]), R_to_quat(self.Rc_g[i])
            d = qmul(r1, r2 * np.array([1, -1, -1, -1]))
            ang = np.degrees(2 * np.arctan2(np.linalg.norm(d[1:]), abs(d[0])))
            huber = 5.0 / ang if ang > 5.0 else 1.0
            w, q = r1[0], r1[1:]
            L = np.zeros((4, 4))
            L[:3, :3] = w * np.eye(3) + skew(q)
            L[:3, 3] = q
            L[3, :3] = -q
            L[3, 3] = w
            rij = R_to_quat(self.Rimu[i])
            w, q = rij[0], rij[1:]
            R = np.zeros((4, 4))
            R[:3, :3] = w * np.eye(3) - skew(q)
            R[:3, 3] = q
            R[3, :3] = -q
            R[3, 3] = w
            A[4 * (i - 1): 4 * i] = huber * (L - R)
        _, sv, vt = np.linalg.svd(A)
        x = vt[3]  # Eigen quaternion coefficients (x, y, z, w)
        self.ric = quat_to_R(np.array([x[3], x[0], x[1], x[2]]) / np.linalg.norm(x)).T
        self.cov1 = sv[2] if len(sv) >= 3 else 0.0
        return bool(self.frame_count >= window_size and self.cov1 > 0.25), self.ric
