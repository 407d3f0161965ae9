"""The product's host-side front-end pieces (vins_mono_b200/csrc/{fm_ransac.cpp,tracker.cu}) checked on a box without
a GPU through handle-free debug entries: the F-matrix RANSAC of rejectWithF (inlier masks bit-identical to
cv2.findFundamentalMat on the golden problems of tests/golden/frontend_ops.npz, plus degenerate inputs), the disc table of
setMask (reproduces cv2's filled circles) and PinholeCamera::liftProjective."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ransac():
    from vins_mono_b200 import build, load_library
    build.build()
    lib = load_library()
    lib.vt_debug_fundamental_ransac.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]

    def run(p1, p2, thr=1.0, conf=0.99):
        p1, p2 = np.ascontiguousarray(p1, np.float32), np.ascontiguousarray(p2, np.float32)
        mask = np.zeros(len(p1), np.uint8)
        ok = lib.vt_debug_fundamental_ransac(p1.ctypes.data, p2.ctypes.data, len(p1), thr, conf, mask.ctypes.data)
        return ok, mask
    return run


def test_masks_equal_cv2_golden(ransac):
    ops = np.load(os.path.join(ROOT, "tests", "golden", "frontend_ops.npz"))
    n = int(ops["fm_count"])
    assert n >= 10
    for t in range(n):
        ok, mask = ransac(ops[f"fm{t}_p1"], ops[f"fm{t}_p2"])
        assert ok == 1 and np.array_equal(mask, ops[f"fm{t}_mask"]), f"trial {t}"


def test_degenerate_inputs(ransac):
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 400, (6, 2)).astype(np.float32)
    ok, mask = ransac(p, p + 1)
    assert ok == 0 and mask.sum() == 0                    # < 7 points: no model
    p = rng.uniform(0, 400, (40, 2)).astype(np.float32)
    q = p.copy()
    q[:, 0] = 5.0                                          # collinear in image 2: no valid subset
    ok, mask = ransac(p, q)
    assert ok == 0 and mask.sum() == 0
    ok, mask = ransac(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert ok == 0


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_disc_table_reproduces_cv2_circle_masks():
    """The half-width table the mask kernel rasterises from, applied on the host to the golden centres, gives cv2's filled
    circles bit for bit (clipping at the image border included)."""
    from vins_mono_b200 import load_library
    lib = load_library()
    ops = np.load(os.path.join(ROOT, "tests", "golden", "frontend_ops.npz"))
    r = 30
    hw = np.zeros(r + 1, np.int32)
    assert lib.vt_debug_disc_half_widths(r, hw.ctypes.data_as(C.c_void_p)) == 0
    assert hw[0] == r and hw[r] >= 0 and (np.diff(hw) <= 0).all()
    for k in (0, 1):
        rows, cols = (int(v) for v in ops[f"img{k}_shape"])
        mask = np.full((rows, cols), 255, np.uint8)
        for cx, cy in ops[f"mask{k}_centres"]:
            for dy in range(-r, r + 1):
                y = int(cy) + dy
                if 0 <= y < rows:
                    w = int(hw[abs(dy)])
                    mask[y, max(int(cx) - w, 0):min(int(cx) + w, cols - 1) + 1] = 0
        assert _sha(mask) == str(ops[f"mask{k}_sha"])


def test_lift_projective_equals_oracle_and_inverts_the_distortion():
    import orc
    from harness import synth
    from vins_mono_b200 import load_library
    lib = load_library()
    cfg = synth.tracker_config_dict()
    K = np.array([cfg[k] for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2")], np.float64)
    rng = np.random.default_rng(5)
    px = np.ascontiguousarray(np.c_[rng.uniform(0, 752, 500), rng.uniform(0, 480, 500)])
    out = np.zeros_like(px)
    assert lib.vt_debug_lift_projective(K.ctypes.data_as(C.c_void_p), px.ctypes.data_as(C.c_void_p), len(px),
                                        out.ctypes.data_as(C.c_void_p)) == 0
    assert np.array_equal(out, orc.lift_projective(cfg, px))          # same operation sequence as the CPU oracle
    x, y = out[:, 0], out[:, 1]
    r2 = x * x + y * y
    rad = cfg["k1"] * r2 + cfg["k2"] * r2 * r2
    xd = x + x * rad + 2 * cfg["p1"] * x * y + cfg["p2"] * (r2 + 2 * x * x)
    yd = y + y * rad + 2 * cfg["p2"] * x * y + cfg["p1"] * (r2 + 2 * y * y)
    back = np.c_[cfg["fx"] * xd + cfg["cx"], cfg["fy"] * yd + cfg["cy"]]
    centre = np.hypot(px[:, 0] - cfg["cx"], px[:, 1] - cfg["cy"]) < 150
    assert np.abs(back - px)[centre].max() < 1e-3                     # 8 fixed-point iterations: exact near the centre
    assert np.abs(back - px).max() < 0.3                              # and to a fraction of a pixel in the corners
