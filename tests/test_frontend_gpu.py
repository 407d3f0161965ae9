"""GPU parity tests of the front end: CUDA path (through the C ABI) vs the CPU oracle on the same inputs."""
import numpy as np
import pytest

import orc
from harness import synth, pipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def FT():
    from vins_mono_b200 import FeatureTracker
    return FeatureTracker


def _in_border(p, rows, cols):
    x, y = np.rint(p[:, 0]).astype(int), np.rint(p[:, 1]).astype(int)
    return (1 <= x) & (x < cols - 1) & (1 <= y) & (y < rows - 1)


@pytest.mark.parametrize("shape", [(480, 752), (123, 157), (240, 376)])
def test_clahe_and_pyramid_bit_exact(FT, shape):
    rows, cols = shape
    img = synth.value_noise_image(rows, cols, 21)
    t = FT(**synth.tracker_config_dict(rows=rows, cols=cols))
    t.readImage(img, 0.0, False)
    ref = orc.clahe(img)
    lvl = 0
    while True:
        got = t.debug_level(lvl)
        assert np.array_equal(got, ref), f"level {lvl}"
        nxt = orc.pyrdown(ref)
        if lvl == 3 or nxt.shape[0] <= 21 or nxt.shape[1] <= 21:
            break
        ref, lvl = nxt, lvl + 1


@pytest.mark.parametrize("shape,seed", [((480, 752), 1), ((480, 752), 2), ((123, 157), 3)])
def test_gftt_bit_exact(FT, shape, seed):
    rows, cols = shape
    img = orc.clahe(synth.value_noise_image(rows, cols, seed))
    rng = np.random.default_rng(seed)
    mask = np.full((rows, cols), 255, np.uint8)
    for _ in range(25):
        orc.circle(mask, rng.integers(-10, cols + 10), rng.integers(-10, rows + 10), 30)
    t = FT(**synth.tracker_config_dict(rows=rows, cols=cols, max_cnt=300))
    for m, maxc in [(mask, 150), (None, 300), (mask, 7)]:
        ref, ncand_ref = orc.gftt(img, maxc, 0.01, 30, m)
        got, ncand, eig = t.debug_gftt(img, m, maxc, want_eig=True)
        assert np.array_equal(eig.view(np.int32), orc.min_eig(img).view(np.int32))   # f32 bit patterns
        assert ncand == ncand_ref
        assert np.array_equal(got, ref)                                                # list and order


def test_gftt_empty_and_flat(FT):
    rows, cols = 240, 376
    t = FT(**synth.tracker_config_dict(rows=rows, cols=cols))
    flat = np.full((rows, cols), 128, np.uint8)
    got, ncand, _ = t.debug_gftt(flat, None, 50)
    assert len(got) == 0 and ncand == 0
    img = synth.value_noise_image(rows, cols, 4)
    got, ncand, _ = t.debug_gftt(img, np.zeros((rows, cols), np.uint8), 50)      # everything masked
    ref, _ = orc.gftt(img, 50, 0.01, 30, np.zeros((rows, cols), np.uint8))
    assert len(got) == len(ref) == 0


@pytest.mark.parametrize("shift", [(3.3, -2.2), (12.5, 7.0), (-25.0, 14.0), (0.4, 0.3)])
def test_lk_bit_exact(FT, shift):
    rows, cols = 480, 752
    big = synth.value_noise_image(rows + 80, cols + 80, 9)
    a = big[40:40 + rows, 40:40 + cols]
    # integer part by slicing, fractional part by a bilinear blend
    ix, iy = int(np.floor(shift[0])), int(np.floor(shift[1]))
    fx, fy = shift[0] - ix, shift[1] - iy
    def crop(dx, dy):
        return big[40 - dy:40 - dy + rows, 40 - dx:40 - dx + cols].astype(np.float32)
    b = ((1 - fx) * (1 - fy) * crop(ix, iy) + fx * (1 - fy) * crop(ix + 1, iy) + (1 - fx) * fy * crop(ix, iy + 1)
         + fx * fy * crop(ix + 1, iy + 1))
    b = np.clip(np.rint(b), 0, 255).astype(np.uint8)
    pts, _ = orc.gftt(a, 150, 0.01, 30)
    extra = np.float32([[1.5, 2.5], [750.2, 478.9], [3.0, 400.0], [700.5, 2.2], [0, 0], [751, 479], [375.5, 0.5]])
    pts = np.vstack([pts, extra]).astype(np.float32)
    t = FT(**synth.tracker_config_dict(rows=rows, cols=cols, max_cnt=200))
    got, st = t.debug_lk(a, b, pts)
    ref, rst = orc.lk(a, b, pts)
    rst = rst & _in_border(ref, rows, cols).astype(np.uint8)     # the CUDA kernel fuses readImage's inBorder cull
    assert np.array_equal(st, rst)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))  # f32 bit patterns, tracked or not
    assert rst.sum() > 100


def test_tracker_sequence_matches_oracle(FT):
    """Whole readImage path, 30 frames of a rendered sequence: ids, track counts, pixel coordinates,
    undistorted coordinates and velocities must be identical (bit patterns) to the CPU oracle."""
    seq = synth.Sequence(seed=3, duration=2.0)
    ts, imgs = seq.images(30)
    cfg = synth.tracker_config_dict()
    gpu, cpu = FT(**cfg), orc.OracleTracker(cfg)
    n_pub = 0
    for i in range(len(ts)):
        rg, _ = gpu.node_image(imgs[i], float(ts[i]))
        rc, _ = cpu.node_image(imgs[i], float(ts[i]))
        assert rg == rc
        if not rc:
            continue
        a, b = gpu.result(), cpu.result()
        for k in ("ids", "track_cnt"):
            assert np.array_equal(a[k], b[k]), f"frame {i} {k}"
        for k in ("cur_pts", "un_pts", "velocity"):
            assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), f"frame {i} {k}"
        n_pub += rc == 2
    assert n_pub >= 10 and len(a["ids"]) > 100 and a["ids"].max() > 160


def test_full_size_properties(FT):
    """BASELINE-size run: properties that need no oracle — ids unique and monotone in age, min distance
    respected between surviving tracks at publish frames, published points all have track_cnt > 1."""
    seq = synth.Sequence(seed=5, duration=3.0)
    ts, imgs = seq.images(40)
    t = FT(**synth.tracker_config_dict())
    for i in range(len(ts)):
        r, _ = t.node_image(imgs[i], float(ts[i]))
        if r == 0:
            continue
        res = t.result()
        assert len(set(res["ids"].tolist())) == len(res["ids"]) <= 150
        assert (res["track_cnt"] >= 1).all()
        if r == 2:
            msg = t.feature_message()
            assert len(msg) == int((res["track_cnt"] > 1).sum())
            p = np.rint(res["cur_pts"]).astype(int)
            d = np.abs(p[:, None, :] - p[None, :, :]).max(-1) + np.eye(len(p), dtype=int) * 1000
            assert d.min() >= 15  # discs of radius 30 keep rounded points apart (Chebyshev >= 21 on the raster)


def _same_result(gpu, cpu, tag):
    a, b = gpu.result(), cpu.result()
    for k in ("ids", "track_cnt"):
        assert np.array_equal(a[k], b[k]), f"{tag} {k}"
    for k in ("cur_pts", "un_pts", "velocity"):
        assert np.array_equal(a[k].view(np.int32), b[k].view(np.int32)), f"{tag} {k}"
    return a


def test_tracker_edge_cases_match_oracle(FT):
    """The gating and bookkeeping branches of img_callback / readImage the plain sequence never takes
    (feature_tracker_node.cpp:36-62, feature_tracker.cpp:106-128, 169-173): frames without any corner, every track lost
    at once (ids keep counting), fewer than 8 tracked points (rejectWithF skipped), a time jump (restart published,
    tracker state dropped), a backwards stamp, and recovery afterwards.  Everything bit-identical to the CPU oracle."""
    cfg = synth.tracker_config_dict()
    gpu, cpu = FT(**cfg), orc.OracleTracker(cfg)
    rows, cols = cfg["rows"], cfg["cols"]
    tex = synth.value_noise_image(rows, cols, seed=21)
    tex2 = synth.value_noise_image(rows, cols, seed=22)
    flat = np.full((rows, cols), 128, np.uint8)
    sparse = np.full((rows, cols), 90, np.uint8)           # five isolated blobs: < 8 trackable points
    for cy, cx in ((100, 120), (200, 500), (350, 300), (420, 650), (60, 700)):
        sparse[cy - 6:cy + 6, cx - 6:cx + 6] = 220
    shift = lambda im, dx: np.ascontiguousarray(np.roll(im, dx, axis=1))
    frames = []
    t = 10.0
    def add(img, dt=0.05):
        nonlocal t
        t += dt
        frames.append((img, t))
    for k in range(6):
        add(shift(tex, 2 * k))                              # normal tracking
    for k in range(3):
        add(flat)                                           # nothing to track, nothing to detect
    for k in range(4):
        add(shift(tex2, 3 * k))                             # all-new features after a total loss
    for k in range(5):
        add(shift(sparse, k))                               # < 8 points: the F-matrix test is skipped
    add(shift(tex, 0), dt=1.5)                              # stamp jumps by > 1 s: restart
    for k in range(1, 4):
        add(shift(tex, 2 * k))
    add(shift(tex, 8), dt=-0.2)                             # stamp goes backwards: restart as well
    for k in range(5, 9):
        add(shift(tex, 2 * k))
    seen_restart, max_id = 0, -1
    for i, (img, stamp) in enumerate(frames):
        rg, sg = gpu.node_image(img, stamp)
        rc, sc = cpu.node_image(img, stamp)
        assert (rg, sg) == (rc, sc), f"frame {i}"
        seen_restart += int(sc)
        if rc:
            a = _same_result(gpu, cpu, f"frame {i}")
            if len(a["ids"]):
                max_id = max(max_id, int(a["ids"].max()))
    assert seen_restart == 2 and max_id > 150


@pytest.mark.parametrize("variant", ["fisheye_mask", "no_equalize", "small_dense", "qvga_rate20", "mei_camera", "kannala_brandt_camera"])
def test_tracker_config_variants_match_oracle(FT, variant):
    """Configurations beyond the EuRoC default (euroc_config.yaml:45-51 / parameters.cpp:37-74): FISHEYE with a circular
    mask as the initial setMask image (feature_tracker.cpp:38-41), EQUALIZE off (:87-95), other MAX_CNT / MIN_DIST, other
    image sizes and FREQ, the MEI (config/3dm) and KANNALA_BRANDT (config/cla) camera models in undistortedPoints / rejectWithF
    (CataCamera.cc:556-625, EquidistantCamera.cc:428-442).  20 frames each, bit-identical to the CPU oracle."""
    rows, cols = 480, 752
    kw, mask = {}, None
    if variant == "fisheye_mask":
        yy, xx = np.mgrid[0:rows, 0:cols]
        mask = np.where((yy - rows / 2) ** 2 + (xx - cols / 2) ** 2 < 210 ** 2, 255, 0).astype(np.uint8)
        kw = dict(fisheye=1)
    elif variant == "no_equalize":
        kw = dict(equalize=0)
    elif variant == "small_dense":
        kw = dict(max_cnt=300, min_dist=12)
    elif variant == "qvga_rate20":
        rows, cols = 240, 320
        kw = dict(max_cnt=80, min_dist=15, freq=20)
    cfg = synth.tracker_config_dict(rows=rows, cols=cols, **{k: v for k, v in kw.items() if k in ("max_cnt", "min_dist", "freq", "equalize")})
    cfg.update({k: v for k, v in kw.items() if k == "fisheye"})
    if variant == "mei_camera":
        cfg.update(camera_model=1, xi=2.057, k1=7.145e-02, k2=5.059e-01, p1=4.727e-05, p2=-5.492e-04, fx=1.115e+03, fy=1.114e+03,
                   cx=3.672e+02, cy=2.385e+02)
    elif variant == "kannala_brandt_camera":
        cfg.update(camera_model=2, xi=0.0, k1=-0.005740195474458931, k2=0.02878252863739417, p1=-0.04010621197185408,
                   p2=0.02008469575876223, fx=472.2863830700696, fy=470.83759684346785, cx=368.8316828103749, cy=232.23688706965652)
    gpu = FT(fisheye_mask=mask, **cfg) if mask is not None else FT(**cfg)
    cpu = orc.OracleTracker(cfg, fisheye_mask=mask)
    tex = synth.value_noise_image(rows + 40, cols + 60, seed=31)
    n_pub, n_feat = 0, 0
    for k in range(20):
        dx, dy = int(round(18 * np.sin(0.3 * k))) + 20, int(round(9 * np.cos(0.25 * k))) + 15
        img = np.ascontiguousarray(tex[dy:dy + rows, dx:dx + cols])
        rg, sg = gpu.node_image(img, 5.0 + 0.05 * k)
        rc, sc = cpu.node_image(img, 5.0 + 0.05 * k)
        assert (rg, sg) == (rc, sc), f"{variant} frame {k}"
        if rc:
            a = _same_result(gpu, cpu, f"{variant} frame {k}")
            n_feat = max(n_feat, len(a["ids"]))
            if mask is not None and len(a["ids"]):
                p = np.rint(a["cur_pts"]).astype(int)
                new = a["track_cnt"] == 1
                assert (mask[p[new, 1], p[new, 0]] == 255).all()      # detections only inside the fisheye mask
        n_pub += rc == 2
    assert n_pub >= 5 and n_feat >= 40


def test_tracker_bit_identical_to_cv2_twin_over_bench_sequence(FT):
    """The CUDA tracker against the golden digests of the cv2-backed (real OpenCV) twin of readImage + img_callback over the
    whole bench sequence (240 frames of seed 0, 119 publishes, > 1000 ids): ids, track counts, pixel / undistorted coordinates
    and velocities bit-identical in every frame (tests/golden/make_frontend_golden.py generated the digests with cv2 4.13)."""
    import hashlib
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend_track.npz"))
    n = int(g["n_frames"])
    seq = synth.Sequence(seed=int(g["seed"]), duration=n / 20.0 + 0.5)
    ts, imgs = pipeline.cached_images(seq, n)
    assert hashlib.sha256(np.ascontiguousarray(imgs).tobytes()).hexdigest() == str(g["images_sha"])
    tr = FT(**synth.tracker_config_dict())
    for i in range(n):
        r, restart = tr.node_image(imgs[i], float(ts[i]))
        assert r == int(g["rets"][i]) and restart == 0
        if not r:
            continue
        res = tr.result()
        h = hashlib.sha256()
        for k in ("ids", "track_cnt", "cur_pts", "un_pts", "velocity"):
            h.update(np.ascontiguousarray(res[k]).tobytes())
        assert len(res["ids"]) == int(g["counts"][i]) and h.hexdigest()[:24] == str(g["digests"][i]), f"diverged from the cv2 twin at frame {i}"
