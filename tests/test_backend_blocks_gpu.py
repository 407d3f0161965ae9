"""GPU parity of single factor blocks: the device functions the solve uses (through ve_debug_*) against the CPU oracle's
factor restatements and against finite differences with the reference's own check() recipe
(vins_estimator/src/factor/projection_factor.cpp:176-224): ProjectionFactor, ProjectionTdFactor, the Cauchy corrector,
IntegrationBase (pre-integration state, Jacobian, covariance, sqrt_info) and the whitened IMUFactor."""
import numpy as np
import pytest

import orc
from test_oracle_backend import E1, e1_params, e3_samples, e4_params, BA, BG

pytestmark = pytest.mark.gpu


def _gpu():
    import vins_mono_b200 as v
    return v


TD_DATA = np.r_[E1["pts_i"], E1["pts_j"], 0.3, -0.1, 0.25, -0.15, 0.001, 0.002, 200, 310]


def _data12(data14):
    """oracle layout pts_i(3) pts_j(3) vel_i vel_j td_i td_j row_i row_j -> device layout without the two z = 1."""
    d = np.asarray(data14, float)
    return np.r_[d[0:2], d[3:5], d[6:14]] if len(d) == 14 else np.r_[d[0:2], d[3:5], np.zeros(8)]


@pytest.mark.parametrize("use_td", [False, True])
def test_projection_factor_block_matches_oracle(use_td):
    v = _gpu()
    p = e1_params(td=0.003 if use_td else 0.0)
    kw = dict(use_td=use_td, TR=0.033, ROW=480.0)
    res, J = orc.projection_factor(p, TD_DATA, **kw)
    r, Jg, half_rho = v.debug_projection_factor(p, _data12(TD_DATA), use_td=use_td, tr=0.033, row=480.0)
    assert np.allclose(r, res, rtol=0, atol=1e-10)
    ref = np.hstack([J[0][:, :6], J[1][:, :6], J[2][:, :6], J[3], J[4] if use_td else np.zeros((2, 1))])
    assert np.abs(Jg - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    assert abs(half_rho - 0.5 * float(res @ res)) <= 1e-9 * float(res @ res)
    if not use_td:  # SURVEY Appendix E1 known answer
        assert np.allclose(r, [47.3616956445, 37.3001252650], rtol=0, atol=1e-8)
    else:           # E2
        assert np.allclose(r, [47.8872029413, 37.0120033366], rtol=0, atol=1e-8)


def test_projection_jacobian_finite_differences_on_device():
    v = _gpu()
    p0 = e1_params(td=0.003)
    d12 = _data12(TD_DATA)
    kw = dict(use_td=True, tr=0.033, row=480.0)
    r0, J, _ = v.debug_projection_factor(p0, d12, **kw)
    eps = 1e-6
    num = np.zeros((2, 20))
    for blk, off in ((0, 0), (1, 7), (2, 14)):
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            p = p0.copy()
            p[off:off + 7] = orc.pose_plus(p0[off:off + 7], d)
            num[:, 6 * blk + k] = (v.debug_projection_factor(p, d12, **kw)[0] - r0) / eps
    for col, idx in ((18, 21), (19, 22)):
        p = p0.copy()
        p[idx] += eps
        num[:, col] = (v.debug_projection_factor(p, d12, **kw)[0] - r0) / eps
    assert np.abs(J - num).max() < 2e-4 * max(1.0, np.abs(num).max())  # forward difference, the reference's eps


def test_cauchy_corrector_on_device():
    """ceres::CauchyLoss(1.0) + Corrector (restated in marginalization_factor.cpp:37-68): rho/2 = log(1 + s)/2, and with
    rho'' < 0 the alpha = 0 branch: r and J scaled by sqrt(rho')."""
    v = _gpu()
    p = e1_params()
    r, J, _ = v.debug_projection_factor(p, _data12(TD_DATA))
    rr, Jr, half_rho = v.debug_projection_factor(p, _data12(TD_DATA), robust=True)
    s = float(r @ r)
    assert abs(half_rho - 0.5 * np.log1p(s)) < 1e-12 * max(1.0, np.log1p(s))
    rho1 = 1.0 / (1.0 + s)  # rho'' = -rho'^2 < 0: Ceres / VINS take the alpha = 0 branch (marginalization_factor.cpp:49-53)
    assert np.allclose(rr, np.sqrt(rho1) * r, rtol=1e-12, atol=0)
    assert np.abs(Jr - np.sqrt(rho1) * J).max() <= 1e-12 * np.abs(J).max()


def test_preintegration_block_matches_oracle():
    v = _gpu()
    dt, acc, gyr = e3_samples()
    ref = orc.preintegrate(orc.be_config(), BA, BG, dt, acc, gyr)
    g = v.debug_imu_factor(BA, BG, dt, acc, gyr)
    assert abs(g["sum_dt"] - ref["sum_dt"]) < 1e-15
    assert np.allclose(g["dp"], ref["delta_p"], rtol=0, atol=1e-15)
    assert np.allclose(g["dq"], ref["delta_q"], rtol=0, atol=1e-15)
    assert np.allclose(g["dv"], ref["delta_v"], rtol=0, atol=1e-15)
    assert np.abs(g["jacobian"] - ref["jacobian"]).max() <= 1e-13 * np.abs(ref["jacobian"]).max()
    assert np.abs(g["covariance"] - ref["covariance"]).max() <= 1e-12 * np.abs(ref["covariance"]).max()
    # sqrt_info is a Cholesky factor of the inverse of a matrix with condition 8e8: compare what it stands for
    Pm = ref["covariance"]
    assert np.allclose(g["sqrt_info"].T @ g["sqrt_info"] @ Pm, np.eye(15), atol=1e-6)
    assert np.abs(g["sqrt_info"] - ref["sqrt_info"]).max() <= 1e-7 * np.abs(ref["sqrt_info"]).max()
    # SURVEY Appendix E3
    assert np.allclose(g["dp"], [0.0006770624474, 0.0010870152918, 0.0491786337654], rtol=0, atol=5e-10)


def test_imu_factor_block_matches_oracle_and_fd():
    v = _gpu()
    dt, acc, gyr = e3_samples()
    cfg = orc.be_config()
    p0 = e4_params()
    res, raw, J = orc.imu_factor(cfg, BA, BG, dt, acc, gyr, p0)
    g = v.debug_imu_factor(BA, BG, dt, acc, gyr, p0)
    # whitened quantities carry sqrt_info (entries up to 2e7, accurate to ~1e-7 relative): compare relative to scale
    assert np.abs(g["residual"] - res).max() <= 1e-6 * np.abs(res).max()
    ref = np.hstack([J[0][:, :6], J[1], J[2][:, :6], J[3]])
    assert np.abs(g["jacobian_w"] - ref).max() <= 1e-6 * np.abs(ref).max()
    # un-whitened residual against the oracle's raw residual and the Jacobian against finite differences
    si = g["sqrt_info"]
    raw_g = np.linalg.solve(si, g["residual"])
    assert np.allclose(raw_g, raw, rtol=0, atol=1e-9)
    eps = 1e-7
    Jraw = np.linalg.solve(si, g["jacobian_w"])
    col = 0
    for off, size, is_pose in [(0, 7, True), (7, 9, False), (16, 7, True), (23, 9, False)]:
        local = 6 if is_pose else size
        for k in range(local):
            p = p0.copy()
            if is_pose:
                d = np.zeros(6)
                d[k] = eps
                p[off:off + 7] = orc.pose_plus(p0[off:off + 7], d)
            else:
                p[off + k] += eps
            rk = np.linalg.solve(si, v.debug_imu_factor(BA, BG, dt, acc, gyr, p)["residual"])
            num = (rk - raw_g) / eps
            assert np.abs(Jraw[:, col] - num).max() < 5e-6 * max(1.0, np.abs(num).max()), (off, k)
            col += 1
