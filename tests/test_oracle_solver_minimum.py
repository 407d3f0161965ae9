"""Pins the MINIMUM the oracle's trust-region solver converges to (oracle/be_solver.cpp restates Ceres' dogleg +
DENSE_SCHUR from its published algorithm; Ceres itself is not available, so its iteration path stays unpinned):
scipy's Levenberg-Marquardt (MINPACK) minimises the same robustified residual vector — every factor of a real window
problem, Cauchy loss folded in as sqrt(rho(s)) r/|r| — starting from the same point, without Schur complement, dogleg
or Jacobi scaling.  Both must reach the same cost, and the oracle's 8-iteration answer must already be close to it."""
import numpy as np
import pytest
from scipy.optimize import least_squares

import orc
from harness import synth, pipeline


def run_until_solve(n_iterations, probe=None, solve_index=0):
    seq = synth.Sequence(seed=21, duration=2.5)
    msgs = synth.track_messages(seq, 12, max_feats=60)          # the 11th processed message fills the window
    est = orc.OracleEstimator(orc.be_config())
    est.set_iterations(n_iterations)
    est.set_seed(pipeline.gt_seed_rows(seq, [m[0] for m in msgs]), seq.ba, seq.bg)
    if probe:
        est.set_probe(probe, solve_index)
    feeder = pipeline.ImuFeeder(*seq.imu())
    for stamp, ids, d in msgs:
        feeder.feed(est, stamp)
        est.processImage(ids, d, stamp)
        if est.info()["n_solves"] > solve_index:
            break
    return est.info()


@pytest.mark.timeout(600)
def test_solver_minimum_matches_scipy():
    found = {}

    def probe(ncols, nres, evaluate):
        r0 = evaluate(np.zeros(ncols))
        found["initial"] = 0.5 * float(r0 @ r0)
        sol = least_squares(evaluate, np.zeros(ncols), method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-12, max_nfev=40 * ncols)
        found["scipy"] = float(sol.cost)
        found["shape"] = (ncols, nres)

    info8 = run_until_solve(8, probe)
    info_conv = run_until_solve(2000)   # runs until Ceres' function tolerance fires (~750 iterations: the weakly
                                        # determined bias / scale directions converge linearly)
    print("columns x residuals", found["shape"], "initial", found["initial"], "scipy minimum", found["scipy"],
          "oracle after 8 iterations", info8["final_cost"], "converged", info_conv["final_cost"], "iterations", info_conv["iterations"],
          "termination", info_conv["termination"])
    assert found["shape"][0] > 150 and found["shape"][1] > 300
    assert abs(info8["initial_cost"] - found["initial"]) <= 1e-9 * found["initial"]      # same cost function
    assert found["scipy"] < found["initial"]
    assert info_conv["termination"] == 2 and info_conv["iterations"] < 2000
    assert abs(info_conv["final_cost"] - found["scipy"]) <= 2e-5 * found["scipy"]        # same minimum
    assert info8["final_cost"] - found["scipy"] <= 0.05 * (found["initial"] - found["scipy"])  # 8 steps: >= 95 % of the way
