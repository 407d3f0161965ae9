"""A short seeded run with two relocalisation messages and marginalisations (compute-sanitizer target: the relocalisation rows,
the relo block pairs of ba_reduce and the candidate update are all exercised within ~16 frames)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

if __name__ == "__main__":
    from harness import pipeline, synth
    from vins_mono_b200 import Estimator
    kw = dict(estimate_td=1, estimate_extrinsic=1) if "--wide" in sys.argv else {}
    seq = synth.Sequence(seed=0, duration=3.0)
    msgs = synth.track_messages(seq, 17, max_feats=60)
    est = Estimator(tic=synth.TIC, ric=synth.RIC, **kw)
    est.set_seed(pipeline.gt_seed_rows(seq, [m[0] for m in msgs]), seq.ba, seq.bg)
    feeder = pipeline.ImuFeeder(*seq.imu())
    for k, (stamp, ids, d) in enumerate(msgs):
        feeder.feed(est, stamp)
        if k in (13, 15):
            j = k - 6
            mp, p_old, R_old = synth.loop_frame_matches(seq, 0.4, msgs[j][1], pixel_sigma=0.3, seed=k)
            est.setReloFrame(msgs[j][0], 3, mp, p_old + 0.1, R_old)
        est.processImage(ids, d, stamp)
    r = est.relo()
    print("relo solves", r["solves"], "factors", r["factors"], "info", est.info()["n_solves"])
    assert r["solves"] == 2
