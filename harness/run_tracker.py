"""Replays a rendered sequence through the CUDA tracker (profiling / timing helper)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import synth  # noqa: E402
from vins_mono_b200 import FeatureTracker  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    seq = synth.Sequence(seed=a.seed, duration=a.frames / 20.0 + 0.1)
    ts, imgs = seq.images(a.frames)
    t = FeatureTracker(**synth.tracker_config_dict())
    wall, dev = [], []
    for i in range(a.frames):
        t0 = time.perf_counter()
        r, _ = t.node_image(imgs[i], float(ts[i]))
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append((r,) + t.timing())
    for i in range(a.frames):
        print(f"frame {i} ret {dev[i][0]} wall_ms {wall[i]:.3f} device_ms {dev[i][1]:.3f} launches {dev[i][2]}")
    print("n features", len(t.result()["ids"]))
