"""Runs a batch of sequences through the batched path (vt_batch / ve_batch / vr_open_batch) for a few published frames:
profiling helper for the batched kernels (profiles/capture.sh)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=64)
    ap.add_argument("--distinct", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import bench
    inputs = bench.distinct_inputs(100, a.distinct, bench.INIT_PUBS + 2 + a.steps)
    r = bench.run_batch_pass(inputs, a.seqs, bench.INIT_PUBS, 2, a.steps, 0)
    print("frames", r["frames"], "fps", r["fps"], "launches", r["launches"])
