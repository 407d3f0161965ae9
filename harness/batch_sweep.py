"""Aggregate frames/s of S concurrent replicas of one sequence on one GPU (native replay driver), S from argv."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402  (sets CUDA_DEVICE_MAX_CONNECTIONS=32 unless the caller chose a value)

if __name__ == "__main__":
    import torch
    sizes = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8, 16]
    steps, warmup = 20, 3
    seq, ts, imgs, imu = bench.sequence_inputs(0, bench.INIT_PUBS + warmup + steps)
    print("host cpus:", os.cpu_count(), flush=True)
    bench.run_replay_pass(seq, ts, imgs, imu, bench.INIT_PUBS, 0, 2, 0, 1)  # shake-out
    for s in sizes:
        r = bench.run_replay_pass(seq, ts, imgs, imu, bench.INIT_PUBS, warmup, steps, 0, s)
        p0 = r["trajs"][0][1]
        dp = max(float(abs(t[1] - p0).max()) for t in r["trajs"])
        print(json.dumps({"sequences": s, "fps": round(r["fps"], 1), "ms_region": round(r["region_ms"], 2), "frames": r["frames"],
                          "launches": r["launches"], "max_position_diff_between_replicas_m": dp}), flush=True)
