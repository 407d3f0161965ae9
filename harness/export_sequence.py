"""Writes a synthetic sequence in the plain-file layout examples/offline_replay.cpp reads."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import synth, pipeline  # noqa: E402

if __name__ == "__main__":
    out = sys.argv[1]
    n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    seq = synth.Sequence(seed=int(sys.argv[3]) if len(sys.argv) > 3 else 0, duration=n_img / 20.0 + 0.5)
    ts, imgs = seq.images(n_img)
    t_imu, acc, gyr = seq.imu()
    os.makedirs(out, exist_ok=True)
    open(os.path.join(out, "meta.txt"), "w").write(f"{imgs.shape[1]} {imgs.shape[2]} {len(ts)}\n")
    np.ascontiguousarray(imgs, np.uint8).tofile(os.path.join(out, "frames.u8"))
    np.savetxt(os.path.join(out, "stamps.txt"), ts, fmt="%.9f")
    np.savetxt(os.path.join(out, "imu.txt"), np.c_[t_imu, acc, gyr], fmt="%.12g")
    rows = pipeline.gt_seed_rows(seq, ts)
    with open(os.path.join(out, "seed.txt"), "w") as f:
        np.savetxt(f, rows, fmt="%.12g")
        f.write("bias " + " ".join(f"{v:.12g}" for v in np.r_[seq.ba, seq.bg]) + "\n")
    print("wrote", out, imgs.shape)
