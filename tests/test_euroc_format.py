"""The EuRoC / ASL directory layout reader and writer (harness/euroc_format.py): the input data format of both hot paths.  CPU only."""
import os

import numpy as np
import pytest

from harness import euroc_format, synth

cv2 = pytest.importorskip("cv2")


def test_asl_round_trip(tmp_path):
    seq = synth.Sequence(seed=1, duration=0.6, rows=120, cols=188)
    root = str(tmp_path / "V_synth")
    ts, imgs = euroc_format.export_synthetic(seq, root, 6)
    assert os.path.exists(os.path.join(root, "mav0", "cam0", "data.csv")) and len(os.listdir(os.path.join(root, "mav0", "cam0", "data"))) == 6
    head = open(os.path.join(root, "mav0", "imu0", "data.csv")).readline().strip()
    assert head.startswith("#timestamp [ns],w_RS_S_x")          # gyroscope columns first, as in the dataset
    d = euroc_format.read_asl(root)
    t_imu, acc, gyr = seq.imu()
    shift = d["stamps"][0] - ts[0]                                # read_asl measures time from the first stamp on disk
    assert np.array_equal(d["images"], np.asarray(imgs))         # PNG is lossless
    assert np.abs(d["stamps"] - shift - ts).max() < 1e-9 and np.abs(d["imu_t"] - shift - t_imu).max() < 1e-9
    assert np.array_equal(d["acc"], acc) and np.array_equal(d["gyr"], gyr)   # repr() round-trips float64
    assert d["gt"].shape == (6, 17) and np.abs(d["gt"][:, 1:4] - np.array([seq.pose(t)[0] for t in ts])).max() < 1e-12
    # a prefix of the images, an explicit time origin
    d2 = euroc_format.read_asl(root, max_images=3, t_origin_ns=d["t_origin_ns"])
    assert len(d2["stamps"]) == 3 and np.array_equal(d2["stamps"], d["stamps"][:3])
