"""The fork-pool renderer of the synthetic sequences (harness/synth.py) must produce exactly the frames of the serial
renderer: bench.py's CPU baseline and GPU arm, and the golden-free GPU tests, all consume these frames."""
import numpy as np

from harness import synth


def test_parallel_render_equals_serial():
    a = synth.Sequence(seed=4, duration=1.5, rows=120, cols=188)
    b = synth.Sequence(seed=4, duration=1.5, rows=120, cols=188)
    ts1, f1 = a.images(26, workers=1)
    ts2, f2 = b.images(26, workers=4)
    assert np.array_equal(ts1, ts2) and f1.dtype == np.uint8 and f1.shape == (26, 120, 188)
    assert np.array_equal(f1, f2)
    assert f1.std() > 10          # textured, not blank
