"""Inputs of Estimator::initialStructure built from synthetic feature messages the way the estimator's own bookkeeping builds
them while solver_flag == INITIAL (test input generator, not product code): the window headers, every image frame with the IMU
samples pre-integrated into it (tmp_pre_integration, estimator.cpp:98-101, 137-140) and the FeatureManager tracks."""
from __future__ import annotations

import numpy as np

from . import pipeline, synth


class _Recorder:
    def __init__(self):
        self.cur = []

    def processIMU(self, dt, a, g):
        self.cur.append(np.r_[dt, a, g])


def first_window(seq, n_window=11, drop=(), max_feats=150, pixel_sigma=0.3, skip_first=True, offset=0):
    """Returns (headers, frames, tracks).  `drop` lists message indices (within the used messages) that were judged
    non-keyframes and left the window (they stay in all_image_frame): the window is the remaining first `n_window` messages."""
    n_msgs = n_window + len(drop) + (1 if skip_first else 0) + offset
    msgs = synth.track_messages(seq, n_msgs, max_feats=max_feats, pixel_sigma=pixel_sigma)
    msgs = msgs[offset + (1 if skip_first else 0):]  # offset: the window starts later (e.g. after a reboot)
    t_imu, acc, gyr = seq.imu()
    feeder, rec = pipeline.ImuFeeder(t_imu, acc, gyr), _Recorder()
    frames, last = [], None
    if offset:  # IMU samples before the first used image were consumed by earlier frames
        feeder.feed(_Recorder(), msgs[0][0] - 1e-9, td=0.0)
    for k, (stamp, ids, d) in enumerate(msgs):
        rec.cur = []
        feeder.feed(rec, stamp, td=0.0)
        imu = np.array(rec.cur).reshape(-1, 7)
        frames.append(dict(t=stamp, ids=np.asarray(ids, np.int32), xy=d[:, :2].copy(), imu=imu if k > 0 else np.zeros((0, 7)),
                           lin=np.zeros(6) if last is None else last[1:7].copy()))
        last = imu[-1]
    keep = [k for k in range(len(msgs)) if k not in set(drop)]
    assert len(keep) == n_window
    headers = [msgs[k][0] for k in keep]
    tracks, index = [], {}
    for w, k in enumerate(keep):
        for i, row in zip(msgs[k][1], msgs[k][2]):
            i = int(i)
            if i in index and tracks[index[i]][1] + len(tracks[index[i]][2]) == w:
                tracks[index[i]][2].append(row[:2].copy())
            elif i not in index:
                index[i] = len(tracks)
                tracks.append([i, w, [row[:2].copy()]])
    tracks = [(i, s, np.array(xy)) for i, s, xy in tracks]
    return headers, frames, tracks


def oracle_frames(frames):
    """The same frames as objects of the oracle twin (oracle/initial.py)."""
    import initial as oi
    out = []
    for k, f in enumerate(frames):
        pre = None
        if k > 0:
            pre = oi.Preint(f["lin"][:3], f["lin"][3:])
            for r in f["imu"]:
                pre.push_back(r[0], r[1:4], r[4:7])
        out.append(oi.ImageFrame(f["t"], f["ids"], f["xy"], pre))
    return out
