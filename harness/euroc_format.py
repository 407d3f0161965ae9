"""EuRoC MAV dataset layout (ASL format) in and out, the data format on the input side of both hot paths:

    <root>/mav0/cam0/data.csv      #timestamp [ns],filename          one row per image
    <root>/mav0/cam0/data/<timestamp>.png                              8-bit grayscale frames (752 x 480)
    <root>/mav0/imu0/data.csv      #timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y,w_RS_S_z,a_RS_S_x [m s^-2],a_RS_S_y,a_RS_S_z
    <root>/mav0/state_groundtruth_estimate0/data.csv   #timestamp,p_RS_R_x,y,z [m],q_RS_w,x,y,z,v_RS_R_x,y,z [m s^-1],b_w_RS_S_x,y,z,b_a_RS_S_x,y,z

read_asl() yields exactly what the replay driver takes (vins_mono_b200.ReplaySession / vr_sequence): image stamps in seconds, the
frames as one uint8 array, IMU stamps, accelerometer and gyroscope samples (note the file's column order: gyroscope first).  The
reference's euroc launch files feed the same data through rosbag topics (/cam0/image_raw, /imu0: config/euroc/euroc_config.yaml:4-5).
write_asl() exports a synthetic sequence in this layout, so that everything downstream can be exercised on files shaped like the
dataset (the dataset itself cannot travel to the GPU box).  PNG coding goes through cv2 (OpenCV is in the image)."""
from __future__ import annotations

import os

import numpy as np

IMU_HEADER = "#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]"
CAM_HEADER = "#timestamp [ns],filename"
GT_HEADER = ("#timestamp, p_RS_R_x [m], p_RS_R_y [m], p_RS_R_z [m], q_RS_w [], q_RS_x [], q_RS_y [], q_RS_z [], v_RS_R_x [m s^-1], "
             "v_RS_R_y [m s^-1], v_RS_R_z [m s^-1], b_w_RS_S_x [rad s^-1], b_w_RS_S_y [rad s^-1], b_w_RS_S_z [rad s^-1], "
             "b_a_RS_S_x [m s^-2], b_a_RS_S_y [m s^-2], b_a_RS_S_z [m s^-2]")


def _ns(t):
    return int(round(float(t) * 1e9))


def write_asl(root, stamps, images, imu_t, acc, gyr, ground_truth=None, t_offset_ns=1403636579000000000):
    """stamps [s], images [n, rows, cols] uint8, IMU arrays; ground_truth = rows of (t, p3, q wxyz, v3, bg3, ba3) or None.
    t_offset_ns shifts the clock to dataset-like epoch stamps (MH_01_easy starts at 1403636579...)."""
    import cv2
    cam = os.path.join(root, "mav0", "cam0")
    os.makedirs(os.path.join(cam, "data"), exist_ok=True)
    os.makedirs(os.path.join(root, "mav0", "imu0"), exist_ok=True)
    with open(os.path.join(cam, "data.csv"), "w") as f:
        f.write(CAM_HEADER + "\n")
        for t, img in zip(stamps, images):
            ns = _ns(t) + t_offset_ns
            name = f"{ns}.png"
            if not cv2.imwrite(os.path.join(cam, "data", name), np.ascontiguousarray(img)):
                raise RuntimeError(f"cannot write {name}")
            f.write(f"{ns},{name}\n")
    with open(os.path.join(root, "mav0", "imu0", "data.csv"), "w") as f:
        f.write(IMU_HEADER + "\n")
        for t, a, w in zip(imu_t, acc, gyr):
            f.write(f"{_ns(t) + t_offset_ns}," + ",".join(repr(float(v)) for v in (*w, *a)) + "\n")
    if ground_truth is not None:
        gt = os.path.join(root, "mav0", "state_groundtruth_estimate0")
        os.makedirs(gt, exist_ok=True)
        with open(os.path.join(gt, "data.csv"), "w") as f:
            f.write(GT_HEADER + "\n")
            for row in ground_truth:
                f.write(f"{_ns(row[0]) + t_offset_ns}," + ",".join(repr(float(v)) for v in row[1:]) + "\n")


def _read_csv(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            rows.append(line.split(","))
    return rows


def read_asl(root, max_images=None, t_origin_ns=None):
    """Returns dict(stamps, images, imu_t, acc, gyr, t_origin_ns[, gt]): seconds relative to t_origin_ns (default: the earlier of the
    first image / first IMU stamp, which keeps double precision for the sub-millisecond arithmetic of the IMU interpolation)."""
    import cv2
    cam = os.path.join(root, "mav0", "cam0")
    cam_rows = _read_csv(os.path.join(cam, "data.csv"))
    if max_images is not None:
        cam_rows = cam_rows[:max_images]
    imu_rows = _read_csv(os.path.join(root, "mav0", "imu0", "data.csv"))
    cam_ns = np.array([int(r[0]) for r in cam_rows], np.int64)
    imu_ns = np.array([int(r[0]) for r in imu_rows], np.int64)
    if t_origin_ns is None:
        t_origin_ns = int(min(cam_ns[0], imu_ns[0]))
    images = []
    for r in cam_rows:
        img = cv2.imread(os.path.join(cam, "data", r[1].strip()), cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise RuntimeError(f"cannot read {r[1]}")
        images.append(img)
    vals = np.array([[float(v) for v in r[1:7]] for r in imu_rows])
    out = dict(stamps=(cam_ns - t_origin_ns) * 1e-9, images=np.ascontiguousarray(np.stack(images)), imu_t=(imu_ns - t_origin_ns) * 1e-9,
               gyr=np.ascontiguousarray(vals[:, 0:3]), acc=np.ascontiguousarray(vals[:, 3:6]), t_origin_ns=t_origin_ns)
    gt_path = os.path.join(root, "mav0", "state_groundtruth_estimate0", "data.csv")
    if os.path.exists(gt_path):
        g = _read_csv(gt_path)
        out["gt"] = np.array([[(int(r[0]) - t_origin_ns) * 1e-9] + [float(v) for v in r[1:]] for r in g])
    return out


def export_synthetic(seq, root, n_images, workers=None):
    """A synth.Sequence written as an ASL directory (images rendered by the sequence itself)."""
    from . import synth
    ts, imgs = seq.images(n_images, workers=workers) if workers else seq.images(n_images)
    t_imu, acc, gyr = seq.imu()
    gt = []
    for t in ts:
        p, R, v, _, _ = seq.pose(t)
        gt.append(np.r_[t, p, synth.rot_to_quat_wxyz(R), v, seq.bg, seq.ba])
    write_asl(root, ts, imgs, t_imu, acc, gyr, ground_truth=gt)
    return ts, imgs
