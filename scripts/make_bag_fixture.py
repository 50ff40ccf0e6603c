"""Builds tests/golden/bag_stream.npz from the reference's shipped dataset (config C1).

Runs HERE only (reads /root/reference/dataset/*.bag, which does not exist on the GPU box); the resulting
fixture is committed.  Contents = exactly what the reference's node would feed its EKF for this bag:
  * a from-scratch rosbag-v2 reader (uncompressed chunks; record layout per the rosbag 2.0 format spec);
  * ROS1 deserialisation of nav_msgs/Odometry and sensor_msgs/LaserScan;
  * a restatement of LaserReflectorDetect::HandleLaserScan (reference
    src/reflector_detect/laser/laser_reflector_detect.cc:23-316) with the launch-file parameters
    (launch/slam.launch:24-27: intensity_min 160, width 0.18 ± 0.06, sensor_to_base_link (0.13686, 0, 0);
    range gate [0.3, 10] m from ros_node.cc:257-265), float32 arithmetic where the reference uses
    Eigen::Vector2f.  The bag has scan_time = 0, so every beam carries the scan stamp and the pose
    extrapolator's motion un-distortion (:239-306) is the identity — it is not restated.
The node's call pattern (ros_node.cc:421-441, :627-660): the first scan only constructs the EKF with
init_time = its stamp (odometry before it is ignored because slam_ is null), later scans go through the
detector into HandleObservationMessage (empty frames included), every odometry message into
HandleOdometryMessage.
"""
import glob
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32


def read_records(buf, pos, end):
    while pos < end:
        hlen = struct.unpack_from("<I", buf, pos)[0]
        pos += 4
        hdr, hend = {}, pos + hlen
        while pos < hend:
            flen = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
            name, _, val = buf[pos:pos + flen].partition(b"=")
            hdr[name.decode()] = val
            pos += flen
        dlen = struct.unpack_from("<I", buf, pos)[0]
        pos += 4
        yield hdr, pos, dlen
        pos += dlen


def read_bag(path):
    buf = open(path, "rb").read()
    assert buf.startswith(b"#ROSBAG V2.0\n")
    conns, msgs = {}, []
    for hdr, dpos, dlen in read_records(buf, 13, len(buf)):
        op = hdr["op"][0]
        if op == 0x05:                                     # chunk
            assert hdr["compression"] == b"none", "only uncompressed chunks are supported"
            for h2, p2, l2 in read_records(buf, dpos, dpos + dlen):
                op2 = h2["op"][0]
                if op2 == 0x07:
                    conns[struct.unpack("<I", h2["conn"])[0]] = h2["topic"].decode()
                elif op2 == 0x02:
                    sec, nsec = struct.unpack("<II", h2["time"])
                    msgs.append((sec + nsec * 1e-9, struct.unpack("<I", h2["conn"])[0], p2, l2))
        elif op == 0x07:
            conns[struct.unpack("<I", hdr["conn"])[0]] = hdr["topic"].decode()
    msgs.sort(key=lambda m: m[0])                          # bag-time order, stable
    return buf, conns, msgs


def parse_header(buf, pos):
    seq, sec, nsec, flen = struct.unpack_from("<IIII", buf, pos)
    return sec + nsec * 1e-9, pos + 16 + flen


def parse_odometry(buf, pos):
    stamp, pos = parse_header(buf, pos)
    flen = struct.unpack_from("<I", buf, pos)[0]
    pos += 4 + flen                                        # child_frame_id
    pos += 8 * 7 + 8 * 36                                  # pose + covariance
    lin = struct.unpack_from("<3d", buf, pos)
    ang = struct.unpack_from("<3d", buf, pos + 24)
    return stamp, lin[0], lin[1], ang[2]                   # what reflector_ekf_slam.cc:216 reads


def parse_scan(buf, pos):
    stamp, pos = parse_header(buf, pos)
    amin, amax, ainc, tinc, stime, rmin, rmax = struct.unpack_from("<7f", buf, pos)
    pos += 28
    n = struct.unpack_from("<I", buf, pos)[0]
    ranges = np.frombuffer(buf, "<f4", n, pos + 4)
    pos += 4 + 4 * n
    k = struct.unpack_from("<I", buf, pos)[0]
    inten = np.frombuffer(buf, "<f4", k, pos + 4)
    return dict(stamp=stamp, angle_min=F32(amin), angle_max=F32(amax), angle_increment=F32(ainc), scan_time=F32(stime),
                range_min=F32(rmin), range_max=F32(rmax), ranges=ranges, intensities=inten)


def detect_reflectors(scan, intensity_min=160.0, width=0.18, width_err=0.06, gate=(0.3, 10.0), tx=F32(0.13686)):
    """LaserReflectorDetect::HandleLaserScan (:23-316) for scan_time == 0 (no motion un-distortion)."""
    assert scan["scan_time"] == 0.0
    ranges, inten = scan["ranges"], scan["intensities"]
    npts = len(ranges)
    # :55 — no fabs: true for every scan whose field of view is <= 2π
    is_circle = (float(scan["angle_max"]) - float(scan["angle_min"]) - 2 * np.pi) < 1e-6

    def to_base(r, a):                                     # :69-71 / :124-125, float32 like Eigen::Vector2f
        return (F32(F32(r) * F32(np.cos(F32(a)))) + tx, F32(F32(r) * F32(np.sin(F32(a)))))

    groups, group_ids = [], []                             # reflector_points / reflector_ids
    cur, cur_ids = [], []
    last_cloud_pt = None
    angle = F32(scan["angle_min"])

    def length(pts):
        return F32(np.hypot(F32(pts[0][0] - pts[-1][0]), F32(pts[0][1] - pts[-1][1])))

    for i in range(npts):
        rng = ranges[i]
        if scan["range_min"] <= rng <= scan["range_max"]:
            last_cloud_pt = to_base(rng, angle)            # point_cloud.back()
        if gate[0] <= rng <= gate[1] and inten[i] > intensity_min:
            if not cur:
                cur.append(last_cloud_pt); cur_ids.append(i)
            else:
                last_id = cur_ids[-1]
                if i - last_id == 1:
                    cur.append(last_cloud_pt); cur_ids.append(i)
                else:
                    gap = (i - last_id < 4 and abs(float(ranges[i]) - float(ranges[last_id])) < 0.3
                           and inten[i + 1 if i + 1 < npts else i] > intensity_min)   # :110
                    if gap:
                        for j in range(last_id + 1, i):    # :114-130
                            if np.isinf(ranges[j]):
                                continue
                            a_gap = F32(angle - F32(scan["angle_increment"] * F32(i - j)))
                            cur.append(to_base(ranges[j], a_gap)); cur_ids.append(j)
                        cur.append(last_cloud_pt); cur_ids.append(i)
                    else:                                  # :140-170 close the current run, start a new one
                        if (is_circle and cur_ids[0] == 0) or abs(float(length(cur)) - width) < width_err:
                            groups.append(cur); group_ids.append(cur_ids)
                        cur, cur_ids = [last_cloud_pt], [i]
        angle = F32(angle + scan["angle_increment"])
    if cur:                                                # :178-224
        if groups:
            first_id, last_id = group_ids[0][0], cur_ids[-1]
            fp, lp = groups[0][0], cur[-1]
            if is_circle and first_id == 0 and last_id == npts - 1 and np.hypot(float(lp[0] - fp[0]), float(lp[1] - fp[1])) < 0.1:
                groups[0] = groups[0] + cur
            elif abs(float(length(cur)) - width) < width_err:
                groups.append(cur)
            # :205-214 can only trigger when last_id == 0, i.e. never together with a non-empty first group
        elif abs(float(length(cur)) - width) < width_err:
            groups.append(cur)
    # :226-236 reads reflector_ids.front() even when it is empty (UB in the reference); it only ever removes the
    # first group when that group starts at beam 0, which the width gate bypass above makes a wrap-around
    # fragment.  Restated for the non-empty case only.
    elif groups and is_circle and group_ids[0][0] == 0:
        a, b = groups[0][0], groups[-1][0]
        if abs(np.hypot(float(a[0] - b[0]), float(a[1] - b[1])) - width) >= width_err:
            groups.pop(0)
    out = []
    for pts in groups:                                     # :297-306 centroid in float32
        cx = cy = F32(0.0)
        for p in pts:
            cx = F32(cx + p[0]); cy = F32(cy + p[1])
        out.append((F32(cx / F32(len(pts))), F32(cy / F32(len(pts)))))
    return np.array(out, F32).reshape(-1, 2)


def main():
    bags = glob.glob("/root/reference/dataset/*.bag")
    assert bags, "reference dataset not found"
    buf, conns, msgs = read_bag(bags[0])
    kind, time, odom_v, obs_start, obs_xy = [], [], [], [], []
    started = False
    n_scans = 0
    for _, conn, pos, _len in msgs:
        topic = conns[conn]
        if topic.endswith("odom"):
            t, vx, vy, wz = parse_odometry(buf, pos)
            if not started:
                continue                                   # slam_ is null before the first scan (ros_node.cc:635)
            kind.append(0); time.append(t); odom_v.append((vx, vy, wz)); obs_start.append(len(obs_xy))
        elif topic.endswith("scan"):
            scan = parse_scan(buf, pos)
            n_scans += 1
            if not started:                                # first scan only constructs the EKF (:424-441)
                started = True
                init_time = scan["stamp"]
                continue
            xy = detect_reflectors(scan)
            kind.append(1); time.append(scan["stamp"]); odom_v.append((0.0, 0.0, 0.0)); obs_start.append(len(obs_xy))
            obs_xy.extend(xy.tolist())
    obs_start.append(len(obs_xy))
    out = os.path.join(ROOT, "tests", "golden", "bag_stream.npz")
    np.savez_compressed(out, init_time=np.float64(init_time), kind=np.array(kind, np.int8), time=np.array(time, np.float64),
                        odom_v=np.array(odom_v, np.float64), obs_start=np.array(obs_start, np.int32),
                        obs_xy=np.array(obs_xy, np.float32).reshape(-1, 2))
    k = np.array(kind)
    cnt = np.diff(np.array(obs_start))[k == 1]
    print(f"{os.path.basename(bags[0])}: {n_scans} scans, {int((k == 0).sum())} odometry msgs after the first scan, "
          f"{int((k == 1).sum())} observation frames, reflectors/frame max {cnt.max()} mean {cnt.mean():.2f} -> {out}")


if __name__ == "__main__":
    main()
