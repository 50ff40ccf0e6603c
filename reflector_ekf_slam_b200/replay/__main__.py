"""python -m reflector_ekf_slam_b200.replay <file.bag> [--out <filebase>] [--max-landmarks N] [--gps]

bag -> detector -> C ABI, with the node's call pattern (reference src/ros_node.cc:421-441, :627-660): the first scan only
constructs the filter (init_time = its stamp; odometry before it is dropped because slam_ is null), every later scan goes through
LaserReflectorDetect into HandleObservationMessage (empty frames included), every odometry message into the detector's
extrapolator AND HandleOdometryMessage.  At the end: the reference's two-line landmark map (Node::SaveReflectorResult)."""
import argparse
import sys

import numpy as np

from . import DetectOptions, LaserReflectorDetect, parse_odometry, parse_scan, read_bag


def replay(path, handle_odometry, handle_observation, create, options=None, limit=None):
    """Drives callbacks with the node's pattern; `create(init_time)` is called on the first scan.  Returns message counts."""
    buf, conns, msgs = read_bag(path)
    det = LaserReflectorDetect(options)
    started, n_od, n_obs, n_ref = False, 0, 0, 0
    for _, conn, pos, _len in msgs:
        topic = conns[conn]
        if topic.endswith("odom"):
            od = parse_odometry(buf, pos)
            if not started:
                continue                                       # slam_ is null before the first scan (ros_node.cc:635)
            det.HandleOdometryData(od["time"], od["position"], od["orientation"], od["linear"], od["angular"])   # :651
            handle_odometry(od["time"], od["linear"][0], od["linear"][1], od["angular"][2])                      # :637
            n_od += 1
        elif topic.endswith("scan"):
            scan = parse_scan(buf, pos)
            if not started:                                    # :424-441
                started = True
                create(scan["stamp"])
                continue
            t, xy = det.HandleLaserScan(scan)
            handle_observation(t, xy)
            n_obs += 1
            n_ref += len(xy)
            if limit and n_obs >= limit:
                break
    return {"odometry": n_od, "observations": n_obs, "reflectors": n_ref}


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m reflector_ekf_slam_b200.replay")
    ap.add_argument("bag")
    ap.add_argument("--out", default=None, help="write <out>.txt in the reference's map format (ros_node.cc:75-140)")
    ap.add_argument("--max-landmarks", type=int, default=64)
    ap.add_argument("--limit", type=int, default=0, help="stop after this many observation frames")
    a = ap.parse_args(argv)
    from ..engine import Observation, OdometryData, ReflectorEKFSLAM
    box = {}

    def create(t0):
        box["ekf"] = ReflectorEKFSLAM(init_time=t0, max_landmarks=a.max_landmarks, max_observations=32)

    stats = replay(a.bag, lambda t, vx, vy, wz: box["ekf"].HandleOdometryMessage(OdometryData(t, vx, vy, wz)),
                   lambda t, xy: box["ekf"].HandleObservationMessage(Observation(t, xy)), create, limit=a.limit or None)
    ekf = box["ekf"]
    ekf.sync()
    mu = ekf.GetStateVector()
    print(f"{a.bag}: {stats['odometry']} odometry messages, {stats['observations']} scans, {stats['reflectors']} reflector detections")
    print("pose  %.4f %.4f %.4f" % tuple(mu[:3]))
    for j, (x, y) in enumerate(mu[3:].reshape(-1, 2)):
        print("landmark %2d  %.4f %.4f" % (j, x, y))
    if a.out:
        ekf.save_map_txt(a.out)
        print("map ->", a.out + ".txt")
    return 0


if __name__ == "__main__":
    sys.exit(main())
