"""rosbag 2.0 reader (uncompressed chunks; record layout per the rosbag 2.0 format) + ROS1 deserialisation of the two message
types the node subscribes to (reference src/ros_node.cc:166-183): nav_msgs/Odometry and sensor_msgs/LaserScan."""
import struct

import numpy as np

F32 = np.float32


def _records(buf, pos, end):
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
    """-> (buffer, {connection id: topic}, [(bag time, connection id, offset, length)] in bag-time order)."""
    buf = open(path, "rb").read()
    if not buf.startswith(b"#ROSBAG V2.0\n"):
        raise ValueError(f"{path}: not a rosbag 2.0 file")
    conns, msgs = {}, []
    for hdr, dpos, dlen in _records(buf, 13, len(buf)):
        op = hdr["op"][0]
        if op == 0x05:                                     # chunk
            if hdr["compression"] != b"none":
                raise ValueError("only uncompressed chunks are supported")
            for h2, p2, l2 in _records(buf, dpos, dpos + dlen):
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


def _header(buf, pos):
    _seq, sec, nsec, flen = struct.unpack_from("<IIII", buf, pos)
    return sec, nsec, pos + 16 + flen


def parse_odometry(buf, pos):
    """nav_msgs/Odometry -> dict(time, position[3], orientation (w, x, y, z), linear[3], angular[3]) — the fields
    Node::ToOdometryData copies (ros_node.cc:662-680)."""
    sec, nsec, pos = _header(buf, pos)
    flen = struct.unpack_from("<I", buf, pos)[0]
    pos += 4 + flen                                        # child_frame_id
    p = struct.unpack_from("<3d", buf, pos)
    qx, qy, qz, qw = struct.unpack_from("<4d", buf, pos + 24)
    pos += 8 * 7 + 8 * 36                                  # pose + covariance
    lin = struct.unpack_from("<3d", buf, pos)
    ang = struct.unpack_from("<3d", buf, pos + 24)
    return dict(time=sec + nsec * 1e-9, position=p, orientation=(qw, qx, qy, qz), linear=lin, angular=ang)


def parse_scan(buf, pos):
    sec, nsec, pos = _header(buf, pos)
    amin, amax, ainc, tinc, stime, rmin, rmax = struct.unpack_from("<7f", buf, pos)
    pos += 28
    n = struct.unpack_from("<I", buf, pos)[0]
    ranges = np.frombuffer(buf, "<f4", n, pos + 4)
    pos += 4 + 4 * n
    k = struct.unpack_from("<I", buf, pos)[0]
    inten = np.frombuffer(buf, "<f4", k, pos + 4)
    return dict(sec=sec, nsec=nsec, stamp=sec + nsec * 1e-9, angle_min=F32(amin), angle_max=F32(amax), angle_increment=F32(ainc),
                time_increment=F32(tinc), scan_time=F32(stime), range_min=F32(rmin), range_max=F32(rmax), ranges=ranges,
                intensities=inten)
