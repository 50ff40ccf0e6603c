"""reflector_detect::LaserReflectorDetect (reference src/reflector_detect/laser/laser_reflector_detect.cc:23-324) without ROS:
intensity-thresholded runs of beams -> reflector candidates -> width gate -> per-beam motion un-distortion through the
PoseExtrapolator (:239-306) -> centroids in base_link at the time of the last beam.  float32 where the reference uses
Eigen::Vector2f / Rigid2f, double where it uses Rigid2d.

Quirks kept: `is_circle_scan` lacks an fabs (:55) so it is true for every scan with FOV <= 2π, which lets a run that
starts at beam 0 bypass the width gate (:151); `reflector_ids.front()` on an empty deque (:226) — undefined in the reference —
is a no-op here."""
import math
from dataclasses import dataclass

import numpy as np

from .pose_extrapolator import OdometrySample, PoseExtrapolator

F32 = np.float32


@dataclass
class DetectOptions:                      # ReflectorDetectOptions as Node::LoadOptions fills it (ros_node.cc:240-283)
    intensity_min: float = 160.0
    reflector_min_length: float = 0.18
    reflector_length_error: float = 0.06
    range_min: float = 0.3
    range_max: float = 10.0
    sensor_to_base_link: tuple = (0.13686, 0.0, 0.0)   # x, y, yaw (launch/slam.launch:27)


def _rot_apply_f32(yaw, tx, ty, px, py):
    """Rigid2f * Vector2f: Rotation2D<float>(yaw) * p + t, every operation rounded to float32."""
    a = F32(yaw)
    c, s = F32(math.cos(float(a))), F32(math.sin(float(a)))
    return F32(F32(F32(c * px) - F32(s * py)) + F32(tx)), F32(F32(F32(s * px) + F32(c * py)) + F32(ty))


class LaserReflectorDetect:
    def __init__(self, options=None):
        self.o = options or DetectOptions()
        self.extrapolator = PoseExtrapolator()
        self.range_returns = np.zeros((0, 2), F32)      # range_data_.returns: the motion-corrected scan (for the grid mapper)

    def HandleOdometryData(self, time, position, orientation, linear, angular):      # :318-322
        self.extrapolator.HandleOdometryData(OdometrySample(time, tuple(position), tuple(orientation), tuple(linear), tuple(angular)))

    def HandleLaserScan(self, scan):
        """scan: dict as returned by replay.bag.parse_scan -> (observation time, (k, 2) float32 reflector centres in base_link)."""
        o = self.o
        ranges, inten = scan["ranges"], scan["intensities"]
        npts = len(ranges)
        t_last = float(scan["stamp"])
        dt_pt = float(scan["scan_time"]) / npts if npts else 0.0               # :47, double arithmetic on the float scan_time
        t_first = t_last - float(scan["scan_time"])
        self.extrapolator.TrimDataByTime(t_first)                              # :50-51
        tx, ty, tyaw = o.sensor_to_base_link
        is_circle = (float(scan["angle_max"]) - float(scan["angle_min"]) - 2 * math.pi) < 1e-6   # :55 (float subtraction widened)
        ainc = F32(scan["angle_increment"])

        def to_base(r, a):                                                     # :67-70
            px, py = F32(F32(r) * F32(math.cos(float(F32(a))))), F32(F32(r) * F32(math.sin(float(F32(a)))))
            return _rot_apply_f32(tyaw, tx, ty, px, py)

        cloud = []                                 # point_cloud: (x, y, time)
        groups, group_ids = [], []                 # reflector_points / reflector_ids
        cur, cur_ids = [], []
        angle = F32(scan["angle_min"])

        def length(pts):
            return F32(np.hypot(F32(pts[0][0] - pts[-1][0]), F32(pts[0][1] - pts[-1][1])))

        def good(pts):
            return abs(float(length(pts)) - o.reflector_min_length) < o.reflector_length_error

        for i in range(npts):
            rng = ranges[i]
            if scan["range_min"] <= rng <= scan["range_max"]:
                x, y = to_base(rng, angle)
                cloud.append((x, y, t_first + i * dt_pt))
            if o.range_min <= rng <= o.range_max and inten[i] > o.intensity_min and cloud:
                if not cur:
                    cur.append(cloud[-1]); cur_ids.append(i)
                else:
                    last_id = cur_ids[-1]
                    if i - last_id == 1:
                        cur.append(cloud[-1]); cur_ids.append(i)
                    else:
                        gap = (i - last_id < 4 and abs(float(ranges[i]) - float(ranges[last_id])) < 0.3
                               and inten[i + 1 if i + 1 < npts else i] > o.intensity_min)          # :110
                        if gap:
                            for j in range(last_id + 1, i):                                          # :114-130
                                if np.isinf(ranges[j]):
                                    continue
                                a_gap = F32(angle - F32(ainc * F32(i - j)))
                                gx, gy = to_base(ranges[j], a_gap)
                                cur.append((gx, gy, t_first + j * dt_pt)); cur_ids.append(j)
                            cur.append(cloud[-1]); cur_ids.append(i)
                        else:                                                                         # :140-170
                            if (is_circle and cur_ids[0] == 0) or good(cur):
                                groups.append(cur); group_ids.append(cur_ids)
                            cur, cur_ids = [cloud[-1]], [i]
            angle = F32(angle + ainc)
        if cur:                                                                                       # :178-224
            if groups:
                first_id, last_id = group_ids[0][0], cur_ids[-1]
                fp, fl = groups[0][0], groups[0][-1]
                lp, lf = cur[-1], cur[0]
                if (is_circle and first_id == 0 and last_id == npts - 1
                        and float(F32(np.hypot(F32(lp[0] - fp[0]), F32(lp[1] - fp[1])))) < 0.1):
                    groups[0] = groups[0] + cur
                elif good(cur):
                    groups.append(cur)
                if is_circle and last_id == 0:                                                       # :205-214
                    if abs(float(F32(np.hypot(F32(fl[0] - lf[0]), F32(fl[1] - lf[1])))) - o.reflector_min_length) >= o.reflector_length_error:
                        groups.pop(0)
            elif good(cur):
                groups.append(cur)
        elif groups and is_circle and group_ids[0][0] == 0:                                          # :226-236 (non-empty case only)
            a, b = groups[0][0], groups[-1][0]
            if abs(float(F32(np.hypot(F32(a[0] - b[0]), F32(a[1] - b[1])))) - o.reflector_min_length) >= o.reflector_length_error:
                groups.pop(0)

        if not cloud:                                    # point_cloud.back() on an empty cloud: undefined in the reference
            self.range_returns = np.zeros((0, 2), F32)
            return t_last, np.zeros((0, 2), F32)
        # ---- motion un-distortion (:239-306): every beam is moved by the odometry pose at ITS time, then everything is
        #      expressed in the base_link of the last beam -----------------------------------------------------------------
        ex = self.extrapolator.ExtrapolatorPose
        mx, my, myaw = ex(cloud[-1][2])                  # max_time_pose = poses.back()
        # last_pose_inverse (double): rotation -yaw, translation -(R(-yaw)·t)
        ci, si = math.cos(-myaw), math.sin(-myaw)
        itx, ity = -(ci * mx - si * my), -(si * mx + ci * my)
        ret = np.zeros((len(cloud), 2), F32)
        for k, (x, y, t) in enumerate(cloud):            # (last_pose_inverse * poses[i]).cast<float>() * p  (:256-260)
            px_, py_, pyaw = ex(t)
            ryaw = -myaw + pyaw
            rtx, rty = ci * px_ - si * py_ + itx, si * px_ + ci * py_ + ity
            ret[k] = _rot_apply_f32(ryaw, rtx, rty, x, y)
        self.range_returns = ret
        out = []
        for pts in groups:
            acc_x = acc_y = F32(0.0)
            for (x, y, t) in pts:
                px_, py_, pyaw = ex(t)
                ox, oy = _rot_apply_f32(pyaw, px_, py_, x, y)                    # pose.cast<float>() * p  (:293)
                bx, by = _rot_apply_f32(-myaw, itx, ity, ox, oy)                 # max_time_pose.inverse().cast<float>() * p  (:299)
                acc_x = F32(acc_x + bx); acc_y = F32(acc_y + by)
            out.append((F32(acc_x / F32(len(pts))), F32(acc_y / F32(len(pts)))))
        return t_last, np.array(out, F32).reshape(-1, 2)                         # USE_CORRECT_TIME is not defined: time = scan stamp
