"""reflector_detect::PoseExtrapolator (reference src/reflector_detect/laser/pose_extrapolator.cc:12-129), the default build
(USE_UNIFORM_VELOCITY is not defined, CMakeLists.txt): constant-velocity extrapolation from ONE odometry sample.

Quirks kept: in the "between" branch the loop at :72-76 does not break, so the sample used is always the NEWEST one whose stamp is
>= the query (i.e. the back of the deque); the backward branch of Interpolator (:115-127) uses `yaw - w*dt` with dt < 0 like the
forward one uses it with dt <= 0."""
import math
from collections import deque, namedtuple

OdometrySample = namedtuple("OdometrySample", "time position orientation linear angular")   # orientation = (w, x, y, z)


class PoseExtrapolator:
    def __init__(self):
        self.data = deque()

    def TrimDataByTime(self, time):                        # :12-28 (#else branch)
        while len(self.data) > 1 and self.data[0].time < time:
            self.data.popleft()

    def HandleOdometryData(self, msg):                     # :30-34
        self.data.append(msg)

    def ExtrapolatorPose(self, time):                      # :36-79 -> (x, y, yaw)
        if not self.data:
            return 0.0, 0.0, 0.0
        if time <= self.data[0].time:
            return self._interpolate(self.data[0], time)
        if time >= self.data[-1].time:
            return self._interpolate(self.data[-1], time)
        sample = None
        for d in self.data:                                # :72-76, no break
            if d.time >= time:
                sample = d
        return self._interpolate(sample, time)

    @staticmethod
    def _interpolate(s, time):                             # :98-127
        w = s.angular[2]
        yaw0 = 2 * math.atan2(s.orientation[3], s.orientation[0])
        vx, vy = s.linear[0], s.linear[1]
        if s.time <= time:
            dt = s.time - time
            yaw = yaw0 - w * dt
            x = s.position[0] - vx * dt * math.cos(yaw) + vy * dt * math.sin(yaw)
            y = s.position[1] - vx * dt * math.sin(yaw) - vy * dt * math.cos(yaw)
            return x, y, yaw
        dt = time - s.time
        yaw = yaw0 - w * dt
        x = s.position[0] + vx * dt * math.cos(yaw) - vy * dt * math.sin(yaw)
        y = s.position[1] + vx * dt * math.sin(yaw) + vy * dt * math.cos(yaw)
        return x, y, yaw
