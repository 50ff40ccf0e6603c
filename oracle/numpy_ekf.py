"""Independent numpy float64 EKF-SLAM — the second opinion that pins the C oracle.

TEST INFRASTRUCTURE.  Written separately from oracle/rekf_oracle.c (shares no code with it):
plain dense numpy matrices, `np.linalg.inv`, textbook equations.  It follows the *behaviour* of
the reference (src/reflector_ekf_slam/reflector_ekf_slam.cc) — including its quirks: Euclidean
gate 0.6 m against the state (:437,:446), sqrt(dᵀΣd) gate 0.05 against the beacon map (:411,:420),
float32 roundings (:389-393, :431, :327-331), simple-form covariance update (:308) and the dense
G_z·Qt·G_zᵀ block for landmarks added in one frame (:354).
"""
import numpy as np

DIFF, OMNI = 0, 1


class NumpyEKF:
    def __init__(self, init_time=0.0, init_pose=(0, 0, 0), odom_model=DIFF, linear_velocity_cov=0.0025,
                 angular_velocity_cov=0.0064, observation_cov=0.0025):
        self.time = float(init_time)
        self.mu = np.array(init_pose, dtype=np.float64)
        self.sigma = np.zeros((3, 3))
        self.model = odom_model
        if odom_model == DIFF:
            self.Qu = np.diag([linear_velocity_cov, angular_velocity_cov])
        else:
            self.Qu = np.diag([linear_velocity_cov, linear_velocity_cov, angular_velocity_cov])
        self.Qt = np.eye(2) * observation_cov
        self.vt = np.zeros(3)
        self.map_xy = np.zeros((0, 2), np.float32)
        self.map_cov = np.zeros((0, 2, 2))
        self.last_match = (np.zeros((0, 2), int), np.zeros((0, 2), int), np.zeros(0, int))

    # -- motion ---------------------------------------------------------------------------
    def _jacobians(self, dt):
        n = self.mu.size
        vx, vy, w = self.vt
        th = self.mu[2]
        G = np.eye(n)
        if self.model == DIFF:
            a = th + w * dt / 2
            step = np.array([vx * dt * np.cos(a), vx * dt * np.sin(a), w * dt])
            G[0, 2] = -vx * dt * np.sin(a)
            G[1, 2] = vx * dt * np.cos(a)
            Gu = np.zeros((n, 2))
            Gu[:3] = [[dt * np.cos(a), -vx * dt * dt * np.sin(a) / 2],
                      [dt * np.sin(a), vx * dt * dt * np.cos(a) / 2],
                      [0, dt]]
        else:
            step = np.array([vx * dt * np.cos(th) - vy * dt * np.sin(th),
                             vx * dt * np.sin(th) + vy * dt * np.cos(th), w * dt])
            G[0, 2] = -vx * dt * np.sin(th) - vy * dt * np.cos(th)
            G[1, 2] = vx * dt * np.cos(th) - vy * dt * np.sin(th)
            Gu = np.zeros((n, 3))
            Gu[:3] = dt * np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
        return G, Gu, step

    def _predict(self, dt):
        G, Gu, step = self._jacobians(dt)
        self.sigma = G @ self.sigma @ G.T + Gu @ self.Qu @ Gu.T
        self.mu[:3] += step
        self.mu[2] = np.arctan2(np.sin(self.mu[2]), np.cos(self.mu[2]))

    def predict_state(self, time):
        G, Gu, step = self._jacobians(time - self.time)
        mu = self.mu.copy()
        mu[:3] += step
        mu[2] = np.arctan2(np.sin(mu[2]), np.cos(mu[2]))
        return mu, G @ self.sigma @ G.T + Gu @ self.Qu @ Gu.T

    def handle_odometry(self, time, vx, vy, wz):
        if time < self.time:
            return
        self.vt = np.array([vx, vy, wz], dtype=np.float64)
        self._predict(time - self.time)
        self.time = time

    # -- association ------------------------------------------------------------------------
    def _to_global(self, p):
        c, s = np.cos(self.mu[2]), np.sin(self.mu[2])
        x = np.float32(np.float64(p[0]) * c - np.float64(p[1]) * s + self.mu[0])
        y = np.float32(np.float64(p[0]) * s + np.float64(p[1]) * c + self.mu[1])
        return np.array([x, y], dtype=np.float32)

    def _match(self, cloud):
        state_pairs, map_pairs, new_ids = [], [], []
        N = (self.mu.size - 3) // 2
        for i, p in enumerate(cloud):
            g = self._to_global(p)
            if len(self.map_xy):
                d = (self.map_xy - g).astype(np.float64)            # float32 subtraction, then widened
                dist = np.sqrt(np.einsum("ja,jab,jb->j", d, self.map_cov, d))
                j = int(np.argmin(dist))
                if dist[j] < 0.05:
                    map_pairs.append((i, j))
                    continue
            if N:
                lm = self.mu[3:].reshape(-1, 2).astype(np.float32)
                d = (g - lm).astype(np.float64)
                dist = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
                j = int(np.argmin(dist))
                if dist[j] < 0.6:
                    state_pairs.append((i, j))
                    continue
            new_ids.append(i)
        return state_pairs, map_pairs, new_ids

    # -- update -----------------------------------------------------------------------------
    def handle_observation(self, time, cloud, gps_pose=None):
        cloud = np.asarray(cloud, dtype=np.float32).reshape(-1, 2)
        self._predict(time - self.time)
        self.time = time
        if len(cloud) == 0:
            return
        sp, mp, new = self._match(cloud)
        self.last_match = (np.array(sp, int).reshape(-1, 2), np.array(mp, int).reshape(-1, 2), np.array(new, int))
        n = self.mu.size
        pairs = [(i, j, True) for i, j in sp] + [(i, j, False) for i, j in mp]
        if pairs:
            c, s = np.cos(self.mu[2]), np.sin(self.mu[2])
            rows = 2 * len(pairs) + (3 if gps_pose is not None else 0)
            H = np.zeros((rows, n))
            innov = np.zeros(rows)
            R = np.zeros((rows, rows))
            for k, (i, j, in_state) in enumerate(pairs):
                l = self.mu[3 + 2 * j: 5 + 2 * j] if in_state else self.map_xy[j].astype(np.float64)
                d = l - self.mu[:2]
                zhat = np.array([d[0] * c + d[1] * s, -d[0] * s + d[1] * c])
                H[2 * k: 2 * k + 2, :3] = [[-c, -s, -d[0] * s + d[1] * c], [s, -c, -d[0] * c - d[1] * s]]
                if in_state:
                    H[2 * k: 2 * k + 2, 3 + 2 * j: 5 + 2 * j] = [[c, s], [-s, c]]
                innov[2 * k: 2 * k + 2] = cloud[i].astype(np.float64) - zhat
                R[2 * k: 2 * k + 2, 2 * k: 2 * k + 2] = self.Qt
            if gps_pose is not None:
                b = 2 * len(pairs)
                H[b: b + 3, :3] = np.eye(3)
                innov[b: b + 2] = np.asarray(gps_pose[:2]) - self.mu[:2]
                dth = gps_pose[2] - self.mu[2]
                innov[b + 2] = np.arctan2(np.sin(dth), np.cos(dth))
                R[b: b + 3, b: b + 3] = np.diag([0.05 ** 2, 0.05 ** 2, 0.017 ** 2])
            S = H @ self.sigma @ H.T + R
            K = self.sigma @ H.T @ np.linalg.inv(S)
            self.mu = self.mu + K @ innov
            self.mu[2] = np.arctan2(np.sin(self.mu[2]), np.cos(self.mu[2]))
            self.sigma = self.sigma - K @ H @ self.sigma
        if new:
            c, s = np.cos(self.mu[2]), np.sin(self.mu[2])
            k2 = len(new)
            mu2 = np.concatenate([self.mu, np.zeros(2 * k2)])
            Gp = np.zeros((2 * k2, 3))
            Gz = np.zeros((2 * k2, 2))
            for q, i in enumerate(new):
                mu2[n + 2 * q: n + 2 * q + 2] = self._to_global(cloud[i]).astype(np.float64)
                rx, ry = np.float64(cloud[i][0]), np.float64(cloud[i][1])
                Gp[2 * q: 2 * q + 2] = [[1, 0, -rx * s - ry * c], [0, 1, rx * c - ry * s]]
                Gz[2 * q: 2 * q + 2] = [[c, -s], [s, c]]
            Gfx = np.zeros((2 * k2, n))
            Gfx[:, :3] = Gp
            sig2 = np.zeros((n + 2 * k2, n + 2 * k2))
            sig2[:n, :n] = self.sigma
            mx = Gfx @ self.sigma
            sig2[n:, :n] = mx
            sig2[:n, n:] = mx.T
            sig2[n:, n:] = Gp @ self.sigma[:3, :3] @ Gp.T + Gz @ self.Qt @ Gz.T
            self.mu, self.sigma = mu2, sig2
