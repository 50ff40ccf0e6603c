// rekf_kernels.cuh — CUDA-core kernels of the EKF step (sm_100a).
//
//   k_odometry            HandleOdometryMessage          reflector_ekf_slam.cc:208-223 (+ Predict :154-206)
//   k_observation_front   Predict + ReflectorMatch + H/z build   :229-304, :370-455
//   k_innovation          S = H·Σ·Hᵀ + Q (5x5 block gathers)     :305
//   k_cholesky            S = L·Lᵀ, L⁻¹ν, diagonal-block inverses (replaces .inverse(), :305)
//   k_solve_w             W = L⁻¹·H·Σ (block gather + blocked TRSM), μ += Wᵀ·L⁻¹ν  (:305-307)
//   k_syrk_f64            Σ −= Wᵀ·W on the fp64 pipe (reference-accuracy mode)     (:308)
//   k_augment             landmark initialisation        :311-364
//   k_pack_* / k_unpack_* layout conversion at the C-ABI boundary
//
// Algebra.  The reference forms K = ΣHᵀS⁻¹ and Σ − K·H·Σ with dense Eigen products.  With S = L·Lᵀ and
// W = L⁻¹·(HΣ):  K·ν = Wᵀ·(L⁻¹ν)  and  K·H·Σ = Wᵀ·W, so the update is one rank-r symmetric downdate.
// H has at most five non-zeros per row (3 pose columns + the landmark's 2, :272-275), so H·Σ is a gather
// of five rows of Σ per measurement row and is never materialised outside shared memory.
#pragma once
#include <cooperative_groups.h>
#include "rekf_device.cuh"
#include "syrk_exact_rows.cuh"

namespace rekf {

// ---------------------------------------------------------------------------------------------
// motion model (Predict, :154-206)
// ---------------------------------------------------------------------------------------------
struct MotionTerms {
  double g02, g12;   // G_xi(0,2), G_xi(1,2)
  double V[9];       // G_u·Qu·G_uᵀ top-left 3x3, row-major (symmetric)
  double d[3];       // pose increment
};

__device__ inline MotionTerms motion_model(const Layout &L, const double vt[3], double theta, double dt) {
  MotionTerms t;
  const double vx = vt[0], vy = vt[1], w = vt[2];
  double Gu[9];  // 3 x q row-major, q = 2 (DIFF) or 3 (OMNI)
  double q[3];
  int nq;
  if (L.odom_model == 0) {  // DIFF :156-183
    const double delta_theta = w * dt;
    const double a = theta + delta_theta / 2;
    double sa, ca;
    sincos(a, &sa, &ca);
    t.d[0] = vx * dt * ca;
    t.d[1] = vx * dt * sa;
    t.d[2] = delta_theta;
    t.g02 = -vx * dt * sa;
    t.g12 = vx * dt * ca;
    Gu[0] = dt * ca; Gu[1] = -vx * dt * dt * sa / 2;
    Gu[3] = dt * sa; Gu[4] = vx * dt * dt * ca / 2;
    Gu[6] = 0;       Gu[7] = dt;
    Gu[2] = Gu[5] = Gu[8] = 0;
    q[0] = L.q_lin; q[1] = L.q_ang; q[2] = 0;
    nq = 2;
  } else {                  // OMNI :184-205
    double sa, ca;
    sincos(theta, &sa, &ca);
    t.d[2] = w * dt;
    t.d[0] = vx * dt * ca - vy * dt * sa;
    t.d[1] = vx * dt * sa + vy * dt * ca;
    t.g02 = -vx * dt * sa - vy * dt * ca;
    t.g12 = vx * dt * ca - vy * dt * sa;
    Gu[0] = dt * ca; Gu[1] = -dt * sa; Gu[2] = 0;
    Gu[3] = dt * sa; Gu[4] = dt * ca;  Gu[5] = 0;
    Gu[6] = 0;       Gu[7] = 0;        Gu[8] = dt;
    q[0] = L.q_lin; q[1] = L.q_lin; q[2] = L.q_ang;
    nq = 3;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < nq; ++k) s += Gu[i * 3 + k] * q[k] * Gu[j * 3 + k];
      t.V[i * 3 + j] = s;
    }
  return t;
}

__device__ inline double wrap_angle(double a) {
  double s, c;
  sincos(a, &s, &c);
  return atan2(s, c);   // :181 atan2(sin θ, cos θ)
}

// Σ ← G_xi·Σ·G_xiᵀ + G_u·Qu·G_uᵀ, μ[0:3] += d, θ wrapped — by one CTA, O(n).  G_xi is the identity
// plus two entries, so only rows/columns 0 and 1 change: row 0 += g02·row 2, row 1 += g12·row 2 (the
// columns are their mirror images and are not stored; the two n³ products of :178/:202 reduce to this
// without changing a sum).
__device__ inline void predict_cta(const Layout &L, int s, double dt) {
  SessionState &st = L.st[s];
  double *mu = L.mu + (size_t)s * L.ld;
  double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  const int n = internal_dim(st.N);
  const int ld = L.ld;
  const double vt[3] = {st.vt[0], st.vt[1], st.vt[2]};
  const double theta = mu[2];
  const MotionTerms t = motion_model(L, vt, theta, dt);
  for (int c = kPoseSlots + threadIdx.x; c < n; c += blockDim.x) {
    const double s2 = Sg[(size_t)2 * ld + c];
    Sg[c] += t.g02 * s2;
    Sg[(size_t)ld + c] += t.g12 * s2;
  }
  __syncthreads();   // everyone has read mu[2] / st before thread 0 rewrites them
  if (threadIdx.x == 0) {
    double P[9], T[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) P[i * 3 + j] = Sg[sym_idx(i, j, ld)];
    // T = G3·P (rows), P' = T·G3ᵀ (columns)
    for (int j = 0; j < 3; ++j) {
      T[0 + j] = P[0 + j] + t.g02 * P[6 + j];
      T[3 + j] = P[3 + j] + t.g12 * P[6 + j];
      T[6 + j] = P[6 + j];
    }
    for (int i = 0; i < 3; ++i) {
      P[i * 3 + 0] = T[i * 3 + 0] + t.g02 * T[i * 3 + 2];
      P[i * 3 + 1] = T[i * 3 + 1] + t.g12 * T[i * 3 + 2];
      P[i * 3 + 2] = T[i * 3 + 2];
    }
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) Sg[(size_t)i * ld + j] = P[i * 3 + j] + t.V[i * 3 + j];
    mu[0] += t.d[0];
    mu[1] += t.d[1];
    mu[2] = wrap_angle(mu[2] + t.d[2]);
  }
  __syncthreads();
}

__device__ inline int current_step(const InputRef &in) { return in.step ? *in.step : 0; }

// HandleOdometryMessage (:208-223): stale drop, latch vt_ BEFORE predicting, predict, set time.
__device__ inline void odometry_cta(const Layout &L, int s, const InputRef &in) {
  SessionState &st = L.st[s];
  const double *msg = in.odom + (size_t)s * in.odom_ss + (size_t)current_step(in) * 4;
  const double time = msg[0];
  const double t_state = st.time;
  __syncthreads();
  if (time < t_state) return;                      // :211 (block-uniform)
  if (threadIdx.x == 0) { st.vt[0] = msg[1]; st.vt[1] = msg[2]; st.vt[2] = msg[3]; }   // :216
  __syncthreads();
  predict_cta(L, s, time - t_state);               // :217-218
  if (threadIdx.x == 0) st.time = time;            // :219
  __syncthreads();
}

__global__ void __launch_bounds__(1024, 1) k_odometry(Layout L, InputRef in_arg) {
  timeline_mark(L, 0);
  const InputRef in = resolve_input(in_arg);
  odometry_cta(L, L.s0 + blockIdx.x, in);
}

// ---------------------------------------------------------------------------------------------
// observation front end: Predict (:232-234) + ReflectorMatch (:370-455) + measurement rows (:248-304)
// ---------------------------------------------------------------------------------------------
__device__ inline void warp_argmin(double &d, int &j) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, off);
    const int oj = __shfl_xor_sync(0xffffffffu, j, off);
    if (od < d || (od == d && oj < j)) { d = od; j = oj; }   // lowest index on exact ties
  }
}

constexpr int kMatchNew = 0, kMatchState = 1, kMatchMap = 2;

// One predict of the 3x3 pose block and the pose itself (thread 0): P ← G3·P·G3ᵀ + V, μ[0:3] += d, θ wrapped.
__device__ inline void predict_pose_block(const MotionTerms &t, double P[9], double pose[3]) {
  double T[9];
  for (int j = 0; j < 3; ++j) {          // T = G3·P (rows), P' = T·G3ᵀ (columns)
    T[0 + j] = P[0 + j] + t.g02 * P[6 + j];
    T[3 + j] = P[3 + j] + t.g12 * P[6 + j];
    T[6 + j] = P[6 + j];
  }
  for (int i = 0; i < 3; ++i) {
    P[i * 3 + 0] = T[i * 3 + 0] + t.g02 * T[i * 3 + 2];
    P[i * 3 + 1] = T[i * 3 + 1] + t.g12 * T[i * 3 + 2];
    P[i * 3 + 2] = T[i * 3 + 2];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      const double v = P[i * 3 + j] + t.V[i * 3 + j];
      P[i * 3 + j] = v;
      P[j * 3 + i] = v;
    }
  pose[0] += t.d[0];
  pose[1] += t.d[1];
  pose[2] = wrap_angle(pose[2] + t.d[2]);
}

// The kernel is latency-bound (one CTA per session, a few KB of traffic), so it is organised around its dependent global
// round trips: everything that does not depend on the message — rows 0..2 of Σ, the landmark means — is requested by
// all threads at entry; meanwhile thread 0 walks the only serial chain (input descriptor → step counter → message →
// both motion models → 3x3 block and pose) and warp 1 stages the frame's observations.  After ONE barrier every thread
// applies both predicts to its columns from registers (the two updates in the reference's order, so no sum changes), and
// association, compaction (ballot + popc prefix) and the measurement rows run out of shared memory.
struct FrontShared {
  double g[4];            // g02, g12 of the odometry predict (0 when there is none), g02, g12 of the observation predict
  double pose[3];         // pose after both predicts
  double sn, cs;
  double time;            // observation stamp
  int n, N, m, flags0, has_odom, t_idx;
  const float *xy;
  const double *gps;
  int counts[3];
  double P[9], vt[3];     // the predicted 3x3 block and the latched velocity: written to global memory after the cluster barrier
  int vt_latched;
};

__global__ void __launch_bounds__(1024, 1) k_observation_front(Layout L, InputRef in_arg) {
  pdl_wait();                        // inside a multi-step graph: the previous step's k_augment is complete
  // first phase of the cluster barrier: a CTA's shared memory may only be written from another CTA once that CTA has started.
  // Arrive here, wait just before the first remote store (the chain and the staging lie in between: nobody actually waits)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  timeline_mark(L, 1);
  extern __shared__ int sm_i[];
  __shared__ FrontShared fs;
  // A cluster of kFrontCluster CTAs per session shares the association (the 10^5 distance tests are instruction-bound on one
  // SM: 8 of this kernel's 15 µs): every CTA walks the (read-only) serial chain and stages the landmark means itself, takes
  // every kFrontCluster-th observation and writes its decisions into rank 0's shared memory; rank 0 alone writes global state —
  // the predicted pose block only AFTER the cluster barrier, so that no other rank can read it half-updated — and goes on.
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int rank = (int)cluster.block_rank(), nrank = (int)cluster.num_blocks();   // 0 / 1 without a cluster launch
  const int s = L.s0 + blockIdx.x / nrank;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int ld = L.ld;
  SessionState &st = L.st[s];
  double *mu = L.mu + (size_t)s * ld;
  double *Sg = L.sigma + (size_t)s * ld * ld;
  int *kind = sm_i;                 // [mcap]
  int *target = sm_i + L.mcap;      // [mcap]
  // landmark means: float32 once per frame (:431 rounds per pair; F2F.F32.F64 is a slow-pipe op) and the doubles for the rows
  float2 *lmf = reinterpret_cast<float2 *>(sm_i + 2 * L.mcap);                    // [Ncap]
  double2 *lmd = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(lmf + L.Ncap) + 15) & ~(uintptr_t)15);   // [Ncap], 16-byte aligned
  float2 *xys = reinterpret_cast<float2 *>(lmd + L.Ncap);                          // [mcap] this frame's observations

  if (L.sync && rank == 0)           // pacing flags of this frame's Cholesky / gather / TRSM (solve_ll.cuh)
    for (int i = tid; i < L.sync_n; i += blockDim.x) L.sync[(size_t)s * L.sync_n + i] = 0;
  // ---- requests that do not depend on the message ------------------------------------------------------------
  constexpr int kPre = 2;           // columns per thread held in registers (covers Ncap <= 1022; the rest goes through a loop)
  double r0[kPre], r1[kPre], r2[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int c = kPoseSlots + tid + u * 1024;
    r0[u] = r1[u] = r2[u] = 0.0;
    if (c < L.ncap && rank == 0) { r0[u] = Sg[c]; r1[u] = Sg[(size_t)ld + c]; r2[u] = Sg[(size_t)2 * ld + c]; }
  }
  for (int j = tid; j < L.Ncap; j += blockDim.x) {          // slots past the live N hold zeros: harmless
    const double2 l = *reinterpret_cast<const double2 *>(mu + kPoseSlots + 2 * j);
    if (rank == 0) lmd[j] = l;
    lmf[j] = make_float2((float)l.x, (float)l.y);
  }

  // ---- the serial chain (thread 0) and the observation staging (warp 1) ----------------------------------------
  if (tid == 0) {
    const InputRef in = resolve_input(in_arg);
    const int t_idx = current_step(in);
    const double *msg = in.odom ? in.odom + (size_t)s * in.odom_ss + (size_t)t_idx * 4 : nullptr;
    const double t_obs = in.obs_time[(size_t)s * in.time_ss + t_idx];
    int m = in.obs_count ? in.obs_count[s] : in.m_fixed;
    double t_state = st.time;
    double vt[3] = {st.vt[0], st.vt[1], st.vt[2]};
    double pose[3] = {mu[0], mu[1], mu[2]};
    double P[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) P[i * 3 + j] = Sg[sym_idx(i, j, ld)];
    const int N = st.N;
    fs.flags0 = st.flags;
    fs.g[0] = fs.g[1] = 0.0;
    fs.has_odom = 0;
    fs.vt_latched = 0;
    if (in.fuse_odom) {                            // replay / step call: this step's HandleOdometryMessage first (:208-223)
      const double t_od = msg[0];
      if (!(t_od < t_state)) {                     // :211 stale messages are dropped
        vt[0] = msg[1]; vt[1] = msg[2]; vt[2] = msg[3];      // :216 latched BEFORE predicting
        const MotionTerms t = motion_model(L, vt, pose[2], t_od - t_state);
        fs.g[0] = t.g02; fs.g[1] = t.g12;
        fs.has_odom = 1;
        predict_pose_block(t, P, pose);
        t_state = t_od;
        fs.vt_latched = 1;
      }
    }
    const MotionTerms t = motion_model(L, vt, pose[2], t_obs - t_state);   // :232-233, no sign check on dt
    fs.g[2] = t.g02; fs.g[3] = t.g12;
    predict_pose_block(t, P, pose);
    for (int i = 0; i < 9; ++i) fs.P[i] = P[i];
    fs.vt[0] = vt[0]; fs.vt[1] = vt[1]; fs.vt[2] = vt[2];
    fs.pose[0] = pose[0]; fs.pose[1] = pose[1]; fs.pose[2] = pose[2];
    sincos(pose[2], &fs.sn, &fs.cs);
    int flag_add = 0;
    if (m > L.mcap) { m = L.mcap; flag_add |= FLAG_OBS_CAPACITY; }
    if (m < 0) m = 0;
    fs.flags0 |= flag_add;
    fs.time = t_obs; fs.N = N; fs.n = internal_dim(N); fs.m = m; fs.t_idx = t_idx;
    fs.gps = in.gps ? in.gps + 4 * s : nullptr;
    fs.counts[0] = fs.counts[1] = fs.counts[2] = 0;
  } else if (warp == 1) {
    const InputRef in = resolve_input(in_arg);
    const int t_idx = current_step(in);
    const float *xy = in.obs_xy + (size_t)s * in.xy_ss + (size_t)t_idx * in.m_stride * 2;
    int m = in.obs_count ? in.obs_count[s] : in.m_fixed;
    m = max(0, min(m, L.mcap));
    for (int i = lane; i < m; i += 32) xys[i] = *reinterpret_cast<const float2 *>(xy + 2 * i);
    if (lane == 0) fs.xy = xy;
  }
  const int Mmap = *L.map_count;
  __syncthreads();

  // ---- both predicts on rows 0 and 1 (the mirrored columns are not stored) ---------------------------------------
  const int n = fs.n, N = fs.N, m = fs.m;
  if (rank == 0) {
    const double ga0 = fs.g[0], ga1 = fs.g[1], gb0 = fs.g[2], gb1 = fs.g[3];
    const bool has_odom = fs.has_odom != 0;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int c = kPoseSlots + tid + u * 1024;
      if (c < n) {
        double v0 = r0[u], v1 = r1[u];
        if (has_odom) { v0 += ga0 * r2[u]; v1 += ga1 * r2[u]; }
        v0 += gb0 * r2[u]; v1 += gb1 * r2[u];
        Sg[c] = v0;
        Sg[(size_t)ld + c] = v1;
      }
    }
    for (int c = kPoseSlots + tid + kPre * 1024; c < n; c += blockDim.x) {
      const double s2 = Sg[(size_t)2 * ld + c];
      double v0 = Sg[c], v1 = Sg[(size_t)ld + c];
      if (has_odom) { v0 += ga0 * s2; v1 += ga1 * s2; }
      v0 += gb0 * s2; v1 += gb1 * s2;
      Sg[c] = v0;
      Sg[(size_t)ld + c] = v1;
    }
  }
  const double px = fs.pose[0], py = fs.pose[1], th = fs.pose[2], sn = fs.sn, cs = fs.cs;

  // --- ReflectorMatch: one warp per observation, lanes stride over landmarks; this CTA takes observations rank, rank + nrank, … -
  int *kind0 = cluster.map_shared_rank(kind, 0), *target0 = cluster.map_shared_rank(target, 0);
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // every CTA of the cluster is running
  for (int i = rank + nrank * warp; i < m; i += nrank * nwarps) {
    // point_transformed_to_global_frame (:389-393): double arithmetic, float32 result
    const float2 o = xys[i];
    const double ox = (double)o.x, oy = (double)o.y;
    const float gx = (float)(ox * cs - oy * sn + px);
    const float gy = (float)(ox * sn + oy * cs + py);
    int k = kMatchNew, tgt = -1;
    if (Mmap > 0) {                                // :401-425
      double best = INFINITY;
      int bj = 0x7fffffff;
      for (int j = lane; j < Mmap; j += 32) {
        const double *C = L.map_cov + 4 * j;
        const float dfx = L.map_xy[2 * j] - gx, dfy = L.map_xy[2 * j + 1] - gy;   // :408 float subtraction
        const double dx = (double)dfx, dy = (double)dfy;
        const double t0 = dx * C[0] + dy * C[2], t1 = dx * C[1] + dy * C[3];
        const double dist = sqrt(t0 * dx + t1 * dy);                               // :411 Σ not inverted
        if (dist < best) { best = dist; bj = j; }
      }
      warp_argmin(best, bj);
      if (best < 0.05) { k = kMatchMap; tgt = bj; }                                // :420
    }
    if (k == kMatchNew && N > 0) {                 // :426-451
      double best = INFINITY;
      int bj = 0x7fffffff;
      // The reference takes the nearest landmark and then asks whether it is within 0.6 m (:446), so only landmarks inside the
      // gate can matter: an fp32 estimate of d² (two FP32 ops per pair) screens the N candidates with a margin far above its
      // rounding error, and the exact comparison — the float differences widened to fp64, d² exact there, lowest index on
      // ties — runs on the one or two survivors.  (Widening every pair cost two quarter-rate F2F per pair: 2·10^5 per frame.)
      for (int j = lane; j < N; j += 32) {
        const float2 l = lmf[j];
        const float dfx = gx - l.x, dfy = gy - l.y;                                // :431, :433
        if (fmaf(dfx, dfx, dfy * dfy) < 0.3601f) {
          const double dx = (double)dfx, dy = (double)dfy;
          const double d2 = dx * dx + dy * dy;                                     // :437 Euclidean
          if (d2 < best) { best = d2; bj = j; }
        }
      }
      warp_argmin(best, bj);
      if (sqrt(best) < 0.6) { k = kMatchState; tgt = bj; }                         // :446
    }
    if (lane == 0) { kind0[i] = k; target0[i] = tgt; }
  }
  cluster.sync();                                  // every rank's decisions are in rank 0's shared memory
  if (rank != 0) return;
  if (tid == 0) {                                  // the chain's global writes, now that no other rank reads the old values
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j) Sg[(size_t)i * ld + j] = fs.P[i * 3 + j];
    mu[0] = fs.pose[0]; mu[1] = fs.pose[1]; mu[2] = fs.pose[2];
    if (fs.vt_latched) { st.vt[0] = fs.vt[0]; st.vt[1] = fs.vt[1]; st.vt[2] = fs.vt[2]; }
  }

  // --- ordered compaction: lists keep observation order like the push_backs at :422/:448/:452.  One warp per list kind:
  //     ballot over 32 observations at a time, position = running count + popc of the lower lanes ----------------------
  int *sp = L.state_pairs + (size_t)s * L.mcap * 2;
  int *mp = L.map_pairs + (size_t)s * L.mcap * 2;
  int *nw = L.new_ids + (size_t)s * L.mcap;
  if (warp < 3) {
    const int mine = warp;                         // kMatchNew = 0, kMatchState = 1, kMatchMap = 2
    int run = 0;
    for (int base = 0; base < m; base += 32) {
      const int i = base + lane;
      const bool hit = i < m && kind[i] == mine;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = run + __popc(bal & ((1u << lane) - 1u));
        if (mine == kMatchState) { sp[2 * pos] = i; sp[2 * pos + 1] = target[i]; }
        else if (mine == kMatchMap) { mp[2 * pos] = i; mp[2 * pos + 1] = target[i]; }
        else nw[pos] = i;
      }
      run += __popc(bal);
    }
    if (lane == 0) fs.counts[mine] = run;
  }
  __syncthreads();
  const int M = fs.counts[kMatchState], Mm = fs.counts[kMatchMap], N2 = fs.counts[kMatchNew];
  const int MM = M + Mm;
  const double *gps = fs.gps;
  const bool has_gps = gps && gps[0] != 0.0 && MM > 0;   // the GPS rows live inside `if (MM > 0)` (gps.cc:246,305)
  const int r = MM > 0 ? 2 * MM + (has_gps ? 3 : 0) : 0;

  pdl_trigger();                                   // k_innovation may be launched: it waits for this kernel's end in pdl_wait()
  // --- measurement rows: A_k (:272-273), B (:255), z − ẑ (:265-270), Q (:276) ----------------
  double *Hp = L.Hp + (size_t)s * L.rcap * 4;
  double *Hl = L.Hl + (size_t)s * L.rcap * 2;
  int *Hslot = L.Hslot + (size_t)s * L.rcap;
  double *innov = L.innov + (size_t)s * L.rcap;
  double *Qd = L.Qd + (size_t)s * L.rcap;
  for (int k = threadIdx.x; k < MM; k += blockDim.x) {
    int i, slot;
    double lx, ly;
    if (k < M) {
      i = sp[2 * k];
      const int j = sp[2 * k + 1];
      slot = kPoseSlots + 2 * j;
      lx = lmd[j].x; ly = lmd[j].y;
    } else {                                       // beacon read as float32 (:281-283), no B block (:300)
      i = mp[2 * (k - M)];
      const int j = mp[2 * (k - M) + 1];
      slot = -1;
      lx = (double)L.map_xy[2 * j]; ly = (double)L.map_xy[2 * j + 1];
    }
    const double dx = lx - px, dy = ly - py;
    const double zh0 = dx * cs + dy * sn;
    const double zh1 = -dx * sn + dy * cs;
    double *a0 = Hp + 8 * k, *a1 = a0 + 4;
    a0[0] = -cs; a0[1] = -sn; a0[2] = -dx * sn + dy * cs; a0[3] = 0;
    a1[0] = sn;  a1[1] = -cs; a1[2] = -dx * cs - dy * sn; a1[3] = 0;
    Hl[4 * k + 0] = cs;  Hl[4 * k + 1] = sn;
    Hl[4 * k + 2] = -sn; Hl[4 * k + 3] = cs;
    Hslot[2 * k] = slot; Hslot[2 * k + 1] = slot;
    innov[2 * k] = (double)xys[i].x - zh0;
    innov[2 * k + 1] = (double)xys[i].y - zh1;
    Qd[2 * k] = L.q_obs; Qd[2 * k + 1] = L.q_obs;
  }
  if (has_gps && threadIdx.x < 3) {                // reflector_ekf_slam_gps.cc:314-334
    const int b = threadIdx.x, q = 2 * MM + b;
    double *a = Hp + 4 * q;
    a[0] = b == 0; a[1] = b == 1; a[2] = b == 2; a[3] = 0;
    Hl[2 * q] = 0; Hl[2 * q + 1] = 0;
    Hslot[q] = -1;
    if (b < 2) {
      innov[q] = gps[1 + b] - (b == 0 ? px : py);
      Qd[q] = 0.05 * 0.05;
    } else {
      const double dth = gps[3] - th;
      double qz, qw;
      sincos(dth / 2, &qz, &qw);
      const double nrm = sqrt(qw * qw + qz * qz);
      qw /= nrm; qz /= nrm;
      if (qw < 0.) { qw = -qw; qz = -qz; }
      const double angle = 2. * atan2(fabs(qz), qw);
      const double scale = angle < 1e-7 ? 2. : angle / sin(angle / 2.);
      innov[q] = scale * qz;
      Qd[q] = 0.017 * 0.017;
    }
  }
  if (threadIdx.x == 0) {
    st.time = fs.time;                             // :234
    st.m = m; st.M = M; st.Mmap = Mm; st.N2 = N2; st.r = r;
    st.flags = fs.flags0;
    st.ticket = 0;
    st.exact_update = 0;
    st.exact_slots = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// S = H·Σ·Hᵀ + Q (lower triangle) and the ν row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_innovation(Layout L) {
  pdl_trigger();
  pdl_wait();
  timeline_mark(L, 2);
  const int s = L.s0 + blockIdx.z;
  const SessionState &st = L.st[s];
  const int r = st.r;
  const int q = blockIdx.x * 16 + threadIdx.x;   // row
  const int p = blockIdx.y * 16 + threadIdx.y;   // column
  if (r == 0 || q >= r || p > q) return;
  const double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  const double *Hp = L.Hp + (size_t)s * L.rcap * 4;
  const double *Hl = L.Hl + (size_t)s * L.rcap * 2;
  const int *Hslot = L.Hslot + (size_t)s * L.rcap;
  const int ld = L.ld;
  int ia[5], ib[5];
  double ha[5], hb[5];
  const int sq = Hslot[q], spp = Hslot[p];
  for (int u = 0; u < 3; ++u) { ia[u] = u; ha[u] = Hp[4 * q + u]; ib[u] = u; hb[u] = Hp[4 * p + u]; }
  const int na = sq >= 0 ? 5 : 3, nb = spp >= 0 ? 5 : 3;
  if (sq >= 0) { ia[3] = sq; ia[4] = sq + 1; ha[3] = Hl[2 * q]; ha[4] = Hl[2 * q + 1]; }
  if (spp >= 0) { ib[3] = spp; ib[4] = spp + 1; hb[3] = Hl[2 * p]; hb[4] = Hl[2 * p + 1]; }
  double acc = 0;
  for (int b = 0; b < nb; ++b) {       // ((H·Σ)·Hᵀ): y_b = Σ_a H[q,a]·Σ[a,b]
    double y = 0;
    for (int a = 0; a < na; ++a) y += ha[a] * Sg[sym_idx(ia[a], ib[b], ld)];
    acc += y * hb[b];
  }
  if (p == q) acc += L.Qd[(size_t)s * L.rcap + q];
  double *Sb = L.Sbuf + (size_t)s * L.rld * L.sld;
  Sb[(size_t)p * L.sld + q] = acc;
  if (p == 0) Sb[(size_t)q * L.sld + r] = L.innov[(size_t)s * L.rcap + q];   // ν as row r: Cholesky turns it into L⁻¹ν
}

// ---------------------------------------------------------------------------------------------
// Y = H·Σ (r x n) into global memory, off the critical path: it only needs the match lists and the predicted Σ, so it runs on a
// side stream (a parallel branch of the step graph) beside k_innovation and the Cholesky, and k_solve_w3 starts from coalesced
// 256-byte rows instead of gathering.  H has <= 5 non-zeros per row (:272-275): Y[q][c] = A_q·Σ[0:3][c] + B_q·Σ[slot:slot+2][c].
// Only the upper triangle of Σ is stored: Σ[slot][c] with slot < c is row `slot` (coalesced across the lanes' columns); for
// columns left of the slot it is Σ[c][slot..slot+1], ONE 16-byte read per lane for both rows of the reflector.
// grid (ld/128, ceil(rcap/2/16), Sg) x 256 threads: lane = column (4 per lane), warps stride 16 row pairs.
// ---------------------------------------------------------------------------------------------
constexpr int kGYPairs = 16;
__global__ void __launch_bounds__(256) k_gather_y(Layout L) {
  timeline_mark(L, 8);
  const int s = L.s0 + blockIdx.z;
  const SessionState &st = L.st[s];
  const int r = st.r;
  const int pair0 = blockIdx.y * kGYPairs;
  if (r == 0) return;
  const int n = internal_dim(st.N);
  const int ld = L.ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double *Sg = L.sigma + (size_t)s * ld * ld;
  const double *Hp = L.Hp + (size_t)s * L.rcap * 4;
  const double *Hl = L.Hl + (size_t)s * L.rcap * 2;
  const int *Hslot = L.Hslot + (size_t)s * L.rcap;
  double *Y = L.Ybuf + (size_t)s * L.rld * ld;
  const int cbase = blockIdx.x * 128;
  const bool work = 2 * pair0 < r && cbase < round_up(n, kSigmaTile);
  double p0[4], p1[4], p2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = cbase + 32 * u + lane;
    p0[u] = p1[u] = p2[u] = 0.0;
    if (work) { p0[u] = Sg[sym_idx(0, c, ld)]; p1[u] = Sg[sym_idx(1, c, ld)]; p2[u] = Sg[sym_idx(2, c, ld)]; }
  }
  for (int pb = pair0 + warp; work && pb < pair0 + kGYPairs && 2 * pb < r; pb += 8) {
    const int q0 = 2 * pb, q1 = q0 + 1;
    const bool two = q1 < r;
    const int slot0 = Hslot[q0], slot1 = two ? Hslot[q1] : -1;
    double xa[4], xb[4], za[4], zb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = cbase + 32 * u + lane;
      xa[u] = xb[u] = za[u] = zb[u] = 0.0;
      if (slot0 >= 0) {
        if (slot0 > c) { const double2 v = *reinterpret_cast<const double2 *>(Sg + (size_t)c * ld + slot0); xa[u] = v.x; xb[u] = v.y; }
        else { xa[u] = Sg[(size_t)slot0 * ld + c]; xb[u] = Sg[sym_idx(slot0 + 1, c, ld)]; }
      }
      if (slot1 == slot0) { za[u] = xa[u]; zb[u] = xb[u]; }
      else if (slot1 >= 0) {                       // never the case for the reference's row layout; kept general
        za[u] = Sg[sym_idx(slot1, c, ld)]; zb[u] = Sg[sym_idx(slot1 + 1, c, ld)];
      }
    }
    const double a0 = Hp[4 * q0], a1 = Hp[4 * q0 + 1], a2 = Hp[4 * q0 + 2], l0 = Hl[2 * q0], l1 = Hl[2 * q0 + 1];
    double b0 = 0, b1 = 0, b2 = 0, m0 = 0, m1 = 0;
    if (two) { b0 = Hp[4 * q1]; b1 = Hp[4 * q1 + 1]; b2 = Hp[4 * q1 + 2]; m0 = Hl[2 * q1]; m1 = Hl[2 * q1 + 1]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = cbase + 32 * u + lane;
      const bool live = c < n;
      double y0 = a0 * p0[u] + a1 * p1[u] + a2 * p2[u];
      if (slot0 >= 0) y0 += l0 * xa[u] + l1 * xb[u];
      Y[(size_t)q0 * ld + c] = live ? y0 : 0.0;
      if (two) {
        double y1 = b0 * p0[u] + b1 * p1[u] + b2 * p2[u];
        if (slot1 >= 0) y1 += m0 * za[u] + m1 * zb[u];
        Y[(size_t)q1 * ld + c] = live ? y1 : 0.0;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// blocked left-looking Cholesky of S in one CTA; row r (ν) rides along and leaves as L⁻¹ν.
// Also emits the inverse of every 32x32 diagonal block of L for the TRSM in k_solve_w.
// ---------------------------------------------------------------------------------------------
constexpr int kPS = kCholNb + 1;   // padded panel pitch in shared memory

// only_oversize: take only the frames the two-level resident path (chol_smem.cuh) cannot (r beyond 2·208 − 16 rows)
__global__ void __launch_bounds__(1024, 1) k_cholesky(Layout L, int only_oversize) {
  extern __shared__ double sm_d[];
  const int s = L.s0 + blockIdx.x;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  if (only_oversize && (r <= 208 || (32 * (r / 64) <= 208 && r - 32 * (r / 64) <= 208))) return;
  const int sld = L.sld;
  double *Sb = L.Sbuf + (size_t)s * L.rld * sld;
  double *P = sm_d;                                  // [(rcap+1)][kPS] current block column (rows J..r)
  double *X = sm_d + (size_t)(L.rcap + 1) * kPS;     // [32][kPS] inverse of the diagonal block
  const int tid = threadIdx.x, NT = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  bool bad = false;

  for (int J = 0; J < r; J += kCholNb) {
    const int jb = min(kCholNb, r - J);
    const int rows = r - J + 1;                      // incl. the ν row
    // load the block column (lower part)
    for (int e = tid; e < rows * jb; e += NT) {
      const int jj = e / rows, i = e - jj * rows;
      P[i * kPS + jj] = (i >= jj) ? Sb[(size_t)(J + jj) * sld + J + i] : 0.0;
    }
    __syncthreads();
    // left-looking update with the finished columns 0..J-1:  P[i][jj] -= Σ_k L[J+i][k]·L[J+jj][k]
    if (J > 0) {
      for (int task = tid; task < rows * 4; task += NT) {
        const int jg = task / rows, i = task - jg * rows;
        double acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.0;
        const double *Li = Sb + J + i;
        const double *Lj = Sb + J + jg * 8;
#pragma unroll 2
        for (int k = 0; k < J; ++k) {
          const double li = Li[(size_t)k * sld];
          const double *lj = Lj + (size_t)k * sld;
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = fma(li, lj[u], acc[u]);   // lj[u] beyond jb hits finished rows: harmless
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int jj = jg * 8 + u;
          if (jj < jb && i >= jj) P[i * kPS + jj] -= acc[u];
        }
      }
      __syncthreads();
    }
    // factor the jb x jb diagonal block (right-looking, two barriers per column)
    {
      const int i = tid & 31, jj = tid >> 5;         // element (i, jj) of the block
      const bool mine = (i < jb && jj < jb && i >= jj);
      for (int j = 0; j < jb; ++j) {
        const double d = P[j * kPS + j];
        if (!(d > 0.0)) bad = true;
        const double inv = rsqrt(d);
        double li = 0, lj = 0;
        if (mine && jj >= j) { li = P[i * kPS + j]; lj = P[jj * kPS + j]; }
        __syncthreads();
        if (mine) {
          if (jj == j) P[i * kPS + j] = (i == j) ? d * inv : li * inv;
          else if (jj > j) P[i * kPS + jj] -= (li * inv) * (lj * inv);
        }
        __syncthreads();
      }
    }
    // X = D⁻¹ (lower triangular): warp c owns column c, lane i owns X[i][c]; column-oriented substitution
    if (warp < jb) {
      const int c = warp;
      const double dii = (lane < jb) ? P[lane * kPS + lane] : 1.0;
      const double invd = 1.0 / dii;
      double t = (lane == c) ? 1.0 : 0.0;            // running rhs for row `lane`
      double x = 0.0;
      for (int k = c; k < jb; ++k) {
        const double xk = __shfl_sync(0xffffffffu, t * invd, k);   // lane k finalises x_k = t_k / D[k][k]
        if (lane == k) x = xk;
        if (lane > k && lane < jb) t = fma(-P[lane * kPS + k], xk, t);
      }
      if (lane < jb) X[lane * kPS + c] = (lane >= c) ? x : 0.0;
    } else if (warp < kCholNb) {
      if (lane < kCholNb) X[lane * kPS + warp] = 0.0;
    }
    __syncthreads();
    // rows below the diagonal block (incl. ν): P[i][:] ← P[i][:]·D⁻ᵀ, i.e. out[jj] = Σ_{k<=jj} P[i][k]·X[jj][k]
    // Four threads per row (8 columns each); results are held in registers across a barrier because a
    // row's threads read entries their neighbours overwrite.
    {
      const int below = rows - jb;
      for (int base = 0; base < below * 4; base += NT) {   // tasks of one row never straddle a pass (NT % 4 == 0)
        const int task = base + tid;
        const bool act = task < below * 4;
        const int i = jb + (task >> 2), jg = task & 3;
        double out[8];
        if (act) {
          const double *row = P + i * kPS;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int jj = jg * 8 + u;
            const double *x = X + jj * kPS;
            double a0 = 0.0, a1 = 0.0;
            if (jj < jb) {
              int k = 0;
              for (; k + 1 <= jj; k += 2) { a0 = fma(row[k], x[k], a0); a1 = fma(row[k + 1], x[k + 1], a1); }
              if (k <= jj) a0 = fma(row[k], x[k], a0);
            }
            out[u] = a0 + a1;
          }
        }
        __syncthreads();
        if (act) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int jj = jg * 8 + u;
            if (jj < jb) P[i * kPS + jj] = out[u];
          }
        }
        __syncthreads();
      }
    }
    // write the finished block column of L and the block inverse
    for (int e = tid; e < rows * jb; e += NT) {
      const int jj = e / rows, i = e - jj * rows;
      if (i >= jj) Sb[(size_t)(J + jj) * sld + J + i] = P[i * kPS + jj];
    }
    double *Dg = L.Dinv + ((size_t)s * (L.rld / kCholNb) + J / kCholNb) * kCholNb * kCholNb;
    for (int e = tid; e < kCholNb * kCholNb; e += NT) {
      const int i = e >> 5, k = e & 31;
      Dg[e] = (i < jb && k < jb) ? X[i * kPS + k] : 0.0;
    }
    __syncthreads();
  }
  if (bad && tid == 0) atomicOr(&st.flags, FLAG_NOT_SPD);
}

// ---------------------------------------------------------------------------------------------
// W = L⁻¹·(H·Σ) for 16 columns per CTA: block-gather from Σ into shared memory, blocked forward
// substitution, then   μ[c] += Σ_k W[k][c]·(L⁻¹ν)[k]   and the tf32 hi/lo (or fp64) panels of Wᵀ.
// ---------------------------------------------------------------------------------------------
constexpr int kYS = kWCols + 1;    // padded pitch of the Y/W block in shared memory

__device__ inline float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) k_solve_w(Layout L) {
  extern __shared__ double sm_d[];
  const int s = L.s0 + blockIdx.z;
  const SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int n = internal_dim(st.N);
  const int c0 = blockIdx.x * kWCols;
  if (c0 >= round_up(n, kSigmaTile)) return;         // beyond the tiles the SYRK will touch
  const int ld = L.ld, sld = L.sld, rld = L.rld;
  const double *Sg = L.sigma + (size_t)s * ld * ld;
  const double *Sb = L.Sbuf + (size_t)s * rld * sld;
  const double *Hp = L.Hp + (size_t)s * L.rcap * 4;
  const double *Hl = L.Hl + (size_t)s * L.rcap * 2;
  const int *Hslot = L.Hslot + (size_t)s * L.rcap;
  double *Y = sm_d;                                  // [rld][kYS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // gather: Y[q][cc] = Σ_b H[q][b]·Σ[b][c]  (Σ symmetric: read row c)
  for (int cc = warp; cc < kWCols; cc += 8) {
    const int c = c0 + cc;
    if (c < n) {
      const double p0 = Sg[sym_idx(0, c, ld)], p1 = Sg[sym_idx(1, c, ld)], p2 = Sg[sym_idx(2, c, ld)];
      for (int q = lane; q < r; q += 32) {
        const double *h = Hp + 4 * q;
        double y = h[0] * p0 + h[1] * p1 + h[2] * p2;
        const int slot = Hslot[q];
        if (slot >= 0) y += Hl[2 * q] * Sg[sym_idx(slot, c, ld)] + Hl[2 * q + 1] * Sg[sym_idx(slot + 1, c, ld)];
        Y[q * kYS + cc] = y;
      }
    } else {
      for (int q = lane; q < r; q += 32) Y[q * kYS + cc] = 0.0;
    }
  }
  __syncthreads();

  // blocked forward substitution L·W = Y
  const double *Dinv = L.Dinv + (size_t)s * (rld / kCholNb) * kCholNb * kCholNb;
  for (int J = 0; J < r; J += kCholNb) {
    const int jb = min(kCholNb, r - J);
    const double *Dg = Dinv + (size_t)(J / kCholNb) * kCholNb * kCholNb;
    // W_J = D_J⁻¹·Y_J : 32x16 outputs, two per thread
    const int cc = tid & 15;
    double w[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = (tid >> 4) + 16 * h;
      double a = 0.0;
      if (i < jb)
        for (int k = 0; k <= i; ++k) a = fma(Dg[i * kCholNb + k], Y[(J + k) * kYS + cc], a);
      w[h] = a;
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = (tid >> 4) + 16 * h;
      if (i < jb) Y[(J + i) * kYS + cc] = w[h];
    }
    __syncthreads();
    // trailing rows: Y[i2][:] -= L[i2][J..J+jb)·W_J
    const int below = r - (J + jb);
    for (int task = tid; task < below * 4; task += blockDim.x) {
      const int cg = task / below, i2 = J + jb + (task - cg * below);
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double *Lr = Sb + (size_t)J * sld + i2;
      for (int k = 0; k < jb; ++k) {
        const double l = Lr[(size_t)k * sld];
        const double *wk = Y + (J + k) * kYS + cg * 4;
        a0 = fma(l, wk[0], a0); a1 = fma(l, wk[1], a1); a2 = fma(l, wk[2], a2); a3 = fma(l, wk[3], a3);
      }
      double *y = Y + i2 * kYS + cg * 4;
      y[0] -= a0; y[1] -= a1; y[2] -= a2; y[3] -= a3;
    }
    __syncthreads();
  }

  // μ += Wᵀ·(L⁻¹ν)  (:306), θ wrapped (:307)
  double *mu = L.mu + (size_t)s * ld;
  for (int cc = warp; cc < kWCols; cc += 8) {
    const int c = c0 + cc;
    double a = 0.0;
    double d2 = 0.0;
    for (int k = lane; k < r; k += 32) {
      const double w = Y[k * kYS + cc];
      a = fma(w, Sb[(size_t)k * sld + r], a);
      d2 = fma(w, w, d2);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      d2 += __shfl_xor_sync(0xffffffffu, d2, off);
    }
    if (lane == 0 && c < n) {
      const double v = mu[c] + a;
      mu[c] = (c == 2) ? wrap_angle(v) : v;
    }
    // The diagonal of the downdate is a sum of squares: every rounding/truncation of the tensor-core
    // product has the same sign there and would accumulate step after step, so it is kept in fp64.
    if (lane == 0 && L.Wdiag) L.Wdiag[(size_t)s * ld + c] = (c < n) ? d2 : 0.0;
  }
  // Wᵀ panels (row c, K contiguous), zero beyond r
  if (L.W64) {
    double *W = L.W64 + (size_t)s * ld * rld;
    for (int e = tid; e < kWCols * rld; e += blockDim.x) {
      const int k = e / kWCols, cc = e - k * kWCols;
      W[(size_t)k * ld + c0 + cc] = (k < r) ? Y[k * kYS + cc] : 0.0;
    }
  }
  if (L.Wq) {
    // exact int8 digit slices for the kind::i8 SYRK: x = w·2^-e, |x| <= 1/2, x ≈ Σ_p d_p·2^(-7(p+1))
    __shared__ int sexp[kWCols];
    for (int cc = warp; cc < kWCols; cc += 8) {
      double mx = 0.0;
      for (int k = lane; k < r; k += 32) mx = fmax(mx, fabs(Y[k * kYS + cc]));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const int e = (mx > 0.0 && c0 + cc < n) ? ilogb(mx) + 2 : 0;
      if (lane == 0) {
        const int c = c0 + cc;
        sexp[cc] = e; L.Wexp[(size_t)s * ld + c] = e;
        L.Wscale[(size_t)s * ld + c] = scalbn(1.0, e);
        // The slices resolve 2^-29 of the row scale 2^e.  When the downdate removes almost all of a
        // state's variance (first update after dead reckoning, loop closure) that is no longer small
        // against the posterior: such slots are flagged for k_syrk_exact_rows (same rule as k_solve_w3);
        // more than kMaxExactSlots of them and the whole frame takes the fp64 SYRK.
        bool exact = false;
        if (mx > 0.0 && c < n) {
          const double post = Sg[(size_t)c * ld + c] - L.Wdiag[(size_t)s * ld + c];
          exact = !(post > 0.0) || scalbn(1.0, 2 * e) > kMaxSliceGain2 * post;
          if (exact) {
            const int pos = atomicAdd(&L.st[s].exact_slots, 1);
            if (pos < kMaxExactSlots) L.exact_list[(size_t)s * kMaxExactSlots + pos] = c;
            else atomicOr(&L.st[s].exact_update, 1);
          }
        }
        L.Wflag[(size_t)s * ld + c] = exact ? 1 : 0;
      }
    }
    __syncthreads();
    const int kq4 = L.kq / 4;
    for (int e4 = tid; e4 < kWCols * kq4; e4 += blockDim.x) {
      const int cc = e4 / kq4, k0 = (e4 - cc * kq4) * 4;
      const int e = sexp[cc];
      const bool live = (c0 + cc < n);
      uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u;
        double rem = (live && k < r) ? scalbn(Y[k * kYS + cc], 7 - e) : 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const double d = rint(rem);
          packed[p] |= ((uint32_t)(int)d & 0xffu) << (8 * u);
          rem = (rem - d) * 128.0;
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
        *reinterpret_cast<uint32_t *>(L.Wq + wq_offset(L, s, p, c0 + cc, k0)) = packed[p];
    }
  } else if (L.Wt_hi) {
    float *Wh = L.Wt_hi + (size_t)s * ld * rld, *Wl = L.Wt_lo + (size_t)s * ld * rld;
    for (int e = tid; e < kWCols * rld; e += blockDim.x) {
      const int cc = e / rld, k = e - cc * rld;
      const double w = (k < r) ? Y[k * kYS + cc] : 0.0;
      const float hi = to_tf32((float)w);
      const float lo = to_tf32((float)(w - (double)hi));
      Wh[(size_t)(c0 + cc) * rld + k] = hi;
      Wl[(size_t)(c0 + cc) * rld + k] = lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Σ −= Wᵀ·W on the fp64 pipe: 64x64 tiles on/above the diagonal, upper elements only
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_syrk_f64(Layout L) {
  pdl_trigger();
  pdl_wait();
  timeline_mark(L, 5);
  const int s = L.s0 + blockIdx.z;
  const SessionState &st = L.st[s];
  if (L.tile_counter && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) *L.tile_counter = 0;   // queue head of the next launch
  const int r = st.r;
  if (r == 0) return;
  if (L.Wq && !st.exact_update) {           // the int8 tensor-core SYRK handles this frame, except flagged slots
    syrk_exact_rows(L, s);
    return;
  }
  const int n = internal_dim(st.N);
  const int Tn = L.ld / 64;
  for (int tile = blockIdx.x; tile < Tn * Tn; tile += gridDim.x) {   // grid-stride over tiles: a cheap no-op launch
  const int ti = tile / Tn, tj = tile - ti * Tn;
  const int i0 = ti * 64, j0 = tj * 64;
  if (ti > tj || j0 >= n) continue;
  const int rld = L.rld, ld = L.ld;
  const double *W = L.W64 + (size_t)s * ld * rld;
  double *Sg = L.sigma + (size_t)s * ld * ld;
  __shared__ double As[16][65], Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < r; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int kk = e >> 6, row = e & 63;                   // measurement-row major panel: coalesced along the slots
      As[kk][row] = W[(size_t)(k0 + kk) * ld + i0 + row];    // rows past r are zero
      Bs[kk][row] = W[(size_t)(k0 + kk) * ld + j0 + row];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty * 4 + u]; b[u] = Bs[kk][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < n && j < n && i <= j) {
        Sg[(size_t)i * ld + j] -= acc[u][v];
      }
    }
  __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// augmentation (:311-364) + end-of-step bookkeeping
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_augment(Layout L, InputRef in_arg) {
  pdl_wait();
  pdl_trigger();                     // the next step's k_observation_front may be launched (it waits in pdl_wait())
  timeline_mark(L, 7);
  const InputRef in = resolve_input(in_arg);
  const int s = L.s0 + blockIdx.z;
  SessionState &st = L.st[s];
  const int t_idx = current_step(in);
  const int N = st.N;
  int N2 = st.N2;
  bool overflow = false;
  if (N + N2 > L.Ncap) { N2 = max(0, L.Ncap - N); overflow = true; }
  const int n = internal_dim(N);
  const int ld = L.ld;
  double *Sg = L.sigma + (size_t)s * ld * ld;
  double *mu = L.mu + (size_t)s * ld;
  const float *xy = in.obs_xy + (size_t)s * in.xy_ss + (size_t)t_idx * in.m_stride * 2;
  const int *nw = L.new_ids + (size_t)s * L.mcap;
  if (N2 > 0) {
    const double px = mu[0], py = mu[1], th = mu[2];   // post-update pose (:323-331)
    double sn, cs;
    sincos(th, &sn, &cs);
    // cross blocks: Σ[new rows][c] = G_p·Σ[0:3][c]  (:355-357), c over the old state
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) {
      const double s0 = Sg[c], s1 = Sg[sym_idx(1, c, ld)], s2 = Sg[sym_idx(2, c, ld)];
      for (int q = 0; q < N2; ++q) {
        const int i = nw[q];
        const double rx = (double)xy[2 * i], ry = (double)xy[2 * i + 1];   // :344-345
        const double g0 = -rx * sn - ry * cs, g1 = rx * cs - ry * sn;       // :347
        const double v0 = s0 + g0 * s2;     // [1 0 g0]·Σ[0:3][c]
        const double v1 = s1 + g1 * s2;     // [0 1 g1]·Σ[0:3][c]
        const int slot = n + 2 * q;               // c < n <= slot: the upper-triangle element is (c, slot)
        *reinterpret_cast<double2 *>(Sg + (size_t)c * ld + slot) = make_double2(v0, v1);
      }
    }
    if (blockIdx.x == 0) {
      // new-new blocks: G_p·Σ_xx·G_pᵀ + G_z·Qt·G_zᵀ (:354); G_z stacks the same rotation, so every
      // 2x2 block — off-diagonal ones too — receives R(θ)·Qt·R(θ)ᵀ (reference quirk kept)
      double P[9];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) P[a * 3 + b] = Sg[sym_idx(a, b, ld)];
      const double qo = L.q_obs;
      const double GQG[4] = {(cs * qo) * cs + (-sn * qo) * (-sn), (cs * qo) * sn + (-sn * qo) * cs,
                             (sn * qo) * cs + (cs * qo) * (-sn), (sn * qo) * sn + (cs * qo) * cs};
      const int R = 2 * N2;
      for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
        const int i = e / R, j = e - i * R;
        if (i > j) continue;                       // upper triangle only
        const int oi = nw[i >> 1], oj = nw[j >> 1];
        const double rxi = (double)xy[2 * oi], ryi = (double)xy[2 * oi + 1];
        const double rxj = (double)xy[2 * oj], ryj = (double)xy[2 * oj + 1];
        double gi[3], gj[3];
        if ((i & 1) == 0) { gi[0] = 1; gi[1] = 0; gi[2] = -rxi * sn - ryi * cs; } else { gi[0] = 0; gi[1] = 1; gi[2] = rxi * cs - ryi * sn; }
        if ((j & 1) == 0) { gj[0] = 1; gj[1] = 0; gj[2] = -rxj * sn - ryj * cs; } else { gj[0] = 0; gj[1] = 1; gj[2] = rxj * cs - ryj * sn; }
        double acc = 0;
        for (int b = 0; b < 3; ++b) {
          double t = 0;
          for (int a = 0; a < 3; ++a) t += gi[a] * P[a * 3 + b];
          acc += t * gj[b];
        }
        const double v = acc + GQG[(i & 1) * 2 + (j & 1)];
        Sg[(size_t)(n + i) * ld + n + j] = v;
      }
      // new means, rounded through float32 (:327-331, :341-342)
      for (int q = threadIdx.x; q < N2; q += blockDim.x) {
        const int i = nw[q];
        const double ox = (double)xy[2 * i], oy = (double)xy[2 * i + 1];
        mu[n + 2 * q] = (double)(float)(ox * cs - oy * sn + px);
        mu[n + 2 * q + 1] = (double)(float)(ox * sn + oy * cs + py);
      }
    }
  }
  // The last block of this session commits the new size and publishes the pose.  Steady state (no new reflectors): nothing
  // was written, so block 0 commits at once — no ticket, no fences (each a ~1 µs round trip on the step's critical path).
  bool commit;
  if (N2 > 0 || overflow) {
    __threadfence();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(&st.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    commit = last && threadIdx.x == 0;
    if (commit) __threadfence();
  } else {
    commit = blockIdx.x == 0 && threadIdx.x == 0;
  }
  if (commit) {
    st.N = N + N2;
    if (overflow) st.flags |= FLAG_LANDMARK_CAPACITY;
    st.ticket = 0;
    if (st.r > 0) {
      st.n_updates += 1;
      if (st.exact_update) st.n_exact_frames += 1;
      else st.n_exact_slots += min(st.exact_slots, kMaxExactSlots);
    }
    if (in.pose_out) {
      double *o = in.pose_out + (size_t)s * in.pose_ss + (size_t)t_idx * 3;
      o[0] = mu[0]; o[1] = mu[1]; o[2] = mu[2];
    }
  }
  // replay: the last block of the whole launch advances the step counter (every block has read it by now)
  if (in.step && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(L.step_ticket, 1u) == gridDim.x * gridDim.z - 1) {
      *L.step_ticket = 0;
      *const_cast<int *>(in.step) += 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// layout conversion at the C-ABI boundary (reference order ↔ internal slots)
// ---------------------------------------------------------------------------------------------
__device__ __host__ inline int ref_to_slot(int i) { return i < 3 ? i : i + 1; }

// out: n_ref x n_ref column-major (ld_out), Eigen::MatrixXd layout
__global__ void k_pack_sigma(Layout L, int s, double *out, int ld_out) {
  const int n_ref = 3 + 2 * L.st[s].N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n_ref || j >= n_ref) return;
  const double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  out[(size_t)j * ld_out + i] = Sg[sym_idx(ref_to_slot(i), ref_to_slot(j), L.ld)];
}
__global__ void k_pack_mu(Layout L, int s, double *out) {
  const int n_ref = 3 + 2 * L.st[s].N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ref) out[i] = L.mu[(size_t)s * L.ld + ref_to_slot(i)];
}
// rekf_get_state: header {n_ref, time, flags}, then mu (n_ref) and the full symmetric column-major sigma (n_ref x n_ref), all
// in one launch; mu / sigma are only written when the caller's idea of the dimension (n_expect) is right
__global__ void k_pack_state(Layout L, int s, int n_expect, double *hdr, double *out_mu, double *out_sigma) {
  const SessionState &st = L.st[s];
  const int n_ref = 3 + 2 * st.N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i == 0 && j == 0) { hdr[0] = (double)n_ref; hdr[1] = st.time; hdr[2] = (double)st.flags; }
  if (n_ref != n_expect || i >= n_ref || j >= n_ref) return;
  const double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  if (out_sigma) out_sigma[(size_t)j * n_ref + i] = Sg[sym_idx(ref_to_slot(i), ref_to_slot(j), L.ld)];
  if (j == 0) out_mu[i] = L.mu[(size_t)s * L.ld + ref_to_slot(i)];
}
// landmark means and diagonal 2x2 blocks (row-major), what the node reads for markers / saving
__global__ void k_pack_landmarks(Layout L, int s, double *xy, double *cov) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L.st[s].N) return;
  const int a = kPoseSlots + 2 * j;
  const double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  const double *mu = L.mu + (size_t)s * L.ld;
  xy[2 * j] = mu[a]; xy[2 * j + 1] = mu[a + 1];
  cov[4 * j + 0] = Sg[(size_t)a * L.ld + a];       cov[4 * j + 1] = Sg[(size_t)a * L.ld + a + 1];
  cov[4 * j + 2] = Sg[(size_t)a * L.ld + a + 1]; cov[4 * j + 3] = Sg[(size_t)(a + 1) * L.ld + a + 1];
}
// Node::ReflectorToRosMarkers (ros_node.cc:736-789) on the device: per landmark the mean and the 95 % covariance ellipse
// (chi-square 5.991, :763-764) of its 2x2 block.  The reference takes the eigen-decomposition from Eigen::EigenSolver,
// whose eigenvalue order and eigenvector sign are implementation details; here the symmetric 2x2 problem is solved in
// closed form with the major axis first, which describes the same ellipse.  out[5j..] = x, y, angle, x_len, y_len.
__global__ void k_pack_markers(Layout L, int s, double *out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L.st[s].N) return;
  const int a = kPoseSlots + 2 * j;
  const double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  const double *mu = L.mu + (size_t)s * L.ld;
  const double sxx = Sg[(size_t)a * L.ld + a], sxy = Sg[(size_t)a * L.ld + a + 1], syy = Sg[(size_t)(a + 1) * L.ld + a + 1];
  const double mean = 0.5 * (sxx + syy), diff = 0.5 * (sxx - syy);
  const double rad = hypot(diff, sxy);
  const double l1 = mean + rad, l2 = mean - rad;             // l1 >= l2
  out[5 * j + 0] = mu[a];
  out[5 * j + 1] = mu[a + 1];
  out[5 * j + 2] = 0.5 * atan2(2.0 * sxy, sxx - syy);        // direction of the l1 eigenvector (:762)
  out[5 * j + 3] = 2.0 * sqrt(l1 * 5.991);                   // :763
  out[5 * j + 4] = 2.0 * sqrt(l2 * 5.991);                   // :764
}
// in: mu (n_ref), sigma n_ref x n_ref column-major (ld_in).  Symmetrised on the way in: (Σ+Σᵀ)/2.
__global__ void k_unpack_state(Layout L, int s, const double *mu_in, const double *sig_in, int ld_in, int n_ref) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n_ref || j >= n_ref) return;
  double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  const double v = 0.5 * (sig_in[(size_t)j * ld_in + i] + sig_in[(size_t)i * ld_in + j]);
  if (i <= j) Sg[(size_t)ref_to_slot(i) * L.ld + ref_to_slot(j)] = v;
  if (j == 0) L.mu[(size_t)s * L.ld + ref_to_slot(i)] = mu_in[i];
}
// PredictState (:97-152) into a packed copy: out_mu (n_ref), out_sigma column-major or nullptr
__global__ void k_predict_state(Layout L, int s, double time, double *out_mu, double *out_sigma, int ld_out) {
  const SessionState &st = L.st[s];
  const int n_ref = 3 + 2 * st.N;
  const double *mu = L.mu + (size_t)s * L.ld;
  const double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
  const int ld = L.ld;
  const double vt[3] = {st.vt[0], st.vt[1], st.vt[2]};
  const MotionTerms t = motion_model(L, vt, mu[2], time - st.time);   // :100
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n_ref || j >= n_ref) return;
  if (j == 0) {
    double v = mu[ref_to_slot(i)];
    if (i < 2) v += t.d[i];
    if (i == 2) v = wrap_angle(v + t.d[2]);
    out_mu[i] = v;
  }
  if (!out_sigma) return;
  const int a = ref_to_slot(i), b = ref_to_slot(j);
  // (G Σ Gᵀ)[i][j] = Σ_ab G[i][a] Σ[a][b] G[j][b], G = I + g02·e0e2ᵀ + g12·e1e2ᵀ
  const double gi = i == 0 ? t.g02 : (i == 1 ? t.g12 : 0.0);
  const double gj = j == 0 ? t.g02 : (j == 1 ? t.g12 : 0.0);
  double v = Sg[sym_idx(a, b, ld)];
  if (i < 2) v += gi * Sg[sym_idx(2, b, ld)];
  if (j < 2) {
    double col2 = Sg[sym_idx(a, 2, ld)];
    if (i < 2) col2 += gi * Sg[(size_t)2 * ld + 2];
    v += gj * col2;
  }
  if (i < 3 && j < 3) v += t.V[i * 3 + j];
  out_sigma[(size_t)j * ld_out + i] = v;
}

}  // namespace rekf
