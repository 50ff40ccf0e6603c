// rekf_device.cuh — device-side data layout of the B200 EKF engine.
//
// One handle = a batch of S independent sessions (filters); every kernel is launched once for the
// whole batch (blockIdx.z or blockIdx.x = session).  All control state that the reference keeps in
// host members (state_.time, vt_, state_.mu.rows(), the ReflectorMatchResult) lives in SessionState
// in HBM, so a whole HandleObservationMessage (reflector_ekf_slam.cc:229-368) is a fixed chain of
// launches with capacity-sized grids and no host round trip; kernels read the live sizes (N, r, N2)
// from SessionState and exit early.
//
// Internal state ordering (differs from the reference's [x y θ l0x l0y …] to keep every landmark
// pair 16-byte aligned): slot 0,1,2 = x,y,θ; slot 3 = zero padding (row and column of Σ stay 0
// under every kernel); landmark j occupies slots 4+2j, 5+2j.  Σ is fp64 with row pitch `ld` (a multiple
// of 128 so tcgen05 tiles never straddle the buffer edge) and ONLY ITS UPPER TRIANGLE (row <= column) IS
// MAINTAINED: every kernel reads Σ[i][j] through sym_idx() and writes the upper element only, so the
// covariance downdate — the HBM-bound part of the step — moves half the bytes and needs no mirrored
// (transposed) stores; what lies below the diagonal is stale and never read.  The reference ordering and
// the full symmetric column-major matrix are restored at the C-ABI boundary (k_pack_sigma / k_pack_mu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rekf {

constexpr int kPoseSlots = 4;
constexpr int kSigmaTile = 128;   // Σ pitch granularity = tcgen05 output tile edge
constexpr int kKBlock = 32;       // measurement-row (GEMM K) granularity: one 128-byte swizzle row of tf32
constexpr int kCholNb = 32;       // Cholesky / TRSM block size
constexpr int kWCols = 16;        // columns of W = L⁻¹·H·Σ solved per CTA
// int8 SYRK admission: (row scale)² / posterior variance above this sends the frame to the fp64 SYRK
constexpr double kMaxSliceGain2 = 1.0;
constexpr int kMaxExactSlots = 64;   // more flagged slots than this: the whole frame goes to the fp64 SYRK

enum : int {
  FLAG_LANDMARK_CAPACITY = 1,     // augmentation would exceed max_landmarks: extra reflectors dropped
  FLAG_NOT_SPD = 2,               // a pivot of S = HΣHᵀ+Q was not positive
  FLAG_OBS_CAPACITY = 4,          // more observations in a frame than max_observations
  FLAG_TCGEN05_TIMEOUT = 8,       // an mbarrier wait in the tensor-core SYRK gave up (should never happen)
  FLAG_SYNC_TIMEOUT = 16          // k_solve_ll gave up waiting for a flag of the Cholesky / gather kernels (should never happen)
};

struct SessionState {
  double time;          // state_.time
  double vt[3];         // vt_ (reflector_ekf_slam.h:52-53)
  double gps_innov_pad; // unused, keeps 8-byte fields together
  int N;                // landmarks in the state: state_.mu.rows() == 3 + 2N
  int m;                // observations in the current frame
  int M;                // result.state_obs_match_ids.size()
  int Mmap;             // result.map_obs_match_ids.size()
  int N2;               // result.new_ids.size()
  int r;                // measurement rows of this frame: 2(M+Mmap) (+3 with a GPS pose), 0 = no update
  int flags;            // sticky FLAG_* bits
  unsigned ticket;      // last-block-done counter (augment kernel)
  int exact_update;     // this frame's downdate cancels too deeply for the int8 slices: use the fp64 SYRK
  int exact_slots;      // slots whose rows/columns of the downdate are done in fp64 this frame (k_syrk_exact_rows)
  // cumulative since creation / rekf_set_state (rekf_get_counters): how often the int8 covariance update left the tensor path
  long long n_updates;        // frames that carried an update (r > 0)
  long long n_exact_frames;   // of those: whole frame on the fp64 SYRK (exact_update)
  long long n_exact_slots;    // flagged slots handled by k_syrk_exact_rows, summed over the other frames
};

// Where the current message comes from: the handle's device mailbox (host path) or device-resident
// replay arrays indexed by a device-side step counter (no host traffic; CUDA-graph friendly).
struct InputRef {
  const double *odom;       // session s, step t: odom + s*odom_ss + t*4   -> (time, vx, vy, wz)
  const double *obs_time;   // obs_time + s*time_ss + t
  const float *obs_xy;      // obs_xy + s*xy_ss + t*m_stride*2
  const int *obs_count;     // per-session counts (host path) or nullptr -> m_fixed
  const double *gps;        // per session 4 doubles (flag, x, y, yaw) or nullptr
  long long odom_ss, time_ss, xy_ss;
  int m_stride, m_fixed;
  const int *step;          // device step counter, nullptr -> 0
  int fuse_odom;            // replay: k_observation_front handles the step's odometry message first (no k_odometry launch)
  double *pose_out;         // optional: pose after the step, pose_out + s*pose_ss + t*3
  long long pose_ss;
  const InputRef *indirect; // non-null: the real descriptor lives at this DEVICE address (replay graphs are captured once
                            // with only this pointer baked in; rekf_replay_device rewrites the descriptor, not the graph)
};
__device__ __forceinline__ InputRef resolve_input(const InputRef &in) { return in.indirect ? *in.indirect : in; }

struct Layout {
  int S;          // sessions of the handle (extent of every [S]-leading array and of the TMA maps)
  int s0, Sg;     // the sessions this launch covers: s0 .. s0+Sg-1 (a pipeline group; the whole batch when Sg == S)
  int Ncap;       // landmark capacity
  int ncap;       // 4 + 2*Ncap
  int ld;         // Σ / μ pitch, multiple of 128
  int mcap;       // observation capacity per frame
  int rcap;       // 2*mcap + 4 (3 GPS rows + pad)
  int rld;        // round_up(rcap, 32): pitch of the Wᵀ panels (GEMM K extent)
  int sld;        // pitch of the S / L buffer (rld + 8: one extra row carries ν)
  int mapcap;
  int odom_model;
  double q_lin, q_ang, q_obs;   // linear_velocity_cov, angular_velocity_cov, observation_cov
  SessionState *st;
  double *mu;     // [S][ld]
  double *sigma;  // [S][ld][ld]
  // pre-loaded beacon map (sensor::Map), shared by all sessions
  float *map_xy;  // [mapcap][2]
  double *map_cov;// [mapcap][4] row-major 2x2
  int *map_count; // device scalar
  // per-frame scratch
  int *state_pairs, *map_pairs, *new_ids;   // [S][mcap*2], [S][mcap*2], [S][mcap]
  double *Hp;     // [S][rcap][4]  pose coefficients of each measurement row (4th unused)
  double *Hl;     // [S][rcap][2]  landmark coefficients
  int *Hslot;     // [S][rcap]     internal slot of the row's landmark, -1 = none
  double *innov;  // [S][rcap]
  double *Qd;     // [S][rcap]     diagonal of Q
  double *Sbuf;   // [S][rld][sld] column-major lower triangle of S, then L; row r carries ν → L⁻¹ν
  double *Dinv;   // [S][rld/32][32][32] inverses of L's diagonal blocks
  float *Wt_hi, *Wt_lo;  // [S][ld][rld] tf32 hi/lo split of Wᵀ (K-major operands of the tcgen05 SYRK)
  double *W64;    // [S][rld][ld] fp64 W, measurement-row major: row k holds W[k][all slots] (fp64 SYRK and the exact rows; nullptr in tf32 mode)
  double *Ybuf;   // [S][rld][ld] Y = H·Σ, written by k_gather_y beside the Cholesky, read by k_solve_w3 as coalesced tiles
                  // (k_solve_ll: scratch for the thread-private accumulated updates of the blocks below the current one)
  int8_t *Wq;     // [S][4][kq/64][ld][64] signed 7-bit digit slices of the row-scaled Wᵀ (REKF_COV_TCGEN05_I8X4), K in
                  // 64-byte chunks OUTSIDE the row index: a TMA box of 128 rows x 64 K-bytes is one contiguous 8 KB block
  int *Wexp;      // [S][ld] per-row power-of-two exponent e_c of Wq
  double *Wscale; // [S][ld] 2^e_c as a double (what the SYRK epilogue multiplies by)
  int kq;         // round_up(rcap, 64): K extent of Wq in bytes
  double *Wdiag;  // [S][ld] exact fp64 diagonal of Wᵀ·W (tensor-core modes use it for Σ[i][i])
  unsigned char *Wflag;  // [S][ld] 1: this slot's row and column of the downdate are computed in fp64
  int *exact_list;       // [S][kMaxExactSlots] the flagged slots of this frame
  int *step;      // device step counter for replay (one per pipeline group)
  int *tile_counter;  // work-queue head of the persistent SYRK (one per pipeline group; reset by k_syrk_f64)
  unsigned *step_ticket;  // blocks of k_augment that have finished (one per pipeline group): the last one advances `step`
  // cross-kernel pacing of the shadowed TRSM (solve_ll.cuh), [S][sync_n] ints cleared by k_observation_front every frame:
  // [b] block column b of L (with X_b and ν_b) is in global memory
  int *sync;
  int sync_n;
  int shadow;     // k_cholesky_smem → k_solve_ll is a programmatic-launch edge and the Cholesky triggers at its start: side by side
  unsigned long long *tlog;  // optional kernel-start timeline (REKF_TIMELINE=1): [0] = entries used, then (globaltimer ns << 12 | kernel id << 8 | first session)
};

constexpr int kTimelineCap = 1 << 16;
// one entry per kernel launch: stamped by the first thread of the first block when the kernel starts running
__device__ __forceinline__ void timeline_mark(const Layout &L, int kernel_id) {
  if (L.tlog && threadIdx.x == 0 && threadIdx.y == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long k = atomicAdd(L.tlog, 1ULL);
    if (k + 1 < (unsigned long long)kTimelineCap) L.tlog[k + 1] = (t << 12) | ((unsigned long long)kernel_id << 8) | (unsigned long long)(L.s0 & 255);
  }
}
// Programmatic dependent launch (the kernels of a step form a chain in one stream / graph branch): pdl_wait() blocks until the
// preceding kernel of the chain has completed and flushed; pdl_trigger() lets the NEXT kernel of the chain be launched (its
// blocks then sit in pdl_wait()), so that its launch latency and prologue overlap this kernel's tail.  Both are no-ops for a
// kernel launched without the attribute.  Triggers sit a few microseconds before a kernel's end, not at its start: blocks of
// a dependent that wait for tens of microseconds would hold the SMs the other pipeline group needs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// bounded wait for a flag another kernel raises (thread 0 of a CTA)
__device__ __forceinline__ bool sync_wait(const int *flag) {
  for (int it = 0; it < (1 << 21); ++it) {
    int v;
    asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v) return true;
    __nanosleep(200);
  }
  return false;
}
// raise a flag after this thread's (and, behind a barrier, its CTA's) global writes
__device__ __forceinline__ void sync_raise(int *flag) {
  __threadfence();
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}

__host__ __device__ inline int round_up(int v, int g) { return (v + g - 1) / g * g; }
// offset of Σ[i][j] in the upper-triangle-only storage
__host__ __device__ __forceinline__ size_t sym_idx(int i, int j, int ld) {
  return i <= j ? (size_t)i * ld + j : (size_t)j * ld + i;
}
// byte offset of digit slice p, state slot `row`, measurement byte `kbyte` in Layout::Wq
__host__ __device__ inline size_t wq_offset(const Layout &L, int s, int p, int row, int kbyte) {
  return ((((size_t)s * 4 + p) * (L.kq >> 6) + (kbyte >> 6)) * L.ld + row) * 64 + (kbyte & 63);
}
__host__ __device__ inline int internal_dim(int N) { return kPoseSlots + 2 * N; }

}  // namespace rekf
