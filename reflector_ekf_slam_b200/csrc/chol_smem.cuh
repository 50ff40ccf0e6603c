// chol_smem.cuh — S = L·Lᵀ with the whole lower triangle resident in shared memory (r <= ~204).
//
// Replaces `(H·Σ·Hᵀ + Q).inverse()` (reflector_ekf_slam.cc:305): S is SPD, so it is factored instead of
// inverted, and ν rides along as an extra row that comes out as L⁻¹ν (the mean update then is Wᵀ·L⁻¹ν).
// This kernel is the serial spine of the step — r pivots, each a dependent rsqrt — so the design goal is
// latency, not throughput:
//   * right-looking, 32-column panels, everything in shared memory (packed block columns, ~215 KB at r=200),
//     filled with cp.async so the global-memory latency is paid once;
//   * phase A: the 32x32 diagonal block is factored by ONE warp with the block in registers (lane = row).
//     Per column the next pivot is formed from registers and shuffled out before the column broadcast, so
//     the rsqrt chain (the critical path) overlaps the rank-1 update; no CTA barrier inside the block.
//     Meanwhile the other seven warps invert the PREVIOUS diagonal block (X = L_bb⁻¹, five interleaved
//     substitution chains per warp) — the TRSM kernel consumes these inverses;
//   * phase B: every row below the block (and the ν row) is owned by one thread, held in registers, and
//     solved against the diagonal block by column-oriented substitution (31−j independent FMAs per step);
//   * phase C: trailing rank-32 update on the fp64 tensor pipe (mma.sync m8n8k4 → DMMA; full rate on
//     B200): a warp task is one 8-row tile against all column tiles left of it, A fragments in registers;
//   * L goes back to global memory once, at the end.
// 256 threads (the per-thread 32-double rows stay in registers), one CTA per session.
// Larger r (config C4) falls back to k_cholesky (global-memory panels).
#pragma once
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"

namespace rekf {

constexpr int kCholSmemThreads = 256;
constexpr int kPS2 = kCholNb + 4;    // pitch 36: DMMA fragment loads (8 rows x 4 k) are bank-conflict free

// packed block-column storage: block column b holds rows 32b..R1-1 (R1 = r+1 incl. the ν row), pitch kPS2
__host__ __device__ inline int chol_panel_off(int R1, int b) { return (b * R1 - 16 * b * (b - 1)) * kPS2; }
inline size_t smem_chol_resident(int rcap) {
  const int R1 = rcap + 1, nb = (rcap + kCholNb - 1) / kCholNb;
  return sizeof(double) * ((size_t)chol_panel_off(R1, nb) + 2 * 32 + 32);
}

// D(8x8) = A(8x4)·B(4x8) + C on the fp64 tensor pipe.  Fragments (lane = 4·g + t): a = A[g][t], b = B[t][g],
// c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// X = D⁻¹ for the jb x jb lower-triangular block at P (pitch kPS2), written to Dg[32][32] (identity past jb).
// Called by `nw` warps (index wi); each interleaves its columns c = wi, wi+nw, ... as independent chains.
__device__ inline void invert_diag_block(const double *P, int jb, double *Dg, int wi, int nw, int lane) {
  constexpr int kMaxCols = 5;
  const double dinv = (lane < jb) ? 1.0 / P[lane * kPS2 + lane] : 1.0;
  double t[kMaxCols], x[kMaxCols];
  int col[kMaxCols];
#pragma unroll
  for (int u = 0; u < kMaxCols; ++u) {
    col[u] = wi + u * nw;
    t[u] = (lane == col[u]) ? 1.0 : 0.0;
    x[u] = 0.0;
  }
  for (int k = 0; k < kCholNb; ++k) {
    const double lk = (lane > k && lane < jb && k < jb) ? P[min(lane, jb - 1) * kPS2 + k] : 0.0;
#pragma unroll
    for (int u = 0; u < kMaxCols; ++u) {
      const double xk = __shfl_sync(0xffffffffu, t[u] * dinv, k);
      if (lane == k) x[u] = xk;
      t[u] = fma(-lk, xk, t[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < kMaxCols; ++u)
    if (col[u] < kCholNb) Dg[lane * kCholNb + col[u]] = x[u];
}

__global__ void __launch_bounds__(kCholSmemThreads, 1) k_cholesky_smem(Layout L) {
  extern __shared__ double sm_d[];
  const int s = L.s0 + blockIdx.x;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int R1 = r + 1;
  const int nblk = (r + kCholNb - 1) / kCholNb;
  const int sld = L.sld;
  double *Sb = L.Sbuf + (size_t)s * L.rld * sld;
  double *Dinv = L.Dinv + (size_t)s * (L.rld / kCholNb) * kCholNb * kCholNb;
  double *A = sm_d;                                         // packed block columns
  double *cb = sm_d + chol_panel_off(R1, nblk);             // [2][32] column broadcast buffer of the panel warp
  double *invd = cb + 64;                                   // [32] reciprocals of the current diagonal
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT / 32;
  bool bad = false;
#ifdef REKF_CHOL_TIMING
  double *tlog = L.Qd + (size_t)s * L.rcap;                 // per-phase cycle stamps (thread 0)
  int tl = 0;
#define REKF_TSTAMP() do { if (tid == 0) tlog[tl++] = (double)clock64(); } while (0)
#else
#define REKF_TSTAMP() do { } while (0)
#endif
  REKF_TSTAMP();

  // ---- load the lower triangle (+ ν row) with cp.async: warp per column, lanes over rows -------------------
  for (int c = warp; c < r; c += NW) {
    const int b = c >> 5, jj = c & 31, base = b << 5;
    double *P = A + chol_panel_off(R1, b) + jj;
    const double *src = Sb + (size_t)c * sld;
    for (int i = base + lane; i < R1; i += 32)
      if (i >= c)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(P + (i - base) * kPS2)), "l"(src + i) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  REKF_TSTAMP();

  for (int b = 0; b < nblk; ++b) {
    const int J = b * kCholNb, jb = min(kCholNb, r - J), rows = R1 - J;
    double *P = A + chol_panel_off(R1, b);

    if (warp == 0) {
      // ---- phase A: one warp factors the jb x jb diagonal block held in registers ----------------------
      double a[kCholNb];
#pragma unroll
      for (int jj = 0; jj < kCholNb; ++jj)
        a[jj] = (lane < jb && jj < jb) ? ((jj <= lane) ? P[min(lane, jb - 1) * kPS2 + jj] : 0.0) : ((jj == lane) ? 1.0 : 0.0);
      double d = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
      for (int j = 0; j < kCholNb; ++j) {
        if (j < jb && !(d > 0.0)) bad = true;
        const double inv = rsqrt(d);
        const double l = (lane == j) ? d * inv : a[j] * inv;
        if (lane == j) invd[j] = inv;
        if (j + 1 < kCholNb) d = __shfl_sync(0xffffffffu, fma(-l, l, a[j + 1]), j + 1);   // next pivot, early
        cb[(j & 1) * 32 + lane] = l;
        __syncwarp();
#pragma unroll
        for (int jj = j + 1; jj < kCholNb; ++jj) a[jj] = fma(-l, cb[(j & 1) * 32 + jj], a[jj]);
        a[j] = l;
      }
      if (lane < jb) {
#pragma unroll
        for (int jj = 0; jj < kCholNb; ++jj)
          if (jj <= lane) P[lane * kPS2 + jj] = a[jj];
      }
    } else if (b > 0) {
      // ---- meanwhile: inverse of the previous (full) diagonal block, for the TRSM kernel --------------------
      invert_diag_block(A + chol_panel_off(R1, b - 1), kCholNb, Dinv + (size_t)(b - 1) * kCholNb * kCholNb, warp - 1, NW - 1, lane);
    }
    __syncthreads();
    REKF_TSTAMP();

    // ---- phase B: rows below the block (incl. ν), one thread per row, substitution in registers ----------
    if (jb == kCholNb) {
      for (int ii = kCholNb + tid; ii < rows; ii += NT) {
        double a[kCholNb];
        double *row = P + ii * kPS2;
#pragma unroll
        for (int jj = 0; jj < kCholNb; jj += 2) {
          const double2 t = *reinterpret_cast<const double2 *>(row + jj);
          a[jj] = t.x; a[jj + 1] = t.y;
        }
#pragma unroll
        for (int j = 0; j < kCholNb; ++j) {
          const double l = a[j] * invd[j];
          const double *lj = P + j;                          // column j of the diagonal block: P[jj][j]
#pragma unroll
          for (int jj = j + 1; jj < kCholNb; ++jj) a[jj] = fma(-l, lj[jj * kPS2], a[jj]);
          a[j] = l;
        }
#pragma unroll
        for (int jj = 0; jj < kCholNb; jj += 2) *reinterpret_cast<double2 *>(row + jj) = make_double2(a[jj], a[jj + 1]);
      }
    } else {                                                 // last, partial block: only the ν row is below it
      for (int ii = jb + tid; ii < rows; ii += NT) {
        double *row = P + ii * kPS2;
        for (int j = 0; j < jb; ++j) {
          const double l = row[j] * invd[j];
          for (int jj = j + 1; jj < jb; ++jj) row[jj] = fma(-l, P[jj * kPS2 + j], row[jj]);
          row[j] = l;
        }
      }
    }
    __syncthreads();
    REKF_TSTAMP();

    // ---- phase C: trailing update A[i][c] −= Σ_k P[i][k]·P[c][k], J+32 <= c <= i, on the fp64 tensor pipe ---------
    const int T = R1 - (J + kCholNb);                       // remaining rows (incl. ν); <= 0 on the last block
    const int Tc = r - (J + kCholNb);                       // remaining columns
    if (T > 0 && Tc > 0) {
      const int nrt = (T + 7) / 8, nct = (Tc + 7) / 8;
      const int g = lane >> 2, t4 = lane & 3;
      for (int q = 0;; ++q) {                               // snake order balances the triangular task sizes
        const int idx = q * NW + ((q & 1) ? NW - 1 - warp : warp);
        if (q * NW >= nrt) break;
        if (idx >= nrt) continue;
        const int rt = nrt - 1 - idx;
        const int irow = kCholNb + 8 * rt + g;              // panel-relative row held by this lane's A fragment
        const double *ap = P + min(irow, rows - 1) * kPS2 + t4;
        double af[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) af[ks] = -ap[4 * ks];
        const int gi = J + irow;
        const int ctmax = min(rt, nct - 1);
        for (int ct = 0; ct <= ctmax; ++ct) {
          const int crow = kCholNb + 8 * ct + g;            // B fragment: column index n = g → row crow of P
          const double *bp = P + min(crow, rows - 1) * kPS2 + t4;
          const int gc = J + kCholNb + 8 * ct + 2 * t4;     // C fragment columns gc, gc+1 of row gi
          const bool ok = (gi < R1) && (gc < r);
          const int cbk = gc >> 5;
          double *cp = A + chol_panel_off(R1, min(cbk, nblk - 1)) + (gi - cbk * kCholNb) * kPS2 + (gc & 31);
          double c0 = 0.0, c1 = 0.0;
          if (ok) { const double2 cv = *reinterpret_cast<const double2 *>(cp); c0 = cv.x; c1 = cv.y; }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) dmma884(c0, c1, af[ks], bp[4 * ks], c0, c1);
          if (ok) {
            if (gc + 1 < r) *reinterpret_cast<double2 *>(cp) = make_double2(c0, c1);
            else cp[0] = c0;
          }
        }
      }
    }
    __syncthreads();
    REKF_TSTAMP();
  }
  // inverse of the last diagonal block (identity-padded past jb), all warps
  {
    const int b = nblk - 1, jb = r - b * kCholNb;
    invert_diag_block(A + chol_panel_off(R1, b), jb, Dinv + (size_t)b * kCholNb * kCholNb, warp, NW, lane);
  }

  // ---- publish L (k_solve_w3 reads it from global / L2): warp per column, lanes over rows -------------------
  for (int c = warp; c < r; c += NW) {
    const int b = c >> 5, jj = c & 31, base = b << 5;
    const double *P = A + chol_panel_off(R1, b) + jj;
    double *dst = Sb + (size_t)c * sld;
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + lane + 32 * u;
      v[u] = (i < R1) ? P[(min(i, R1 - 1) - base) * kPS2] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + lane + 32 * u;
      if (i >= c && i < R1) dst[i] = v[u];
    }
  }
  REKF_TSTAMP();
  if (bad && lane == 0) atomicOr(&st.flags, FLAG_NOT_SPD);
}

}  // namespace rekf
