// chol_smem.cuh — S = L·Lᵀ with the whole lower triangle resident in shared memory (r <= ~204).
//
// Replaces `(H·Σ·Hᵀ + Q).inverse()` (reflector_ekf_slam.cc:305): S is SPD, so it is factored instead of
// inverted, and ν rides along as an extra row that comes out as L⁻¹ν (the mean update then is Wᵀ·L⁻¹ν).
// This kernel is the serial spine of the step — r pivots, each a dependent rsqrt — so the design goal is
// latency, not throughput:
//   * right-looking, 32-column panels, everything in shared memory (packed block columns, ~190 KB at r=200);
//   * the 32x32 diagonal block is factored by ONE warp with the block in registers (lane = row): per
//     column one shuffle (pivot), one rsqrt, one shared-memory column broadcast, no CTA barrier;
//   * the diagonal block's inverse X (needed anyway by the TRSM in k_solve_w) turns the panel below the
//     block into independent dot products (no substitution chain);
//   * the trailing rank-32 update is register-tiled 4x4 over the remaining triangle by all 16 warps.
// One CTA per session.  Larger r (config C4) falls back to k_cholesky (global-memory panels).
#pragma once
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"

namespace rekf {

constexpr int kCholSmemThreads = 512;

// packed block-column storage: block column b holds rows 32b..R1-1 (R1 = r+1 incl. the ν row), pitch kPS
__host__ __device__ inline int chol_panel_off(int R1, int b) { return (b * R1 - 16 * b * (b - 1)) * kPS; }
inline size_t smem_chol_resident(int rcap) {
  const int R1 = rcap + 1, nb = (rcap + kCholNb - 1) / kCholNb;
  return sizeof(double) * ((size_t)chol_panel_off(R1, nb) + (size_t)kCholNb * kPS + 2 * 32);
}

__global__ void __launch_bounds__(kCholSmemThreads, 1) k_cholesky_smem(Layout L) {
  extern __shared__ double sm_d[];
  const int s = blockIdx.x;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int R1 = r + 1;
  const int nblk = (r + kCholNb - 1) / kCholNb;
  const int sld = L.sld;
  double *Sb = L.Sbuf + (size_t)s * L.rld * sld;
  double *A = sm_d;                                         // packed block columns
  double *X = sm_d + chol_panel_off(R1, nblk);              // [32][kPS] inverse of the current diagonal block
  double *cb = X + kCholNb * kPS;                           // [2][32] column broadcast buffer of the panel warp
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
  bool bad = false;

  // ---- load the lower triangle (+ ν row) ---------------------------------------------------------
  for (int b = 0; b < nblk; ++b) {
    const int J = b * kCholNb, jb = min(kCholNb, r - J), rows = R1 - J;
    double *P = A + chol_panel_off(R1, b);
    for (int e = tid; e < rows * jb; e += NT) {
      const int jj = e / rows, ii = e - jj * rows;
      P[ii * kPS + jj] = (ii >= jj) ? Sb[(size_t)(J + jj) * sld + J + ii] : 0.0;
    }
  }
  __syncthreads();

  for (int b = 0; b < nblk; ++b) {
    const int J = b * kCholNb, jb = min(kCholNb, r - J), rows = R1 - J;
    double *P = A + chol_panel_off(R1, b);

    // ---- phase A: one warp factors the jb x jb diagonal block held in registers ----------------------
    if (warp == 0) {
      double a[kCholNb];
#pragma unroll
      for (int jj = 0; jj < kCholNb; ++jj)
        a[jj] = (lane < jb && jj < jb) ? ((jj <= lane) ? P[lane * kPS + jj] : 0.0) : ((jj == lane) ? 1.0 : 0.0);
#pragma unroll
      for (int j = 0; j < kCholNb; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (j < jb && !(d > 0.0)) bad = true;
        const double inv = rsqrt(d);
        const double l = (lane == j) ? d * inv : a[j] * inv;
        cb[(j & 1) * 32 + lane] = l;
        __syncwarp();
#pragma unroll
        for (int jj = j + 1; jj < kCholNb; ++jj) a[jj] = fma(-l, cb[(j & 1) * 32 + jj], a[jj]);
        a[j] = l;
      }
      if (lane < jb) {
#pragma unroll
        for (int jj = 0; jj < kCholNb; ++jj)
          if (jj <= lane) P[lane * kPS + jj] = a[jj];
      }
    }
    __syncthreads();

    // ---- X = D⁻¹: warp c handles columns c, c+16; lane = row; column-oriented substitution -------------
    for (int c = warp; c < kCholNb; c += NT / 32) {
      if (c < jb) {
        const double invd = (lane < jb) ? 1.0 / P[lane * kPS + lane] : 1.0;
        double t = (lane == c) ? 1.0 : 0.0, x = 0.0;
        for (int k = c; k < jb; ++k) {
          const double xk = __shfl_sync(0xffffffffu, t * invd, k);
          if (lane == k) x = xk;
          if (lane > k && lane < jb) t = fma(-P[lane * kPS + k], xk, t);
        }
        X[lane * kPS + c] = (lane >= c && lane < jb) ? x : 0.0;
      } else {
        X[lane * kPS + c] = 0.0;
      }
    }
    __syncthreads();

    // ---- rows below the block (incl. ν): P[i][:] ← P[i][:]·D⁻ᵀ, two threads per row, 16 columns each ------
    {
      const int below = rows - jb;
      for (int base = 0; base < below * 2; base += NT) {
        const int task = base + tid;
        const bool act = task < below * 2;
        const int ii = jb + (task >> 1), jg = task & 1;
        double out[16];
        if (act) {
          const double *row = P + ii * kPS;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int jj = jg * 16 + u;
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
          for (int u = 0; u < 16; ++u) {
            const int jj = jg * 16 + u;
            if (jj < jb) P[ii * kPS + jj] = out[u];
          }
        }
        __syncthreads();
      }
    }

    // ---- publish this block column: L to global (k_solve_w reads it) and the block inverse --------------
    for (int e = tid; e < rows * jb; e += NT) {
      const int jj = e / rows, ii = e - jj * rows;
      if (ii >= jj) Sb[(size_t)(J + jj) * sld + J + ii] = P[ii * kPS + jj];
    }
    double *Dg = L.Dinv + ((size_t)s * (L.rld / kCholNb) + b) * kCholNb * kCholNb;
    for (int e = tid; e < kCholNb * kCholNb; e += NT) {
      const int i = e >> 5, k = e & 31;
      Dg[e] = (i < jb && k < jb) ? X[i * kPS + k] : 0.0;
    }

    // ---- trailing update: A[i][c] −= Σ_k P[i][k]·P[c][k] for J+32 <= c <= i, 4x4 register tiles ----------
    const int T = R1 - (J + kCholNb);                       // remaining rows (incl. ν); <= 0 on the last block
    if (T > 0) {
      const int nt = (T + 3) / 4;
      const int ntiles = nt * (nt + 1) / 2;
      for (int tile = tid; tile < ntiles; tile += NT) {
        int ti = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
        while (ti * (ti + 1) / 2 > tile) --ti;
        const int tc = tile - ti * (ti + 1) / 2;            // tc <= ti
        const int i0 = kCholNb + 4 * ti, c0 = kCholNb + 4 * tc;   // panel-relative rows
        double acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
        const double *pi[4], *pc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pi[u] = P + min(i0 + u, rows - 1) * kPS;
          pc[u] = P + min(c0 + u, rows - 1) * kPS;
        }
#pragma unroll 4
        for (int k = 0; k < kCholNb; ++k) {
          double av[4], bv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { av[u] = pi[u][k]; bv[u] = pc[u][k]; }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int gi = J + i0 + u, gc = J + c0 + v;     // global row / column
            if (gi < R1 && gc < r && gc <= gi) {
              const int cbk = gc >> 5;
              A[chol_panel_off(R1, cbk) + (gi - cbk * kCholNb) * kPS + (gc & 31)] -= acc[u][v];
            }
          }
      }
    }
    __syncthreads();
  }
  if (bad && lane == 0) atomicOr(&st.flags, FLAG_NOT_SPD);
}

}  // namespace rekf
