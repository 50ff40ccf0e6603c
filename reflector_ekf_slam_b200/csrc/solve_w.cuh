// solve_w.cuh — W = L⁻¹·(H·Σ), μ += Wᵀ·(L⁻¹ν), and the operand panels of the covariance SYRK.
//
// Reference: K_t = Σ·Hᵀ·S⁻¹ and mu += K_t·(z − ẑ) (reflector_ekf_slam.cc:305-307).  With S = L·Lᵀ the gain
// never has to be formed: W = L⁻¹·H·Σ gives K·ν = Wᵀ·(L⁻¹ν) and K·H·Σ = Wᵀ·W.
//
// One CTA = 32 columns of W (32 state slots), 256 threads = 8 warps, lane = column:
//   gather   Y[q][c] = Σ_b H[q][b]·Σ[b][c]: H has <= 5 non-zeros per row (:272-275), so each entry is a
//            16-byte read from row c of the (symmetric) Σ — H·Σ never touches HBM as a matrix;
//   solve    blocked forward substitution, right-looking, 32-row blocks: W_J = D_J⁻¹·Y_J with the block
//            inverses from the Cholesky kernel (independent dot products instead of a substitution chain),
//            then Y_I −= L_IJ·W_J for the rows below.  Each thread keeps its column of W_J in registers;
//            L is staged through shared memory in 64-row chunks and read as broadcast LDS.128, which is
//            what lets the fp64 pipe rather than shared-memory bandwidth set the pace;
//   epilogue μ update (θ wrapped, :307), exact diag(WᵀW), int8 digit slices / tf32 hi-lo / fp64 panels.
#pragma once
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"

namespace rekf {

constexpr int kW2Cols = 32;
constexpr int kW2YS = kW2Cols + 1;       // padded pitch of the Y/W block
constexpr int kW2Chunk = 64;             // rows of L staged per pass

inline size_t smem_solve_w2(int rld) {
  return sizeof(double) * ((size_t)rld * kW2YS + (size_t)kW2Chunk * 32 + 32 * 32) + 64 * sizeof(int);
}

__global__ void __launch_bounds__(256, 2) k_solve_w2(Layout L) {
  extern __shared__ double sm_d[];
  const int s = blockIdx.z;
  const SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int n = internal_dim(st.N);
  const int c0 = blockIdx.x * kW2Cols;
  if (c0 >= round_up(n, kSigmaTile)) return;
  const int ld = L.ld, sld = L.sld, rld = L.rld;
  const double *Sg = L.sigma + (size_t)s * ld * ld;
  const double *Sb = L.Sbuf + (size_t)s * rld * sld;
  const double *Hp = L.Hp + (size_t)s * L.rcap * 4;
  const double *Hl = L.Hl + (size_t)s * L.rcap * 2;
  const int *Hslot = L.Hslot + (size_t)s * L.rcap;
  double *Y = sm_d;                                   // [rld][kW2YS]
  double *Lp = Y + (size_t)rld * kW2YS;               // [64][32] swizzled chunk of an L panel
  double *Xs = Lp + kW2Chunk * 32;                    // [32][32] inverse of the current diagonal block
  int *sexp = reinterpret_cast<int *>(Xs + 32 * 32);  // [32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r32 = round_up(r, 32);

  // ---- gather ----------------------------------------------------------------------------------------
  for (int cc = warp; cc < kW2Cols; cc += 8) {
    const int c = c0 + cc;
    if (c < n) {
      const double *rowc = Sg + (size_t)c * ld;
      const double p0 = rowc[0], p1 = rowc[1], p2 = rowc[2];
      for (int q = lane; q < r32; q += 32) {
        double y = 0.0;
        if (q < r) {
          const double *h = Hp + 4 * q;
          y = h[0] * p0 + h[1] * p1 + h[2] * p2;
          const int slot = Hslot[q];
          if (slot >= 0) {
            const double2 v = *reinterpret_cast<const double2 *>(rowc + slot);
            y += Hl[2 * q] * v.x + Hl[2 * q + 1] * v.y;
          }
        }
        Y[q * kW2YS + cc] = y;
      }
    } else {
      for (int q = lane; q < r32; q += 32) Y[q * kW2YS + cc] = 0.0;
    }
  }

  // ---- blocked forward substitution L·W = Y ---------------------------------------------------------------
  const double *Dinv = L.Dinv + (size_t)s * (rld / kCholNb) * kCholNb * kCholNb;
  for (int J = 0; J < r; J += kCholNb) {
    const int jb = min(kCholNb, r - J);
    const double *Dg = Dinv + (size_t)(J / kCholNb) * kCholNb * kCholNb;
    for (int e = tid; e < 32 * 32; e += 256) Xs[e] = Dg[e];          // zero above the diagonal and past jb
    __syncthreads();                                                 // also orders the gather / previous update
    double w[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) w[k] = Y[(J + k) * kW2YS + lane];
    // W_J = D⁻¹·Y_J : warp handles rows 4·warp .. 4·warp+3 of the block
    double out[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double2 *x = reinterpret_cast<const double2 *>(Xs + (4 * warp + u) * 32);
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double2 xv = x[k];
        a0 = fma(xv.x, w[2 * k], a0);
        a1 = fma(xv.y, w[2 * k + 1], a1);
      }
      out[u] = a0 + a1;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) Y[(J + 4 * warp + u) * kW2YS + lane] = out[u];
    __syncthreads();
    if (J + jb >= r) break;
#pragma unroll
    for (int k = 0; k < 32; ++k) w[k] = Y[(J + k) * kW2YS + lane];   // this thread's column of W_J
    // rows below: Y[i][:] −= L[i][J..J+32)·W_J, L staged 64 rows at a time (k pairs XOR-swizzled by row)
    for (int i0 = J + jb; i0 < r; i0 += kW2Chunk) {
      const int nrows = min(kW2Chunk, r - i0);
      for (int e = tid; e < kW2Chunk * 32; e += 256) {
        const int ii = e & (kW2Chunk - 1), k = e >> 6;
        const double v = (ii < nrows && k < jb) ? Sb[(size_t)(J + k) * sld + i0 + ii] : 0.0;
        Lp[ii * 32 + (k ^ ((ii & 15) << 1))] = v;
      }
      __syncthreads();
      for (int ii = warp; ii < nrows; ii += 8) {
        const int sw = (ii & 15) << 1;
        const double *lrow = Lp + ii * 32;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const double2 lv = *reinterpret_cast<const double2 *>(lrow + (k ^ sw));
          a0 = fma(lv.x, w[k], a0);
          a1 = fma(lv.y, w[k + 1], a1);
        }
        Y[(i0 + ii) * kW2YS + lane] -= a0 + a1;
      }
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- μ += Wᵀ·(L⁻¹ν) (:306), θ wrapped (:307); exact diagonal of the downdate -------------------------------
  double *mu = L.mu + (size_t)s * ld;
  for (int cc = warp; cc < kW2Cols; cc += 8) {
    const int c = c0 + cc;
    double a = 0.0, d2 = 0.0, mx = 0.0;
    for (int k = lane; k < r; k += 32) {
      const double wv = Y[k * kW2YS + cc];
      a = fma(wv, Sb[(size_t)k * sld + r], a);
      d2 = fma(wv, wv, d2);
      mx = fmax(mx, fabs(wv));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      d2 += __shfl_xor_sync(0xffffffffu, d2, off);
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) {
      if (c < n) {
        const double v = mu[c] + a;
        mu[c] = (c == 2) ? wrap_angle(v) : v;
      }
      // The diagonal of the downdate is a sum of squares: every truncation of a tensor-core product has
      // the same sign there and would accumulate step after step, so it is kept in fp64.
      if (L.Wdiag) L.Wdiag[(size_t)s * ld + c] = (c < n) ? d2 : 0.0;
      if (L.Wq) {
        const int e = (mx > 0.0 && c < n) ? ilogb(mx) + 2 : 0;
        sexp[cc] = e;
        L.Wexp[(size_t)s * ld + c] = e;
        // int8 slices resolve 2^-29 of the row scale 2^e; when the downdate removes almost all of a state's
        // variance that is no longer small against the posterior → this frame takes the fp64 SYRK.
        if (mx > 0.0 && c < n) {
          const double post = Sg[(size_t)c * ld + c] - d2;
          if (!(post > 0.0) || scalbn(1.0, 2 * e) > kMaxSliceGain2 * post) atomicOr(&L.st[s].exact_update, 1);
        }
      }
    }
  }
  __syncthreads();

  // ---- operand panels of Wᵀ (row c, K contiguous), zero beyond r ---------------------------------------------------
  if (L.W64) {
    double *W = L.W64 + (size_t)s * ld * rld;
    for (int e = tid; e < kW2Cols * rld; e += 256) {
      const int cc = e / rld, k = e - cc * rld;
      W[(size_t)(c0 + cc) * rld + k] = (k < r) ? Y[k * kW2YS + cc] : 0.0;
    }
  }
  if (L.Wq) {
    const int kq4 = L.kq / 4;
    for (int e4 = tid; e4 < kW2Cols * kq4; e4 += 256) {
      const int cc = e4 / kq4, k0 = (e4 - cc * kq4) * 4;
      const int e = sexp[cc];
      const bool live = (c0 + cc < n);
      uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u;
        double rem = (live && k < r) ? scalbn(Y[k * kW2YS + cc], 7 - e) : 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const double d = rint(rem);
          packed[p] |= ((uint32_t)(int)d & 0xffu) << (8 * u);
          rem = (rem - d) * 128.0;
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
        *reinterpret_cast<uint32_t *>(L.Wq + (((size_t)s * 4 + p) * ld + c0 + cc) * L.kq + k0) = packed[p];
    }
  } else if (L.Wt_hi) {
    float *Wh = L.Wt_hi + (size_t)s * ld * rld, *Wl = L.Wt_lo + (size_t)s * ld * rld;
    for (int e = tid; e < kW2Cols * rld; e += 256) {
      const int cc = e / rld, k = e - cc * rld;
      const double wv = (k < r) ? Y[k * kW2YS + cc] : 0.0;
      const float hi = to_tf32((float)wv);
      const float lo = to_tf32((float)(wv - (double)hi));
      Wh[(size_t)(c0 + cc) * rld + k] = hi;
      Wl[(size_t)(c0 + cc) * rld + k] = lo;
    }
  }
}

}  // namespace rekf
