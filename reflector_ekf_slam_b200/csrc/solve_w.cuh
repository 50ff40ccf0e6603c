// solve_w.cuh — W = L⁻¹·(H·Σ), μ += Wᵀ·(L⁻¹ν), and the operand panels of the covariance SYRK.
//
// Reference: K_t = Σ·Hᵀ·S⁻¹ and mu += K_t·(z − ẑ) (reflector_ekf_slam.cc:305-307).  With S = L·Lᵀ the gain
// never has to be formed: W = L⁻¹·H·Σ gives K·ν = Wᵀ·(L⁻¹ν) and K·H·Σ = Wᵀ·W.
//
// One CTA = 32 columns of W (32 state slots), 256 threads = 8 warps:
//   load     the CTA's 32 columns of Y = H·Σ (k_gather_y computed it on a side stream beside the Cholesky: H has <= 5
//            non-zeros per row, :272-275, so Y is a block gather of Σ, 3.6 MB per session that stay in L2);
//   solve    blocked forward substitution, right-looking, 32-row blocks, entirely on the fp64 tensor pipe
//            (mma.sync m8n8k4 → DMMA, full rate on B200): W_J = X_J·Y_J with the block inverses X_J = L_JJ⁻¹
//            the Cholesky kernel emits (no substitution chain), then Y_I −= L_IJ·W_J for the rows below with
//            W_J's B fragments pinned in registers and L streamed through a swizzled shared-memory chunk;
//   epilogue μ update (θ wrapped, :307), exact diag(WᵀW), int8 digit slices / tf32 hi-lo / fp64 panels.
#pragma once
#include "chol_smem.cuh"
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"

namespace rekf {

constexpr int kW3Cols = 32;
constexpr int kW3YS = 36;                // pitch of the Y/W block: B-fragment loads (4 k x 8 n) conflict free
constexpr int kW3Chunk = 64;             // rows of L staged per pass
constexpr int kW3LP = kW3Chunk + 4;      // pitch of a staged L chunk, k-major [32 k][68]: A-fragment loads conflict free

inline size_t smem_solve_w3(int rld) {
  return sizeof(double) * ((size_t)(rld + 8) * kW3YS + (size_t)2 * 32 * kW3LP + 32 * 32 + rld + 32) + 64 * sizeof(int);
}

// column swizzle of the 32-wide operand tiles: conflict-free both for row-contiguous staging stores and for the
// DMMA A-fragment loads (8 rows x 4 k)
__device__ __forceinline__ int swz(int row) { return ((row & 3) << 2) | ((row >> 2) & 3); }

// Blocked forward substitution L·W = Y on the fp64 tensor pipe, in place in the shared-memory tile Y ([rows][kW3YS], 32 columns).
// L (column-major lower triangle, pitch sld) and the inverses of its 32x32 diagonal blocks (Dinv) are read from global / L2:
// L chunks and block inverses arrive by cp.async one stage ahead of their use (double-buffered chunks), so the L2 latency of
// the operand stream is hidden behind the DMMAs of the previous stage.  All 256 threads of the CTA call it; Y must be complete
// (or in flight in an earlier cp.async group); on return every thread may read W from Y.
__device__ __forceinline__ void trsm_forward_tile(double *Y, double *Lp, double *Xs, const double *Sb, const double *Dinv, int r,
                                                  int sld, int rld) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;             // DMMA fragment coordinates
  const int nt = warp & 3, rp = warp >> 2;            // this warp's 8-column tile and row-tile parity
  auto cp8 = [](double *dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  auto stage_X = [&](int Jx) {
    const double *Dg = Dinv + (size_t)(Jx / kCholNb) * kCholNb * kCholNb;
    for (int e = tid; e < 32 * 32; e += 256) {
      const int i = e >> 5, k = e & 31;
      cp8(Xs + i * 32 + (k ^ swz(i)), Dg + e);
    }
  };
  // rows i0s.. of panel Js (always a full 32-column panel), k-major like L's own column-major storage: 16-byte copies
  auto stage_L = [&](int Js, int i0s, double *buf) {
    const int nr = min(kW3Chunk, r - i0s);
    for (int e = tid; e < 32 * (kW3Chunk / 2); e += 256) {
      const int k = e >> 5, ii = (e & 31) * 2;
      double *dst = buf + k * kW3LP + ii;
      const double *src = Sb + (size_t)(Js + k) * sld + i0s + ii;
      if (ii + 1 < nr) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
      } else {
        dst[0] = (ii < nr) ? src[0] : 0.0;
        dst[1] = 0.0;
      }
    }
  };
  int cur = 0;                                        // chunk buffer holding the stage about to be consumed
  stage_X(0);
  if (r > kCholNb) stage_L(0, kCholNb, Lp);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int J = 0; J < r; J += kCholNb) {
    const int jb = min(kCholNb, r - J);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                  // X_J landed; also orders the gather / previous trailing update
    // W_J = X_J·Y_J : 4 row tiles x 4 column tiles of 8x8; this warp: column tile nt, row tiles rp and rp+2
    double w0[2], w1[2], u0[2], u1[2];
    w0[0] = w0[1] = w1[0] = w1[1] = u0[0] = u0[1] = u1[0] = u1[1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                  // four independent chains; X is zero above its diagonal
      const double b = Y[(J + 4 * ks + t4) * kW3YS + 8 * nt + g];
      const double b2 = Y[(J + 4 * (ks + 4) + t4) * kW3YS + 8 * nt + g];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int xi = 8 * (rp + 2 * h) + g;
        dmma884(w0[h], w1[h], Xs[xi * 32 + ((4 * ks + t4) ^ swz(xi))], b, w0[h], w1[h]);
        dmma884(u0[h], u1[h], Xs[xi * 32 + ((4 * (ks + 4) + t4) ^ swz(xi))], b2, u0[h], u1[h]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) { w0[h] += u0[h]; w1[h] += u1[h]; }
    __syncthreads();                                  // every warp has read Y_J and X_J
    if (J + kCholNb < r) {                            // X of the next block: lands during this block's trailing update
      stage_X(J + kCholNb);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rt = rp + 2 * h;
      *reinterpret_cast<double2 *>(Y + (J + 8 * rt + g) * kW3YS + 8 * nt + 2 * t4) = make_double2(w0[h], w1[h]);
    }
    __syncthreads();
    if (J + jb >= r) break;
    double wb[8];                                     // B fragments of W_J for this warp's column tile
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) wb[ks] = Y[(J + 4 * ks + t4) * kW3YS + 8 * nt + g];
    // rows below: Y[i][:] −= L[i][J..J+32)·W_J, 64 rows of L per stage
    for (int i0 = J + jb; i0 < r; i0 += kW3Chunk) {
      const int nrows = min(kW3Chunk, r - i0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                // this stage's chunk landed; the other buffer is free again
      {                                               // prefetch the next stage into the other buffer
        int Jn = J, in = i0 + kW3Chunk;
        if (in >= r) { Jn = J + kCholNb; in = Jn + kCholNb; }
        if (in < r) {
          stage_L(Jn, in, Lp + (cur ^ 1) * 32 * kW3LP);
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
      }
      const double *Lc = Lp + cur * 32 * kW3LP;
      const int ntile = (nrows + 7) >> 3;
      // this warp's (up to) four row tiles of the chunk as four independent DMMA chains
      double d0[4], d1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int rt = rp + 2 * q;                      // rows past nrows are zero-filled in Lc: harmless
        const double2 cv = *reinterpret_cast<const double2 *>(Y + min(i0 + 8 * rt + g, rld + 7) * kW3YS + 8 * nt + 2 * t4);
        d0[q] = cv.x; d1[q] = cv.y;
      }
      double e0[4], e1[4];                            // second half of K as its own chain: half the dependent depth
#pragma unroll
      for (int q = 0; q < 4; ++q) e0[q] = e1[q] = 0.0;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int li = 8 * (rp + 2 * q) + g;
          dmma884(d0[q], d1[q], -Lc[(4 * ks + t4) * kW3LP + li], wb[ks], d0[q], d1[q]);
          dmma884(e0[q], e1[q], -Lc[(4 * (ks + 4) + t4) * kW3LP + li], wb[ks + 4], e0[q], e1[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { d0[q] += e0[q]; d1[q] += e1[q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int rt = rp + 2 * q;
        if (rt < ntile) *reinterpret_cast<double2 *>(Y + (i0 + 8 * rt + g) * kW3YS + 8 * nt + 2 * t4) = make_double2(d0[q], d1[q]);
      }
      cur ^= 1;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

// Split frames (chol_smem.cuh), stage 2: rows r1+1..r of S21 (and ν) against L11 — L21[i][:] = S21[i][:]·L11⁻ᵀ.  Row i of S21 is a
// right-hand side of L11·x = S21[i][:]ᵀ, so 32 rows form one tile of trsm_forward_tile; S21[i][k] = Sb[k][i] is contiguous in i.
// grid (ceil((rcap - 32) / 32), 1, Sg).
__global__ void __launch_bounds__(256, 1) k_chol_trsm_rows(Layout L) {
  extern __shared__ double sm_d[];
  const int s = L.s0 + blockIdx.z;
  const int r = L.st[s].r;
  const int r1 = chol_split(r);
  if (r1 <= 0) return;
  const int c0 = r1 + 1 + blockIdx.x * kW3Cols;       // first absolute row of this tile (row r1 itself rode along in pass 1)
  if (c0 > r) return;
  const int sld = L.sld, rld = L.rld;
  double *Sb = L.Sbuf + (size_t)s * rld * sld;
  double *Y = sm_d;                                   // [r1 + 8][kW3YS]
  double *Lp = Y + (size_t)(rld + 8) * kW3YS;
  double *Xs = Lp + 2 * 32 * kW3LP;
  const int tid = threadIdx.x;
  for (int e = tid; e < r1 * 32; e += 256) {
    const int k = e >> 5, cc = e & 31;
    Y[k * kW3YS + cc] = (c0 + cc <= r) ? Sb[(size_t)k * sld + c0 + cc] : 0.0;
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  trsm_forward_tile(Y, Lp, Xs, Sb, L.Dinv + (size_t)s * (rld / kCholNb) * kCholNb * kCholNb, r1, sld, rld);
  for (int e = tid; e < r1 * 32; e += 256) {
    const int k = e >> 5, cc = e & 31;
    if (c0 + cc <= r) Sb[(size_t)k * sld + c0 + cc] = Y[k * kW3YS + cc];
  }
}

__global__ void __launch_bounds__(256, 2) k_solve_w3(Layout L) {
  pdl_wait();
  timeline_mark(L, 4);
  extern __shared__ double sm_d[];
  const int s = L.s0 + blockIdx.z;
  const SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int n = internal_dim(st.N);
  const int c0 = blockIdx.x * kW3Cols;
  if (c0 >= round_up(n, kSigmaTile)) return;
  const int ld = L.ld, sld = L.sld, rld = L.rld;
  const double *Sg = L.sigma + (size_t)s * ld * ld;
  const double *Sb = L.Sbuf + (size_t)s * rld * sld;
  double *Y = sm_d;                                   // [rld + 8][kW3YS] (the last row tile may overhang r by 7 rows)
  double *Lp = Y + (size_t)(rld + 8) * kW3YS;         // [2][64][32] swizzled chunks of an L panel (double buffer)
  double *Xs = Lp + 2 * 32 * kW3LP;                   // [32][32] swizzled inverse of the current diagonal block
  double *nu = Xs + 32 * 32;                          // [rld] L⁻¹ν (row r of the factor), staged once
  int *sexp = reinterpret_cast<int *>(nu + rld);      // [32]
  double *sdiag = reinterpret_cast<double *>(sexp + 64);   // [32] prior Σ[c][c] of this CTA's columns, fetched up front
  if (threadIdx.x < kW3Cols) sdiag[threadIdx.x] = Sg[(size_t)min(c0 + (int)threadIdx.x, ld - 1) * (ld + 1)];
  for (int k = threadIdx.x; k < r; k += 256) nu[k] = Sb[(size_t)k * sld + r];   // visible after the gather's barriers
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r32 = round_up(r, 32);
#ifdef REKF_SOLVE_TIMING
  double *tlog = L.innov + (size_t)s * L.rcap;        // cycle stamps of CTA 0 (thread 0)
  int tl = 0;
#define REKF_WSTAMP() do { if (tid == 0 && blockIdx.x == 0) tlog[tl++] = (double)clock64(); } while (0)
#else
#define REKF_WSTAMP() do { } while (0)
#endif
#ifdef REKF_SOLVE_TIMING2
#define REKF_WSTAMP2() REKF_WSTAMP()
#else
#define REKF_WSTAMP2() do { } while (0)
#endif
  REKF_WSTAMP();

  // ---- Y tile: rows 0..r-1 of Y = H·Σ (k_gather_y wrote them beside the Cholesky), this CTA's 32 columns = 256 contiguous
  //      bytes per row, 16-byte cp.async; rows r..r32-1 of the tile are zero --------------------------------------------
  {
    const double *Yg = L.Ybuf + (size_t)s * rld * ld + c0;
    for (int e = tid; e < r32 * 16; e += 256) {
      const int q = e >> 4, cc = (e & 15) * 2;
      double *dst = Y + q * kW3YS + cc;
      if (q < r) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(Yg + (size_t)q * ld + cc) : "memory");
      } else {
        dst[0] = 0.0; dst[1] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // waited for together with the first L / X stage below
  }
  REKF_WSTAMP();

  trsm_forward_tile(Y, Lp, Xs, Sb, L.Dinv + (size_t)s * (rld / kCholNb) * kCholNb * kCholNb, r, sld, rld);
  pdl_trigger();
  REKF_WSTAMP();

  // ---- μ += Wᵀ·(L⁻¹ν) (:306), θ wrapped (:307); exact diagonal of the downdate -------------------------------
  double *mu = L.mu + (size_t)s * ld;
  double *red = Lp;                                   // [3][8][32] partial sums (the L chunk buffers are dead)
  {                                                   // lane = column, warps stride the rows: conflict-free reads
    double a = 0.0, d2 = 0.0, mx = 0.0;
    for (int k = warp; k < r; k += 8) {
      const double wv = Y[k * kW3YS + lane];
      a = fma(wv, nu[k], a);
      d2 = fma(wv, wv, d2);
      mx = fmax(mx, fabs(wv));
    }
    red[warp * 32 + lane] = a;
    red[(8 + warp) * 32 + lane] = d2;
    red[(16 + warp) * 32 + lane] = mx;
  }
  __syncthreads();
  if (warp == 0) {
    const int cc = lane, c = c0 + cc;
    double a = 0.0, d2 = 0.0, mx = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += red[w * 32 + cc];
      d2 += red[(8 + w) * 32 + cc];
      mx = fmax(mx, red[(16 + w) * 32 + cc]);
    }
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
      L.Wscale[(size_t)s * ld + c] = scalbn(1.0, e);
      // int8 slices resolve 2^-29 of the row scale 2^e; when the downdate removes almost all of a state's
      // variance that is no longer small against the posterior, so such slots are flagged: the tensor kernel
      // skips their rows/columns and k_syrk_exact_rows does them in fp64.  More than kMaxExactSlots of them
      // (first update after map building: everything collapses) and the whole frame goes to the fp64 SYRK.
      bool exact = false;
      if (mx > 0.0 && c < n) {
        const double post = sdiag[cc] - d2;
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
  REKF_WSTAMP();

  // ---- operand panels of Wᵀ (row c, K contiguous), zero beyond r ---------------------------------------------------
  if (L.W64) {
    // measurement-row major panel: row k of the tile is 256 contiguous bytes of W64 row k (lane = column: conflict-free reads,
    // full-line writes)
    double *Wg = L.W64 + (size_t)s * ld * rld + c0 + lane;
    for (int k = warp; k < rld; k += 8) Wg[(size_t)k * ld] = (k < r) ? Y[k * kW3YS + lane] : 0.0;
  }
  REKF_WSTAMP2();
  if (L.Wq) {
    // x = w·2^-e, |x| < 1/2, x ≈ Σ_p d_p·2^(-7(p+1)).  q = rint(w·2^(28-e)) by magic-number addition (one fp64 FMA,
    // |q| <= 2^27), then four balanced base-128 digits d_p ∈ [-64, 64] by integer arithmetic: q = Σ d_p·128^(3-p)
    const double kMagic = 6755399441055744.0;          // 2^52 + 2^51
    constexpr int kQP = 68;                            // words per staged row: 64 K-groups of 4, 16-byte aligned rows
    uint32_t *Qs = reinterpret_cast<uint32_t *>(Lp);   // [4 planes][32 columns][kQP]
    const int kq4 = L.kq / 4;
    const bool live = (c0 + lane < n);
    const double sc = __longlong_as_double((long long)(1023 + 28 - sexp[lane]) << 52);  // 2^(28-e)
    for (int g0 = 0; g0 < kq4; g0 += 64) {             // 256 K values per pass
      const int ng = min(64, kq4 - g0);
      for (int kg = warp; kg < ng; kg += 8) {
        const int k0 = (g0 + kg) * 4;
        uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + u;
          const double wv = (live && k < r) ? Y[k * kW3YS + lane] : 0.0;
          int q = __double2loint(fma(wv, sc, kMagic));
#pragma unroll
          for (int p = 3; p > 0; --p) {
            const int d = ((q + 64) & 127) - 64;
            packed[p] |= ((uint32_t)d & 0xffu) << (8 * u);
            q = (q - d) >> 7;
          }
          packed[0] |= ((uint32_t)q & 0xffu) << (8 * u);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) Qs[(p * 32 + lane) * kQP + kg] = packed[p];
      }
      __syncthreads();
      REKF_WSTAMP2();
      for (int e = tid; e < 128 * 16; e += 256) {      // row = plane * 32 + column, 16 bytes per thread
        const int row = e >> 4, j = (e & 15) * 4;
        if (j < ng) {                                  // kq is a multiple of 64 bytes: ng is a multiple of 16 words
          const int p = row >> 5, cc = row & 31;
          uint32_t *dst = reinterpret_cast<uint32_t *>(L.Wq + wq_offset(L, s, p, c0 + cc, 4 * (g0 + j)));
          *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(Qs + row * kQP + j);
        }
      }
      if (g0 + 64 < kq4) __syncthreads();
    }
  } else if (L.Wt_hi) {
    float *Wh = L.Wt_hi + (size_t)s * ld * rld, *Wl = L.Wt_lo + (size_t)s * ld * rld;
    for (int e = tid; e < kW3Cols * rld; e += 256) {
      const int cc = e / rld, k = e - cc * rld;
      const double wv = (k < r) ? Y[k * kW3YS + cc] : 0.0;
      const float hi = to_tf32((float)wv);
      const float lo = to_tf32((float)(wv - (double)hi));
      Wh[(size_t)(c0 + cc) * rld + k] = hi;
      Wl[(size_t)(c0 + cc) * rld + k] = lo;
    }
  }
  REKF_WSTAMP();
}

}  // namespace rekf
