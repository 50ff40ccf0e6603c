// solve_ll.cuh — W = L⁻¹·(H·Σ) in the SHADOW of the Cholesky: a left-looking, flag-paced TRSM.
//
// Reference: K_t = Σ·Hᵀ·S⁻¹, mu += K_t·(z − ẑ) (reflector_ekf_slam.cc:305-307), as in solve_w.cuh: W = L⁻¹·H·Σ, K·ν = Wᵀ·(L⁻¹ν).
//
// k_solve_w3 (solve_w.cuh) starts when the factor is complete and then walks the same seven 32-row blocks the Cholesky has just
// walked, one dependent step after the other: two serial spines back to back (~60 µs + ~20-70 µs at config C3).  This kernel runs
// BESIDE k_cholesky_smem instead and consumes block column J of L the moment the factorisation publishes it:
//   * left-looking: step J forms Ỹ_J = Y_J − Σ_{K<J} L_JK·W_K and then W_J = X_J·Ỹ_J (X_J = L_JJ⁻¹ from the Cholesky kernel).
//     The sum only needs block columns K < J, which were published at earlier steps — it is done BEFORE block column J arrives;
//     what is left once the Cholesky raises flag J is one 32x32x32 product.  After the last flag: that product and the epilogue.
//   * W lives in global memory / L2 (Layout::W64, the fp64 panel the exact paths consume anyway), not in shared memory: a CTA
//     stages 32x32 blocks of L and of its own earlier W_K through a double buffer (47 KB), so FOUR CTAs fit an SM and the
//     544 column tiles of 8 sessions are all resident at once beside the 8 Cholesky CTAs (the resident-tile kernel needs 113 KB:
//     two waves at 8 sessions, and the second wave would start after the factorisation).
//   * pacing: per-session flags in global memory (Layout::sync; cleared by k_observation_front): k_cholesky_smem sets flag J
//     when block column J, X_J and ν_J have landed in global memory (bulk copies complete → proxy fence → release), k_gather_y
//     sets one flag per (32-row block, 128-column tile) of Y.  One thread of the CTA polls (ld.acquire.gpu), bounded: a wait
//     that gives up sets FLAG_SYNC_TIMEOUT and carries on (garbage instead of a hang; never seen).
//   * co-residency (no deadlock): the three kernels form one programmatic-dependent-launch chain in one stream,
//     k_cholesky_smem → k_gather_y → k_solve_ll, each triggering at its start: a dependent grid is launched only when EVERY
//     block of its primary has triggered, i.e. is resident, so no polling block can hold an SM a Cholesky or gather block still
//     needs.  Without the attribute (several pipeline groups, profiling, ncu/compute-sanitizer serialisation) the kernels run
//     one after the other and every flag is already up: same results, no overlap.
// 128 threads = 4 warps; warp = 8-column tile of the CTA's 32 columns, four 8-row DMMA tiles each.
#pragma once
#include "chol_smem.cuh"
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"

namespace rekf {

constexpr int kLLThreads = 128;
constexpr int kLLP = 36;                 // pitch of a staged 32x32 block: DMMA fragment loads (8 x 4) conflict free
constexpr int kLLBlk = 32 * kLLP;
constexpr int kLLQP = 20;                // words per staged (digit plane, column) row of 64 K-bytes

inline size_t smem_solve_ll() { return sizeof(double) * (5 * kLLBlk + 4 * 32) + 64 * sizeof(int); }

__global__ void __launch_bounds__(kLLThreads, 4) k_solve_ll(Layout L) {
  if (!L.shadow) pdl_wait();
  timeline_mark(L, 4);
  extern __shared__ __align__(16) double sm_d[];
  const int s = L.s0 + blockIdx.z;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int n = internal_dim(st.N);
  const int c0 = blockIdx.x * 32;
  if (c0 >= round_up(n, kSigmaTile)) return;
  const int ld = L.ld, sld = L.sld, rld = L.rld;
  const double *Sg = L.sigma + (size_t)s * ld * ld;
  const double *Sb = L.Sbuf + (size_t)s * rld * sld;
  const double *Dinv = L.Dinv + (size_t)s * (rld / kCholNb) * kCholNb * kCholNb;
  const double *Yg = L.Ybuf + (size_t)s * rld * ld + c0;
  double *Wg = L.W64 + (size_t)s * rld * ld + c0;
  const int *flags = L.sync + (size_t)s * L.sync_n;
  const int gx = ld / 128;
  const int *gyf = flags + 8 + c0 / 128;               // + y·gx: Y rows 32y.. of this CTA's 128-column tile are in global memory
  double *stg = sm_d;                                  // [2 stages][L block, W block][32 k][kLLP]
  double *T = stg;                                     // Ỹ_J, k-major (B operand of the diagonal product); aliases stage 0
  double *Xs = sm_d + 4 * kLLBlk;                      // X_J, row-major
  double *cstat = Xs + kLLBlk;                         // [3][32] per column: Wᵀ·(L⁻¹ν), Σ W², max |W|
  double *sdiag = cstat + 96;                          // [32] prior Σ[c][c]
  int *sexp = reinterpret_cast<int *>(sdiag + 32);     // [32]
  int &s_ok = sexp[32];                                // every flag wait succeeded (thread 0's)
  const int tid = threadIdx.x, lane = tid & 31, nt = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  if (tid < 32) sdiag[tid] = Sg[(size_t)min(c0 + tid, ld - 1) * (ld + 1)];
  if (tid == 0) s_ok = 1;
  const int nblk = (r + kCholNb - 1) / kCholNb;

  auto cp16 = [](double *dst, const double *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  // L_JK (rows J0.., columns 32kb..) and W_K (rows 32kb.., this CTA's columns), both k-major, into stage `sg`
  auto stage_chunk = [&](int J0, int kb, int sg) {
    double *Ls = stg + (size_t)(2 * sg) * kLLBlk, *Ws = Ls + kLLBlk;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + kLLThreads * u, k = e >> 4, q = (e & 15) * 2;
      cp16(Ls + k * kLLP + q, Sb + (size_t)(32 * kb + k) * sld + J0 + q);
      cp16(Ws + k * kLLP + q, Wg + (size_t)(32 * kb + k) * ld + q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  double pa[2] = {0.0, 0.0}, pd[2] = {0.0, 0.0}, pm[2] = {0.0, 0.0};   // this thread's columns 8nt+2t4, +1: running sums
  for (int jb = 0; jb < nblk; ++jb) {
    const int J0 = kCholNb * jb;
    if (tid == 0 && !sync_wait(gyf + jb * gx)) s_ok = 0;
    __syncthreads();                                   // Y_J is there; the previous step's W_J stores and T reads are done
    if (jb > 0) stage_chunk(J0, 0, 0);
    double acc[4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int row = J0 + 8 * mt + g;
      double2 v = make_double2(0.0, 0.0);
      if (row < r) v = __ldcg(reinterpret_cast<const double2 *>(Yg + (size_t)row * ld + 8 * nt + 2 * t4));
      acc[mt][0] = v.x; acc[mt][1] = v.y;
    }
    // ---- Ỹ_J = Y_J − Σ_{K<J} L_JK·W_K: block columns K < J were published at earlier steps --------------------------
    for (int kb = 0; kb < jb; ++kb) {
      if (kb + 1 < jb) {
        stage_chunk(J0, kb + 1, (kb + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                                 // chunk kb has landed for every thread
      const double *Lc = stg + (size_t)(2 * (kb & 1)) * kLLBlk, *Wc = Lc + kLLBlk;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const double b = Wc[(4 * ks + t4) * kLLP + 8 * nt + g];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          dmma884(acc[mt][0], acc[mt][1], -Lc[(4 * ks + t4) * kLLP + 8 * mt + g], b, acc[mt][0], acc[mt][1]);
      }
      __syncthreads();                                 // stage kb&1 may be refilled
    }
    // ---- the diagonal block: wait for block column J of the factor (X_J, ν_J), then W_J = X_J·Ỹ_J -------------------
    if (tid == 0 && !sync_wait(flags + jb)) s_ok = 0;
    if (jb == nblk - 1) timeline_mark(L, 10);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + kLLThreads * u, j = e >> 4, q = (e & 15) * 2;
      cp16(Xs + j * kLLP + q, Dinv + (size_t)jb * kCholNb * kCholNb + j * kCholNb + q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    double nuv[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int row = J0 + 8 * mt + g;
      nuv[mt] = (row < r) ? __ldcg(Sb + (size_t)row * sld + r) : 0.0;
      double2 v = make_double2(acc[mt][0], acc[mt][1]);
      if (row >= r) v = make_double2(0.0, 0.0);        // rows past r: the factor's rows there are not L (ν row, stale data)
      *reinterpret_cast<double2 *>(T + (8 * mt + g) * kLLP + 8 * nt + 2 * t4) = v;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double w[4][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) w[mt][0] = w[mt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {                   // X is lower triangular: row tile mt needs k < 8(mt+1)
      const double b = T[(4 * ks + t4) * kLLP + 8 * nt + g];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        if (ks < 2 * (mt + 1)) dmma884(w[mt][0], w[mt][1], Xs[(8 * mt + g) * kLLP + 4 * ks + t4], b, w[mt][0], w[mt][1]);
    }
    if (jb == nblk - 1) pdl_trigger();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int row = J0 + 8 * mt + g;                  // < rld; rows r.. of a partial last block come out as zeros
      *reinterpret_cast<double2 *>(Wg + (size_t)row * ld + 8 * nt + 2 * t4) = make_double2(w[mt][0], w[mt][1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        pa[e] = fma(w[mt][e], nuv[mt], pa[e]);
        pd[e] = fma(w[mt][e], w[mt][e], pd[e]);
        pm[e] = fmax(pm[e], fabs(w[mt][e]));
      }
    }
  }

  // ---- μ += Wᵀ·(L⁻¹ν) (:306), θ wrapped (:307); exact diagonal of the downdate; row scales -----------------------------
#pragma unroll
  for (int e = 0; e < 2; ++e) {
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      pa[e] += __shfl_xor_sync(0xffffffffu, pa[e], off);
      pd[e] += __shfl_xor_sync(0xffffffffu, pd[e], off);
      pm[e] = fmax(pm[e], __shfl_xor_sync(0xffffffffu, pm[e], off));
    }
    if (g == 0) {
      const int cc = 8 * nt + 2 * t4 + e;
      cstat[cc] = pa[e]; cstat[32 + cc] = pd[e]; cstat[64 + cc] = pm[e];
    }
  }
  for (int k = kCholNb * nblk + nt; k < rld; k += 4) Wg[(size_t)k * ld + lane] = 0.0;   // measurement rows this frame does not have
  __syncthreads();
  double *mu = L.mu + (size_t)s * ld;
  if (nt == 0) {
    const int cc = lane, c = c0 + cc;
    const double a = cstat[cc], d2 = cstat[32 + cc], mx = cstat[64 + cc];
    if (c < n) {
      const double v = mu[c] + a;
      mu[c] = (c == 2) ? wrap_angle(v) : v;
    }
    // The diagonal of the downdate is a sum of squares: every truncation of a tensor-core product has the same sign there and
    // would accumulate step after step, so it is kept in fp64.
    if (L.Wdiag) L.Wdiag[(size_t)s * ld + c] = (c < n) ? d2 : 0.0;
    if (L.Wq) {
      const int e = (mx > 0.0 && c < n) ? ilogb(mx) + 2 : 0;
      sexp[cc] = e;
      L.Wexp[(size_t)s * ld + c] = e;
      L.Wscale[(size_t)s * ld + c] = scalbn(1.0, e);
      // admission of the int8 slices (see k_solve_w3): slots whose downdate removes almost all of their variance go to fp64
      bool exact = false;
      if (mx > 0.0 && c < n) {
        const double post = sdiag[cc] - d2;
        exact = !(post > 0.0) || scalbn(1.0, 2 * e) > kMaxSliceGain2 * post;
        if (exact) {
          const int pos = atomicAdd(&st.exact_slots, 1);
          if (pos < kMaxExactSlots) L.exact_list[(size_t)s * kMaxExactSlots + pos] = c;
          else atomicOr(&st.exact_update, 1);
        }
      }
      L.Wflag[(size_t)s * ld + c] = exact ? 1 : 0;
    }
    if (cc == 0 && !s_ok) atomicOr(&st.flags, FLAG_SYNC_TIMEOUT);
  }
  __syncthreads();

  // ---- int8 digit slices of the row-scaled Wᵀ (see k_solve_w3), 64 measurement rows per pass, W re-read from L2 ----------
  if (L.Wq) {
    const double kMagic = 6755399441055744.0;          // 2^52 + 2^51
    uint32_t *Qs = reinterpret_cast<uint32_t *>(stg);  // [4 planes][32 columns][kLLQP]
    const bool live = (c0 + lane < n);
    const double sc = __longlong_as_double((long long)(1023 + 28 - sexp[lane]) << 52);  // 2^(28-e)
    for (int ch = 0; ch < (L.kq >> 6); ++ch) {
      double wv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int k = 64 * ch + 4 * (nt + 4 * u) + v;
          wv[u][v] = (live && k < r) ? __ldcg(Wg + (size_t)k * ld + lane) : 0.0;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          int q = __double2loint(fma(wv[u][v], sc, kMagic));
#pragma unroll
          for (int p = 3; p > 0; --p) {
            const int d = ((q + 64) & 127) - 64;
            packed[p] |= ((uint32_t)d & 0xffu) << (8 * v);
            q = (q - d) >> 7;
          }
          packed[0] |= ((uint32_t)q & 0xffu) << (8 * v);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) Qs[(p * 32 + lane) * kLLQP + nt + 4 * u] = packed[p];
      }
      __syncthreads();
      for (int e = tid; e < 128 * 4; e += kLLThreads) {  // row = plane * 32 + column, 16 bytes per thread
        const int row = e >> 2, j = (e & 3) * 4;
        const int p = row >> 5, cc = row & 31;
        uint32_t *dst = reinterpret_cast<uint32_t *>(L.Wq + wq_offset(L, s, p, c0 + cc, 64 * ch + 4 * j));
        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(Qs + row * kLLQP + j);
      }
      __syncthreads();
    }
  }
  timeline_mark(L, 11);
}

}  // namespace rekf
