// solve_ll.cuh — W = L⁻¹·(H·Σ) in the SHADOW of the Cholesky: a flag-paced TRSM whose working set lives in L2.
//
// Reference: K_t = Σ·Hᵀ·S⁻¹, mu += K_t·(z − ẑ) (reflector_ekf_slam.cc:305-307), as in solve_w.cuh: W = L⁻¹·H·Σ, K·ν = Wᵀ·(L⁻¹ν).
//
// k_solve_w3 (solve_w.cuh) starts when the factor is complete and then walks the same seven 32-row blocks the Cholesky has just
// walked, one dependent step after the other: two serial spines back to back (~60 µs + ~20-70 µs at config C3).  This kernel runs
// BESIDE k_cholesky_smem instead and consumes block column J of L the moment the factorisation publishes it:
//   * right-looking, so that the work is FRONT-loaded: when flag J comes up, W_J = X_J·Ỹ_J (X_J = L_JJ⁻¹ from the Cholesky
//     kernel) and then U_I −= L_IJ·W_J for the blocks I below, the next diagonal block first.  After the LAST flag only one
//     32x32x32 product and the epilogue are left (a left-looking order leaves the largest sum, 6 of 21 block products, for the end);
//   * Y = H·Σ is gathered by this kernel itself, one 32-row block ahead of its use (H has <= 5 non-zeros per row, :272-275:
//     Y[q][c] = A_q·Σ[0:3][c] + B_q·Σ[slot:slot+2][c], upper-triangle storage read through sym_idx) — no gather kernel, no Y buffer;
//     the accumulated updates U_I (Ỹ_I = Y_I + U_I) are THREAD-PRIVATE: a thread owns the same 8 elements of every block (its
//     DMMA accumulator fragment), keeps them in Layout::Ybuf between steps and re-reads only its own stores — no fences;
//   * nothing but staging in shared memory (49 KB: two L blocks, X_J, Ỹ_J / W_J, the next block's Σ values and H rows), so FOUR CTAs fit an SM and the 544 column tiles
//     of 8 sessions are all resident at once beside the 8 Cholesky CTAs (the resident-tile kernel needs 113 KB: two waves at 8
//     sessions, the second one after the factorisation); W goes to Layout::W64 (the fp64 panel the exact paths consume anyway)
//     block by block and is re-read from L2 for the int8 digit slices once the row scales are known;
//   * pacing: eight flags per session in global memory (Layout::sync; cleared by k_observation_front): k_cholesky_smem raises
//     flag J when block column J, X_J and ν_J are in global memory (bulk copies complete → proxy fence → release).  One thread
//     of the CTA polls (ld.acquire.gpu), bounded: a wait that gives up sets FLAG_SYNC_TIMEOUT and carries on (never seen);
//   * co-residency (no deadlock): k_cholesky_smem → this kernel is a programmatic-dependent-launch edge and the Cholesky
//     triggers at its start: the dependent grid is launched only when EVERY block of the primary is resident, so no polling
//     block can hold the SM a Cholesky block still needs.  Without the attribute (several pipeline groups, profiling,
//     ncu / compute-sanitizer serialisation) the kernels run one after the other and every flag is up: same results, no overlap.
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
constexpr int kLLGP = 33;                // columns per row pair of the staged Σ values (+1: the fragment reads spread over the banks)

inline size_t smem_solve_ll() { return sizeof(double) * (5 * kLLBlk + 4 * 32 + 32 * 6) + (64 + kCholResidentMax) * sizeof(int); }

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
  double *Wg = L.W64 + (size_t)s * rld * ld + c0;
  const int *flags = L.sync + (size_t)s * L.sync_n;
  double *Ls = sm_d;                                   // [2][32 k][kLLP] staged L_IJ, k-major (L's own column-major order)
  double *Xs = sm_d + 2 * kLLBlk;                      // X_J, row-major
  double *T = sm_d + 3 * kLLBlk;                       // Ỹ_J, k-major: B operand of the diagonal product ...
  double *Wt = T;                                      // ... then W_J, k-major: B operand of the trailing updates
  double *Gs = sm_d + 4 * kLLBlk;                      // [16 row pairs][kLLGP columns][2] Σ[slot][c], Σ[slot+1][c] of the NEXT block's rows
  double *cstat = sm_d + 5 * kLLBlk;                   // [3][32] per column: Wᵀ·(L⁻¹ν), Σ W², max |W|
  double *sdiag = cstat + 96;                          // [32] prior Σ[c][c]
  int *sexp = reinterpret_cast<int *>(sdiag + 32);     // [32]
  int &s_ok = sexp[32];                                // every flag wait succeeded (thread 0's)
  int *Hs = sexp + 64;                                 // [kCholResidentMax] landmark slot of every row of H
  double *Hr = reinterpret_cast<double *>(Hs + kCholResidentMax);   // [32][6] coefficients of the NEXT block's rows: A_q (3), pad, B_q (2)
  const int tid = threadIdx.x, lane = tid & 31, nt = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  if (tid < 32) sdiag[tid] = Sg[(size_t)min(c0 + tid, ld - 1) * (ld + 1)];
  if (tid == 0) s_ok = 1;
  const int nblk = (r + kCholNb - 1) / kCholNb;
  const int cme = c0 + 8 * nt + 2 * t4;                // this thread's columns cme, cme + 1 of every block (its accumulator fragment)
  double *Up = L.Ybuf + (size_t)s * rld * ld + cme;    // its private elements of the accumulated updates: row q at Up + q·ld
  double *Wp = Wg + 8 * nt + 2 * t4;

  auto cp16 = [](double *dst, const double *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  // L_IJ (rows 32I.., columns 32J..), k-major, into stage buffer `sg`
  auto stage_L = [&](int I, int J, int sg) {
    double *dst = Ls + (size_t)sg * kLLBlk;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + kLLThreads * u, k = e >> 4, q = (e & 15) * 2;
      cp16(dst + k * kLLP + q, Sb + (size_t)(32 * J + k) * sld + 32 * I + q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // Y = H·Σ, one 32-row block ahead of its use.  Row q of H: A_q on the pose slots, B_q on the slots of its landmark (:272-275);
  // rows 2k and 2k+1 belong to one reflector.  The landmark's two rows of Σ at this CTA's 32 columns come in by cp.async while
  // the previous step's updates run: row `slot` of the upper triangle where slot <= c (8-byte copies, coalesced across the
  // columns), Σ[c][slot..slot+1] where slot > c (one 16-byte copy per column).
  const double *Hp = L.Hp + (size_t)s * L.rcap * 4;
  const double *Hl = L.Hl + (size_t)s * L.rcap * 2;
  const int *Hslot = L.Hslot + (size_t)s * L.rcap;
  auto issue_gather = [&](int jb) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pair = nt + 4 * u, q0 = kCholNb * jb + 2 * pair, c = c0 + lane;
      double *dst = Gs + (size_t)(pair * kLLGP + lane) * 2;
      const int slot = (q0 < r) ? Hs[q0] : -1;
      if (slot < 0) {
        dst[0] = 0.0; dst[1] = 0.0;
      } else if (slot > c) {
        cp16(dst, Sg + (size_t)c * ld + slot);
      } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(Sg + (size_t)slot * ld + c) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 1)), "l"(Sg + sym_idx(slot + 1, c, ld)) : "memory");
      }
    }
    if (tid < 96) {                                    // the rows' coefficients: Hp[q][0..3] as two 16-byte halves, Hl[q][0..1] as one
      const int row = tid / 3, part = tid - 3 * row, q = kCholNb * jb + row;
      if (q < r) cp16(Hr + row * 6 + 2 * part, part < 2 ? Hp + 4 * q + 2 * part : Hl + 2 * q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double p0[2], p1[2], p2[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    p0[e] = Sg[sym_idx(0, cme + e, ld)]; p1[e] = Sg[sym_idx(1, cme + e, ld)]; p2[e] = Sg[sym_idx(2, cme + e, ld)];
  }
  // rows 32jb + 8mt + g at this thread's two columns, from the staged Σ values (the copies of issue_gather(jb) have landed)
  auto gather = [&](int jb, double (&y)[4][2]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int q = kCholNb * jb + 8 * mt + g;
      y[mt][0] = y[mt][1] = 0.0;
      if (q < r) {
        const int slot = Hs[q];
        const bool staged = slot == Hs[q & ~1];        // always, for the reference's row layout; kept general
        const double *hr = Hr + (8 * mt + g) * 6;
        const double a0 = hr[0], a1 = hr[1], a2 = hr[2], l0 = hr[4], l1 = hr[5];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = cme + e;
          double v = a0 * p0[e] + a1 * p1[e] + a2 * p2[e];
          if (slot >= 0) {
            double2 x;
            if (staged) x = *reinterpret_cast<const double2 *>(Gs + (size_t)((4 * mt + (g >> 1)) * kLLGP + 8 * nt + 2 * t4 + e) * 2);
            else x = make_double2(Sg[sym_idx(slot, c, ld)], Sg[sym_idx(slot + 1, c, ld)]);
            v += l0 * x.x + l1 * x.y;
          }
          y[mt][e] = (c < n) ? v : 0.0;
        }
      }
    }
  };

  double pa[2] = {0.0, 0.0}, pd[2] = {0.0, 0.0}, pm[2] = {0.0, 0.0};   // this thread's columns: running sums over the rows
  double acc[4][2];                                    // Ỹ_J of the coming step
  for (int q = tid; q < r; q += kLLThreads) Hs[q] = Hslot[q];
  __syncthreads();
  issue_gather(0);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  gather(0, acc);
  for (int jb = 0; jb < nblk; ++jb) {
    const int J0 = kCholNb * jb;
    // ---- the diagonal block: wait for block column J of the factor (X_J, ν_J), then W_J = X_J·Ỹ_J --------------------
    if (tid == 0 && !sync_wait(flags + jb)) s_ok = 0;
    if (jb == nblk - 1) timeline_mark(L, 10);
    __syncthreads();                                   // also: the previous step's reads of Xs / T / Wt / Ls / Gs are done
    if (jb + 1 < nblk) issue_gather(jb + 1);           // oldest group of this step: landed whenever a later group is waited for
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + kLLThreads * u, j = e >> 4, q = (e & 15) * 2;
      cp16(Xs + j * kLLP + q, Dinv + (size_t)jb * kCholNb * kCholNb + j * kCholNb + q);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (jb + 1 < nblk) stage_L(jb + 1, jb, 0);
    double nuv[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int row = J0 + 8 * mt + g;
      nuv[mt] = (row < r) ? __ldcg(Sb + (size_t)row * sld + r) : 0.0;
      double2 v = make_double2(acc[mt][0], acc[mt][1]);
      if (row >= r) v = make_double2(0.0, 0.0);        // rows past r: the factor's rows there are not L (ν row, stale data)
      *reinterpret_cast<double2 *>(T + (8 * mt + g) * kLLP + 8 * nt + 2 * t4) = v;
    }
    if (jb + 1 < nblk) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // X_J and Ỹ_J are in shared memory
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
    // the first block below: its accumulated updates, in flight while W_J goes out
    double u[4][2];
    auto load_u = [&](int ib, double (&x)[4][2]) {
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        double2 v = make_double2(0.0, 0.0);
        if (jb > 0 && ib < nblk) v = __ldcg(reinterpret_cast<const double2 *>(Up + (size_t)(kCholNb * ib + 8 * mt + g) * ld));
        x[mt][0] = v.x; x[mt][1] = v.y;
      }
    };
    load_u(jb + 1, u);
    if (jb + 1 < nblk) __syncthreads();                // every warp has read Ỹ_J: the buffer takes W_J
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int row = J0 + 8 * mt + g;                  // < rld; rows r.. of a partial last block come out as zeros
      *reinterpret_cast<double2 *>(Wp + (size_t)row * ld) = make_double2(w[mt][0], w[mt][1]);
      if (jb + 1 < nblk) *reinterpret_cast<double2 *>(Wt + (8 * mt + g) * kLLP + 8 * nt + 2 * t4) = make_double2(w[mt][0], w[mt][1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        pa[e] = fma(w[mt][e], nuv[mt], pa[e]);
        pd[e] = fma(w[mt][e], w[mt][e], pd[e]);
        pm[e] = fmax(pm[e], fabs(w[mt][e]));
      }
    }
    if (jb + 1 == nblk) break;
    __syncthreads();                                   // W_J is in shared memory
    double wb[8];                                      // −W_J: its B fragments for this warp's column tile
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) wb[ks] = -Wt[(4 * ks + t4) * kLLP + 8 * nt + g];
    // ---- U_I −= L_IJ·W_J for the blocks below, the next diagonal block first (it stays in registers) --------------------
    for (int ib = jb + 1; ib < nblk; ++ib) {
      const int sg = (ib - jb - 1) & 1;
      double un[4][2];                                 // the next block's, one iteration ahead
      load_u(ib + 1, un);
      if (ib + 1 < nblk) {
        stage_L(ib + 1, jb, sg ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                                 // L_IJ has landed for every thread
      const double *Lc = Ls + (size_t)sg * kLLBlk;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          dmma884(u[mt][0], u[mt][1], Lc[(4 * ks + t4) * kLLP + 8 * mt + g], wb[ks], u[mt][0], u[mt][1]);
      }
      if (ib == jb + 1) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) { acc[mt][0] = u[mt][0]; acc[mt][1] = u[mt][1]; }
      } else {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          *reinterpret_cast<double2 *>(Up + (size_t)(kCholNb * ib + 8 * mt + g) * ld) = make_double2(u[mt][0], u[mt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) { u[mt][0] = un[mt][0]; u[mt][1] = un[mt][1]; }
      __syncthreads();                                 // this stage buffer may be refilled
    }
    {                                                  // Ỹ of the next step = its rows of Y + the updates so far
      double y[4][2];
      gather(jb + 1, y);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) { acc[mt][0] += y[mt][0]; acc[mt][1] += y[mt][1]; }
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
    uint32_t *Qs = reinterpret_cast<uint32_t *>(sm_d);  // [4 planes][32 columns][kLLQP] (the staging buffers are dead)
    const bool live = (c0 + lane < n);
    const double sc = __longlong_as_double((long long)(1023 + 28 - sexp[lane]) << 52);  // 2^(28-e)
    double wv[4][4], wn[4][4];
    auto load_w = [&](int ch, double (&x)[4][4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int k = 64 * ch + 4 * (nt + 4 * u) + v;
          x[u][v] = (live && k < r) ? __ldcg(Wg + (size_t)k * ld + lane) : 0.0;
        }
    };
    load_w(0, wv);
    for (int ch = 0; ch < (L.kq >> 6); ++ch) {
      load_w(ch + 1, wn);                              // rows past r read as zeros: no bound on ch + 1 needed
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
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) wv[u][v] = wn[u][v];
      __syncthreads();
    }
  }
  timeline_mark(L, 11);
}

}  // namespace rekf
