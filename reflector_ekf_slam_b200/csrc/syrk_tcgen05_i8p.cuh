// syrk_tcgen05_i8p.cuh — persistent, fully overlapped form of the exact int8-slice covariance SYRK
// (Σ −= Wᵀ·W, reflector_ekf_slam.cc:308; see syrk_tcgen05_i8.cuh for the digit-slice scheme).  The kernel is bound by
// data movement, so the design is about bytes:
//
//   * only the UPPER triangle of Σ exists (rekf_device.cuh): 128x64 tiles on/above the diagonal, no mirrored stores;
//   * the SM never READS Σ.  The epilogue turns the s32 accumulators into the fp64 downdate −G·2^(e_i+e_j) (an exact
//     product: G is an integer below 2^51, the scales are powers of two), writes it into a shared-memory half-tile and
//     the store warp hands it to the TMA as cp.reduce.async.bulk.tensor .add.f64: the read-modify-write of Σ happens in
//     the L2 (one correctly rounded fp64 add per element — the same bits as an FMA on the SM, since the product is
//     exact).  There is no HBM load latency on the SM's critical path and no load ring; measured alone
//     (scripts/probe_tma_reduce.cu) 148 CTAs reduce-add the upper triangles of 4 C3 sessions in 26 µs, the same as an
//     ideal TMA load + store pipeline with nothing else in the loop;
//   * one CTA per SM pulls (session, tile) work items from a device-wide atomic queue (the operand producer fetches
//     one item ahead and publishes it to the other roles through a 4-entry shared-memory ring), so a CTA that starts
//     late — its SM was still running another pipeline group's Cholesky — simply takes fewer tiles;
//   * the s32 accumulators are double-buffered in TMEM (2 x 4 x 64 columns = all 512), so the tensor pipe works on
//     tile t+1 while the epilogue warps drain tile t;
//   * warp roles: 0-15 epilogue, 16 operand TMA producer (one 5-D box per operand per stage: 4 digit slices x rows x
//     64 K-bytes, contiguous in the chunk-tiled Wq layout), 17 MMA issuer (tcgen05.mma.kind::i8), 18 column
//     scales/flags of the tile, 19 Σ reduce-add issuer.
// Tiles that touch the diagonal (34 of 306 at C3) keep a direct global-memory epilogue: element predicates (i <= j) and
// the exact fp64 diagonal from k_solve_w3.
#pragma once
#include "syrk_tcgen05_i8.cuh"

namespace rekf {

constexpr int kPEpiWarps = 16;                           // epilogue warps: warp w owns TMEM lanes 32·(w%4).. and 8 of each half-tile's 32 columns
constexpr int kPThreads = (kPEpiWarps + 4) * 32;         // + operand TMA, MMA, Σ load, Σ store
constexpr int kPSigHalf = 128 * 32 * 8;                  // one half-tile of Σ: 128 rows x 32 columns fp64 = 32 KB
constexpr int kPSigSlots = 4;                            // downdate half-tile ring (two whole tiles) between the epilogue and the TMA reduce
constexpr int kPMaxSess = 32;                           // sessions whose (r, n) are cached in shared memory
constexpr int kPQ = 4;                                   // work-item ring entries
// Operand ring: K = 64 per stage (two MMA k-steps, 64-byte swizzle), two stages.  A 4-stage ring of 32-K boxes (32-byte
// swizzle) was tried when the epilogue's top stall was the wait for acc_full: no gain — the kernel is bound by L2 sector
// throughput (operand re-reads), not by operand latency.
constexpr int kPKBox = 64;
constexpr int kPBoxA = 128 * kPKBox;                     // 8 KB
constexpr int kPBoxB = kI8TileN * kPKBox;                // 4 KB
constexpr int kPStageBytes = kI8Slices * (kPBoxA + kPBoxB);   // 48 KB
constexpr int kPStages = 128 / kPKBox;
// operand ring + Σ ring + alignment slack + barriers + [2][64] column scales + [2][64] column flags + session table
constexpr int kPSmemBytes = kPStages * kPStageBytes + kPSigSlots * kPSigHalf + 1024 + 512 + 2 * 64 * 8 + 2 * 64 + kPMaxSess * 8 + 64;
struct SyrkI8P {
  CUtensorMap map_a, map_b, map_sig;
  int num_sms = 148;
  int reserve_sms = 0;          // SMs left to the other pipeline groups' latency-bound kernels
  bool ready = false;
};

// Σ[box] += smem tile, performed by the L2 (SASS: UTMAREDG.3D.ADD); the element type (fp64) comes from the tensor map
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// waits of the single-thread roles back off between probes so that they do not steal issue slots from the epilogue
__device__ __forceinline__ bool mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0; it < (kSpinLimit >> 4); ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return true;
    __nanosleep(64);
  }
  return false;
}
// exact int64 → double for |g| < 2^51 without the slow I2F.F64.S64: add to the bits of 2^52+2^51, subtract it back
__device__ __forceinline__ double i64_to_f64(long long g) {
  return __longlong_as_double(g + 0x4338000000000000LL) - 6755399441055744.0;
}

__global__ void __launch_bounds__(kPThreads, 1)
k_syrk_tcgen05_i8p(Layout L, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_sig) {
  timeline_mark(L, 6);
#ifdef REKF_SYRK_TIMING
  // per-CTA stamps (thread 0 = an epilogue thread): start, after setup, first tile done, last tile done, exit, tiles
  auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (double)t; };
  double *tlog = L.innov + (size_t)blockIdx.x * 6;
  if (threadIdx.x == 0) { tlog[0] = gtime(); tlog[2] = 0; tlog[3] = 0; tlog[5] = 0; }
#endif
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *ops = base;                                   // [2][48 KB] int8 slice boxes
  uint8_t *sig = base + kPStages * kPStageBytes;         // [4][32 KB] downdate half-tiles on their way to the TMA reduce
  uint64_t *bars = reinterpret_cast<uint64_t *>(sig + kPSigSlots * kPSigHalf);
  uint64_t *op_full = bars, *op_empty = bars + 4, *acc_full = bars + 8, *acc_empty = bars + 10,
           *sig_empty = bars + 16, *sig_done = bars + 20;
  uint64_t *sc_full = bars + 24;
  uint64_t *q_full = bars + 26, *q_empty = bars + 30;                                      // work-item ring
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 34);
  double *sc_tab = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(bars) + 512);   // [2][64] Wscale of the tile's columns
  unsigned char *fl_tab = reinterpret_cast<unsigned char *>(sc_tab + 2 * 64);             // [2][64] Wflag of the tile's columns
  int *sess_r = reinterpret_cast<int *>(fl_tab + 2 * 64);                                  // [kPMaxSess] r, 0 = nothing to do
  int *sess_n = sess_r + kPMaxSess;                                                        // [kPMaxSess] internal dimension
  volatile int *q_item = sess_n + kPMaxSess;                                               // [kPQ] item index, -1 = no more work
  for (int q = threadIdx.x; q < min(L.Sg, kPMaxSess); q += kPThreads) {
    const SessionState &st = L.st[L.s0 + q];
    sess_r[q] = (st.r > 0 && !st.exact_update) ? st.r : 0;
    sess_n[q] = internal_dim(st.N);
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tn64 = L.ld / kI8TileN;
  const int tiles = (L.ld / 128) * (L.ld / 128 + 1);
  const int total = tiles * L.Sg;

  if (warp == kPEpiWarps) {
    if (lane == 0) {
      for (int i = 0; i < kPStages; ++i) { mbar_init(&op_full[i], 1); mbar_init(&op_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kPEpiWarps * 32); mbar_init(&sc_full[i], 32); }
      for (int i = 0; i < kPSigSlots; ++i) {
        mbar_init(&sig_empty[i], 1); mbar_init(&sig_done[i], kPEpiWarps * 32);
      }
      for (int i = 0; i < kPQ; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], kPEpiWarps + 3); }   // consumer warps
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  bool timeout = false;
#ifdef REKF_SYRK_TIMING
  if (threadIdx.x == 0) tlog[1] = gtime();
#endif

  // item → (session, tile); false = nothing to do for it (the fetcher skips those, the other roles never see them)
  auto decode = [&](int item, int &s, int &i0, int &j0, int &r, int &n, bool &inA) -> bool {
    // queue order: every session's diagonal tiles first (direct global-memory epilogue: the slow ones), then the
    // off-diagonal tiles — the tail of the launch then consists of uniform, fast tiles
    const int ndiag = Tn64 * L.Sg;                       // two diagonal 128x64 tiles per 128-row block
    int sl, rem, ti = 0;
    if (item < ndiag) {
      sl = item / Tn64;
      const int d = item - sl * Tn64;
      ti = d >> 1; rem = d & 1;
    } else {
      const int noff = tiles - Tn64, it2 = item - ndiag;
      sl = it2 / noff;
      rem = it2 - sl * noff;
      while (rem >= Tn64 - 2 * ti - 2) { rem -= Tn64 - 2 * ti - 2; ++ti; }
      rem += 2;
    }
    s = L.s0 + sl;
    const int tj = 2 * ti + rem;
    i0 = ti * 128; j0 = tj * kI8TileN;
    inA = rem < 2;
    if (sl < kPMaxSess) {
      r = sess_r[sl]; n = sess_n[sl];
    } else {
      const SessionState &st = L.st[s];
      r = (st.r > 0 && !st.exact_update) ? st.r : 0; n = internal_dim(st.N);
    }
    return r > 0 && j0 < n;
  };
  // consumer side of the work-item ring: entry q of the sequence, -1 = the queue is drained
  auto next_item = [&](uint32_t q, bool whole_warp) -> int {
    const int slot = q & (kPQ - 1);
    if (!mbar_wait_backoff(&q_full[slot], (q / kPQ) & 1)) { timeout = true; return -1; }
    const int item = q_item[slot];
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(&q_empty[slot]);
    return item;
  };

  if (warp == kPEpiWarps) {
    // ===== operand TMA producer =====
    if (lane == 0) {
      uint32_t kbc = 0;
      int nxt = atomicAdd(L.tile_counter, 1);              // fetched one item ahead: the round trip hides behind the loads
      for (uint32_t q = 0; !timeout; ++q) {
        int s, i0, j0, r, n; bool inA;
        int item = nxt;
        while (item < total && !decode(item, s, i0, j0, r, n, inA)) item = atomicAdd(L.tile_counter, 1);
        const int slot = q & (kPQ - 1);
        if (!mbar_wait_backoff(&q_empty[slot], ((q / kPQ) & 1) ^ 1)) { timeout = true; break; }
        q_item[slot] = item < total ? item : -1;
        mbar_arrive(&q_full[slot]);
        if (item >= total) break;
        nxt = atomicAdd(L.tile_counter, 1);
        const int nkb = (r + kPKBox - 1) / kPKBox;
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int stage = kbc & (kPStages - 1);
          if (!mbar_wait_backoff(&op_empty[stage], ((kbc / kPStages) & 1) ^ 1)) { timeout = true; break; }
          uint8_t *sa = ops + (size_t)stage * kPStageBytes, *sb = sa + kI8Slices * kPBoxA;
          mbar_expect_tx(&op_full[stage], kI8Slices * (kPBoxA + (inA ? 0 : kPBoxB)));
          // one box per operand: all four digit slices ride in the box's third dimension (a single thread issues
          // these, and eight small boxes per stage made the issue rate the bottleneck)
          tma_load_5d(sa, &map_a, &op_full[stage], 0, i0, kb, 0, s);
          if (!inA) tma_load_5d(sb, &map_b, &op_full[stage], 0, j0, kb, 0, s);
        }
      }
    }
  } else if (warp == kPEpiWarps + 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t kbc = 0, iter = 0;
      for (uint32_t q = 0; !timeout; ++q) {
        const int item = next_item(q, false);
        if (item < 0) break;
        int s, i0, j0, r, n; bool inA;
        decode(item, s, i0, j0, r, n, inA);
        const int set = iter & 1;
        if (!mbar_wait_backoff(&acc_empty[set], ((iter >> 1) & 1) ^ 1)) { timeout = true; break; }
        tc_fence_after();
        const uint32_t acc = tmem + set * 256;
        const int nkb = (r + kPKBox - 1) / kPKBox;
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int stage = kbc & (kPStages - 1);
          if (!mbar_wait_backoff(&op_full[stage], (kbc / kPStages) & 1)) { timeout = true; break; }
          tc_fence_after();
          const uint32_t sa = smem_u32(ops + (size_t)stage * kPStageBytes);
          const uint32_t sb = inA ? sa + (uint32_t)(j0 - i0) * kPKBox : sa + kI8Slices * kPBoxA;
          const uint32_t bstride = inA ? kPBoxA : kPBoxB;
          const int steps = min(kPKBox / 32, (r + 31) / 32 - kb * (kPKBox / 32));
          for (int ks = 0; ks < steps; ++ks) {
            const uint32_t koff = ks * 32;
            const bool first = (kb | ks) == 0;
#pragma unroll
            for (int sgrp = 0; sgrp < kI8Slices; ++sgrp) {
#pragma unroll
              for (int p = 0; p <= sgrp; ++p) {
                const int q = sgrp - p;
                const uint32_t aa = sa + p * kPBoxA + koff, bb = sb + q * bstride + koff;
                tc_mma_i8(acc + sgrp * kI8TileN, make_kmajor_sw64_desc(aa), make_kmajor_sw64_desc(bb), kIdescI8, (first && p == 0) ? 0u : 1u);
              }
            }
          }
          tc_commit(&op_empty[stage]);
        }
        tc_commit(&acc_full[set]);
        ++iter;
      }
    }
  } else if (warp == kPEpiWarps + 2) {
    // ===== column scales / flags of every tile → shared memory (whole warp).  The slot is the accumulator set's: free
    //       once the epilogue released that set.  Every lane arrives for its own two entries. =====
    uint32_t iter = 0;
    for (uint32_t q = 0; !timeout; ++q) {
      const int item = next_item(q, true);
      if (item < 0) break;
      int s, i0, j0, r, n; bool inA;
      decode(item, s, i0, j0, r, n, inA);
      const int set = iter & 1;
      if (!mbar_wait_backoff(&acc_empty[set], ((iter >> 1) & 1) ^ 1)) timeout = true;
      {
        const size_t off = (size_t)s * L.ld + j0 + 2 * lane;
        const double2 v = *reinterpret_cast<const double2 *>(L.Wscale + off);
        const unsigned short f = *reinterpret_cast<const unsigned short *>(L.Wflag + off);
        *reinterpret_cast<double2 *>(sc_tab + set * 64 + 2 * lane) = v;
        *reinterpret_cast<unsigned short *>(fl_tab + set * 64 + 2 * lane) = f;
      }
      mbar_arrive(&sc_full[set]);
      timeout = __any_sync(0xffffffffu, timeout);
      ++iter;
    }
  } else if (warp == kPEpiWarps + 3) {
    // ===== Σ reduce-add issuer: waits until the 512 epilogue threads have written a downdate half-tile, hands it to the
    //       TMA (the L2 adds it into Σ), and frees the slot one group later, when the TMA has read it out =====
    if (lane == 0) {
      uint32_t sit = 0;
      int pending = -1;                                    // slot whose reduce has been issued but not yet released
      for (uint32_t q = 0; !timeout; ++q) {
        const int item = next_item(q, false);
        if (item < 0) break;
        int s, i0, j0, r, n; bool inA;
        decode(item, s, i0, j0, r, n, inA);
        if (inA) {
          if (pending >= 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&sig_empty[pending]);
            pending = -1;
          }
          continue;
        }
        for (int h = 0; h < 2; ++h) {
          const int slot = ((sit & 1) << 1) | h;
          if (!mbar_wait_backoff(&sig_done[slot], (sit >> 1) & 1)) { timeout = true; break; }
          const uint8_t *src = sig + (size_t)slot * kPSigHalf;
          tma_reduce_add_3d(&map_sig, src, j0 + 32 * h, i0, s);
          tma_reduce_add_3d(&map_sig, src + kPSigHalf / 2, j0 + 32 * h + 16, i0, s);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (pending >= 0) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // every group but the one just issued
            mbar_arrive(&sig_empty[pending]);
          }
          pending = slot;
        }
        ++sit;
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all reductions performed
    }
  } else if (warp < kPEpiWarps) {
    // ===== epilogue: warp w owns TMEM lanes 32·(w%4).. ; within a 32-column half-tile, columns 8·(w/4).. .  Sixteen warps
    //       (four per scheduler) because the per-element chain — TMEM load, integer recombination, int→fp64, scale, FMA —
    //       is latency-bound: with eight warps the epilogue, not HBM, set the tile period =====
    const int quad = warp & 3, cgp = warp >> 2;
    const int il = quad * 32 + lane;                     // row inside the tile
    uint32_t iter = 0, sit = 0;
    const int ld = L.ld;
    int row_s = -1, row_i0 = -1;                         // row scale / flag are reloaded only when the row block changes
    double si = 0.0;
    bool row_ok = false;
    for (uint32_t q = 0;; ++q) {
      const int item = next_item(q, true);
      if (item < 0) break;
      int s, i0, j0, r, n; bool inA;
      decode(item, s, i0, j0, r, n, inA);
      const int set = iter & 1;
      const int i = i0 + il;
      double *Sg = L.sigma + (size_t)s * ld * ld;
      if (s != row_s || i0 != row_i0) {
        si = L.Wscale[(size_t)s * ld + i] * 0x1p-35;
        row_ok = !L.Wflag[(size_t)s * ld + i];
        row_s = s; row_i0 = i0;
      }
      const double *sct = sc_tab + set * 64;
      const unsigned char *flt = fl_tab + set * 64;
      if (!mbar_wait(&sc_full[set], (iter >> 1) & 1)) timeout = true;
      if (!mbar_wait(&acc_full[set], (iter >> 1) & 1)) timeout = true;
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int col0 = 32 * h + 8 * cgp;               // first of this thread's 8 tile columns
        const int jbase = j0 + col0;
        const uint32_t taddr = tmem + set * 256 + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
        // s32 accumulators of the four digit groups, all four loads in flight; groups are recombined pairwise in 32 bits
        // (|acc| < 2^23 for K <= 512, so acc·2^7 + acc' fits), then once in 64 bits: G = Σ_g acc_g · 2^(7·(3-g))
        long long G[8];
        {
          uint32_t g0[8], g1[8], g2[8], g3[8];
          tc_ld8(taddr, g0);
          tc_ld8(taddr + kI8TileN, g1);
          tc_ld8(taddr + 2 * kI8TileN, g2);
          tc_ld8(taddr + 3 * kI8TileN, g3);
          tc_wait_ld();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            G[u] = ((long long)(((int)g0[u] << 7) + (int)g1[u]) << 14) + (long long)(((int)g2[u] << 7) + (int)g3[u]);
        }
        const uint2 cf = *reinterpret_cast<const uint2 *>(flt + col0);
        const double2 *scj = reinterpret_cast<const double2 *>(sct + col0);
        double cur[8];
        if (!inA) {
          // ---- downdate half-tile: −G·2^(e_i+e_j) (exact) into the swizzled staging slot; the TMA reduce adds it into Σ ----
          const int slot = ((sit & 1) << 1) | h;
          if (!mbar_wait(&sig_empty[slot], ((sit >> 1) & 1) ^ 1)) timeout = true;   // the previous reduce has read it out
          uint8_t *rowp = sig + (size_t)slot * kPSigHalf + (size_t)(cgp >> 1) * (kPSigHalf / 2) + (size_t)il * 128;
          const int ch0 = 4 * (cgp & 1);                  // this thread's four 16-byte chunks of the 128-byte row
          const bool no_flags = row_ok && (cf.x | cf.y) == 0u;   // the common case, warp-uniform but for row_ok
          if (no_flags) {
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const double2 sj = scj[u >> 1];
              cur[u] = -i64_to_f64(G[u]) * (si * sj.x);
              cur[u + 1] = -i64_to_f64(G[u + 1]) * (si * sj.y);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const unsigned cfw = (u < 4) ? cf.x : cf.y;
              const bool skip = !row_ok || ((cfw >> (8 * (u & 3))) & 0xffu);   // flagged slots: k_syrk_exact_rows did them
              cur[u] = skip ? 0.0 : -i64_to_f64(G[u]) * (si * sct[col0 + u]);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<double2 *>(rowp + (((ch0 + c) ^ (il & 7)) << 4)) = make_double2(cur[2 * c], cur[2 * c + 1]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&sig_done[slot]);                     // hand the half-tile to the reduce issuer
        } else {
          // ---- diagonal tile: direct global accesses, element predicates (i <= j), exact diagonal ----
          const bool want = row_ok && i < n && jbase < n && !(jbase + 7 < i);
          if (want) {
            double *row = Sg + (size_t)i * ld + jbase;
            ldg256(row, cur);
            ldg256(row + 4, cur + 4);
            const int ud = i - jbase;
            double old_diag = 0.0;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const double2 sj = scj[u >> 1];
              if (u == ud) old_diag = cur[u];
              if (u + 1 == ud) old_diag = cur[u + 1];
              cur[u] = fma(-i64_to_f64(G[u]), si * sj.x, cur[u]);
              cur[u + 1] = fma(-i64_to_f64(G[u + 1]), si * sj.y, cur[u + 1]);
            }
            if (ud >= 0 && ud < 8) {
              const double dd = old_diag - L.Wdiag[(size_t)s * ld + i];
#pragma unroll
              for (int u = 0; u < 8; ++u) if (u == ud) cur[u] = dd;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = jbase + u;
              const unsigned cfw = (u < 4) ? cf.x : cf.y;
              if (j < n && i <= j && !((cfw >> (8 * (u & 3))) & 0xffu)) row[u] = cur[u];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[set]);                      // this thread is done with the accumulator set
#ifdef REKF_SYRK_TIMING
      if (threadIdx.x == 0) { const double t = gtime(); if (iter == 0) tlog[2] = t; tlog[3] = t; tlog[5] += 1; }
#endif
      ++iter;
      if (!inA) ++sit;
    }
  }
  if (timeout) atomicOr(&L.st[L.s0].flags, FLAG_TCGEN05_TIMEOUT);
  __syncthreads();
#ifdef REKF_SYRK_TIMING
  if (threadIdx.x == 0) tlog[4] = gtime();
#endif
  if (warp == kPEpiWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

inline const char *syrk_i8p_init(SyrkI8P &tc, const Layout &L) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return "cuTensorMapEncodeTiled entry point not available";
  PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(fn);
  // Wq is [S][4 slices][kq/64 chunks][ld rows][64 bytes]: a box = 64 K-bytes x rows x 1 chunk x 4 slices, each slice's part contiguous
  static_assert(kPKBox == 64, "the Wq layout is tiled in 64-byte K chunks");
  const cuuint64_t nch = (cuuint64_t)(L.kq / 64);
  const cuuint64_t dims[5] = {64u, (cuuint64_t)L.ld, nch, (cuuint64_t)kI8Slices, (cuuint64_t)L.S};
  const cuuint64_t strides[4] = {64u, (cuuint64_t)L.ld * 64, nch * L.ld * 64, (cuuint64_t)kI8Slices * nch * L.ld * 64};
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const cuuint32_t box_a[5] = {64u, 128u, 1u, (cuuint32_t)kI8Slices, 1u};
  const cuuint32_t box_b[5] = {64u, (cuuint32_t)kI8TileN, 1u, (cuuint32_t)kI8Slices, 1u};
  if (encode(&tc.map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, L.Wq, dims, strides, box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wq, A box) failed";
  if (encode(&tc.map_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, L.Wq, dims, strides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wq, B box) failed";
  const cuuint64_t sdims[3] = {(cuuint64_t)L.ld, (cuuint64_t)L.ld, (cuuint64_t)L.S};
  const cuuint64_t sstrides[2] = {(cuuint64_t)L.ld * sizeof(double), (cuuint64_t)L.ld * L.ld * sizeof(double)};
  const cuuint32_t sbox[3] = {16u, 128u, 1u};
  const cuuint32_t sestr[3] = {1u, 1u, 1u};
  if (encode(&tc.map_sig, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, L.sigma, sdims, sstrides, sbox, sestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Sigma) failed";
  if (cudaFuncSetAttribute(k_syrk_tcgen05_i8p, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemBytes) != cudaSuccess)
    return "cudaFuncSetAttribute(k_syrk_tcgen05_i8p, smem) failed";
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&tc.num_sms, cudaDevAttrMultiProcessorCount, dev);
  tc.ready = true;
  return nullptr;
}

inline int syrk_i8p_launch(const SyrkI8P &tc, const Layout &L, cudaStream_t stream) {
  if (!tc.ready) return -1;
  const int total = (L.ld / 128) * (L.ld / 128 + 1) * L.Sg;
  const int ctas = tc.num_sms - tc.reserve_sms > 1 ? tc.num_sms - tc.reserve_sms : 1;
  const int grid = total < ctas ? total : ctas;
  k_syrk_tcgen05_i8p<<<grid, kPThreads, kPSmemBytes, stream>>>(L, tc.map_a, tc.map_b, tc.map_sig);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace rekf
