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
//   * operand reuse: work is handed out as STRIPS — up to kPMaxStrip consecutive 128x64 tiles of one 128-row block.
//     The row block's A panel (4 digit slices x 128 rows x K <= 256 bytes = 128 KB) is loaded once per strip and stays
//     resident in shared memory; only the 64-row B boxes (16 KB per 64 K-bytes, 4-stage ring) stream per tile.  Per
//     tile the SM pulls 64 KB (+128 KB / strip length) of int8 panels from L2 instead of 192 KB — the round-1 kernel
//     was bound by exactly that L2→SM stream.  (K > 256 bytes, config C4: the panel does not fit; both operands stream
//     through a 4-stage ring — template parameter kRes.)
//   * one CTA per SM pulls strips from a device-wide atomic cursor over the linearised tile order with guided
//     self-scheduling (strip length = remaining / (2 x CTAs), clamped to [1, kPMaxStrip]): long strips while there is
//     plenty of work, single tiles at the end, so the launch has no tail; a CTA that starts late — its SM was still
//     running another pipeline group's Cholesky — simply takes fewer.  The operand producer fetches one span ahead and
//     publishes tiles to the other roles through a 4-entry shared-memory ring;
//   * the s32 accumulators are double-buffered in TMEM (2 x 4 x 64 columns = all 512), so the tensor pipe works on
//     tile t+1 while the epilogue warps drain tile t;
//   * warp roles: 0-15 epilogue (16 columns = one TMA box at a time, double-buffered), 16 operand TMA producer (one
//     5-D box per operand per 64 K-bytes: 4 digit slices x rows x 64 B, contiguous in the chunk-tiled Wq layout),
//     17 MMA issuer (tcgen05.mma.kind::i8), 18 column scales/flags of the tile, 19 Σ reduce-add issuer.
// Tiles that touch the diagonal take the same path: elements below the diagonal add 0, the diagonal itself adds the
// exact fp64 −Σ_k W[k][i]² from k_solve_w3 instead of the truncated digit product.
#pragma once
#include "syrk_tcgen05_i8.cuh"

namespace rekf {

constexpr int kPEpiWarps = 16;                           // epilogue warps: warp w owns TMEM lanes 32·(w%4).. and 4 of a box's 16 columns
constexpr int kPThreads = (kPEpiWarps + 4) * 32;         // + operand TMA, MMA, scales, Σ reduce
constexpr int kPSigBox = 128 * 16 * 8;                   // one downdate box: 128 rows x 16 columns fp64 = 16 KB (128-byte swizzle)
constexpr int kPSigSlots = 2;                            // downdate boxes in flight between the epilogue and the TMA reduce
constexpr int kPMaxSess = 32;                            // sessions whose (r, n) are cached in shared memory
constexpr int kPQ = 4;                                   // tile ring entries
constexpr int kPMaxStrip = 6;                            // longest strip (tiles sharing one resident A panel)
constexpr int kPKBox = 64;                               // K bytes per TMA box (64-byte swizzle, two MMA k-steps)
constexpr int kPBoxA = 128 * kPKBox;                     // 8 KB per slice
constexpr int kPBoxB = kI8TileN * kPKBox;                // 4 KB per slice
constexpr int kPChunkA = kI8Slices * kPBoxA;             // 32 KB: all four slices of 128 rows x 64 K-bytes
constexpr int kPChunkB = kI8Slices * kPBoxB;             // 16 KB
constexpr int kPResChunks = 4;                           // resident A panel: K <= 256 bytes
constexpr int kPResBStages = 4;                          // resident mode: B ring
constexpr int kPStrStages = 4;                           // streaming mode: (A + B) ring, 4 x 48 KB (2 stages left the MMA waiting for operands: C4 203 us)
constexpr int kPTables = 2048;                           // barriers, scale/flag tables, session table, tile ring
constexpr int kPSmemRes = kPResChunks * kPChunkA + kPResBStages * kPChunkB + kPSigSlots * kPSigBox + 1024 + kPTables;   // 227 KB exactly
constexpr int kPSmemStr = kPStrStages * (kPChunkA + kPChunkB) + kPSigSlots * kPSigBox + 1024 + kPTables;
static_assert(kPSmemRes <= 232448, "resident-panel SYRK exceeds the 227 KB shared-memory limit");

struct SyrkI8P {
  CUtensorMap map_a, map_b, map_sig;
  int num_sms = 148;
  int reserve_sms = 0;          // SMs left to the other pipeline groups' latency-bound kernels
  bool resident = true;         // the A panel fits (kq <= 256)
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
    __nanosleep(32);
  }
  return false;
}
// one lane of a converged warp (elect.sync): ptxas then knows a single thread issues the tcgen05/TMA instructions inside and
// moves their operands to uniform registers once, instead of wrapping every instruction in a per-active-thread loop
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
// exact int64 → double for |g| < 2^51 without the slow I2F.F64.S64: add to the bits of 2^52+2^51, subtract it back
__device__ __forceinline__ double i64_to_f64(long long g) {
  return __longlong_as_double(g + 0x4338000000000000LL) - 6755399441055744.0;
}

// -DREKF_SYRK_TIMING: every role's elected thread accumulates the cycles it spends in each wait (scripts/syrk_timing.py);
// per CTA 24 doubles in the (dead by now) S/L buffer of the launch's first session
#ifdef REKF_SYRK_TIMING
#define REKF_T(slot, expr) do { const long long t_ = clock64(); expr; tacc[slot] += (double)(clock64() - t_); } while (0)
#else
#define REKF_T(slot, expr) do { expr; } while (0)
#endif

template <bool kRes>
__global__ void __launch_bounds__(kPThreads, 1)
k_syrk_tcgen05_i8p(Layout L, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_sig) {
  timeline_mark(L, 6);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // resident: [4 chunks][32 KB] A panel, then [4][16 KB] B ring.  streaming: [2][48 KB] (A chunk | B chunk) ring.
  uint8_t *opsA = base;
  uint8_t *opsB = base + (kRes ? kPResChunks * kPChunkA : 0);
  constexpr int kOpsBytes = kRes ? kPResChunks * kPChunkA + kPResBStages * kPChunkB : kPStrStages * (kPChunkA + kPChunkB);
  uint8_t *sig = base + kOpsBytes;                       // [2][16 KB] downdate boxes on their way to the TMA reduce
  uint64_t *bars = reinterpret_cast<uint64_t *>(sig + kPSigSlots * kPSigBox);
  uint64_t *a_full = bars, *a_empty = bars + 4, *b_full = bars + 8, *b_empty = bars + 12, *acc_full = bars + 16,
           *acc_empty = bars + 18, *sig_empty = bars + 20, *sig_done = bars + 22, *sc_full = bars + 24, *q_full = bars + 26,
           *q_empty = bars + 30;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 34);
  double *sc_tab = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(bars) + 512);   // [2][64] Wscale of the tile's columns
  unsigned char *fl_tab = reinterpret_cast<unsigned char *>(sc_tab + 2 * 64);             // [2][64] Wflag of the tile's columns
  int *sess_r = reinterpret_cast<int *>(fl_tab + 2 * 64);                                  // [kPMaxSess] r, 0 = nothing to do
  int *sess_n = sess_r + kPMaxSess;                                                        // [kPMaxSess] internal dimension
  volatile int4 *q_tile = reinterpret_cast<volatile int4 *>(sess_n + kPMaxSess);           // [kPQ] {session, i0, j0, flags}; session < 0: drained
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tn64 = L.ld / kI8TileN;
  const int tiles = (L.ld / 128) * (L.ld / 128 + 1);     // 128x64 tiles on/above the diagonal, per session
  const int total = tiles * L.Sg;

  if (warp == kPEpiWarps) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kPEpiWarps * 32); mbar_init(&sc_full[i], 32); }
      for (int i = 0; i < kPSigSlots; ++i) { mbar_init(&sig_empty[i], 1); mbar_init(&sig_done[i], kPEpiWarps); }
      for (int i = 0; i < kPQ; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], kPEpiWarps + 3); }   // consumer warps
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // barrier init and the TMEM allocation above overlap the tail of the kernel in front (programmatic dependent launch);
  // everything below reads what the chain produced
  pdl_trigger();
  pdl_wait();
  for (int q = threadIdx.x; q < min(L.Sg, kPMaxSess); q += kPThreads) {
    const SessionState &st = L.st[L.s0 + q];
    sess_r[q] = (st.r > 0 && !st.exact_update) ? st.r : 0;
    sess_n[q] = internal_dim(st.N);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  bool timeout = false;
#ifdef REKF_SYRK_TIMING
  double tacc[6] = {0, 0, 0, 0, 0, 0};
  double *tlog = L.Sbuf + (size_t)L.s0 * L.rld * L.sld + (size_t)blockIdx.x * 24;
  const long long t_begin = clock64();
#endif

  // session-local r / n (0: the session has nothing for this kernel)
  auto sess_rn = [&](int sl, int &r, int &n) {
    if (sl < kPMaxSess) { r = sess_r[sl]; n = sess_n[sl]; }
    else { const SessionState &st = L.st[L.s0 + sl]; r = (st.r > 0 && !st.exact_update) ? st.r : 0; n = internal_dim(st.N); }
  };
  // consumer side of the tile ring
  auto next_tile = [&](uint32_t q, bool whole_warp, int &s, int &i0, int &j0, int &flags) -> bool {
    const int slot = q & (kPQ - 1);
    bool ok;
    REKF_T(0, ok = mbar_wait_backoff(&q_full[slot], (q / kPQ) & 1));
    if (!ok) { timeout = true; return false; }
    s = q_tile[slot].x; i0 = q_tile[slot].y; j0 = q_tile[slot].z; flags = q_tile[slot].w;
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(&q_empty[slot]);
    return s >= 0;
  };

  if (warp == kPEpiWarps) {
    // ===== operand TMA producer + work distribution =====
    if (lane == 0) {
      // guided self-scheduling over the linear tile order (session-major, row-block-major inside a session)
      // The first span of every CTA is static (CTA b takes tiles [b·w0, (b+1)·w0): no atomic round trip before the first load);
      // the shared cursor (zeroed by the k_syrk_f64 launch in front) counts the tiles handed out after those.
      const int w0 = min(kPMaxStrip, max(1, total / (2 * (int)gridDim.x)));
      const int base0 = w0 * (int)gridDim.x;
      auto grab = [&](int &start, int &cnt) {
        const int seen = base0 + *reinterpret_cast<volatile int *>(L.tile_counter);
        const int rem = total - seen;
        if (rem <= 0) { start = total; cnt = 0; return; }
        const int want = min(kPMaxStrip, max(1, rem / (2 * (int)gridDim.x)));
        start = base0 + atomicAdd(L.tile_counter, want);
        cnt = start < total ? min(want, total - start) : 0;
      };
      uint32_t qn = 0, bcnt = 0, acnt[4] = {0u, 0u, 0u, 0u};
      auto publish = [&](int s, int i0, int j0, int flags) -> bool {
        const int slot = qn & (kPQ - 1);
        bool ok;
        REKF_T(1, ok = mbar_wait_backoff(&q_empty[slot], ((qn / kPQ) & 1) ^ 1));
        if (!ok) { timeout = true; return false; }
        q_tile[slot].x = s; q_tile[slot].y = i0; q_tile[slot].z = j0; q_tile[slot].w = flags;
        mbar_arrive(&q_full[slot]);
        ++qn;
        return true;
      };
      int start = w0 * (int)blockIdx.x, cnt = start < total ? min(w0, total - start) : 0;
      if (cnt == 0) grab(start, cnt);
      while (cnt > 0 && !timeout) {
        int ts[kPMaxStrip], ti0[kPMaxStrip], tj0[kPMaxStrip], tr[kPMaxStrip], nt = 0;
        for (int l = start; l < start + cnt; ++l) {
          const int sl = l / tiles;
          int t = l - sl * tiles, ib = 0;
          while (t >= Tn64 - 2 * ib) { t -= Tn64 - 2 * ib; ++ib; }
          int r, n;
          sess_rn(sl, r, n);
          const int j0 = (2 * ib + t) * kI8TileN;
          if (r > 0 && j0 < n) { ts[nt] = L.s0 + sl; ti0[nt] = ib * 128; tj0[nt] = j0; tr[nt] = r; ++nt; }
        }
        bool grabbed = false;
        for (int k = 0; k < nt && !timeout; ++k) {
          const bool first = k == 0 || ts[k] != ts[k - 1] || ti0[k] != ti0[k - 1];
          const bool last = k == nt - 1 || ts[k + 1] != ts[k] || ti0[k + 1] != ti0[k];
          const int s = ts[k], i0 = ti0[k], j0 = tj0[k];
          if (!publish(s, i0, j0, (first ? 1 : 0) | (last ? 2 : 0))) break;
          const int nkb = (tr[k] + kPKBox - 1) / kPKBox;
          for (int kb = 0; kb < nkb; ++kb) {
            if (kRes) {
              if (first) {                                 // this strip's A panel, chunk by chunk (each chunk has its own barriers)
                bool ok;
                REKF_T(2, ok = mbar_wait_backoff(&a_empty[kb], (acnt[kb] & 1) ^ 1));
                if (!ok) { timeout = true; break; }
                mbar_expect_tx(&a_full[kb], kPChunkA);
                tma_load_5d(opsA + (size_t)kb * kPChunkA, &map_a, &a_full[kb], 0, i0, kb, 0, s);
                ++acnt[kb];
              }
              {
                const int stage = bcnt & (kPResBStages - 1);
                bool ok;
                REKF_T(3, ok = mbar_wait_backoff(&b_empty[stage], ((bcnt / kPResBStages) & 1) ^ 1));
                if (!ok) { timeout = true; break; }
                mbar_expect_tx(&b_full[stage], kPChunkB);
                tma_load_5d(opsB + (size_t)stage * kPChunkB, &map_b, &b_full[stage], 0, j0, kb, 0, s);
                ++bcnt;
              }
            } else {
              const int stage = bcnt & (kPStrStages - 1);
              if (!mbar_wait_backoff(&b_empty[stage], ((bcnt / kPStrStages) & 1) ^ 1)) { timeout = true; break; }
              uint8_t *sa = opsA + (size_t)stage * (kPChunkA + kPChunkB);
              mbar_expect_tx(&b_full[stage], kPChunkA + kPChunkB);
              tma_load_5d(sa, &map_a, &b_full[stage], 0, i0, kb, 0, s);
              tma_load_5d(sa + kPChunkA, &map_b, &b_full[stage], 0, j0, kb, 0, s);
              ++bcnt;
            }
          }
          if (!grabbed) { grab(start, cnt); grabbed = true; }   // one span ahead, behind this span's first loads: the atomic's round trip is hidden
        }
        if (!grabbed) grab(start, cnt);
      }
      if (!timeout) publish(-1, 0, 0, 0);
#ifdef REKF_SYRK_TIMING
      tlog[0] = tacc[1]; tlog[1] = tacc[2]; tlog[2] = tacc[3]; tlog[3] = (double)(clock64() - t_begin); tlog[4] = (double)qn;
#endif
    }
  } else if (warp == kPEpiWarps + 1) {
    // ===== MMA issuer: the whole warp walks the (warp-uniform) loop, one elected lane issues the MMAs and commits.  With the
    //       loop under `if (lane == 0)` ptxas wrapped each of the 70 MMAs of a tile in a per-active-thread loop of five
    //       R2UR moves (~100 cycles per MMA: the issue thread, not the tensor pipe, set the tile period). =====
    uint32_t bcnt = 0, iter = 0, ause[4] = {0u, 0u, 0u, 0u};
    for (uint32_t q = 0; !timeout; ++q) {
      int s, i0, j0, flags;
      if (!next_tile(q, true, s, i0, j0, flags)) break;
      int r, n;
      sess_rn(s - L.s0, r, n);
      const bool first = flags & 1, last = flags & 2;
      const int set = iter & 1;
      {
        bool ok;
        REKF_T(1, ok = mbar_wait_backoff(&acc_empty[set], ((iter >> 1) & 1) ^ 1));
        if (!ok) timeout = true;
      }
      tc_fence_after();
      const uint32_t acc = tmem + set * 256;
      const int nkb = (r + kPKBox - 1) / kPKBox;
      for (int kb = 0; kb < nkb; ++kb) {
        uint32_t sa, sb;
        int stage;
        if (kRes) {
          if (first) {
            bool ok;
            REKF_T(2, ok = mbar_wait_backoff(&a_full[kb], ause[kb] & 1));
            if (!ok) timeout = true;
          }
          sa = smem_u32(opsA + (size_t)kb * kPChunkA);
          stage = bcnt & (kPResBStages - 1);
          bool ok;
          REKF_T(3, ok = mbar_wait_backoff(&b_full[stage], (bcnt / kPResBStages) & 1));
          if (!ok) timeout = true;
          sb = smem_u32(opsB + (size_t)stage * kPChunkB);
        } else {
          stage = bcnt & (kPStrStages - 1);
          if (!mbar_wait_backoff(&b_full[stage], (bcnt / kPStrStages) & 1)) timeout = true;
          sa = smem_u32(opsA + (size_t)stage * (kPChunkA + kPChunkB));
          sb = sa + kPChunkA;
        }
        timeout = __any_sync(0xffffffffu, timeout);
        if (timeout) break;
        tc_fence_after();
        const int steps = min(kPKBox / 32, (r + 31) / 32 - kb * (kPKBox / 32));
        if (elect_one_sync()) {
          // Per k-step FOUR MMAs instead of ten: the B chunk holds the four digit slices back to back ([slice][64 rows][64 B]),
          // i.e. it IS a 256-row K-major operand.  A-slice p times its first 64·(4−p) rows (slices q = 0..3−p) with D starting at
          // accumulator p drops every product d_pᵀ·d_q into accumulator p+q — the same ten 128x64x32 products, but the A
          // slice is fetched from shared memory once per p, not once per (p, q): 36 KB instead of 60 KB of operand reads per
          // k-step.  (With ten N=64 MMAs the tile period was set by shared-memory bandwidth: ~87 cycles per MMA against 32.)
          // Descriptors differ only in the 14-bit start-address field (16-byte units).
          const uint64_t da = make_kmajor_sw64_desc(sa), db = make_kmajor_sw64_desc(sb);
          for (int ks = 0; ks < steps; ++ks) {
            const uint32_t zero = (kb | ks) == 0 ? 0u : 1u;
            const uint64_t dbk = db + (uint64_t)((ks * 32) >> 4);
#pragma unroll
            for (int p = 0; p < kI8Slices; ++p)
              tc_mma_i8(acc + p * kI8TileN, da + (uint64_t)((p * kPBoxA + ks * 32) >> 4), dbk, idesc_i8(kI8TileN * (kI8Slices - p)),
                        p == 0 ? zero : 1u);
          }
          if (kRes) {
            tc_commit(&b_empty[stage]);
            if (last) tc_commit(&a_empty[kb]);             // the panel chunk may be overwritten once these MMAs retire
          } else {
            tc_commit(&b_empty[stage]);
          }
          if (kb == nkb - 1) tc_commit(&acc_full[set]);
        }
        __syncwarp();
        ++bcnt;
        if (kRes && last) ++ause[kb];
      }
      ++iter;
    }
#ifdef REKF_SYRK_TIMING
    if (lane == 0) { tlog[5] = tacc[0]; tlog[6] = tacc[1]; tlog[7] = tacc[2]; tlog[8] = tacc[3]; tlog[9] = (double)(clock64() - t_begin); }
#endif
  } else if (warp == kPEpiWarps + 2) {
    // ===== column scales / flags of every tile → shared memory (whole warp).  The slot is the accumulator set's: free
    //       once the epilogue released that set.  Every lane arrives for its own two entries. =====
    uint32_t iter = 0;
    for (uint32_t q = 0; !timeout; ++q) {
      int s, i0, j0, flags;
      if (!next_tile(q, true, s, i0, j0, flags)) break;
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
    // ===== Σ reduce-add issuer: waits until the epilogue warps have written a downdate box, hands it to the TMA (the L2
    //       adds it into Σ) and frees the slot as soon as the TMA has read it out of shared memory =====
    if (lane == 0) {
      uint32_t u = 0;
      for (uint32_t q = 0; !timeout; ++q) {
        int s, i0, j0, flags;
        if (!next_tile(q, false, s, i0, j0, flags)) break;
        for (int qd = 0; qd < 4; ++qd, ++u) {
          const int slot = u & (kPSigSlots - 1);
          bool ok;
          REKF_T(1, ok = mbar_wait_backoff(&sig_done[slot], (u / kPSigSlots) & 1));
          if (!ok) { timeout = true; break; }
          tma_reduce_add_3d(&map_sig, sig + (size_t)slot * kPSigBox, j0 + 16 * qd, i0, s);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          REKF_T(2, asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"));
          mbar_arrive(&sig_empty[slot]);
        }
      }
      REKF_T(3, asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"));   // all reductions performed
#ifdef REKF_SYRK_TIMING
      tlog[10] = tacc[0]; tlog[11] = tacc[1]; tlog[12] = tacc[2]; tlog[13] = tacc[3]; tlog[14] = (double)(clock64() - t_begin);
#endif
    }
  } else if (warp < kPEpiWarps) {
    // ===== epilogue: warp w owns TMEM lanes 32·(w%4).. and columns 4·(w/4).. of every 16-column box.  Sixteen warps (four per
    //       scheduler) because the per-element chain — TMEM load, integer recombination, int→fp64, scale — is latency-bound =====
    const int quad = warp & 3, cg = warp >> 2;
    const int il = quad * 32 + lane;                     // row inside the tile
    uint32_t iter = 0, u = 0;
    const int ld = L.ld;
    int row_s = -1, row_i0 = -1;                         // row scale / flag / exact diagonal: reloaded when the row block changes
    double si = 0.0, wd = 0.0;
    bool row_ok = false;
    for (uint32_t q = 0;; ++q) {
      int s, i0, j0, flags;
      if (!next_tile(q, true, s, i0, j0, flags)) break;
      const int set = iter & 1;
      const int i = i0 + il;
      const bool inA = j0 < i0 + 128;
      if (s != row_s || i0 != row_i0) {
        si = L.Wscale[(size_t)s * ld + i] * 0x1p-35;
        row_ok = !L.Wflag[(size_t)s * ld + i];
        wd = L.Wdiag[(size_t)s * ld + i];
        row_s = s; row_i0 = i0;
      }
      const double *sct = sc_tab + set * 64;
      const unsigned char *flt = fl_tab + set * 64;
      REKF_T(1, if (!mbar_wait(&sc_full[set], (iter >> 1) & 1)) timeout = true);
      REKF_T(2, if (!mbar_wait(&acc_full[set], (iter >> 1) & 1)) timeout = true);
      tc_fence_after();
      // the TMEM loads of box qd+1 are in flight while box qd is converted, written and handed over (TMEM reads run at
      // 64 B/clk per SM: 512 cycles per box for the four s32 accumulators — the floor of this loop)
      uint32_t g[2][4][4];
      const uint32_t tbase = tmem + set * 256 + ((uint32_t)(quad * 32) << 16) + (uint32_t)(4 * cg);
      auto issue_ld = [&](int qd, int buf) {
#pragma unroll
        for (int a = 0; a < 4; ++a) tc_ld4(tbase + 16 * qd + a * kI8TileN, g[buf][a]);
      };
      issue_ld(0, 0);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd, ++u) {
        const int buf = qd & 1;
        const int col0 = 16 * qd + 4 * cg;               // first of this thread's 4 tile columns
        tc_wait_ld();
        if (qd < 3) issue_ld(qd + 1, buf ^ 1);
        const uint32_t cf = *reinterpret_cast<const uint32_t *>(flt + col0);
        const double2 sj0 = *reinterpret_cast<const double2 *>(sct + col0), sj1 = *reinterpret_cast<const double2 *>(sct + col0 + 2);
        const double sj[4] = {sj0.x, sj0.y, sj1.x, sj1.y};
        double cur[4];
        // groups are recombined pairwise in 32 bits (|acc| < 2^23 for K <= 512, so acc·2^7 + acc' fits), then once in 64 bits:
        // G = Σ_a acc_a · 2^(7·(3-a))
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const long long G = ((long long)(((int)g[buf][0][e] << 7) + (int)g[buf][1][e]) << 14) +
                              (long long)(((int)g[buf][2][e] << 7) + (int)g[buf][3][e]);
          cur[e] = -i64_to_f64(G) * (si * sj[e]);          // exact: an integer below 2^51 times a power of two
        }
        if (!row_ok || cf != 0u || inA) {                  // rare: flagged slots (k_syrk_exact_rows did them), diagonal tiles
          const int jb = j0 + col0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool skip = !row_ok || ((cf >> (8 * e)) & 0xffu) || i > jb + e;
            if (i == jb + e) cur[e] = -wd;                 // the diagonal: exact fp64 sum of squares
            if (skip) cur[e] = 0.0;
          }
        }
        const int slot = u & (kPSigSlots - 1);
        REKF_T(3, if (!mbar_wait(&sig_empty[slot], ((u / kPSigSlots) & 1) ^ 1)) timeout = true);   // the previous reduce has read it out
        uint8_t *rowp = sig + (size_t)slot * kPSigBox + (size_t)il * 128;
        *reinterpret_cast<double2 *>(rowp + (((2 * cg) ^ (il & 7)) << 4)) = make_double2(cur[0], cur[1]);
        *reinterpret_cast<double2 *>(rowp + (((2 * cg + 1) ^ (il & 7)) << 4)) = make_double2(cur[2], cur[3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&sig_done[slot]);       // hand the box to the reduce issuer
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[set]);                        // this thread is done with the accumulator set
      ++iter;
    }
#ifdef REKF_SYRK_TIMING
    if (threadIdx.x == 0) {
      tlog[15] = tacc[0]; tlog[16] = tacc[1]; tlog[17] = tacc[2]; tlog[18] = tacc[3]; tlog[19] = (double)(clock64() - t_begin); tlog[20] = (double)iter;
    }
#endif
  }
  if (timeout) atomicOr(&L.st[L.s0].flags, FLAG_TCGEN05_TIMEOUT);
  __syncthreads();
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
  tc.resident = L.kq <= kPResChunks * kPKBox;
  if (cudaFuncSetAttribute(k_syrk_tcgen05_i8p<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemRes) != cudaSuccess ||
      cudaFuncSetAttribute(k_syrk_tcgen05_i8p<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemStr) != cudaSuccess)
    return "cudaFuncSetAttribute(k_syrk_tcgen05_i8p, smem) failed";
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&tc.num_sms, cudaDevAttrMultiProcessorCount, dev);
  tc.ready = true;
  return nullptr;
}

inline int syrk_i8p_launch(const SyrkI8P &tc, const Layout &L, cudaStream_t stream, bool pdl = false) {
  if (!tc.ready) return -1;
  const int total = (L.ld / 128) * (L.ld / 128 + 1) * L.Sg;
  const int ctas = tc.num_sms - tc.reserve_sms > 1 ? tc.num_sms - tc.reserve_sms : 1;
  const int grid = total < ctas ? total : ctas;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = tc.resident ? kPSmemRes : kPSmemStr;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // barrier init / TMEM allocation overlap the kernel in front
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const cudaError_t e = tc.resident ? cudaLaunchKernelEx(&cfg, k_syrk_tcgen05_i8p<true>, L, tc.map_a, tc.map_b, tc.map_sig)
                                    : cudaLaunchKernelEx(&cfg, k_syrk_tcgen05_i8p<false>, L, tc.map_a, tc.map_b, tc.map_sig);
  return e == cudaSuccess ? 0 : -1;
}

}  // namespace rekf
