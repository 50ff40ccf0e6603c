// syrk_tcgen05_i8.cuh — Σ −= Wᵀ·W (reflector_ekf_slam.cc:308) on the INT8 tensor cores, exactly.
//
// The tf32x3 kernel (syrk_tcgen05.cuh) is fp32-class: its TMEM accumulation rounds, and the error it
// leaves in Σ random-walks into μ over hundreds of steps (measured: 2.5e-4 m after 220 steps at C3,
// outside the 1e-4 m parity bar).  This kernel keeps the tensor cores and removes the rounding:
//
//   each row c of Wᵀ is scaled by a power of two 2^e_c so that |x| <= 1/2 and cut into P = 4 signed
//   7-bit digits,  x = Σ_p d_p·2^(-7(p+1)) + ρ,  d_p ∈ [-64, 64] (int8),  |ρ| <= 2^-29;
//   then  (Wᵀ·W)[i][j] = 2^(e_i+e_j) · Σ_s 2^(-7(s+2)) · G_s[i][j],   G_s = Σ_{p+q=s} Σ_k d_p,i[k]·d_q,j[k].
//
// Every G_s is an integer dot product: tcgen05.mma.kind::i8 computes it with s32 accumulation in TMEM
// *exactly* (|G_s| <= 4·256·64² < 2^23), one TMEM accumulator per s, and the epilogue recombines the
// four accumulators in fp64 — also exactly.  The only approximation is dropping the digit pairs with
// p+q >= 4 (measured 1.5e-10 relative Frobenius per step against the fp64 SYRK); the diagonal, where that
// truncation would be one-signed, is taken from the exact fp64 sums of squares k_solve_w3 provides, and
// frames whose downdate cancels deeply are routed to the fp64 SYRK (SessionState::exact_update).
// Ten int8 MMAs of K=32 replace three tf32 MMAs of K=8, so the tensor time per tile is unchanged.
// (kind::i8 exists on sm_100a; B300/sm_103a dropped it.)
//
// Operands: Wq[session][p][c][k] int8, K-major, 64-byte swizzle; TMA boxes of 128 (A) / 64 (B) rows x 64 k.
// CTA = one 128x64 tile on or above the diagonal: warp 8 TMA producer (2-stage ring), warp 9 MMA issuer
// (M=128, N=64, four 64-column s32 accumulators = 256 TMEM columns), warps 0-7 epilogue (tcgen05.ld → fp64 →
// Σ[i][j] and the mirrored Σ[j][i]).  96 KB of shared memory and 256 TMEM columns per CTA, so two CTAs
// share an SM and one's epilogue (HBM-bound) overlaps the other's TMA + MMA phase.
#pragma once
#include "syrk_tcgen05.cuh"

namespace rekf {

constexpr int kI8Slices = 4;
constexpr int kI8KBox = 64;                              // k per TMA box (bytes)
constexpr int kI8TileN = 64;
constexpr int kI8BoxA = 128 * kI8KBox;                   // 8 KB
constexpr int kI8BoxB = kI8TileN * kI8KBox;              // 4 KB
constexpr int kI8StageBytes = kI8Slices * (kI8BoxA + kI8BoxB);   // 48 KB
constexpr int kI8Stages = 2;
constexpr int kI8SmemBytes = kI8Stages * kI8StageBytes + 1024 + 256;
constexpr uint32_t kI8TmemCols = 256;                    // four 64-column s32 accumulators

struct SyrkI8 {
  CUtensorMap map_a, map_b;
  bool ready = false;
};

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per instruction per lane
__device__ __forceinline__ void ldg256(const double *p, double *v) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void stg256(double *p, const double *v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
// K-major, SWIZZLE_64B: rows of 64 bytes, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                               // SWIZZLE_64B
  return d;
}
// D = s32, A = B = signed int8, K-major, N = 64, M = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)kI8TileN >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(kTcThreads, 2)
k_syrk_tcgen05_i8(Layout L, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b) {
  extern __shared__ uint8_t smem_raw[];
  const int s = L.s0 + blockIdx.y;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0 || st.exact_update) return;                // deep-cancellation frames go to k_syrk_f64
  const int n = internal_dim(st.N);
  // tile decode: row block ti (128 rows) x column block tj (64 columns), tj >= 2·ti
  const int Tn64 = L.ld / kI8TileN;
  int ti = 0, rem = (int)blockIdx.x;
  while (rem >= Tn64 - 2 * ti) { rem -= Tn64 - 2 * ti; ++ti; }
  const int tj = 2 * ti + rem;
  const int i0 = ti * 128, j0 = tj * kI8TileN;
  if (j0 >= n) return;                                  // (i0 <= j0, so i0 < n as well)

  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *stages = base;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + kI8Stages * kI8StageBytes);
  uint64_t *empty_bar = full_bar + kI8Stages;
  uint64_t *accum_bar = empty_bar + kI8Stages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (r + kI8KBox - 1) / kI8KBox;          // TMA K-boxes of 64
  const int nk32 = (r + 31) / 32;                       // MMA K-steps of 32 carrying data
  const bool inA = (tj - 2 * ti) < 2;                   // the 64 B-rows are a half of the 128 A-rows: no B load
  const bool diag = inA;                                // such tiles touch the diagonal

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < kI8Stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kI8TmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  bool timeout = false;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % kI8Stages;
        const uint32_t phase = (kb / kI8Stages) & 1;
        if (!mbar_wait(&empty_bar[stage], phase ^ 1)) { timeout = true; break; }
        uint8_t *sa = stages + (size_t)stage * kI8StageBytes;
        uint8_t *sb = sa + kI8Slices * kI8BoxA;
        mbar_expect_tx(&full_bar[stage], kI8Slices * (kI8BoxA + (inA ? 0 : kI8BoxB)));
#pragma unroll
        for (int p = 0; p < kI8Slices; ++p) {
          tma_load_4d(sa + p * kI8BoxA, &map_a, &full_bar[stage], kb * kI8KBox, i0, p, s);
          if (!inA) tma_load_4d(sb + p * kI8BoxB, &map_b, &full_bar[stage], kb * kI8KBox, j0, p, s);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % kI8Stages;
        const uint32_t phase = (kb / kI8Stages) & 1;
        if (!mbar_wait(&full_bar[stage], phase)) { timeout = true; break; }
        tc_fence_after();
        const uint32_t sa = smem_u32(stages + (size_t)stage * kI8StageBytes);
        // B operand: own boxes, or rows (j0 − i0)..+63 of the A boxes (64 rows x 64 B = 4096 B further in)
        const uint32_t sb = inA ? sa + (uint32_t)(j0 - i0) * kI8KBox : sa + kI8Slices * kI8BoxA;
        const uint32_t bstride = inA ? kI8BoxA : kI8BoxB;
        const int steps = min(2, nk32 - kb * 2);
        for (int ks = 0; ks < steps; ++ks) {
          const uint32_t koff = ks * 32;                // 32 int8 = 32 bytes inside the 64-byte swizzle row
          const bool first = (kb | ks) == 0;
#pragma unroll
          for (int sgrp = 0; sgrp < kI8Slices; ++sgrp) {
#pragma unroll
            for (int p = 0; p <= sgrp; ++p) {
              const int q = sgrp - p;
              const uint64_t da = make_kmajor_sw64_desc(sa + p * kI8BoxA + koff);
              const uint64_t db = make_kmajor_sw64_desc(sb + q * bstride + koff);
              tc_mma_i8(tmem + sgrp * kI8TileN, da, db, kIdescI8, (first && p == 0) ? 0u : 1u);
            }
          }
        }
        tc_commit(&empty_bar[stage]);
      }
      tc_commit(accum_bar);
    }
  } else {
    // ===== epilogue: 8 warps; warp w owns TMEM lanes 32·(w%4).. and tile columns 32·(w/4).. =====
    const int quad = warp & 3, half = warp >> 2;
    const int i = i0 + quad * 32 + lane;                // row of Σ owned by this thread
    double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
    const double *Wsc = L.Wscale + (size_t)s * L.ld;
    const int ld = L.ld;
    const double si = Wsc[min(i, ld - 1)] * 0x1p-35;    // 2^(e_i − 35): the weight of the combined integer sum
    const unsigned char *flag = L.Wflag + (size_t)s * L.ld;
    const bool row_ok = !flag[min(i, ld - 1)];           // flagged slots are k_syrk_exact_rows' (fp64)
    const bool above = (i0 + 127 < j0);                 // whole tile strictly above the diagonal
    // the first 16 columns of this thread's Σ row are fetched while the tensor pipe is still busy
    double cur[16];
    {
      const int jbase = j0 + half * 32;
      const bool want = row_ok && i < n && jbase < n && !(diag && jbase + 15 < i);
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        cur[u] = cur[u + 1] = cur[u + 2] = cur[u + 3] = 0.0;
        if (want) ldg256(Sg + (size_t)i * ld + jbase + u, cur + u);
      }
    }
    if (!mbar_wait(accum_bar, 0)) timeout = true;
    tc_fence_after();
#pragma unroll 1
    for (int chunk = 0; chunk < 2; ++chunk) {
      const int col0 = half * 32 + chunk * 16;
      const int jbase = j0 + col0;
      const bool want = row_ok && i < n && jbase < n && !(diag && jbase + 15 < i);
      double *row = Sg + (size_t)i * ld + jbase;
      if (chunk == 1 && want) {
#pragma unroll
        for (int u = 0; u < 16; u += 4) ldg256(row + u, cur + u);
      }
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
      // G = g0·2^21 + g1·2^14 + g2·2^7 + g3 in exact 64-bit integer arithmetic (|g_s| < 2^23), one conversion
      long long G[16];
      {
        uint32_t gq[16];
        tc_ld16(taddr, gq);
        tc_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) G[u] = (long long)(int)gq[u] << 21;
        tc_ld16(taddr + kI8TileN, gq);
        tc_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) G[u] += (long long)(int)gq[u] << 14;
        tc_ld16(taddr + 2 * kI8TileN, gq);
        tc_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) G[u] += (long long)(int)gq[u] << 7;
        tc_ld16(taddr + 3 * kI8TileN, gq);
        tc_wait_ld();
#pragma unroll
        for (int u = 0; u < 16; ++u) G[u] += (long long)(int)gq[u];
      }
      if (want) {
        double old_diag = 0.0;
        const int ud = i - jbase;                       // position of the diagonal element in this chunk, if any
        const double2 *scj = reinterpret_cast<const double2 *>(Wsc + min(jbase, ld - 16));   // 2^e_j, warp-uniform loads
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
          const double2 sj = scj[u >> 1];
          if (u == ud) old_diag = cur[u];
          if (u + 1 == ud) old_diag = cur[u + 1];
          cur[u] = fma(-(double)G[u], si * sj.x, cur[u]);          // (2^-35·2^e_i·2^e_j)·G: one rounding, like the fp64 SYRK
          cur[u + 1] = fma(-(double)G[u + 1], si * sj.y, cur[u + 1]);
        }
        if (diag && ud >= 0 && ud < 16) {               // exact fp64 sum of squares from k_solve_w3
          const double dd = old_diag - L.Wdiag[(size_t)s * ld + i];
#pragma unroll
          for (int u = 0; u < 16; ++u) if (u == ud) cur[u] = dd;
        }
        const uint4 cf = *reinterpret_cast<const uint4 *>(flag + min(jbase, ld - 16));   // 16 column flags, warp-uniform
        const bool cols_ok = (cf.x | cf.y | cf.z | cf.w) == 0u;
        if (above && cols_ok && jbase + 15 < n) {
#pragma unroll
          for (int u = 0; u < 16; u += 4) stg256(row + u, cur + u);     // full 32-byte sectors
#pragma unroll
          for (int u = 0; u < 16; ++u) Sg[(size_t)(jbase + u) * ld + i] = cur[u];
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = jbase + u;
            const unsigned cfw = (u < 4) ? cf.x : (u < 8) ? cf.y : (u < 12) ? cf.z : cf.w;
            if (j < n && i <= j && !((cfw >> (8 * (u & 3))) & 0xffu)) {
              row[u] = cur[u];
              if (i != j) Sg[(size_t)j * ld + i] = cur[u];
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  if (timeout) atomicOr(&st.flags, FLAG_TCGEN05_TIMEOUT);
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kI8TmemCols) : "memory");
  }
}

inline const char *syrk_i8_init(SyrkI8 &tc, const Layout &L) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return "cuTensorMapEncodeTiled entry point not available";
  PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(fn);
  const cuuint64_t dims[4] = {(cuuint64_t)L.kq, (cuuint64_t)L.ld, (cuuint64_t)kI8Slices, (cuuint64_t)L.S};
  const cuuint64_t strides[3] = {(cuuint64_t)L.kq, (cuuint64_t)L.ld * L.kq, (cuuint64_t)kI8Slices * L.ld * L.kq};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const cuuint32_t box_a[4] = {(cuuint32_t)kI8KBox, 128u, 1u, 1u};
  const cuuint32_t box_b[4] = {(cuuint32_t)kI8KBox, (cuuint32_t)kI8TileN, 1u, 1u};
  if (encode(&tc.map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, L.Wq, dims, strides, box_a, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wq, A box) failed";
  if (encode(&tc.map_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, L.Wq, dims, strides, box_b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wq, B box) failed";
  if (cudaFuncSetAttribute(k_syrk_tcgen05_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes) != cudaSuccess)
    return "cudaFuncSetAttribute(k_syrk_tcgen05_i8, smem) failed";
  tc.ready = true;
  return nullptr;
}

inline int syrk_i8_launch(const SyrkI8 &tc, const Layout &L, cudaStream_t stream) {
  if (!tc.ready) return -1;
  const int Tn = L.ld / 128;
  k_syrk_tcgen05_i8<<<dim3(Tn * (Tn + 1), L.Sg), kTcThreads, kI8SmemBytes, stream>>>(L, tc.map_a, tc.map_b);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace rekf
