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
// p+q >= 4, a relative error of ~1e-8 of the downdate (5e-11 of Σ per step; measured below 1e-9
// relative Frobenius against the fp64 oracle).  Ten int8 MMAs of K=32 replace three tf32 MMAs of K=8,
// so the tensor time per tile is unchanged.  (kind::i8 exists on sm_100a; B300/sm_103a dropped it.)
//
// Operands: Wq[session][p][c][k] int8, K-major, 64-byte swizzle; TMA boxes of 128 rows x 64 k.
// CTA = one 128x128 upper-triangular tile: warp 8 TMA producer (3-stage ring), warp 9 MMA issuer,
// warps 0-7 epilogue (tcgen05.ld → fp64 → Σ[i][j] and the mirrored Σ[j][i]).
#pragma once
#include "syrk_tcgen05.cuh"

namespace rekf {

constexpr int kI8Slices = 4;
constexpr int kI8KBox = 64;                              // k per TMA box (bytes)
constexpr int kI8BoxBytes = 128 * kI8KBox;               // 8 KB
constexpr int kI8StageBytes = 2 * kI8Slices * kI8BoxBytes;   // A slices + B slices = 64 KB
constexpr int kI8Stages = 3;
constexpr int kI8SmemBytes = kI8Stages * kI8StageBytes + 1024 + 256;
constexpr uint32_t kI8TmemCols = 512;                    // four 128-column s32 accumulators

struct SyrkI8 {
  CUtensorMap map;
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
// D = s32, A = B = signed int8, K-major, N = 128, M = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(kTcThreads, 1)
k_syrk_tcgen05_i8(Layout L, const __grid_constant__ CUtensorMap map) {
  extern __shared__ uint8_t smem_raw[];
  const int s = blockIdx.y;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0 || st.exact_update) return;   // deep-cancellation frames go to k_syrk_f64
  const int n = internal_dim(st.N);
  int tj = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((tj + 1) * (tj + 2) / 2 <= (int)blockIdx.x) ++tj;
  while (tj * (tj + 1) / 2 > (int)blockIdx.x) --tj;
  const int ti = (int)blockIdx.x - tj * (tj + 1) / 2;
  const int i0 = ti * 128, j0 = tj * 128;
  if (j0 >= n) return;

  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *stages = base;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + kI8Stages * kI8StageBytes);
  uint64_t *empty_bar = full_bar + kI8Stages;
  uint64_t *accum_bar = empty_bar + kI8Stages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (r + kI8KBox - 1) / kI8KBox;          // TMA K-boxes of 64
  const int nk32 = (r + 31) / 32;                       // MMA K-steps of 32 carrying data
  const bool diag = (ti == tj);

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
        mbar_expect_tx(&full_bar[stage], (diag ? 1 : 2) * kI8Slices * kI8BoxBytes);
#pragma unroll
        for (int p = 0; p < kI8Slices; ++p) {
          tma_load_4d(sa + p * kI8BoxBytes, &map, &full_bar[stage], kb * kI8KBox, i0, p, s);
          if (!diag) tma_load_4d(sa + (kI8Slices + p) * kI8BoxBytes, &map, &full_bar[stage], kb * kI8KBox, j0, p, s);
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
        const uint32_t sb = diag ? sa : sa + kI8Slices * kI8BoxBytes;
        const int steps = min(2, nk32 - kb * 2);
        for (int ks = 0; ks < steps; ++ks) {
          const uint32_t koff = ks * 32;                // 32 int8 = 32 bytes inside the 64-byte swizzle row
          const bool first = (kb | ks) == 0;
#pragma unroll
          for (int sgrp = 0; sgrp < kI8Slices; ++sgrp) {
#pragma unroll
            for (int p = 0; p <= sgrp; ++p) {
              const int q = sgrp - p;
              const uint64_t da = make_kmajor_sw64_desc(sa + p * kI8BoxBytes + koff);
              const uint64_t db = make_kmajor_sw64_desc(sb + q * kI8BoxBytes + koff);
              tc_mma_i8(tmem + sgrp * 128, da, db, kIdescI8, (first && p == 0) ? 0u : 1u);
            }
          }
        }
        tc_commit(&empty_bar[stage]);
      }
      tc_commit(accum_bar);
    }
  } else {
    // ===== epilogue =====
    const int quad = warp & 3, half = warp >> 2;
    if (!mbar_wait(accum_bar, 0)) timeout = true;
    tc_fence_after();
    const int i = i0 + quad * 32 + lane;
    double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
    const int *We = L.Wexp + (size_t)s * L.ld;
    const int ld = L.ld;
    const int ei = We[min(i, ld - 1)];
#pragma unroll 1
    for (int chunk = 0; chunk < 4; ++chunk) {
      const int col0 = half * 64 + chunk * 16;
      uint32_t g0[16], g1[16], g2[16], g3[16];
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
      tc_ld16(taddr, g0);
      tc_ld16(taddr + 128, g1);
      tc_ld16(taddr + 256, g2);
      tc_ld16(taddr + 384, g3);
      tc_wait_ld();
      const int jbase = j0 + col0;
      if (i < n && jbase < n && !(diag && jbase + 15 < i)) {
        double *row = Sg + (size_t)i * ld + jbase;
        double cur[16];
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
          const double2 t = *reinterpret_cast<const double2 *>(row + u);
          cur[u] = t.x; cur[u + 1] = t.y;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          // exact: integers below 2^23 weighted by powers of two spanning 21 bits
          const double v = (double)(int)g0[u] * 0x1p-14 + (double)(int)g1[u] * 0x1p-21 + (double)(int)g2[u] * 0x1p-28 +
                           (double)(int)g3[u] * 0x1p-35;
          cur[u] -= scalbn(v, ei + We[min(jbase + u, ld - 1)]);
        }
        if (diag) {
          const int u = i - jbase;                      // diagonal element: exact fp64 sum of squares from k_solve_w
          if (u >= 0 && u < 16) {
            const double dd = L.Wdiag[(size_t)s * ld + i];
#pragma unroll
            for (int v = 0; v < 16; ++v) if (v == u) cur[v] = row[v] - dd;
          }
        }
        if (!diag && jbase + 15 < n) {
#pragma unroll
          for (int u = 0; u < 16; u += 2) *reinterpret_cast<double2 *>(row + u) = make_double2(cur[u], cur[u + 1]);
#pragma unroll
          for (int u = 0; u < 16; ++u) Sg[(size_t)(jbase + u) * ld + i] = cur[u];
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = jbase + u;
            if (j < n && (!diag || i <= j)) {
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
  const cuuint32_t box[4] = {(cuuint32_t)kI8KBox, 128u, 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  if (encode(&tc.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, L.Wq, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wq) failed";
  if (cudaFuncSetAttribute(k_syrk_tcgen05_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes) != cudaSuccess)
    return "cudaFuncSetAttribute(k_syrk_tcgen05_i8, smem) failed";
  tc.ready = true;
  return nullptr;
}

inline int syrk_i8_launch(const SyrkI8 &tc, const Layout &L, cudaStream_t stream) {
  if (!tc.ready) return -1;
  const int Tn = L.ld / 128;
  k_syrk_tcgen05_i8<<<dim3(Tn * (Tn + 1) / 2, L.S), kTcThreads, kI8SmemBytes, stream>>>(L, tc.map);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace rekf
