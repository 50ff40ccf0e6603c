// syrk_tcgen05.cuh — Σ −= Wᵀ·W (reflector_ekf_slam.cc:308, Σ − K·H·Σ) on the 5th-generation tensor cores.
//
// The one dense contraction of the EKF step: M = N = n (state dimension), K = r (measurement rows).
// Operands are the two tf32 panels Wt_hi / Wt_lo written by k_solve_w (row c of Wᵀ, K contiguous —
// "K-major" for both A and B, since the product is Wᵀ·(Wᵀ)ᵀ).  3xTF32: W = hi + lo with hi, lo exactly
// representable in tf32, and
//     Wᵀ·W ≈ hiᵀ·hi  +  (hiᵀ·lo + loᵀ·hi)
// accumulated in two separate fp32 TMEM accumulators (main, correction) that are added in fp64 in the
// epilogue, so the small correction terms are not rounded away against the large main sum.
//
// One CTA per 128x128 upper-triangular tile (ti <= tj) per session:
//   warp 8  : TMA producer — cp.async.bulk.tensor (128B swizzle) of the four 128x32 operand boxes of a
//             K-block into a 3-stage shared-memory ring, mbarrier complete_tx signalling
//   warp 9  : MMA issuer   — one elected thread issues tcgen05.mma.cta_group::1.kind::tf32
//             (M=128, N=128, K=8), tcgen05.commit releases ring slots and finally publishes the accumulators
//   warps 0-7: epilogue    — tcgen05.ld 32x32b.x16, Σ[i][j] ← Σ[i][j] − (main+corr) in fp64, upper triangle only
//             (the lower triangle is not stored, rekf_device.cuh).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "rekf_device.cuh"

namespace rekf {

constexpr int kTcStages = 3;
constexpr int kTcBoxBytes = 128 * 128;                 // 128 rows x 32 tf32
constexpr int kTcStageBytes = 4 * kTcBoxBytes;         // A_hi, A_lo, B_hi, B_lo
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kTcThreads = 320;
constexpr uint32_t kTcTmemCols = 256;                  // main (128) + correction (128)
constexpr uint32_t kSpinLimit = 1u << 26;

struct SyrkTc {
  CUtensorMap map_hi, map_lo;
  bool ready = false;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug must not hang the GPU (returns false on timeout)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t it = 0; it < kSpinLimit; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem]·B[smem]ᵀ, kind::tf32, M=128 N=128 K=8
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (matches the TMA box)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, 16-byte units
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, N=128, M=128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// ---- kernel --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
k_syrk_tcgen05(Layout L, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ uint8_t smem_raw[];
  const int s = L.s0 + blockIdx.y;
  SessionState &st = L.st[s];
  const int r = st.r;
  if (r == 0) return;
  const int n = internal_dim(st.N);
  // upper-triangular tile decode: blockIdx.x = tj(tj+1)/2 + ti, ti <= tj
  int tj = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((tj + 1) * (tj + 2) / 2 <= (int)blockIdx.x) ++tj;
  while (tj * (tj + 1) / 2 > (int)blockIdx.x) --tj;
  const int ti = (int)blockIdx.x - tj * (tj + 1) / 2;
  const int i0 = ti * 128, j0 = tj * 128;
  if (j0 >= n) return;                                  // whole CTA leaves before touching barriers / TMEM

  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-byte alignment
  uint8_t *stages = base;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + kTcStages * kTcStageBytes);
  uint64_t *empty_bar = full_bar + kTcStages;
  uint64_t *accum_bar = empty_bar + kTcStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (r + kKBlock - 1) / kKBlock;           // K-blocks of 32
  const int nk8 = (r + 7) / 8;                          // K-steps of 8 actually carrying data
  const bool diag = (ti == tj);

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < kTcStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTcTmemCols) : "memory");
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
      for (int kb = 0; kb < nk; ++kb) {
        const int stage = kb % kTcStages;
        const uint32_t phase = (kb / kTcStages) & 1;
        if (!mbar_wait(&empty_bar[stage], phase ^ 1)) { timeout = true; break; }
        uint8_t *sa = stages + (size_t)stage * kTcStageBytes;
        mbar_expect_tx(&full_bar[stage], diag ? 2 * kTcBoxBytes : 4 * kTcBoxBytes);
        tma_load_3d(sa, &map_hi, &full_bar[stage], kb * kKBlock, i0, s);
        tma_load_3d(sa + kTcBoxBytes, &map_lo, &full_bar[stage], kb * kKBlock, i0, s);
        if (!diag) {
          tma_load_3d(sa + 2 * kTcBoxBytes, &map_hi, &full_bar[stage], kb * kKBlock, j0, s);
          tma_load_3d(sa + 3 * kTcBoxBytes, &map_lo, &full_bar[stage], kb * kKBlock, j0, s);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t d_main = tmem, d_corr = tmem + 128;
      for (int kb = 0; kb < nk; ++kb) {
        const int stage = kb % kTcStages;
        const uint32_t phase = (kb / kTcStages) & 1;
        if (!mbar_wait(&full_bar[stage], phase)) { timeout = true; break; }
        tc_fence_after();
        const uint32_t sa = smem_u32(stages + (size_t)stage * kTcStageBytes);
        const uint32_t a_hi = sa, a_lo = sa + kTcBoxBytes;
        const uint32_t b_hi = diag ? a_hi : sa + 2 * kTcBoxBytes, b_lo = diag ? a_lo : sa + 3 * kTcBoxBytes;
        const int steps = min(4, nk8 - kb * 4);
        for (int ks = 0; ks < steps; ++ks) {
          const uint32_t koff = ks * 32;                // 8 tf32 = 32 bytes inside the 128-byte swizzle row
          const uint64_t dah = make_kmajor_sw128_desc(a_hi + koff), dal = make_kmajor_sw128_desc(a_lo + koff);
          const uint64_t dbh = make_kmajor_sw128_desc(b_hi + koff), dbl = make_kmajor_sw128_desc(b_lo + koff);
          const uint32_t acc = (kb | ks) ? 1u : 0u;
          tc_mma_tf32(d_main, dah, dbh, kIdescTf32, acc);      // hiᵀ·hi
          tc_mma_tf32(d_corr, dah, dbl, kIdescTf32, acc);      // hiᵀ·lo
          tc_mma_tf32(d_corr, dal, dbh, kIdescTf32, 1u);       // loᵀ·hi
        }
        tc_commit(&empty_bar[stage]);                   // ring slot free once these MMAs retire
      }
      tc_commit(accum_bar);                             // accumulators complete
    }
  } else {
    // ===== epilogue: 8 warps; warp w owns TMEM lanes 32·(w%4).. and columns 64·(w/4).. =====
    const int quad = warp & 3, half = warp >> 2;
    if (!mbar_wait(accum_bar, 0)) timeout = true;
    tc_fence_after();
    const int i = i0 + quad * 32 + lane;                // row of Σ owned by this thread
    double *Sg = L.sigma + (size_t)s * L.ld * L.ld;
    const int ld = L.ld;
#pragma unroll 1
    for (int chunk = 0; chunk < 4; ++chunk) {
      const int col0 = half * 64 + chunk * 16;
      uint32_t vm[16], vc[16];
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0;
      tc_ld16(taddr, vm);
      tc_ld16(taddr + 128, vc);
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
        for (int u = 0; u < 16; ++u)
          cur[u] -= (double)__uint_as_float(vm[u]) + (double)__uint_as_float(vc[u]);
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
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = jbase + u;
            if (j < n && (!diag || i <= j)) row[u] = cur[u];
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// returns nullptr on success, else a static description of what failed
inline const char *syrk_tc_init(SyrkTc &tc, const Layout &L) {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
      qres != cudaDriverEntryPointSuccess)
    return "cuTensorMapEncodeTiled entry point not available";
  PFN_encodeTiled encode = reinterpret_cast<PFN_encodeTiled>(fn);
  const cuuint64_t dims[3] = {(cuuint64_t)L.rld, (cuuint64_t)L.ld, (cuuint64_t)L.S};
  const cuuint64_t strides[2] = {(cuuint64_t)L.rld * sizeof(float), (cuuint64_t)L.ld * L.rld * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)kKBlock, 128u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (encode(&tc.map_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, L.Wt_hi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wt_hi) failed";
  if (encode(&tc.map_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, L.Wt_lo, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return "cuTensorMapEncodeTiled(Wt_lo) failed";
  if (cudaFuncSetAttribute(k_syrk_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes) != cudaSuccess)
    return "cudaFuncSetAttribute(k_syrk_tcgen05, smem) failed";
  tc.ready = true;
  return nullptr;
}

inline int syrk_tc_launch(const SyrkTc &tc, const Layout &L, cudaStream_t stream) {
  if (!tc.ready) return -1;
  const int Tn = L.ld / 128;
  k_syrk_tcgen05<<<dim3(Tn * (Tn + 1) / 2, L.Sg), kTcThreads, kTcSmemBytes, stream>>>(L, tc.map_hi, tc.map_lo);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

inline void syrk_tc_destroy(SyrkTc &tc) { tc.ready = false; }

}  // namespace rekf
