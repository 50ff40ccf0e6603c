// probe_tma_reduce.cu — micro-benchmark behind the covariance SYRK's store path:
// how fast can 148 persistent CTAs update an fp64 matrix that is larger than L2
//   (a) by TMA load -> (register add in shared memory) -> TMA store   (the round-1 data path), and
//   (b) by cp.reduce.async.bulk.tensor .add.f64 of a delta tile from shared memory (the L2 does the read-modify-write)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probe_tma_reduce scripts/probe_tma_reduce.cu -lcuda
// Run on the GPU box: ./scripts/probe_tma_reduce [S=8] [ld=2176]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int kHalf = 128 * 32 * 8;   // 128 rows x 32 columns fp64
constexpr int kSlots = 4;

// mode 0: load+store (in place, one elected thread; the "update" is skipped: pure data movement)
// mode 1: reduce-add of a constant delta tile
// work: upper-triangular 128x64 tiles of S matrices (like the SYRK), static round-robin over CTAs
__global__ void __launch_bounds__(128, 1) k_probe(const __grid_constant__ CUtensorMap map, int S, int ld, int mode, int upper) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *full = reinterpret_cast<uint64_t *>(base + kSlots * kHalf);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double *d = reinterpret_cast<double *>(base);
  for (int e = threadIdx.x; e < kSlots * kHalf / 8; e += blockDim.x) d[e] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int Tr = ld / 128, Tc = ld / 64;
  const int per = Tr * Tc;
  uint32_t it = 0;
  if (mode == 0) {
    // software pipeline: keep kSlots-1 half-tile loads in flight; store a slot when its load lands
    int pend_s[kSlots] = {0}, pend_i[kSlots] = {0}, pend_j[kSlots] = {0};
    uint32_t head = 0, tail = 0;
    auto issue_store = [&]() {
      const int slot = tail % kSlots;
      mbar_wait(&full[slot], (tail / kSlots) & 1);
      uint8_t *src = base + slot * kHalf;
      tma_store_3d(&map, src, pend_j[slot], pend_i[slot], pend_s[slot]);
      tma_store_3d(&map, src + kHalf / 2, pend_j[slot] + 16, pend_i[slot], pend_s[slot]);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      ++tail;
    };
    for (int item = blockIdx.x; item < per * S; item += gridDim.x) {
      const int s = item / per, t = item - s * per, ti = t / Tc, tj = t - ti * Tc;
      if (upper && tj < 2 * ti) continue;
      for (int h = 0; h < 2; ++h) {
        if (head - tail == kSlots) issue_store();
        // the slot about to be refilled must have been read by its store
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // head - tail <= 2 here: the slot's old store is not among the last one
        const int slot = head % kSlots;
        uint8_t *dst = base + slot * kHalf;
        mbar_expect_tx(&full[slot], kHalf);
        tma_load_3d(dst, &map, &full[slot], tj * 64 + 32 * h, ti * 128, s);
        tma_load_3d(dst + kHalf / 2, &map, &full[slot], tj * 64 + 32 * h + 16, ti * 128, s);
        pend_s[slot] = s; pend_i[slot] = ti * 128; pend_j[slot] = tj * 64 + 32 * h;
        ++head;
        if (head - tail >= kSlots - 1) issue_store();
      }
    }
    while (tail != head) issue_store();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    for (int item = blockIdx.x; item < per * S; item += gridDim.x) {
      const int s = item / per, t = item - s * per, ti = t / Tc, tj = t - ti * Tc;
      if (upper && tj < 2 * ti) continue;
      for (int h = 0; h < 2; ++h) {
        const int slot = it % kSlots;
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kSlots - 1) : "memory");
        const uint8_t *src = base + slot * kHalf;
        tma_reduce_add_3d(&map, src, tj * 64 + 32 * h, ti * 128, s);
        tma_reduce_add_3d(&map, src + kHalf / 2, tj * 64 + 32 * h + 16, ti * 128, s);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++it;
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 8, ld = argc > 2 ? atoi(argv[2]) : 2176;
  double *sig;
  const size_t elems = (size_t)S * ld * ld;
  CK(cudaMalloc(&sig, elems * 8));
  CK(cudaMemset(sig, 0, elems * 8));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)ld, (cuuint64_t)S};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * ld * 8};
  const cuuint32_t box[3] = {16u, 128u, 1u}, estr[3] = {1u, 1u, 1u};
  if (((PFN_encodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, sig, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n");
    return 1;
  }
  const int smem = kSlots * kHalf + 1024 + 256;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int upper = 0; upper < 2; ++upper)
    for (int mode = 0; mode < 2; ++mode) {
      const int Tr = ld / 128, Tc = ld / 64;
      size_t tiles = 0;
      for (int ti = 0; ti < Tr; ++ti) for (int tj = 0; tj < Tc; ++tj) if (!upper || tj >= 2 * ti) ++tiles;
      const double bytes = (double)tiles * S * 128 * 64 * 8;
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a));
        k_probe<<<sms, 128, smem>>>(map, S, ld, mode, upper);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
      }
      CK(cudaGetLastError());
      printf("S=%d ld=%d %s %-12s: %8.1f us  matrix bytes updated %.1f MB -> %.0f GB/s updated, %.0f GB/s read+write equivalent\n", S, ld,
             upper ? "upper" : "full ", mode ? "reduce.add" : "load+store", best * 1e3, bytes / 1e6, bytes / (best * 1e-3) / 1e9,
             2 * bytes / (best * 1e-3) / 1e9);
    }
  // correctness of the reduce: every element of the upper tiles received +1.0 per launch (2 modes x 6 reps of mode 1 => count)
  std::vector<double> hrow(ld);
  CK(cudaMemcpy(hrow.data(), sig + (size_t)5 * ld, ld * 8, cudaMemcpyDeviceToHost));
  printf("row 5: [0]=%g [100]=%g [%d]=%g (expect 12, 12, 12: six full + six upper reduce launches)\n", hrow[0], hrow[100], ld - 1, hrow[ld - 1]);
  CK(cudaMemcpy(hrow.data(), sig + (size_t)(ld - 3) * ld, ld * 8, cudaMemcpyDeviceToHost));
  printf("row %d: [0]=%g (expect 6: below the diagonal, full launches only) [%d]=%g (expect 12)\n", ld - 3, hrow[0], ld - 1, hrow[ld - 1]);
  return 0;
}
