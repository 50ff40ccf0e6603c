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
// This header holds the scheme's shared pieces (slice constants, the kind::i8 MMA and descriptor helpers, 256-bit global
// accesses); the kernel itself is the persistent k_syrk_tcgen05_i8p in syrk_tcgen05_i8p.cuh (the first, one-tile-per-CTA
// form of it was retired when the persistent one overtook it).
// Operands: Wq[session][slice][k/64][c][64] int8, K-major, 64-byte swizzle; TMA boxes of 128 (A) / 64 (B) rows x 64 k.
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

__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
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
// D = s32, A = B = signed int8, K-major, M = 128, N = n (a multiple of 16 up to 256)
__host__ __device__ constexpr uint32_t idesc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
}
constexpr uint32_t kIdescI8 = idesc_i8(kI8TileN);

}  // namespace rekf
