// syrk_exact_rows.cuh — the rows/columns of Σ −= Wᵀ·W that the int8-slice tensor kernel must not touch.
//
// The int8 digit slices resolve 2^-29 of a row's scale.  For a state whose variance collapses in this frame
// (a landmark seen again after a long time, the pose after dead reckoning) that is not small against the
// posterior, so k_solve_w3 flags such slots (Wflag / exact_list) and the tensor kernel skips every element
// whose row or column is flagged.  This kernel computes exactly those elements in fp64 from the fp64 panel
// W64: Σ[a][j] −= Σ_k W[k][a]·W[k][j] for every flagged slot a and every column j (stored once, in the upper triangle).
// A pair of flagged slots (a, a') is owned by the smaller index so that it is subtracted once.
// One warp per element (lanes stride k, shuffle reduction): a handful of flagged slots per frame cost ~n·r
// FMAs each — microseconds — where routing the whole frame to the fp64 SYRK would cost a millisecond.
#pragma once
#include "rekf_device.cuh"

namespace rekf {

// Called by every CTA of k_syrk_f64's (148, 1, S) launch in int8 mode: warps stride the columns, so a frame
// without flagged slots costs one read of SessionState per CTA and no second launch.
__device__ __forceinline__ void syrk_exact_rows(const Layout &L, int s) {
  const SessionState &st = L.st[s];
  const int r = st.r;
  const int cnt = min(st.exact_slots, kMaxExactSlots);
  if (r == 0 || cnt == 0 || st.exact_update) return;       // exact_update: the whole frame is done by k_syrk_f64
  const int n = internal_dim(st.N);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int ld = L.ld, rld = L.rld;
  const double *W = L.W64 + (size_t)s * ld * rld;
  const unsigned char *flag = L.Wflag + (size_t)s * ld;
  const int *list = L.exact_list + (size_t)s * kMaxExactSlots;
  double *Sg = L.sigma + (size_t)s * ld * ld;
  for (int j = blockIdx.x * nwarp + warp; j < n; j += gridDim.x * nwarp) {   // column handled by this warp
    const double *wj = W + (size_t)j * rld;
    for (int q = 0; q < cnt; ++q) {
      const int a = list[q];
      if (flag[j] && j < a) continue;                       // (j, a) is owned by row j
      const double *wa = W + (size_t)a * rld;
      double acc = 0.0;
      for (int k = lane; k < r; k += 32) acc = fma(wa[k], wj[k], acc);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) Sg[sym_idx(a, j, ld)] -= acc;
    }
  }
}

}  // namespace rekf
