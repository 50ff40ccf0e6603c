// syrk_exact_rows.cuh — the rows/columns of Σ −= Wᵀ·W that the int8-slice tensor kernel must not touch.
//
// The int8 digit slices resolve 2^-29 of a row's scale.  For a state whose variance collapses in this frame
// (a landmark seen again after a long time, the pose after dead reckoning) that is not small against the
// posterior, so k_solve_w3 flags such slots (Wflag / exact_list) and the tensor kernel skips every element
// whose row or column is flagged.  This kernel computes exactly those elements in fp64 from the fp64 panel
// W64: Σ[a][j] −= Σ_k W[k][a]·W[k][j] for every flagged slot a and every column j (stored once, in the upper triangle).
// A pair of flagged slots (a, a') is owned by the smaller index so that it is subtracted once.
// W is measurement-row major, so with one lane per column j, W[k][a] is a broadcast and W[k][j] a coalesced read; the k range
// is split over the 8 warps of the block (partial sums meet in shared memory) so that every thread has only ~r/8 dependent
// loads.  A handful of flagged slots per frame cost ~n·r FMAs each — microseconds — where routing the whole frame to the
// fp64 SYRK would cost a millisecond.
#pragma once
#include "rekf_device.cuh"

namespace rekf {

// Called by every CTA of k_syrk_f64's (148, 1, S) x 256-thread launch in int8 mode, so a frame without flagged slots costs
// one read of SessionState per CTA and no second launch.
__device__ __forceinline__ void syrk_exact_rows(const Layout &L, int s) {
  const SessionState &st = L.st[s];
  const int r = st.r;
  const int cnt = min(st.exact_slots, kMaxExactSlots);
  if (r == 0 || cnt == 0 || st.exact_update) return;       // exact_update: the whole frame is done by k_syrk_f64 (block-uniform)
  __shared__ double part[8][33];
  const int n = internal_dim(st.N);
  const int ld = L.ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // blockDim.x == 256
  const double *W = L.W64 + (size_t)s * ld * L.rld;
  const unsigned char *flag = L.Wflag + (size_t)s * ld;
  const int *list = L.exact_list + (size_t)s * kMaxExactSlots;
  double *Sg = L.sigma + (size_t)s * ld * ld;
  for (int j0 = blockIdx.x * 32; j0 < n; j0 += gridDim.x * 32) {
    const int j = j0 + lane;
    const int jc = min(j, n - 1);
    for (int q = 0; q < cnt; ++q) {
      const int a = list[q];
      double acc0 = 0.0, acc1 = 0.0;
      int k = warp;
      for (; k + 8 < r; k += 16) {
        acc0 = fma(W[(size_t)k * ld + a], W[(size_t)k * ld + jc], acc0);
        acc1 = fma(W[(size_t)(k + 8) * ld + a], W[(size_t)(k + 8) * ld + jc], acc1);
      }
      if (k < r) acc0 = fma(W[(size_t)k * ld + a], W[(size_t)k * ld + jc], acc0);
      part[warp][lane] = acc0 + acc1;
      __syncthreads();
      if (warp == 0 && j < n && !(flag[j] && j < a)) {      // (j, a) with both flagged is owned by the smaller index
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += part[w][lane];
        Sg[sym_idx(a, j, ld)] -= acc;
      }
      __syncthreads();
    }
  }
}

}  // namespace rekf
