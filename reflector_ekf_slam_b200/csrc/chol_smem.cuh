// chol_smem.cuh — S = L·Lᵀ with the whole lower triangle resident in shared memory (r <= ~204).
//
// Replaces `(H·Σ·Hᵀ + Q).inverse()` (reflector_ekf_slam.cc:305): S is SPD, so it is factored instead of
// inverted, and ν rides along as an extra row that comes out as L⁻¹ν (the mean update then is Wᵀ·L⁻¹ν).
// This kernel is the serial spine of the step — r pivots, each a dependent rsqrt — so the design goal is
// latency, not throughput:
//   * right-looking with look-ahead, 32-column panels, everything in shared memory: packed block columns, ~215 KB at
//     r=200, COLUMN-major like the global buffer, so the triangle comes in and goes out as one cp.async.bulk per
//     column (the L1 left beside 215 KB of shared memory is ~13 KB, and element-wise LDG/cp.async/STG through it ran
//     at ~12 B/clk: 13k + 16k cycles for the two transfers; the bulk engine does them in ~3k + ~6k);
//   * phase 1: the 32x32 diagonal block is factored by ONE warp with the block in registers (lane = row).
//     Per column the next pivot is formed from registers and shuffled out before the column broadcast, so
//     the rsqrt chain (the critical path) overlaps the rank-1 update; no CTA barrier inside the block.
//     In its shadow warp 1 inverts the PREVIOUS diagonal block (X = L_bb⁻¹, lane = column of X held in registers,
//     broadcast reads of L, no shuffles) — the TRSM kernel consumes these inverses — and warps 2-7 apply the previous
//     panel's rank-32 update to everything right of the current block column;
//   * phase 2: the rows below the block (and the ν row) times the block inverse on the fp64 tensor pipe, P ← P·Xᵀ in 8-row
//     DMMA tiles (one thread per row solving by substitution was issue-bound: 3-9 k cycles per block against ~1.5 k).
//     X = L_bb⁻¹ — the inverse the TRSM kernel consumes anyway — costs no extra pass: the panel warp carries the
//     substitution of the identity along with the factorisation (lane = column of X; step j needs exactly the column of L
//     that the rank-1 update has just broadcast, one more FMA per value read), in the issue slots the rsqrt chain leaves idle;
//   * phase 3: the panel's update of the NEXT block column only (all warps) — the one thing the next factorisation
//     waits for.  Updates run on the fp64 tensor pipe (mma.sync m8n8k4 → DMMA): a warp task is one 8-row tile
//     against its column tiles, A fragments in registers;
//   * L goes back to global memory once, at the end.
// 256 threads (the per-thread 32-double rows stay in registers), one CTA per session.
//
// Larger r (config C4: r = 400, a 641 KB triangle) is factored in two levels with the same kernel (chol_split):
//   S = [S11 S21ᵀ; S21 S22],  r1 = 32·⌊r/64⌋ rows in the leading block, both halves <= kCholResidentMax:
//   (1) this kernel on S11 (part 1) — the row that follows (row r1 of S21) rides along in the ν position and comes out solved;
//   (2) k_chol_trsm_rows: the other rows of S21 and ν against L11 (DMMA TRSM, the same tile routine as k_solve_w3);
//   (3) k_chol_syrk: S22 −= L21·L21ᵀ (incl. the ν row);  (4) this kernel on the updated S22 with ν (part 2).
// Frames with r beyond 2·kCholResidentMax − 16 (only GPS rows on top of 200 reflectors) fall back to k_cholesky.
#pragma once
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"

namespace rekf {

constexpr int kCholSmemThreads = 256;
constexpr int kCholXP = 36;              // pitch of the block inverse in shared memory: DMMA A-fragment loads (8 rows x 4 k) 2-way at worst
constexpr int kCholResidentMax = 208;   // rows one CTA's shared memory holds (219 KB of packed panels)

// two-level split of a frame with r measurement rows: 0 = single pass, else r1 (a multiple of 32), -1 = too large for two levels
__host__ __device__ inline int chol_split(int r) {
  if (r <= kCholResidentMax) return 0;
  const int r1 = 32 * (r / 64);
  return (r1 <= kCholResidentMax && r - r1 <= kCholResidentMax) ? r1 : -1;
}

// Packed COLUMN-major block columns — the same orientation as the global S / L buffer, so that the triangle moves in
// and out as one bulk copy (cp.async.bulk) per column.  Block column b holds columns 32b..32b+31, rows 32b..R1-1
// (R1 = r+1 incl. the ν row) at pitch lda(b) ≡ 8 (mod 16) doubles: the DMMA fragment loads (8 rows x 4 k) then hit
// every bank pair exactly twice per 256-byte request (the minimum), and lane = row accesses are contiguous.
// The last block column always has room for a full 32x32 diagonal block + the ν row (a partial last block is padded
// with the identity in shared memory, so every block runs the same code).
__host__ __device__ inline int chol_lda(int R1, int b) {
  const int rows = R1 - 32 * b;
  return ((max(rows, kCholNb + 1) + 7) & ~15) + 8;
}
__host__ __device__ inline int chol_col_off(int R1, int b) {
  int off = 0;
  for (int q = 0; q < b; ++q) off += kCholNb * chol_lda(R1, q);
  return off;
}
inline size_t smem_chol_resident(int rcap) {
  const int R1 = rcap + 1, nb = (rcap + kCholNb - 1) / kCholNb;
  return sizeof(double) * ((size_t)chol_col_off(R1, nb) + 2 * 32 + 2 * 32 + 2 + 8 + 32 * kCholXP);
}

// D(8x8) = A(8x4)·B(4x8) + C on the fp64 tensor pipe.  Fragments (lane = 4·g + t): a = A[g][t], b = B[t][g],
// c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// 1/sqrt(d) for a pivot of an SPD matrix, branch-free: the hardware's 23-bit seed (MUFU.RSQ64H) and kNewton Newton steps
// (one: ~1e-13, two: full double).  libdevice's rsqrt() carries special-case branches; inside the fully unrolled pivot loop they
// end the basic block, and the scheduler could no longer start the next pivot's chain under the current column's rank-1 update —
// the diagonal blocks ran at ~240 cycles per column instead of the ~130 of the dependent chain.
// The second step is three dependent fp64 operations on the factorisation's serial spine (200 pivots per frame at C3: 2 µs of
// 57).  The tensor-core covariance modes resolve the downdate to 2^-29 of a row scale anyway and take ONE step (the factor then
// satisfies S = L·Lᵀ to ~2e-13 relative instead of 2e-16; measured on the 80-step C3 run: see DESIGN.md §5); the fp64 mode — the
// engine's own reference mode — keeps two.
template <int kNewton>
__device__ __forceinline__ double pivot_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  y = y * fma(-h * y, y, 1.5);
  if (kNewton > 1) y = y * fma(-h * y, y, 1.5);
  return y;
}

// trailing update A[i][c] −= Σ_k P[i][k]·P[c][k] on the fp64 tensor pipe for the 8-column tiles ct_lo..ct_hi (counted
// from column Jp+32) and every row tile at or below them; Pp = block column Jp/32 (rows Jp.., already solved).
// The product is formed transposed — D[m][n] with m = column, n = row — so that a lane's accumulator pair is two
// consecutive ROWS of one column: one 16-byte access in the column-major panels.  A warp task is one 8-row tile against
// its column tiles, the row tile's fragments in registers; two accumulator chains per tile (k halves).
__device__ __forceinline__ void chol_trailing(double *A, const int *tab, const double *Pp, int lda_p, int Jp, int R1, int r,
                                              int ct_lo, int ct_hi, int wi, int nw, int lane) {
  const int rows = R1 - Jp;
  const int T = R1 - (Jp + kCholNb);                        // remaining rows (incl. ν)
  const int Tc = r - (Jp + kCholNb);                        // remaining columns
  if (T <= 0 || Tc <= 0) return;
  const int nrt = (T + 7) / 8, nct = (Tc + 7) / 8;
  const int ntask = nrt - ct_lo;                            // row tiles ct_lo .. nrt-1
  if (ntask <= 0 || ct_lo >= nct) return;
  const int g = lane >> 2, t4 = lane & 3;
  for (int q = 0; q * nw < ntask; ++q) {                    // snake order balances the triangular task sizes
    const int idx = q * nw + ((q & 1) ? nw - 1 - wi : wi);
    if (idx >= ntask) continue;
    const int rt = nrt - 1 - idx;
    const int irow = kCholNb + 8 * rt + g;                  // panel-relative row held by this lane's row-tile fragment
    const double *rp = Pp + t4 * lda_p + min(irow, rows - 1);
    double rf[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) rf[ks] = -rp[4 * ks * lda_p];
    const int gi = Jp + kCholNb + 8 * rt + 2 * t4;          // accumulator rows gi, gi+1 ...
    const int ctmax = min(min(rt, nct - 1), ct_hi);
    for (int ct = ct_lo; ct <= ctmax; ++ct) {
      const int crow = kCholNb + 8 * ct + g;                // column tile: its columns are rows crow of the panel
      const double *cf = Pp + t4 * lda_p + min(crow, rows - 1);
      const int gc = Jp + crow;                             // ... of column gc
      const bool ok = (gc < r) && (gi < R1);
      const int cbk = min(gc >> 5, 7);
      double *cp = A + tab[cbk] + (gc - cbk * kCholNb) * tab[8 + cbk] + (gi - cbk * kCholNb);
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;        // two accumulator chains (k halves)
      if (ok) { const double2 cv = *reinterpret_cast<const double2 *>(cp); c0 = cv.x; c1 = cv.y; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        dmma884(c0, c1, cf[4 * ks * lda_p], rf[ks], c0, c1);
        dmma884(e0, e1, cf[4 * (ks + 4) * lda_p], rf[ks + 4], e0, e1);
      }
      c0 += e0; c1 += e1;
      if (ok) {
        if (gi + 1 < R1) *reinterpret_cast<double2 *>(cp) = make_double2(c0, c1);
        else cp[0] = c0;
      }
    }
  }
}

// part 0: the whole matrix (frames with r <= kCholResidentMax; larger frames: nothing).  part 1 / 2: the leading / trailing
// block of a split frame (nothing for frames that are not split).
template <int kNewton>
__global__ void __launch_bounds__(kCholSmemThreads, 1) k_cholesky_smem(Layout L, int part) {
  pdl_wait();
  if (L.shadow) pdl_trigger();   // k_gather_y and k_solve_ll start beside this kernel (solve_ll.cuh): every block of it is resident by then
  timeline_mark(L, 3);
  extern __shared__ __align__(128) double sm_d[];
  const int s = L.s0 + blockIdx.x;
  SessionState &st = L.st[s];
  const int r_frame = st.r;
  if (r_frame == 0) return;
  const int split = chol_split(r_frame);
  if (split < 0 || (split == 0) != (part == 0)) return;
  const int row0 = part == 2 ? split : 0;                    // first row / column of this pass's block
  const int r = part == 1 ? split : r_frame - row0;          // its size; row r of the block rides along as "ν"
  const int R1 = r + 1;
  const int nblk = (r + kCholNb - 1) / kCholNb;
  const int sld = L.sld;
  double *Sb = L.Sbuf + (size_t)s * L.rld * sld + (size_t)row0 * sld + row0;
  double *Dinv = L.Dinv + ((size_t)s * (L.rld / kCholNb) + row0 / kCholNb) * kCholNb * kCholNb;
  double *A = sm_d;                                         // packed block columns
  double *cb = sm_d + chol_col_off(R1, nblk);               // [2][32] column broadcast buffer of the panel warp
  double *invd = cb + 64;                                   // [2][32] reciprocals of the diagonal, ping-pong by block parity
  uint64_t *bar = reinterpret_cast<uint64_t *>(invd + 64);  // transaction barrier of the bulk load
  int *tab = reinterpret_cast<int *>(bar + 2);              // [8] block-column offsets, [8] pitches
  double *Xs = reinterpret_cast<double *>(tab + 16);        // [32][kCholXP] inverse of the current diagonal block, X[j][k]
  if (threadIdx.x < 8) { tab[threadIdx.x] = chol_col_off(R1, threadIdx.x); tab[8 + threadIdx.x] = chol_lda(R1, threadIdx.x); }
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT / 32;
  bool bad = false;
#ifdef REKF_CHOL_TIMING
  double *tlog = L.Qd + (size_t)s * L.rcap;                 // per-phase cycle stamps (thread 0)
  int tl = 0;
#define REKF_TSTAMP() do { if (tid == 0) tlog[tl++] = (double)clock64(); } while (0)
#else
#define REKF_TSTAMP() do { } while (0)
#endif
  REKF_TSTAMP();

  // ---- load the lower triangle (+ ν row): one bulk copy per column, rows 32b.. (the few entries above the diagonal
  //      inside the diagonal block come along and are never read) -------------------------------------------------------
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned total = 0;
    for (int b = 0; b < nblk; ++b) total += (unsigned)min(kCholNb, r - b * kCholNb) * (unsigned)((R1 - b * kCholNb + 1) & ~1) * 8u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total) : "memory");
  }
  __syncthreads();
  // this thread's column (r <= 224 < 256 threads): shared-memory home, global home, bytes (a multiple of 16)
  const int cb_b = tid >> 5;
  double *col_s = A + chol_col_off(R1, min(cb_b, nblk - 1)) + (tid & 31) * chol_lda(R1, min(cb_b, nblk - 1));
  double *col_g = Sb + (size_t)tid * sld + cb_b * kCholNb;
  const unsigned col_bytes = (tid < r) ? (unsigned)((R1 - cb_b * kCholNb + 1) & ~1) * 8u : 0u;
  if (col_bytes)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(col_s)), "l"(col_g), "r"(col_bytes), "r"(bar_a) : "memory");
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
  }
  REKF_TSTAMP();

  const int jb_last = r - (nblk - 1) * kCholNb;             // true width of the last block
  for (int b = 0; b < nblk; ++b) {
    if (b == nblk - 1 && !L.shadow) pdl_trigger();          // the last block: the TRSM kernel may be launched (it waits in pdl_wait())
    const int J = b * kCholNb;
    const int lda = chol_lda(R1, b);
    double *P = A + chol_col_off(R1, b);                    // element (J + i, J + c) at P[c * lda + i]
    double *rinv = invd + (b & 1) * kCholNb;
    if (b == nblk - 1 && jb_last < kCholNb) {
      // partial last block → identity-padded full block: ν moves from relative row jb to row 32, rows/columns
      // jb..31 become the identity (all trailing updates into this block column are done: see phase 3 / phase 1)
      if (tid < kCholNb) {
        const int c = tid;
        double *col = P + c * lda;
        const double nu_c = (c < jb_last) ? col[jb_last] : 0.0;
        for (int i = (c < jb_last ? jb_last : 0); i < kCholNb; ++i) col[i] = (i == c) ? 1.0 : 0.0;
        col[kCholNb] = nu_c;
      }
      __syncthreads();
    }
    const int rows = (b == nblk - 1) ? kCholNb + 1 : R1 - J;   // rows of this block column (the last one: block + ν)

    // ---- phase 1: warp 0 factors the diagonal block; meanwhile warp 1 inverts the previous diagonal block (for the
    //      TRSM kernel) and warps 2.. finish the previous panel's trailing update right of block column b ------------
    if (warp == 0) {
      double a[kCholNb], x[kCholNb];                        // row `lane` of the block; column `lane` of X = L_bb⁻¹
#pragma unroll
      for (int jj = 0; jj < kCholNb; ++jj) { a[jj] = (jj <= lane) ? P[jj * lda + lane] : 0.0; x[jj] = (jj == lane) ? 1.0 : 0.0; }
      double d = __shfl_sync(0xffffffffu, a[0], 0);
#pragma unroll
      for (int j = 0; j < kCholNb; ++j) {
        bad |= !(d > 0.0);
        const double inv = pivot_rsqrt<kNewton>(d);
        const double l = (lane == j) ? d * inv : a[j] * inv;
        if (lane == j) rinv[j] = inv;
        if (j + 1 < kCholNb) d = __shfl_sync(0xffffffffu, fma(-l, l, a[j + 1]), j + 1);   // next pivot, early
        cb[(j & 1) * 32 + lane] = l;
        const double xj = x[j] * inv;                        // X[j][lane]: forward substitution of e_lane, step j
        __syncwarp();
        {
          const double2 *cb2 = reinterpret_cast<const double2 *>(cb + (j & 1) * 32);   // column j of L: 16-byte broadcast reads
#pragma unroll
          for (int p = (j + 1) >> 1; p < kCholNb / 2; ++p) {
            const double2 v = cb2[p];
            if (2 * p > j) { a[2 * p] = fma(-l, v.x, a[2 * p]); x[2 * p] = fma(-xj, v.x, x[2 * p]); }
            a[2 * p + 1] = fma(-l, v.y, a[2 * p + 1]);
            x[2 * p + 1] = fma(-xj, v.y, x[2 * p + 1]);
          }
        }
        a[j] = l;
        x[j] = xj;
      }
      double *Dg = Dinv + (size_t)b * kCholNb * kCholNb;     // X[j][c], row-major: what the TRSM kernel consumes
#pragma unroll
      for (int jj = 0; jj < kCholNb; ++jj) {
        if (jj <= lane) P[jj * lda + lane] = a[jj];
        Xs[jj * kCholXP + lane] = x[jj];
        Dg[jj * kCholNb + lane] = x[jj];
      }
    } else if (b > 0) {
      if (L.sync && part == 0 && warp == 1 + (b - 1) % (NW - 1)) {
        // this warp issued block column b-1's copies before phase 3: complete → visible to the generic proxy → released:
        // k_solve_ll may consume block column b-1, X_{b-1} (warp 0's stores, ordered by the barriers since) and ν_{b-1}.
        // Here, not at the end of phase 3: the wait would sit on the critical path there (measured: +6 µs per factorisation)
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) sync_raise(L.sync + (size_t)s * L.sync_n + b - 1);
      }
      chol_trailing(A, tab, A + chol_col_off(R1, b - 1), chol_lda(R1, b - 1), J - kCholNb, R1, r, 4, 1 << 30, warp - 1, NW - 1, lane);
    }
    __syncthreads();
    REKF_TSTAMP();

    // ---- phase 2: rows below the block (incl. ν): P ← P·Xᵀ, i.e. out[i][j] = Σ_{k<=j} P[i][k]·X[j][k], on the fp64 tensor pipe.
    //      Formed transposed like the trailing update — D[m = column j][n = row i] — so that a lane's accumulator pair is two
    //      consecutive rows of one column: one 16-byte access in the column-major panel.  A warp task is one 8-row tile: all of
    //      its 32 old values per row are in registers (B fragments) before the first result is written, so it runs in place ----
    {
      const int g = lane >> 2, t4 = lane & 3;
      const int below = rows - kCholNb;
      const int nrt2 = (below + 7) >> 3;
      for (int rt = warp; rt < nrt2; rt += NW) {
        const int irow = kCholNb + 8 * rt + g;
        const double *bp = P + t4 * lda + min(irow, rows - 1);
        double bf[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) bf[ks] = bp[4 * ks * lda];
        __syncwarp();
        const int r0 = kCholNb + 8 * rt + 2 * t4;            // this lane's two result rows
#pragma unroll
        for (int jt = 0; jt < 4; ++jt) {
          const double *xf = Xs + (8 * jt + g) * kCholXP + t4;
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
          for (int ks = 0; ks < 2 * (jt + 1); ks += 2) {       // X is lower triangular: k < 8(jt+1)
            dmma884(c0, c1, xf[4 * ks], bf[ks], c0, c1);
            dmma884(e0, e1, xf[4 * (ks + 1)], bf[ks + 1], e0, e1);
          }
          c0 += e0; c1 += e1;
          double *op = P + (8 * jt + g) * lda + r0;
          if (r0 + 1 < rows) *reinterpret_cast<double2 *>(op) = make_double2(c0, c1);
          else if (r0 < rows) op[0] = c0;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's panel writes → visible to the bulk engine
    __syncthreads();
    REKF_TSTAMP();
    // block column b of L is final: it goes back to global memory now, in the shadow of the remaining blocks (the last block
    // column waits for the ν fix-up after the loop)
    // (issued by a warp that is never the panel warp: it is also the one that later waits for the copies and tells k_solve_ll)
    if (warp == 1 + b % (NW - 1) && b < nblk - 1) {
      const unsigned bytes = (unsigned)((R1 - J + 1) & ~1) * 8u;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Sb + (size_t)(J + lane) * sld + J),
                   "r"((uint32_t)__cvta_generic_to_shared(P + lane * lda)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }

    // ---- phase 3: the panel's update of block column b+1 only (what the next diagonal block and its rows need);
    //      the rest of the trailing matrix is updated in the shadow of the next factorisation (phase 1) ------------
    chol_trailing(A, tab, P, lda, J, R1, r, 0, 3, warp, NW, lane);
    __syncthreads();
    REKF_TSTAMP();
  }
  if (jb_last < kCholNb && tid < jb_last) {                  // ν of the padded last block back to its own row
    double *col = A + chol_col_off(R1, nblk - 1) + tid * chol_lda(R1, nblk - 1);
    col[jb_last] = col[kCholNb];
  }

  // ---- publish the last block column of L (the others left inside the loop): the same per-column bulk copies, the other way ----
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the panels → visible to the bulk engine
  __syncthreads();
  if (col_bytes && cb_b == nblk - 1) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(col_g), "r"((uint32_t)__cvta_generic_to_shared(col_s)), "r"(col_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's copies (issued here or inside the loop) have landed
  REKF_TSTAMP();
#ifdef REKF_CHOL_TIMING
  if (lane == 0) tlog[64 + warp] = (double)clock64();       // per-warp finish stamps
#endif
  if (bad && lane == 0) atomicOr(&st.flags, FLAG_NOT_SPD);
  if (L.sync && part == 0) {                                 // the last block column (and everything before it) is out
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid == 0) sync_raise(L.sync + (size_t)s * L.sync_n + nblk - 1);
  }
  timeline_mark(L, 9);
}

// ---- split frames, stage 3: S22 −= L21·L21ᵀ for rows / columns r1..r (row r = ν; its column does not exist), lower triangle.
//      L21[i][k] = Sb[k][i] (column-major): a 32-row slice of a column is one contiguous 256-byte read.  grid (tiles, 1, Sg) ----
__global__ void __launch_bounds__(256) k_chol_syrk(Layout L) {
  const int s = L.s0 + blockIdx.z;
  const int r = L.st[s].r;
  const int r1 = chol_split(r);
  if (r1 <= 0) return;
  const int nt = (r + 1 - r1 + 31) / 32;                     // 32-row tiles of the trailing block incl. ν
  int t = blockIdx.x, ti = 0;
  while (t > ti) { t -= ti + 1; ++ti; }                      // tile (ti, tj), tj <= ti
  const int tj = t;
  if (ti >= nt) return;
  const int sld = L.sld;
  double *Sb = L.Sbuf + (size_t)s * L.rld * sld;
  __shared__ double As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // thread: column tj*32 + tx of rows ti*32 + ty + 8u
  const int i0 = r1 + 32 * ti, j0 = r1 + 32 * tj;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < r1; k0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int kk = e >> 5, ii = e & 31;
      As[kk][ii] = (i0 + ii <= r) ? Sb[(size_t)(k0 + kk) * sld + i0 + ii] : 0.0;
      Bs[kk][ii] = (j0 + ii <= r) ? Sb[(size_t)(k0 + kk) * sld + j0 + ii] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const double b = Bs[kk][tx];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(As[kk][ty + 8 * u], b, acc[u]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = i0 + ty + 8 * u, j = j0 + tx;
    if (i <= r && j < r && j <= i) Sb[(size_t)j * sld + i] -= acc[u];
  }
}

}  // namespace rekf
