/*
 * rekf_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See rekf_oracle.h.
 *
 * Restates reference src/reflector_ekf_slam/reflector_ekf_slam.cc function by function; every
 * routine cites the reference lines it follows.  PARITY UNPINNED by the reference (no tests /
 * golden vectors exist there); pinned by known answers, oracle/numpy_ekf.py and the bag replay.
 *
 * Matrices are column-major doubles (Eigen::MatrixXd).  No dependencies beyond libc/libm.
 */
#include "rekf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* dense kernels                                                                               */
/* ------------------------------------------------------------------------------------------ */

#define MR 8
#define NR 4
#define MC 128
#define KC 256
#define NC 1024

typedef double v4d __attribute__((vector_size(32), aligned(8)));

static void *xmalloc(size_t bytes) {
  void *p = NULL;
  if (bytes == 0) bytes = 64;
  if (posix_memalign(&p, 64, bytes) != 0 || !p) {
    fprintf(stderr, "oracle: out of memory (%zu bytes)\n", bytes);
    abort();
  }
  return p;
}
static double *zeros(size_t count) {
  double *p = (double *)xmalloc(count * sizeof(double));
  memset(p, 0, count * sizeof(double));
  return p;
}

/* element (i,k) of op(A) */
static inline double opA(const double *A, int lda, int trans, int i, int k) {
  return trans ? A[(size_t)i * lda + k] : A[(size_t)k * lda + i];
}

/* pack an mc x kc block of op(A) into MR-row panels (zero padded) */
static void pack_a(const double *A, int lda, int trans, int i0, int k0, int mc, int kc, double *Ap) {
  for (int ir = 0; ir < mc; ir += MR) {
    const int mr = mc - ir < MR ? mc - ir : MR;
    for (int k = 0; k < kc; ++k) {
      for (int i = 0; i < mr; ++i) Ap[i] = opA(A, lda, trans, i0 + ir + i, k0 + k);
      for (int i = mr; i < MR; ++i) Ap[i] = 0.0;
      Ap += MR;
    }
  }
}
/* pack a kc x nc block of op(B) into NR-column panels (zero padded) */
static void pack_b(const double *B, int ldb, int trans, int k0, int j0, int kc, int nc, double *Bp) {
  for (int jr = 0; jr < nc; jr += NR) {
    const int nr = nc - jr < NR ? nc - jr : NR;
    for (int k = 0; k < kc; ++k) {
      for (int j = 0; j < nr; ++j) {
        const int kk = k0 + k, jj = j0 + jr + j;
        Bp[j] = trans ? B[(size_t)kk * ldb + jj] : B[(size_t)jj * ldb + kk];
      }
      for (int j = nr; j < NR; ++j) Bp[j] = 0.0;
      Bp += NR;
    }
  }
}

#ifndef __AVX__
/* Baseline x86-64 (SSE2, what the reference's flag-less Release build targets): 16-byte vectors, the
 * 8x4 block as two 4x4 halves so the eight accumulators stay in the sixteen xmm registers. */
typedef double v2d __attribute__((vector_size(16), aligned(8)));
static inline void micro_kernel(int kc, const double *Ap, const double *Bp, double *acc /*MR*NR col-major*/) {
  for (int half = 0; half < 2; ++half) {
    const double *a = Ap + 4 * half, *b = Bp;
    v2d c00 = {0, 0}, c10 = c00, c01 = c00, c11 = c00, c02 = c00, c12 = c00, c03 = c00, c13 = c00;
    for (int k = 0; k < kc; ++k) {
      v2d a0, a1;
      memcpy(&a0, a, 16);
      memcpy(&a1, a + 2, 16);
      const v2d b0 = {b[0], b[0]}, b1 = {b[1], b[1]}, b2 = {b[2], b[2]}, b3 = {b[3], b[3]};
      c00 += a0 * b0; c10 += a1 * b0;
      c01 += a0 * b1; c11 += a1 * b1;
      c02 += a0 * b2; c12 += a1 * b2;
      c03 += a0 * b3; c13 += a1 * b3;
      a += MR;
      b += NR;
    }
    double *o = acc + 4 * half;
    memcpy(o + 0, &c00, 16);  memcpy(o + 2, &c10, 16);
    memcpy(o + 8, &c01, 16);  memcpy(o + 10, &c11, 16);
    memcpy(o + 16, &c02, 16); memcpy(o + 18, &c12, 16);
    memcpy(o + 24, &c03, 16); memcpy(o + 26, &c13, 16);
  }
}
#else
static inline void micro_kernel(int kc, const double *Ap, const double *Bp, double *acc /*MR*NR col-major*/) {
  v4d c00 = {0, 0, 0, 0}, c10 = c00, c01 = c00, c11 = c00, c02 = c00, c12 = c00, c03 = c00, c13 = c00;
  for (int k = 0; k < kc; ++k) {
    v4d a0, a1;
    memcpy(&a0, Ap, 32);
    memcpy(&a1, Ap + 4, 32);
    const double b0 = Bp[0], b1 = Bp[1], b2 = Bp[2], b3 = Bp[3];
    c00 += a0 * b0; c10 += a1 * b0;
    c01 += a0 * b1; c11 += a1 * b1;
    c02 += a0 * b2; c12 += a1 * b2;
    c03 += a0 * b3; c13 += a1 * b3;
    Ap += MR;
    Bp += NR;
  }
  memcpy(acc + 0, &c00, 32);  memcpy(acc + 4, &c10, 32);
  memcpy(acc + 8, &c01, 32);  memcpy(acc + 12, &c11, 32);
  memcpy(acc + 16, &c02, 32); memcpy(acc + 20, &c12, 32);
  memcpy(acc + 24, &c03, 32); memcpy(acc + 28, &c13, 32);
}
#endif

/* C = op(A) op(B); the stand-in for Eigen's GEBP product kernel (blocked, packed, vectorised) */
void oracle_dgemm(int transA, int transB, int M, int N, int K, const double *A, int lda,
                  const double *B, int ldb, double *C, int ldc) {
  for (int j = 0; j < N; ++j) memset(C + (size_t)j * ldc, 0, sizeof(double) * (size_t)M);
  if (M <= 0 || N <= 0 || K <= 0) return;
  double *Ap = (double *)xmalloc(sizeof(double) * (size_t)(MC + MR) * KC);
  double *Bp = (double *)xmalloc(sizeof(double) * (size_t)(NC + NR) * KC);
  double acc[MR * NR];
  for (int jc = 0; jc < N; jc += NC) {
    const int nc = N - jc < NC ? N - jc : NC;
    for (int pc = 0; pc < K; pc += KC) {
      const int kc = K - pc < KC ? K - pc : KC;
      pack_b(B, ldb, transB, pc, jc, kc, nc, Bp);
      for (int ic = 0; ic < M; ic += MC) {
        const int mc = M - ic < MC ? M - ic : MC;
        pack_a(A, lda, transA, ic, pc, mc, kc, Ap);
        for (int jr = 0; jr < nc; jr += NR) {
          const int nr = nc - jr < NR ? nc - jr : NR;
          for (int ir = 0; ir < mc; ir += MR) {
            const int mr = mc - ir < MR ? mc - ir : MR;
            micro_kernel(kc, Ap + (size_t)(ir / MR) * kc * MR, Bp + (size_t)(jr / NR) * kc * NR, acc);
            for (int j = 0; j < nr; ++j) {
              double *c = C + (size_t)(jc + jr + j) * ldc + ic + ir;
              for (int i = 0; i < mr; ++i) c[i] += acc[j * MR + i];
            }
          }
        }
      }
    }
  }
  free(Ap);
  free(Bp);
}

/* Eigen's dynamic-size MatrixXd::inverse() is PartialPivLU(...).inverse(): row-pivoted LU, then the
 * solve against the identity (reflector_ekf_slam.cc:305).  In place; returns -1 on a zero pivot. */
int oracle_lu_inverse(int n, double *A, int lda) {
  if (n <= 0) return 0;
  int *piv = (int *)xmalloc(sizeof(int) * (size_t)n);
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(A[(size_t)k * lda + k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = fabs(A[(size_t)k * lda + i]);
      if (v > best) { best = v; p = i; }
    }
    piv[k] = p;
    if (best == 0.0) { free(piv); return -1; }
    if (p != k)
      for (int j = 0; j < n; ++j) {
        const double t = A[(size_t)j * lda + k];
        A[(size_t)j * lda + k] = A[(size_t)j * lda + p];
        A[(size_t)j * lda + p] = t;
      }
    const double inv = 1.0 / A[(size_t)k * lda + k];
    for (int i = k + 1; i < n; ++i) A[(size_t)k * lda + i] *= inv;
    for (int j = k + 1; j < n; ++j) {
      const double ukj = A[(size_t)j * lda + k];
      if (ukj == 0.0) continue;
      double *cj = A + (size_t)j * lda;
      const double *lk = A + (size_t)k * lda;
      for (int i = k + 1; i < n; ++i) cj[i] -= lk[i] * ukj;
    }
  }
  /* X = U^-1 L^-1 P : solve column by column */
  double *X = zeros((size_t)n * n);
  for (int c = 0; c < n; ++c) {
    double *x = X + (size_t)c * n;
    /* right-hand side = P e_c */
    x[c] = 1.0;
  }
  /* apply the row interchanges to the identity (rows of the rhs matrix) */
  for (int k = 0; k < n; ++k)
    if (piv[k] != k)
      for (int c = 0; c < n; ++c) {
        const double t = X[(size_t)c * n + k];
        X[(size_t)c * n + k] = X[(size_t)c * n + piv[k]];
        X[(size_t)c * n + piv[k]] = t;
      }
  for (int c = 0; c < n; ++c) {
    double *x = X + (size_t)c * n;
    for (int k = 0; k < n; ++k) { /* forward, unit lower */
      const double xk = x[k];
      if (xk == 0.0) continue;
      const double *lk = A + (size_t)k * lda;
      for (int i = k + 1; i < n; ++i) x[i] -= lk[i] * xk;
    }
    for (int k = n - 1; k >= 0; --k) { /* backward, upper */
      const double *uk = A + (size_t)k * lda;
      x[k] /= uk[k];
      const double xk = x[k];
      for (int i = 0; i < k; ++i) x[i] -= uk[i] * xk;
    }
  }
  for (int c = 0; c < n; ++c) memcpy(A + (size_t)c * lda, X + (size_t)c * n, sizeof(double) * (size_t)n);
  free(X);
  free(piv);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* the filter                                                                                  */
/* ------------------------------------------------------------------------------------------ */

struct rekf_oracle {
  int algebra;
  int map_loader;
  int odom_model;
  int use_imu;
  double Qu[9]; /* qu x qu, column-major (diagonal) */
  int qu;       /* 2 (DIFF) or 3 (OMNI) */
  double Qt[4];
  double vt[3];
  /* ekf::State (ekf_slam_interface.h:43-48) */
  double time;
  int n;
  double *mu;
  double *sigma; /* n x n, ld = n */
  /* sensor::Map (sensor_data.h:30-37) */
  int map_n;
  float *map_xy;
  double *map_cov; /* row-major 2x2 per beacon */
  /* ReflectorMatchResult of the last frame (ekf_slam_interface.h:18-26) */
  int n_state, n_map, n_new, match_cap;
  int *state_pairs, *map_pairs, *new_ids;
};

static double wrap_angle(double a) { return atan2(sin(a), cos(a)); }

/* ctor, reflector_ekf_slam.cc:6-37 */
rekf_oracle *oracle_create(const rekf_options *opts, int algebra) {
  rekf_oracle *o = (rekf_oracle *)calloc(1, sizeof(rekf_oracle));
  o->algebra = algebra;
  o->map_loader = opts->map_loader;
  o->odom_model = opts->odom_model;
  o->use_imu = opts->use_imu;
  o->time = opts->init_time;                         /* :8 */
  o->n = 3;
  o->mu = zeros(3);
  memcpy(o->mu, opts->init_pose, sizeof(double) * 3); /* :9 */
  o->sigma = zeros(9);                                /* :10-11 */
  memset(o->Qu, 0, sizeof(o->Qu));
  if (opts->odom_model == REKF_ODOM_DIFF) {           /* :15-19 */
    o->qu = 2;
    o->Qu[0] = opts->linear_velocity_cov;
    o->Qu[3] = opts->angular_velocity_cov;
  } else {                                            /* :20-31 (OMNI and default) */
    o->qu = 3;
    o->Qu[0] = opts->linear_velocity_cov;
    o->Qu[4] = opts->linear_velocity_cov;
    o->Qu[8] = opts->angular_velocity_cov;
  }
  o->Qt[0] = opts->observation_cov;                   /* :33-34 */
  o->Qt[1] = 0.0;
  o->Qt[2] = 0.0;
  o->Qt[3] = opts->observation_cov;
  if (opts->map_path && opts->map_path[0]) oracle_load_map_txt(o, opts->map_path); /* :36 */
  return o;
}

void oracle_destroy(rekf_oracle *o) {
  if (!o) return;
  free(o->mu); free(o->sigma); free(o->map_xy); free(o->map_cov);
  free(o->state_pairs); free(o->map_pairs); free(o->new_ids);
  free(o);
}

int oracle_dim(const rekf_oracle *o) { return o->n; }
double oracle_time(const rekf_oracle *o) { return o->time; }
const double *oracle_mu(const rekf_oracle *o) { return o->mu; }
const double *oracle_sigma(const rekf_oracle *o) { return o->sigma; }

void oracle_set_state(rekf_oracle *o, double time, const double vt[3], const double *mu, int n,
                      const double *sigma, int ld) {
  free(o->mu); free(o->sigma);
  o->n = n;
  o->time = time;
  if (vt) memcpy(o->vt, vt, sizeof(double) * 3);
  o->mu = zeros((size_t)n);
  memcpy(o->mu, mu, sizeof(double) * (size_t)n);
  o->sigma = zeros((size_t)n * n);
  for (int j = 0; j < n; ++j) memcpy(o->sigma + (size_t)j * n, sigma + (size_t)j * ld, sizeof(double) * (size_t)n);
}

void oracle_set_map(rekf_oracle *o, const float *xy, const double *cov2x2, int count) {
  free(o->map_xy); free(o->map_cov);
  o->map_n = count;
  o->map_xy = (float *)xmalloc(sizeof(float) * 2 * (size_t)(count > 0 ? count : 1));
  o->map_cov = zeros(4 * (size_t)(count > 0 ? count : 1));
  if (count > 0) {
    memcpy(o->map_xy, xy, sizeof(float) * 2 * (size_t)count);
    memcpy(o->map_cov, cov2x2, sizeof(double) * 4 * (size_t)count);
  }
}

int oracle_get_map(const rekf_oracle *o, float *xy, double *cov2x2, int cap) {
  const int c = o->map_n < cap ? o->map_n : cap;
  if (xy && c > 0) memcpy(xy, o->map_xy, sizeof(float) * 2 * (size_t)c);
  if (cov2x2 && c > 0) memcpy(cov2x2, o->map_cov, sizeof(double) * 4 * (size_t)c);
  return o->map_n;
}

/* ---- map persistence ---------------------------------------------------------------------- */

/* common.cc:5-16 SplitString: std::getline on the delimiter — empty tokens are kept, a trailing
 * empty token is dropped; then std::stod on every token (:58-62).  std::stod("") throws in the
 * reference (uncaught → abort); here a token that does not parse makes the whole load a no-op. */
static int parse_csv_line(const char *line, double **out, int *count) {
  int cap = 16, n = 0;
  double *v = (double *)xmalloc(sizeof(double) * (size_t)cap);
  const char *p = line;
  while (*p) {
    const char *q = strchr(p, ',');
    size_t len = q ? (size_t)(q - p) : strlen(p);
    char tok[128];
    if (len >= sizeof(tok)) { free(v); return -1; }
    memcpy(tok, p, len);
    tok[len] = 0;
    char *end = NULL;
    const double d = strtod(tok, &end);
    if (end == tok) { free(v); return -1; }
    if (n == cap) {
      cap *= 2;
      double *nv = (double *)xmalloc(sizeof(double) * (size_t)cap);
      memcpy(nv, v, sizeof(double) * (size_t)n);
      free(v);
      v = nv;
    }
    v[n++] = d;
    if (!q) break;
    p = q + 1;
  }
  *out = v;
  *count = n;
  return 0;
}

/* LoadMapFromTxtFile, reflector_ekf_slam.cc:43-95 */
void oracle_load_map_txt(rekf_oracle *o, const char *path) {
  if (!path || !path[0]) return;            /* :45 */
  FILE *f = fopen(path, "r");
  if (!f) return;                           /* :45-46 / :67-72 */
  double *rows[3] = {NULL, NULL, NULL};
  int counts[3] = {0, 0, 0};
  int nrows = 0, bad = 0;
  char *line = NULL;
  size_t cap = 0;
  ssize_t got;
  while ((got = getline(&line, &cap, f)) >= 0) {      /* :52 */
    while (got > 0 && (line[got - 1] == '\n' || line[got - 1] == '\r')) line[--got] = 0;
    if (got == 0) continue;                 /* :55 empty lines skipped */
    if (nrows >= 2) { nrows = 3; break; }   /* more than 2 non-empty lines → :74 rejects */
    if (parse_csv_line(line, &rows[nrows], &counts[nrows]) != 0) { bad = 1; break; }
    ++nrows;
  }
  free(line);
  fclose(f);
  if (bad || nrows != 2 || counts[1] != 2 * counts[0]) { /* :74 */
    free(rows[0]); free(rows[1]);
    return;
  }
  const int M_ = counts[0] / 2;             /* :83 */
  const int C_ = counts[1] / 4;             /* :87 */
  float *xy = (float *)xmalloc(sizeof(float) * 2 * (size_t)(M_ > 0 ? M_ : 1));
  double *cov = zeros(4 * (size_t)(M_ > 0 ? M_ : 1));
  for (int i = 0; i < M_; ++i) {            /* :85 (double → float) */
    xy[2 * i] = (float)rows[0][2 * i];
    xy[2 * i + 1] = (float)rows[0][2 * i + 1];
  }
  for (int i = 0; i < C_ && i < M_; ++i) {
    for (int e = 0; e < 4; ++e) {
      double v;
      if (o->map_loader == REKF_MAP_LOADER_REFERENCE) {
        /* :90 reads result[0] (the positions line) — out-of-range reads are UB there, 0.0 here */
        const int idx = 4 * i + e;
        v = idx < counts[0] ? rows[0][idx] : 0.0;
      } else {
        v = rows[1][4 * i + e];
      }
      cov[4 * i + e] = v;                   /* p << a, b, c, d : row-major fill */
    }
  }
  oracle_set_map(o, xy, cov, M_);           /* :93-94 */
  free(xy); free(cov); free(rows[0]); free(rows[1]);
}

/* Node::SaveReflectorResult, ros_node.cc:75-140 (default ostream formatting = %g, 6 significant digits) */
int oracle_save_map_txt(const rekf_oracle *o, const char *filebase) {
  char path[4096];
  snprintf(path, sizeof(path), "%s.txt", filebase);   /* :80 */
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  const int N = (o->n - 3) / 2;
  for (int i = 0; i < o->map_n; ++i)                   /* :87-97 */
    fprintf(f, i != o->map_n - 1 ? "%g,%g," : "%g,%g", (double)o->map_xy[2 * i], (double)o->map_xy[2 * i + 1]);
  if (o->n > 3) {                                      /* :98-110: the leading comma is unconditional */
    fprintf(f, ",");
    for (int i = 0; i < N; ++i)
      fprintf(f, i != N - 1 ? "%g,%g," : "%g,%g", o->mu[3 + 2 * i], o->mu[4 + 2 * i]);
  }
  fprintf(f, "\n");
  for (int i = 0; i < o->map_n; ++i) {                 /* :112-123 */
    const double *c = o->map_cov + 4 * i;
    fprintf(f, i != o->map_n - 1 ? "%g,%g,%g,%g," : "%g,%g,%g,%g", c[0], c[1], c[2], c[3]);
  }
  if (o->n > 3) {                                      /* :124-137 */
    fprintf(f, ",");
    for (int i = 0; i < N; ++i) {
      const int a = 3 + 2 * i;
      const double s00 = o->sigma[(size_t)a * o->n + a], s01 = o->sigma[(size_t)(a + 1) * o->n + a];
      const double s10 = o->sigma[(size_t)a * o->n + a + 1], s11 = o->sigma[(size_t)(a + 1) * o->n + a + 1];
      fprintf(f, i != N - 1 ? "%g,%g,%g,%g," : "%g,%g,%g,%g", s00, s01, s10, s11);
    }
  }
  fprintf(f, "\n");
  fclose(f);
  return 0;
}

/* ---- motion model -------------------------------------------------------------------------- */

/* The quantities Predict()/PredictState() derive from (mu, vt, dt): reflector_ekf_slam.cc:156-176
 * (DIFF) and :184-200 (OMNI).  g02/g12 are G_xi(0,2), G_xi(1,2); Gu is the top 3 x qu block of G_u
 * (column-major); d is the pose increment. */
typedef struct { double g02, g12; double Gu[9]; double d[3]; } motion_terms;

static motion_terms motion_model(const rekf_oracle *o, const double *mu, double dt) {
  motion_terms t;
  memset(&t, 0, sizeof(t));
  const double vx = o->vt[0], vy = o->vt[1], w = o->vt[2];
  if (o->odom_model == REKF_ODOM_DIFF) {
    const double delta_theta = w * dt;                                   /* :158 */
    const double a = mu[2] + delta_theta / 2;                            /* :165 */
    t.d[0] = vx * dt * cos(mu[2] + delta_theta / 2);                     /* :159 */
    t.d[1] = vx * dt * sin(mu[2] + delta_theta / 2);                     /* :160 */
    t.d[2] = delta_theta;
    t.g02 = -vx * dt * sin(a);                                           /* :167 */
    t.g12 = vx * dt * cos(a);                                            /* :168 */
    /* G_u_2 (3x2), :173-175 */
    t.Gu[0] = dt * cos(a);  t.Gu[3] = -vx * dt * dt * sin(a) / 2;
    t.Gu[1] = dt * sin(a);  t.Gu[4] = vx * dt * dt * cos(a) / 2;
    t.Gu[2] = 0;            t.Gu[5] = dt;
  } else {
    const double th = mu[2];
    t.d[2] = w * dt;                                                     /* :184 */
    t.d[0] = vx * dt * cos(th) - vy * dt * sin(th);                      /* :185 */
    t.d[1] = vx * dt * sin(th) + vy * dt * cos(th);                      /* :186 */
    t.g02 = -vx * dt * sin(th) - vy * dt * cos(th);                      /* :191 */
    t.g12 = vx * dt * cos(th) - vy * dt * sin(th);                       /* :192 */
    /* G_u_2 (3x3), :197-199 */
    t.Gu[0] = dt * cos(th); t.Gu[3] = -dt * sin(th); t.Gu[6] = 0;
    t.Gu[1] = dt * sin(th); t.Gu[4] = dt * cos(th);  t.Gu[7] = 0;
    t.Gu[2] = 0;            t.Gu[5] = 0;             t.Gu[8] = dt;
  }
  return t;
}

/* V = G_u_2 · Qu · G_u_2ᵀ (3x3, column-major): the only non-zero block of G_u·Qu·G_uᵀ (:178/:202) */
static void control_noise(const rekf_oracle *o, const motion_terms *t, double V[9]) {
  const int q = o->qu;
  double T[9] = {0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < q; ++j) {
      double s = 0;
      for (int k = 0; k < q; ++k) s += t->Gu[k * 3 + i] * o->Qu[j * q + k];
      T[j * 3 + i] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < q; ++k) s += T[k * 3 + i] * t->Gu[k * 3 + j];
      V[j * 3 + i] = s;
    }
}

/* sigma_out = G_xi·sigma·G_xiᵀ + G_u·Qu·G_uᵀ (:178 / :202); sigma_out may alias nothing */
static void propagate_covariance(const rekf_oracle *o, const motion_terms *t, const double *sigma, double *out) {
  const int n = o->n;
  double V[9];
  control_noise(o, t, V);
  if (o->algebra == ORACLE_ALGEBRA_AS_WRITTEN) {
    /* dense G_xi = I with two extra entries (:166-168); two n x n x n products, left to right */
    double *G = zeros((size_t)n * n);
    for (int i = 0; i < n; ++i) G[(size_t)i * n + i] = 1.0;
    G[(size_t)2 * n + 0] = t->g02;
    G[(size_t)2 * n + 1] = t->g12;
    double *T1 = zeros((size_t)n * n);
    oracle_dgemm(0, 0, n, n, n, G, n, sigma, n, T1, n);
    oracle_dgemm(0, 1, n, n, n, T1, n, G, n, out, n);
    free(G);
    free(T1);
  } else {
    /* rows 0,1 += g·row 2, then cols 0,1 += g·col 2 — the same sums without the exact zeros */
    memcpy(out, sigma, sizeof(double) * (size_t)n * n);
    for (int c = 0; c < n; ++c) {
      double *col = out + (size_t)c * n;
      const double s2 = col[2];
      col[0] = col[0] + t->g02 * s2;
      col[1] = col[1] + t->g12 * s2;
    }
    double *c0 = out, *c1 = out + n, *c2 = out + (size_t)2 * n;
    for (int r = 0; r < n; ++r) {
      c0[r] = c0[r] + t->g02 * c2[r];
      c1[r] = c1[r] + t->g12 * c2[r];
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[(size_t)j * n + i] += V[j * 3 + i];
}

/* Predict(dt), reflector_ekf_slam.cc:154-206 */
static void predict(rekf_oracle *o, double dt) {
  const motion_terms t = motion_model(o, o->mu, dt);
  double *next = zeros((size_t)o->n * o->n);
  propagate_covariance(o, &t, o->sigma, next);
  free(o->sigma);
  o->sigma = next;
  o->mu[0] += t.d[0];                                 /* :180 / :204 */
  o->mu[1] += t.d[1];
  o->mu[2] += t.d[2];
  o->mu[2] = wrap_angle(o->mu[2]);                    /* :181 / :205 */
}

/* PredictState(time), :97-152 — same formulas on a copy */
void oracle_predict_state(const rekf_oracle *o, double time, double *mu, double *sigma) {
  const double dt = time - o->time;                   /* :100 */
  const motion_terms t = motion_model(o, o->mu, dt);
  if (sigma) propagate_covariance(o, &t, o->sigma, sigma);
  memcpy(mu, o->mu, sizeof(double) * (size_t)o->n);
  mu[0] += t.d[0];
  mu[1] += t.d[1];
  mu[2] += t.d[2];
  mu[2] = wrap_angle(mu[2]);                          /* :126 / :150 */
}

/* HandleOdometryMessage, :208-223 */
void oracle_handle_odometry(rekf_oracle *o, double time, double vx, double vy, double wz) {
  if (time < o->time) return;                         /* :211 */
  if (!o->use_imu) {
    o->vt[0] = vx; o->vt[1] = vy; o->vt[2] = wz;      /* :216 (latched BEFORE predicting) */
    predict(o, time - o->time);                       /* :217-218 */
    o->time = time;                                   /* :219 */
  }
}

/* ---- association --------------------------------------------------------------------------- */

static void ensure_match_capacity(rekf_oracle *o, int m) {
  if (m <= o->match_cap) return;
  free(o->state_pairs); free(o->map_pairs); free(o->new_ids);
  o->match_cap = m;
  o->state_pairs = (int *)xmalloc(sizeof(int) * 2 * (size_t)m);
  o->map_pairs = (int *)xmalloc(sizeof(int) * 2 * (size_t)m);
  o->new_ids = (int *)xmalloc(sizeof(int) * (size_t)m);
}

/* point_transformed_to_global_frame, :389-393 and :327-331: double arithmetic, float32 result */
static void to_global_f32(const double *mu, const float *p, float out[2]) {
  out[0] = (float)((double)p[0] * cos(mu[2]) - (double)p[1] * sin(mu[2]) + mu[0]);
  out[1] = (float)((double)p[0] * sin(mu[2]) + (double)p[1] * cos(mu[2]) + mu[1]);
}

/* ReflectorMatch, :370-455.  The reference sorts {dist, id} pairs with a `<=` comparator (:417,
 * :443 — not a strict weak order, ties unspecified) and takes front(); this restatement takes the
 * minimum with the lowest index on exact ties. */
static void reflector_match(rekf_oracle *o, const float *xy, int m) {
  ensure_match_capacity(o, m);
  o->n_state = o->n_map = o->n_new = 0;
  if (o->n == 3 && o->map_n == 0) {                   /* :379-387 */
    for (int i = 0; i < m; ++i) o->new_ids[o->n_new++] = i;
    return;
  }
  const int M = (o->n - 3) / 2;                       /* :395 */
  const int M_ = o->map_n;                            /* :396 */
  for (int i = 0; i < m; ++i) {
    float g[2];
    to_global_f32(o->mu, xy + 2 * i, g);              /* :399 */
    if (M_ > 0) {                                     /* :401-425 */
      double best = INFINITY;
      int best_j = -1;
      for (int j = 0; j < M_; ++j) {
        const double *S = o->map_cov + 4 * j;         /* row-major 2x2 */
        const float dfx = o->map_xy[2 * j] - g[0];    /* :408 float subtraction */
        const float dfy = o->map_xy[2 * j + 1] - g[1];
        const double dx = (double)dfx, dy = (double)dfy;
        /* (δᵀ·Σ)·δ, Σ not inverted (:411) */
        const double t0 = dx * S[0] + dy * S[2];
        const double t1 = dx * S[1] + dy * S[3];
        const double dist = sqrt(t0 * dx + t1 * dy);
        if (dist < best) { best = dist; best_j = j; }
      }
      if (best_j >= 0 && best < 0.05) {               /* :420 */
        o->map_pairs[2 * o->n_map] = i;
        o->map_pairs[2 * o->n_map + 1] = best_j;
        ++o->n_map;
        continue;
      }
    }
    if (M > 0) {                                      /* :426-451 */
      double best = INFINITY;
      int best_j = -1;
      for (int j = 0; j < M; ++j) {
        const float lx = (float)o->mu[3 + 2 * j], ly = (float)o->mu[4 + 2 * j]; /* :431 */
        const float dfx = g[0] - lx, dfy = g[1] - ly; /* :433 */
        const double dx = (double)dfx, dy = (double)dfy;
        const double dist = sqrt(dx * dx + dy * dy);  /* :437 Euclidean (Mahalanobis commented out) */
        if (dist < best) { best = dist; best_j = j; }
      }
      if (best_j >= 0 && best < 0.6) {                /* :446 */
        o->state_pairs[2 * o->n_state] = i;
        o->state_pairs[2 * o->n_state + 1] = best_j;
        ++o->n_state;
        continue;
      }
    }
    o->new_ids[o->n_new++] = i;                       /* :452 */
  }
}

void oracle_get_match_result(const rekf_oracle *o, int *state_pairs, int *n_state, int *map_pairs,
                             int *n_map, int *new_ids, int *n_new, int cap) {
  if (n_state) *n_state = o->n_state;
  if (n_map) *n_map = o->n_map;
  if (n_new) *n_new = o->n_new;
  if (state_pairs) memcpy(state_pairs, o->state_pairs, sizeof(int) * 2 * (size_t)(o->n_state < cap ? o->n_state : cap));
  if (map_pairs) memcpy(map_pairs, o->map_pairs, sizeof(int) * 2 * (size_t)(o->n_map < cap ? o->n_map : cap));
  if (new_ids) memcpy(new_ids, o->new_ids, sizeof(int) * (size_t)(o->n_new < cap ? o->n_new : cap));
}

/* ---- measurement update -------------------------------------------------------------------- */

/* Measurement rows of one frame: for row-pair k, the 2x3 pose block A_k (row-major), the landmark
 * slot (state landmark id or -1 for a map beacon / the GPS rows), z and z_hat (:248-304). */
typedef struct {
  int rows;        /* 2·MM (+3 with a GPS pose) */
  int MM, M;
  double *A;       /* MM x 6 */
  int *lm;         /* MM */
  double B[4];     /* row-major [[c, s], [-s, c]] (:255) */
  double *innov;   /* rows */
  double *Qdiag;   /* rows: Q is diagonal (block-diag of Qt_, :276/:302; GPS block gps.cc:330-334) */
  int gps;
} meas_t;

static void build_measurements(const rekf_oracle *o, const float *xy, const double *gps_pose, meas_t *z) {
  const int M = o->n_state, M_ = o->n_map, MM = M + M_;
  z->M = M;
  z->MM = MM;
  z->gps = gps_pose != NULL;
  z->rows = 2 * MM + (z->gps ? 3 : 0);
  z->A = zeros(6 * (size_t)(MM > 0 ? MM : 1));
  z->lm = (int *)xmalloc(sizeof(int) * (size_t)(MM > 0 ? MM : 1));
  z->innov = zeros((size_t)z->rows);
  z->Qdiag = zeros((size_t)z->rows);
  const double c = cos(o->mu[2]), s = sin(o->mu[2]);  /* :252-253 */
  z->B[0] = c; z->B[1] = s; z->B[2] = -s; z->B[3] = c; /* :255 */
  for (int k = 0; k < MM; ++k) {
    int local_id, global_id;
    double lx, ly;
    if (k < M) {                                      /* :261-277 */
      local_id = o->state_pairs[2 * k];
      global_id = o->state_pairs[2 * k + 1];
      lx = o->mu[3 + 2 * global_id];
      ly = o->mu[4 + 2 * global_id];
      z->lm[k] = global_id;
    } else {                                          /* :285-303: beacon read as float32, no B block */
      local_id = o->map_pairs[2 * (k - M)];
      global_id = o->map_pairs[2 * (k - M) + 1];
      lx = (double)o->map_xy[2 * global_id];
      ly = (double)o->map_xy[2 * global_id + 1];
      z->lm[k] = -1;
    }
    const double dx = lx - o->mu[0], dy = ly - o->mu[1];      /* :267-268 */
    const double zh0 = dx * c + dy * s;                        /* :269 */
    const double zh1 = -dx * s + dy * c;                       /* :270 */
    double *A = z->A + 6 * k;                                  /* :272-273 */
    A[0] = -c; A[1] = -s; A[2] = -dx * s + dy * c;
    A[3] = s;  A[4] = -c; A[5] = -dx * c - dy * s;
    z->innov[2 * k] = (double)xy[2 * local_id] - zh0;          /* :265, :306 */
    z->innov[2 * k + 1] = (double)xy[2 * local_id + 1] - zh1;
    z->Qdiag[2 * k] = o->Qt[0];
    z->Qdiag[2 * k + 1] = o->Qt[3];
  }
  if (z->gps) { /* reflector_ekf_slam_gps.cc:314-334 */
    const int b = 2 * MM;
    z->innov[b] = gps_pose[0] - o->mu[0];
    z->innov[b + 1] = gps_pose[1] - o->mu[1];
    const double dth = gps_pose[2] - o->mu[2];
    /* dq = (cos(dθ/2), 0, 0, sin(dθ/2)) → angle-axis z component (transform.h:46-70) */
    double qw = cos(dth / 2), qz = sin(dth / 2);
    const double nrm = sqrt(qw * qw + qz * qz);
    qw /= nrm; qz /= nrm;
    if (qw < 0.) { qw = -qw; qz = -qz; }
    const double angle = 2. * atan2(fabs(qz), qw);
    const double scale = angle < 1e-7 ? 2. : angle / sin(angle / 2.);
    z->innov[b + 2] = scale * qz;
    z->Qdiag[b] = 0.05 * 0.05;
    z->Qdiag[b + 1] = 0.05 * 0.05;
    z->Qdiag[b + 2] = 0.017 * 0.017;
  }
}

static void free_measurements(meas_t *z) { free(z->A); free(z->lm); free(z->innov); free(z->Qdiag); }

/* dense H_t (rows x n, column-major), :248, :274-275, :300, gps.cc:314-316 */
static double *dense_H(const rekf_oracle *o, const meas_t *z) {
  const int r = z->rows, n = o->n;
  double *H = zeros((size_t)r * n);
  for (int k = 0; k < z->MM; ++k) {
    const double *A = z->A + 6 * k;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 3; ++b) H[(size_t)b * r + 2 * k + a] = A[a * 3 + b];
    if (z->lm[k] >= 0) {
      const int col = 3 + 2 * z->lm[k];
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) H[(size_t)(col + b) * r + 2 * k + a] = z->B[a * 2 + b];
    }
  }
  if (z->gps)
    for (int b = 0; b < 3; ++b) H[(size_t)b * r + 2 * z->MM + b] = 1.0;
  return H;
}

/* K_t = Σ·Hᵀ·(H·Σ·Hᵀ + Q)⁻¹ evaluated the way Eigen evaluates the expression at :305
 * ((Σ·Hᵀ) · inverse((H·Σ)·Hᵀ + Q)).  Returns n x r. */
static double *dense_gain(const rekf_oracle *o, const double *H, const meas_t *z) {
  const int r = z->rows, n = o->n;
  double *SHt = zeros((size_t)n * r);
  oracle_dgemm(0, 1, n, r, n, o->sigma, n, H, r, SHt, n);        /* Σ·Hᵀ */
  double *HS = zeros((size_t)r * n);
  oracle_dgemm(0, 0, r, n, n, H, r, o->sigma, n, HS, r);         /* H·Σ */
  double *S = zeros((size_t)r * r);
  oracle_dgemm(0, 1, r, r, n, HS, r, H, r, S, r);                /* ·Hᵀ */
  for (int i = 0; i < r; ++i) S[(size_t)i * r + i] += z->Qdiag[i];
  oracle_lu_inverse(r, S, r);                                    /* .inverse() */
  double *K = zeros((size_t)n * r);
  oracle_dgemm(0, 0, n, r, r, SHt, n, S, r, K, n);
  free(SHt); free(HS); free(S);
  return K;
}

/* Y = H·Σ without the zeros of H: row-pair k = A_k·Σ[0:3,:] + B·Σ[3+2j:5+2j,:] (:274-275), GPS rows =
 * Σ[0:3,:].  Y is rows x n column-major. */
static double *structured_HSigma(const rekf_oracle *o, const meas_t *z) {
  const int r = z->rows, n = o->n;
  double *Y = zeros((size_t)r * n);
  for (int c = 0; c < n; ++c) {
    const double *col = o->sigma + (size_t)c * n;
    double *y = Y + (size_t)c * r;
    for (int k = 0; k < z->MM; ++k) {
      const double *A = z->A + 6 * k;
      double y0 = A[0] * col[0] + A[1] * col[1] + A[2] * col[2];
      double y1 = A[3] * col[0] + A[4] * col[1] + A[5] * col[2];
      if (z->lm[k] >= 0) {
        const int j = 3 + 2 * z->lm[k];
        y0 += z->B[0] * col[j] + z->B[1] * col[j + 1];
        y1 += z->B[2] * col[j] + z->B[3] * col[j + 1];
      }
      y[2 * k] = y0;
      y[2 * k + 1] = y1;
    }
    if (z->gps)
      for (int b = 0; b < 3; ++b) y[2 * z->MM + b] = col[b];
  }
  return Y;
}

/* the measurement update, :305-308 */
static void measurement_update(rekf_oracle *o, const meas_t *z) {
  const int r = z->rows, n = o->n;
  if (o->algebra == ORACLE_ALGEBRA_AS_WRITTEN) {
    double *H = dense_H(o, z);
    /* (i) mu += K_t·(zt − zt_hat), :306 — first evaluation of the lazy K_t */
    double *K = dense_gain(o, H, z);
    for (int j = 0; j < r; ++j) {
      const double v = z->innov[j];
      const double *kj = K + (size_t)j * n;
      for (int i = 0; i < n; ++i) o->mu[i] += kj[i] * v;
    }
    free(K);
    o->mu[2] = wrap_angle(o->mu[2]);                             /* :307 */
    /* (ii) sigma = sigma − (K_t·H_t)·sigma, :308 — K_t evaluated again, then two dense products */
    K = dense_gain(o, H, z);
    double *KH = zeros((size_t)n * n);
    oracle_dgemm(0, 0, n, n, r, K, n, H, r, KH, n);
    double *KHS = zeros((size_t)n * n);
    oracle_dgemm(0, 0, n, n, n, KH, n, o->sigma, n, KHS, n);
    for (size_t e = 0; e < (size_t)n * n; ++e) o->sigma[e] -= KHS[e];
    free(K); free(KH); free(KHS); free(H);
  } else {
    double *Y = structured_HSigma(o, z);                         /* H·Σ (= (Σ·Hᵀ)ᵀ up to Σ's asymmetry) */
    /* Σ·Hᵀ needs the columns of Σ weighted by H's rows; Σ is not exactly symmetric in floating
     * point, so form it from Σ's rows the way the dense product would */
    double *SHt = zeros((size_t)n * r);
    for (int k = 0; k < z->MM; ++k) {
      const double *A = z->A + 6 * k;
      for (int a = 0; a < 2; ++a) {
        double *out = SHt + (size_t)(2 * k + a) * n;
        for (int b = 0; b < 3; ++b) {
          const double h = A[a * 3 + b];
          const double *col = o->sigma + (size_t)b * n;
          for (int i = 0; i < n; ++i) out[i] += col[i] * h;
        }
        if (z->lm[k] >= 0)
          for (int b = 0; b < 2; ++b) {
            const double h = z->B[a * 2 + b];
            const double *col = o->sigma + (size_t)(3 + 2 * z->lm[k] + b) * n;
            for (int i = 0; i < n; ++i) out[i] += col[i] * h;
          }
      }
    }
    if (z->gps)
      for (int b = 0; b < 3; ++b) memcpy(SHt + (size_t)(2 * z->MM + b) * n, o->sigma + (size_t)b * n, sizeof(double) * (size_t)n);
    /* S = (H·Σ)·Hᵀ + Q */
    double *S = zeros((size_t)r * r);
    for (int l = 0; l < z->MM; ++l) {
      const double *A = z->A + 6 * l;
      for (int a = 0; a < 2; ++a) {
        double *out = S + (size_t)(2 * l + a) * r;
        for (int b = 0; b < 3; ++b) {
          const double h = A[a * 3 + b];
          const double *ycol = Y + (size_t)b * r;
          for (int i = 0; i < r; ++i) out[i] += ycol[i] * h;
        }
        if (z->lm[l] >= 0)
          for (int b = 0; b < 2; ++b) {
            const double h = z->B[a * 2 + b];
            const double *ycol = Y + (size_t)(3 + 2 * z->lm[l] + b) * r;
            for (int i = 0; i < r; ++i) out[i] += ycol[i] * h;
          }
      }
    }
    if (z->gps)
      for (int b = 0; b < 3; ++b) {
        double *out = S + (size_t)(2 * z->MM + b) * r;
        const double *ycol = Y + (size_t)b * r;
        for (int i = 0; i < r; ++i) out[i] += ycol[i];
      }
    for (int i = 0; i < r; ++i) S[(size_t)i * r + i] += z->Qdiag[i];
    oracle_lu_inverse(r, S, r);
    double *K = zeros((size_t)n * r);
    oracle_dgemm(0, 0, n, r, r, SHt, n, S, r, K, n);
    for (int j = 0; j < r; ++j) {
      const double v = z->innov[j];
      const double *kj = K + (size_t)j * n;
      for (int i = 0; i < n; ++i) o->mu[i] += kj[i] * v;
    }
    o->mu[2] = wrap_angle(o->mu[2]);
    /* Σ −= K·(H·Σ): identical to (K·H)·Σ in exact arithmetic, n·n·r instead of n³ */
    double *KY = zeros((size_t)n * n);
    oracle_dgemm(0, 0, n, n, r, K, n, Y, r, KY, n);
    for (size_t e = 0; e < (size_t)n * n; ++e) o->sigma[e] -= KY[e];
    free(Y); free(SHt); free(S); free(K); free(KY);
  }
}

/* state augmentation, :311-364 */
static void augment(rekf_oracle *o, const float *xy) {
  const int N2 = o->n_new, N = o->n, Me = N + 2 * N2;           /* :311, :316 */
  double *mu2 = zeros((size_t)Me);
  memcpy(mu2, o->mu, sizeof(double) * (size_t)N);               /* :318 */
  double *sig2 = zeros((size_t)Me * Me);                        /* :320 */
  for (int j = 0; j < N; ++j) memcpy(sig2 + (size_t)j * Me, o->sigma + (size_t)j * N, sizeof(double) * (size_t)N); /* :321 */
  const double s = sin(o->mu[2]), c = cos(o->mu[2]);            /* :323-324 */
  double *Gp = zeros((size_t)2 * N2 * 3);                       /* (2N2) x 3 column-major, :332 */
  const int R = 2 * N2;
  for (int i = 0; i < N2; ++i) {
    const int local_id = o->new_ids[i];
    float g[2];
    to_global_f32(o->mu, xy + 2 * local_id, g);                 /* :339 */
    mu2[N + 2 * i] = (double)g[0];                              /* :341-342 */
    mu2[N + 2 * i + 1] = (double)g[1];
    const double rx = (double)xy[2 * local_id], ry = (double)xy[2 * local_id + 1]; /* :344-345 */
    Gp[0 * R + 2 * i] = 1.;  Gp[1 * R + 2 * i] = 0.;  Gp[2 * R + 2 * i] = -rx * s - ry * c;      /* :347 */
    Gp[0 * R + 2 * i + 1] = 0.;  Gp[1 * R + 2 * i + 1] = 1.;  Gp[2 * R + 2 * i + 1] = rx * c - ry * s;
  }
  /* G_z·Qt·G_zᵀ with G_z the same rotation stacked N2 times (:349, :354): every 2x2 block of the
   * (2N2)x(2N2) result — off-diagonal ones included — equals G_zi·Qt·G_ziᵀ */
  double GQG[4];
  {
    const double Gz[4] = {c, -s, s, c}; /* row-major [[c,-s],[s,c]] (:326) */
    double T[4];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) T[a * 2 + b] = Gz[a * 2 + 0] * o->Qt[0 * 2 + b] + Gz[a * 2 + 1] * o->Qt[1 * 2 + b];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) GQG[a * 2 + b] = T[a * 2 + 0] * Gz[b * 2 + 0] + T[a * 2 + 1] * Gz[b * 2 + 1];
  }
  /* sigma_mm = G_p·Σ_xx·G_pᵀ + G_z·Qt·G_zᵀ (:354) */
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < R; ++j) {
      double acc = 0;
      for (int b = 0; b < 3; ++b) {
        double t = 0;
        for (int a = 0; a < 3; ++a) t += Gp[a * R + i] * o->sigma[(size_t)b * N + a];
        acc += t * Gp[b * R + j];
      }
      sig2[(size_t)(N + j) * Me + N + i] = acc + GQG[(i & 1) * 2 + (j & 1)];              /* :358 */
    }
  /* sigma_mx = G_fx·Σ (:355): rows of G_fx are [Gp_i, 0] → 2N2 x N.  As written this is a dense
   * product evaluated twice (lazy `auto`, used at :356 and :357); the second use transposes it. */
  if (o->algebra == ORACLE_ALGEBRA_AS_WRITTEN) {
    double *Gfx = zeros((size_t)R * N);
    for (int b = 0; b < 3; ++b) memcpy(Gfx + (size_t)b * R, Gp + (size_t)b * R, sizeof(double) * (size_t)R);
    double *mx = zeros((size_t)R * N);
    for (int pass = 0; pass < 2; ++pass) {
      oracle_dgemm(0, 0, R, N, N, Gfx, R, o->sigma, N, mx, R);
      for (int cidx = 0; cidx < N; ++cidx)
        for (int i = 0; i < R; ++i) {
          if (pass == 0) sig2[(size_t)cidx * Me + N + i] = mx[(size_t)cidx * R + i];     /* :356 */
          else sig2[(size_t)(N + i) * Me + cidx] = mx[(size_t)cidx * R + i];             /* :357 */
        }
    }
    free(Gfx); free(mx);
  } else {
    for (int cidx = 0; cidx < N; ++cidx) {
      const double *col = o->sigma + (size_t)cidx * N;
      for (int i = 0; i < R; ++i) {
        const double v = Gp[0 * R + i] * col[0] + Gp[1 * R + i] * col[1] + Gp[2 * R + i] * col[2];
        sig2[(size_t)cidx * Me + N + i] = v;
        sig2[(size_t)(N + i) * Me + cidx] = v;
      }
    }
  }
  free(Gp);
  free(o->mu); free(o->sigma);
  o->mu = mu2;                                                  /* :360-363 */
  o->sigma = sig2;
  o->n = Me;
}

/* HandleObservationMessage, :229-368 (and reflector_ekf_slam_gps.cc:305-340 when gps_pose != NULL) */
void oracle_handle_observation(rekf_oracle *o, double time, const float *xy, int m, const double *gps_pose) {
  predict(o, time - o->time);                         /* :232-233, no sign check on dt */
  o->time = time;                                     /* :234 */
  if (m <= 0) {                                       /* :235 */
    o->n_state = o->n_map = o->n_new = 0;
    return;
  }
  reflector_match(o, xy, m);                          /* :237 */
  if (o->n_state + o->n_map > 0) {                    /* :246 */
    meas_t z;
    build_measurements(o, xy, gps_pose, &z);
    measurement_update(o, &z);
    free_measurements(&z);
  }
  if (o->n_new > 0) augment(o, xy);                   /* :312 */
}
