/*
 * rekf_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of ekf::ReflectorEKFSLAM
 * (reference src/reflector_ekf_slam/reflector_ekf_slam.cc, whole file) used only as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing in
 * reflector_ekf_slam_b200/ may link, import or call it.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, golden vectors or known-answer
 * files for this path and cannot be compiled here (Eigen/ROS/glog absent).
 * The oracle is pinned instead by (1) hand-derived known answers, (2) an independently written
 * numpy float64 EKF (oracle/numpy_ekf.py), (3) replay of the reference's shipped rosbag
 * (tests/golden/bag_stream.npz) landing on the 7 landmarks of dataset/bag_2d_*.png.
 *
 * Two algebra modes, both fp64, column-major like Eigen::MatrixXd:
 *   ORACLE_ALGEBRA_AS_WRITTEN  — every dense product the reference forms, in the reference's
 *       association order: dense G_xi·Σ·G_xiᵀ (:178/:202), dense H_t/Q (:248-304), K_t evaluated
 *       twice because it is a lazy `const auto` (:305-308), PartialPivLU inverse, (K·H)·Σ (:308),
 *       G_fx·Σ evaluated twice (:355-357).  This is the CPU baseline that gets timed.
 *   ORACLE_ALGEBRA_STRUCTURED  — the same equations with the exact-zero work removed (O(n) predict,
 *       H·Σ by row gathers, Σ −= K·(HΣ)); used as the checker at sizes where as-written takes
 *       minutes per step.  tests/test_oracle.py holds the two modes to 1e-12 of each other.
 */
#ifndef REKF_ORACLE_H
#define REKF_ORACLE_H

#include "../include/rekf.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_ALGEBRA_AS_WRITTEN = 0, ORACLE_ALGEBRA_STRUCTURED = 1 };

typedef struct rekf_oracle rekf_oracle;

/* Uses the EKFOptions part of rekf_options plus map_loader; engine-only fields are ignored. */
rekf_oracle *oracle_create(const rekf_options *opts, int algebra);
void oracle_destroy(rekf_oracle *o);

void oracle_handle_odometry(rekf_oracle *o, double time, double vx, double vy, double wz);
/* gps_pose: NULL, or (x, y, yaw) — the USE_GPS variant's pose pseudo-measurement
 * (reflector_ekf_slam_gps.cc:305-340) */
void oracle_handle_observation(rekf_oracle *o, double time, const float *xy, int m,
                               const double *gps_pose);

int oracle_dim(const rekf_oracle *o);
double oracle_time(const rekf_oracle *o);
const double *oracle_mu(const rekf_oracle *o);
const double *oracle_sigma(const rekf_oracle *o); /* n x n column-major, ld = n */
void oracle_get_match_result(const rekf_oracle *o, int *state_pairs, int *n_state, int *map_pairs,
                             int *n_map, int *new_ids, int *n_new, int cap);
/* PredictState (:97-152): mu n doubles, sigma n x n (ld = n) */
void oracle_predict_state(const rekf_oracle *o, double time, double *mu, double *sigma);
void oracle_set_state(rekf_oracle *o, double time, const double vt[3], const double *mu, int n,
                      const double *sigma, int ld);
void oracle_set_map(rekf_oracle *o, const float *xy, const double *cov2x2, int count);
int oracle_get_map(const rekf_oracle *o, float *xy, double *cov2x2, int cap);
void oracle_load_map_txt(rekf_oracle *o, const char *path);
int oracle_save_map_txt(const rekf_oracle *o, const char *filebase);

/* C = op(A)·op(B), column-major, overwrite; exposed for the unit test of the blocked kernel */
void oracle_dgemm(int transA, int transB, int M, int N, int K, const double *A, int lda,
                  const double *B, int ldb, double *C, int ldc);
/* in-place inverse through partial-pivot LU (what Eigen's dynamic .inverse() does); 0 on success */
int oracle_lu_inverse(int n, double *A, int lda);

#ifdef __cplusplus
}
#endif
#endif
