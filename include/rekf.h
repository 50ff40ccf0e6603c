/*
 * rekf.h — C ABI of the B200-native reflector EKF-SLAM engine (librekf_b200.so).
 *
 * This is the drop-in boundary for the one hot path of ShihanWang/reflector_ekf_slam:
 * ekf::ReflectorEKFSLAM::{Predict, ReflectorMatch, HandleOdometryMessage,
 * HandleObservationMessage} (reference src/reflector_ekf_slam/reflector_ekf_slam.cc:154-455)
 * behind ekf::ReflectorEKFSLAMInterface (reference include/reflector_ekf_slam/ekf_slam_interface.h:50-67).
 *
 * Plain C types only: no Eigen, no torch, no CUDA types in any signature.  Every entry point
 * names the reference interface it replaces.  All functions return 0 on success or a negative
 * rekf_status; the message for the last failure is available through rekf_last_error().
 * The reference's own error convention is LOG(ERROR)+exit(-1) (reflector_ekf_slam.cc:376-377);
 * this library never exits the process.
 *
 * A handle owns one *batch* of S >= 1 independent EKF sessions that advance through the same
 * kernel launches (grid.z = session).  The reference-shaped calls (rekf_handle_odometry, ...)
 * address session 0 of a batch created with rekf_create() (S = 1).  A handle is not
 * thread-safe (the reference class has no internal locking either, ros_node.cc:637); calls are
 * asynchronous on the handle's CUDA stream, getters synchronise.
 *
 * Layout conventions at this boundary are the reference's: the state vector is
 * [x, y, theta, l0x, l0y, l1x, l1y, ...] (n = 3 + 2N doubles) and the covariance is an
 * n x n column-major double matrix, exactly Eigen::VectorXd / Eigen::MatrixXd
 * (ekf_slam_interface.h:43-48).  Observations are float32 (x, y) pairs in base_link
 * (sensor/sensor_data.h:15,20-28).
 */
#ifndef REKF_H
#define REKF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define REKF_API
#else
#define REKF_API __attribute__((visibility("default")))
#endif

typedef struct rekf_handle rekf_handle;

typedef enum rekf_status {
  REKF_OK = 0,
  REKF_ERR_BAD_ARGUMENT = -1,
  REKF_ERR_CUDA = -2,            /* a CUDA runtime/driver call failed (no CPU fallback exists) */
  REKF_ERR_CAPACITY = -3,        /* more landmarks / observations than the handle was created for */
  REKF_ERR_NOT_SPD = -4,         /* innovation matrix S lost positive definiteness on the device */
  REKF_ERR_IO = -5,              /* map file could not be read / written */
  REKF_ERR_NO_DEVICE = -6,       /* no sm_100 device visible */
  REKF_ERR_UNSUPPORTED = -7
} rekf_status;

/* sensor::OdometryModel (sensor/sensor_data.h:56-60) */
enum { REKF_ODOM_DIFF = 0, REKF_ODOM_OMNI = 1 };

/* How Sigma <- Sigma - K H Sigma (reflector_ekf_slam.cc:308) is evaluated on the device. */
enum {
  REKF_COV_TCGEN05_TF32X3 = 0,   /* tcgen05.mma kind::tf32, 3-term split, fp32 TMEM accumulate (fp32-class: drifts) */
  REKF_COV_SIMT_F64 = 1,         /* fp64 CUDA-core SYRK (reference-accuracy mode) */
  REKF_COV_TCGEN05_I8X4 = 2      /* tcgen05.mma kind::i8 over four exact 7-bit digit slices, s32 TMEM accumulate:
                                    integer-exact products; frames whose downdate cancels deeply are routed to the
                                    fp64 SYRK on the device.  DEFAULT. */
};

/* Loader for the 2-line landmark map file (reflector_ekf_slam.cc:43-95). */
enum {
  REKF_MAP_LOADER_FIXED = 0,      /* covariances read from line 2 (what the author meant) */
  REKF_MAP_LOADER_REFERENCE = 1   /* reads them from line 1 like :87-91; indices past the end read 0.0 */
};

/*
 * ekf::EKFOptions verbatim (ekf_slam_interface.h:28-41) followed by engine-only fields.
 * The three *_cov fields are variances (the node squares the sigmas, ros_node.cc:207-237).
 */
typedef struct rekf_options {
  /* --- ekf::EKFOptions --- */
  int use_imu;                   /* always false in the reference node (ros_node.cc:186) */
  double init_time;
  double init_pose[3];           /* x, y, yaw */
  const char *map_path;          /* may be NULL / "" : pure SLAM mode */
  int odom_model;                /* REKF_ODOM_DIFF | REKF_ODOM_OMNI */
  double linear_velocity_cov;
  double angular_velocity_cov;
  double observation_cov;
  /* --- engine-only --- */
  int max_landmarks;             /* capacity N_cap of the state (landmarks); default 1024 if 0 */
  int max_observations;          /* capacity of one observation frame; default 128 if 0.  Bounded by the shared memory of
                                    the Cholesky / TRSM kernels: about 400; rekf_create fails with REKF_ERR_CAPACITY above */
  int max_map_landmarks;         /* capacity of the pre-loaded beacon map; default 1024 if 0 */
  int device;                    /* CUDA device ordinal */
  int cov_update;                /* REKF_COV_* */
  int map_loader;                /* REKF_MAP_LOADER_* */
  void *stream;                  /* optional cudaStream_t to run on (NULL: the handle creates one) */
  int use_graphs;                /* 1: replay steps through CUDA graphs where possible */
  int pipeline_groups;           /* batch handles only: split the S sessions into this many groups, each advancing on
                                    its own stream, so that one group's latency-bound kernels (association, Cholesky)
                                    overlap another group's bandwidth-bound ones (TRSM, covariance SYRK).  Sessions
                                    never interact, so results are bit-identical to 1.  0 or 1: one group — the fast
                                    setting: with one group the triangular solve runs beside the Cholesky on a
                                    programmatic-launch chain, which covers the same idle time without a second
                                    group competing for the SMs. */
  int syrk_reserve_sms;          /* with pipeline_groups > 1: SMs the persistent covariance SYRK leaves free for the
                                    other groups' kernels (0: default) */
} rekf_options;

/* Fill with the reference defaults (launch/slam.launch:21-23 sigmas squared, DIFF model). */
REKF_API void rekf_default_options(rekf_options *opts);

/* ---- lifetime ------------------------------------------------------------------------- */
/* ReflectorEKFSLAM(const EKFOptions&) (reflector_ekf_slam.cc:6-37), incl. map load (:36). */
REKF_API int rekf_create(const rekf_options *opts, rekf_handle **out);
/* Same, for a batch of `sessions` independent filters sharing every launch. */
REKF_API int rekf_create_batch(const rekf_options *opts, int sessions, rekf_handle **out);
/* ~ReflectorEKFSLAM (:39-41) */
REKF_API int rekf_destroy(rekf_handle *h);
REKF_API const char *rekf_last_error(const rekf_handle *h);
REKF_API const char *rekf_version(void);
REKF_API int rekf_sessions(const rekf_handle *h);

/* ---- the hot path ----------------------------------------------------------------------- */
/* HandleOdometryMessage(const sensor::OdometryData&) (:208-223): uses time, linear_velocity.x/.y,
 * angular_velocity.z (:216).  Stale messages (time < state time) are a silent no-op (:211). */
REKF_API int rekf_handle_odometry(rekf_handle *h, double time, double vx, double vy, double wz);
/* HandleObservationMessage(const sensor::Observation&) (:229-368): predict to `time`, associate,
 * update, augment.  xy = m float32 pairs in base_link; m may be 0 (predict only, :235).
 * gps_pose_or_null: NULL, or the (x, y, yaw) pose of the USE_GPS variant (ekf::ReflectorEKFSLAMGPS,
 * reflector_ekf_slam_gps.cc:305-340): three extra measurement rows on the pose, applied when the frame has matches. */
REKF_API int rekf_handle_observation(rekf_handle *h, double time, const float *xy, int m,
                                     const double *gps_pose_or_null);
/* HandleImuMessage (:224-227) is empty in every reference implementation. */
REKF_API int rekf_handle_imu(rekf_handle *h, double time);

/* Batched forms: one message per session, arrays indexed by session.
 * odom: S x 4 doubles (time, vx, vy, wz).  obs: times S doubles, xy S x m_stride x 2 floats,
 * counts S ints (each <= m_stride <= max_observations). */
REKF_API int rekf_batch_handle_odometry(rekf_handle *h, const double *odom);
REKF_API int rekf_batch_handle_observation(rekf_handle *h, const double *times, const float *xy,
                                           const int *counts, int m_stride);

/* One whole step per session in one call: HandleOdometryMessage(odom[s]) followed by HandleObservationMessage(times[s],
 * xy[s]) — the message pair SURVEY.md §8(d) defines as a step.  Same results as the two calls above; the pair travels
 * in one host-to-device copy and, with use_graphs, runs as one CUDA graph per pipeline group. */
REKF_API int rekf_batch_handle_step(rekf_handle *h, const double *odom, const double *times, const float *xy,
                                    const int *counts, int m_stride);

/* Whole-sequence replay from DEVICE-resident streams (no host traffic inside): for t in [0,T):
 * HandleOdometryMessage(odom[s][t]) then HandleObservationMessage(obs_time[s][t], obs_xy[s][t]).
 * d_odom: S x T x 4 doubles; d_obs_time: S x T doubles; d_obs_xy: S x T x m x 2 floats;
 * d_pose_out (may be NULL): S x T x 3 doubles, the pose after each step.  All device pointers. */
REKF_API int rekf_replay_device(rekf_handle *h, const void *d_odom, const void *d_obs_time,
                                const void *d_obs_xy, int T, int m, void *d_pose_out);

/* ---- accessors (synchronise) -------------------------------------------------------------- */
/* state_.mu.rows() (= 3 + 2N) */
REKF_API int rekf_dim(rekf_handle *h, int session);
/* GetLatestTime() (reflector_ekf_slam.h:33-36) */
REKF_API int rekf_time(rekf_handle *h, int session, double *time_out);
/* GetStateVector() (reflector_ekf_slam.h:25-28): copies min(n, cap) doubles, returns n via *n_out. */
REKF_API int rekf_get_mu(rekf_handle *h, int session, double *mu, int cap, int *n_out);
/* pose + its 3x3 covariance block (what ros_node.cc:802-817 publishes); cov33 column-major. */
REKF_API int rekf_get_pose(rekf_handle *h, int session, double pose[3], double cov33[9]);
/* poses of all sessions in one read: S x 3 doubles (the per-step result of a batched replay) */
REKF_API int rekf_batch_get_pose(rekf_handle *h, double *poses);
/* Streaming form of the per-step result read for batch replay: request_poses enqueues, behind the work already
 * issued, a device-to-host copy of every session's pose into a pinned ring owned by the handle and returns a ticket
 * without waiting; fetch_poses waits for that ticket only and copies the S x 3 doubles out.  At most 64 tickets may
 * be outstanding.  (The node's GetState() right after each message — ros_node.cc:478,515,592,638 — is
 * rekf_get_pose / rekf_batch_get_pose; this pair is for consumers that can lag a few messages behind.) */
REKF_API int rekf_batch_request_poses(rekf_handle *h, int64_t *ticket_out);
REKF_API int rekf_batch_fetch_poses(rekf_handle *h, int64_t ticket, double *poses);
/* landmark means + diagonal 2x2 blocks (what ros_node.cc:97-136,739-765 reads); cov row-major
 * (0,0),(0,1),(1,0),(1,1) per landmark like the save format. */
REKF_API int rekf_get_landmarks(rekf_handle *h, int session, double *xy, double *cov2x2, int cap,
                                int *count_out);
/* Node::ReflectorToRosMarkers (ros_node.cc:736-789) evaluated on the device: 5 doubles per landmark — x, y, ellipse
 * angle, x_len, y_len of the 95 % covariance ellipse (2*sqrt(eigenvalue*5.991), :763-764), major axis first — so the node
 * can publish its markers without pulling Sigma to the host.  Copies min(N, cap) landmarks, returns N via *count_out. */
REKF_API int rekf_get_markers(rekf_handle *h, int session, double *markers, int cap, int *count_out);
/* GetCoviarance() (reflector_ekf_slam.h:29-32): full n x n, column-major, leading dimension ld >= n. */
REKF_API int rekf_get_sigma(rekf_handle *h, int session, double *sigma, int ld);
/* GetState() (reflector_ekf_slam.h:37-40; the node calls it after every message, ros_node.cc:478,515,592,638) in one
 * stream synchronisation: time, mu (n doubles), sigma (n x n column-major, ld >= n; may be NULL) and the sticky device
 * flags (*flags_out != 0: capacity overflow / S not SPD / tensor-kernel timeout — see rekf_sync).  n_expect is the
 * dimension the caller's buffers are sized for; if the state has a different dimension nothing is copied, *n_out
 * reports it and the call returns REKF_ERR_CAPACITY (resize and repeat). */
REKF_API int rekf_get_state(rekf_handle *h, int session, int n_expect, double *time_out, double *mu, double *sigma,
                            int ld, int *n_out, int *flags_out);
/* Page-lock / release a caller-owned host buffer that rekf_get_state / rekf_get_sigma copy into (the adapter's
 * covariance mirror): device-to-host copies then run at PCIe speed.  Optional. */
REKF_API int rekf_host_register(rekf_handle *h, void *ptr, size_t bytes);
REKF_API int rekf_host_unregister(rekf_handle *h, void *ptr);
/* Cumulative per-session counters since creation / rekf_set_state: out[0] frames with an update, out[1] of those taken
 * by the fp64 SYRK as a whole (int8 mode: deep cancellation), out[2] flagged slots redone in fp64 in the other frames. */
REKF_API int rekf_get_counters(rekf_handle *h, int session, int64_t out[3]);
/* ReflectorMatchResult of the last observation frame (ekf_slam_interface.h:18-26); pairs are
 * {observation index, landmark index} as the reference stores them (:422, :448).  Each array may be
 * NULL; cap is the capacity (in pairs / ids) of every non-NULL array. */
REKF_API int rekf_get_match_result(rekf_handle *h, int session, int *state_pairs, int *n_state,
                                   int *map_pairs, int *n_map, int *new_ids, int *n_new, int cap);
/* PredictState(time) (:97-152): non-mutating look-ahead; full predicted mean and covariance
 * (mu: n doubles, sigma: n x n column-major, ld >= n; sigma may be NULL). */
REKF_API int rekf_predict_state(rekf_handle *h, int session, double time, double *mu, int cap,
                                double *sigma, int ld);

/* ---- state injection / persistence --------------------------------------------------------- */
/* Overwrite a session's state (bench warm start, tests).  mu: n doubles, sigma n x n column-major. */
REKF_API int rekf_set_state(rekf_handle *h, int session, double time, const double vt[3],
                            const double *mu, int n, const double *sigma, int ld);
/* sensor::Map (sensor/sensor_data.h:30-37) for all sessions: count beacons, xy float32 pairs,
 * cov row-major 2x2 doubles.  GetGlobalMap() counterpart below. */
REKF_API int rekf_set_map(rekf_handle *h, const float *xy, const double *cov2x2, int count);
REKF_API int rekf_get_map(rekf_handle *h, float *xy, double *cov2x2, int cap, int *count_out);
/* LoadMapFromTxtFile (:43-95).  A missing / malformed file is a silent no-op like the reference. */
REKF_API int rekf_load_map_txt(rekf_handle *h, const char *path);
/* Node::SaveReflectorResult (ros_node.cc:75-140): writes `<filebase>.txt`. */
REKF_API int rekf_save_map_txt(rekf_handle *h, int session, const char *filebase);

/* ---- stream / timing helpers (device time on the handle's own stream) --------------------- */
REKF_API int rekf_sync(rekf_handle *h);
REKF_API void *rekf_stream(rekf_handle *h);                 /* cudaStream_t */
REKF_API int rekf_timer_start(rekf_handle *h);              /* cudaEventRecord on the stream */
REKF_API int rekf_timer_stop(rekf_handle *h, float *ms);    /* record + synchronise + elapsed */
/* Per-kernel device time: when enabled every launch of the update chain is bracketed by events.
 * names/us/calls: up to cap entries (kernel name, mean microseconds per launch, launches). */
REKF_API int rekf_profile_enable(rekf_handle *h, int enable);
REKF_API int rekf_profile_read(rekf_handle *h, const char **names, double *mean_us, int *calls,
                               int cap, int *count_out);
/* number of kernel launches issued by this handle so far (graph nodes count individually) */
REKF_API int64_t rekf_launch_count(const rekf_handle *h);
/* raw device pointers for tests / profiling (Sigma is stored in the engine's internal layout) */
REKF_API int rekf_device_error_flags(rekf_handle *h, int session, int *flags_out);
/* Raw copy of a named internal device buffer of one session, for tests and profiling ("sbuf": S/L buffer
 * incl. the L^-1 nu row, "dinv": block inverses / per-phase cycle counters, "wdiag", "mu", "sigma").
 * Copies min(bytes, size of the buffer); the layout is internal (csrc/rekf_device.cuh). */
REKF_API int rekf_debug_copy(rekf_handle *h, int session, const char *name, void *out, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* REKF_H */
