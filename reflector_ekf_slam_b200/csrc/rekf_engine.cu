// rekf_engine.cu — host side of librekf_b200.so: handle, memory, launch chains and the C ABI of
// include/rekf.h.  The only CUDA-free code in here is the landmark-map text I/O (cold path,
// reference reflector_ekf_slam.cc:43-95 and ros_node.cc:75-140).  There is no CPU fallback: without a
// CUDA device rekf_create() fails with REKF_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/rekf.h"
#include "rekf_device.cuh"
#include "rekf_kernels.cuh"
#include "chol_smem.cuh"
#include "solve_w.cuh"
#include "solve_ll.cuh"
#include "syrk_exact_rows.cuh"
#include "syrk_tcgen05.cuh"
#include "syrk_tcgen05_i8.cuh"
#include "syrk_tcgen05_i8p.cuh"

using namespace rekf;

namespace {

enum KernelId { K_ODOM = 0, K_FRONT, K_GATHER, K_INNOV, K_CHOL, K_SOLVE, K_SYRK_EXACT, K_SYRK, K_AUGMENT, K_COUNT };
const char *kKernelNames[K_COUNT] = {"k_odometry", "k_observation_front", "k_gather_y(side stream)", "k_innovation", "k_cholesky",
                                     "k_solve_w", "k_syrk_exact_rows", "k_syrk", "k_augment"};

struct ProfRecord { int id; cudaEvent_t a, b; };

// A pipeline group: a contiguous range of the handle's sessions that advances on its own stream through its own
// launches (Layout::s0 / Sg select the range).  Groups never exchange data; running several of them lets one
// group's latency-bound kernels (association, innovation, Cholesky: a few CTAs) execute while another group's
// bandwidth-bound ones (TRSM, covariance SYRK) own the rest of the GPU.  `whole` (all sessions, the handle's main
// stream) is the only group when pipeline_groups <= 1 and the one used while per-kernel profiling is on.
struct Group {
  int s0 = 0, Sg = 0;
  cudaStream_t stream = nullptr;       // where this group's work is issued (the handle's main stream while profiling)
  cudaStream_t home_stream = nullptr;  // the group's own stream
  bool own_stream = false;
  cudaStream_t side_stream = nullptr;  // parallel branch of the step: Y = H·Σ beside innovation + Cholesky
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Layout L{};
  // device mailbox for host-delivered messages of this group's sessions + pinned staging ring
  char *mb_dev = nullptr;
  char *mb_host = nullptr;
  size_t mb_bytes = 0, off_odom = 0, off_time = 0, off_gps = 0, off_count = 0, off_xy = 0;
  static constexpr int kSlots = 32;
  cudaEvent_t slot_done[kSlots]{};
  int slot = 0;
  InputRef host_in{};
  InputRef *replay_in_dev = nullptr;   // the replay graphs read their input descriptor from here (InputRef::indirect)
  // replay graph (one step of this group)
  // one step (odometry + observation message) as CUDA graph(s): for the device-resident replay and for the host-message
  // step call (the mailbox is at a fixed address, so that chain is static too)
  struct StepGraphs {
    cudaGraphExec_t step = nullptr;          // whole step, or its latency-bound half when the groups are staggered
    cudaGraphExec_t wide = nullptr;          // staggered groups: the bandwidth-bound half (TRSM, SYRK, augmentation)
    cudaGraphExec_t multi = nullptr;         // one group: kMultiSteps steps in one graph (replay): the programmatic-launch chain runs
                                             // across step boundaries and the graph-to-graph gap is paid once per kMultiSteps steps
    InputRef in{};
    int64_t launches = 0;                    // kernel nodes per step
  } replay_gs, host_gs;
  cudaEvent_t done = nullptr;  // join marker
  cudaEvent_t phase_ev = nullptr;   // recorded when this group's latency-bound half of a step has been issued
  bool phase_recorded = false;
  // asynchronous pose delivery: pinned ring written by stream-ordered copies
  double *pose_host = nullptr;
  static constexpr int kPoseSlots = 64;
  cudaEvent_t pose_done[kPoseSlots]{};
};

}  // namespace

struct rekf_handle {
  rekf_options opts{};
  Layout L{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int64_t launches = 0;
  Group whole;                 // every session, on `stream`
  std::vector<Group> groups;   // pipeline groups (empty: `whole` is the only one)
  int64_t pose_ticket = 0;     // asynchronous pose requests issued so far
  int stagger_mode = 2;        // 0: streams free-running, 1: latency-bound halves alternate, 2: bandwidth-bound halves alternate
  // staging for getters / setters
  double *stage_dev = nullptr;
  size_t stage_elems = 0;
  double *hdr_host = nullptr;   // pinned: header of rekf_get_state
  // host copy of the beacon map
  std::vector<float> map_xy;
  std::vector<double> map_cov;
  // timing
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool profiling = false;
  std::vector<ProfRecord> prof;
  std::vector<cudaEvent_t> event_pool;
  double prof_us[K_COUNT]{};
  int prof_calls[K_COUNT]{};
  // tcgen05 SYRK resources
  bool chol_resident = false, solve_w2 = false;
  int front_cluster = 4;       // CTAs per session in k_observation_front (a thread-block cluster shares the association; REKF_FRONT_CLUSTER)
  int chol_newton = 2;         // Newton steps of the pivot rsqrt (chol_smem.cuh): 1 in the tensor-core covariance modes, 2 in fp64
  bool solve_ll = false;       // the flag-paced left-looking TRSM (solve_ll.cuh) replaces k_solve_w3: frames always fit the resident Cholesky
  bool pdl = true;             // programmatic dependent launch along the step's kernel chain (REKF_PDL=0 turns it off)
  SyrkTc tc{};
  SyrkI8P tc8p{};
  std::vector<void *> allocations;
};

namespace {

// Every entry point runs on the handle's device, whatever the calling thread's current device is (a ROS node's
// callback and service threads default to device 0), and leaves the caller's current device untouched.
struct DeviceGuard {
  int prev = -1, want = -1;
  explicit DeviceGuard(const rekf_handle *h) {
    if (!h) return;
    want = h->device;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) cudaSetDevice(want);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != want) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};

int fail(rekf_handle *h, int code, const char *fmt, ...) {
  if (h) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    h->err = buf;
  }
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(h, REKF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
int dev_alloc(rekf_handle *h, T **p, size_t count, bool zero = true) {
  void *q = nullptr;
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return fail(h, REKF_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  if (zero) {
    e = cudaMemsetAsync(q, 0, bytes, h->stream);
    if (e != cudaSuccess) return fail(h, REKF_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  }
  h->allocations.push_back(q);
  *p = static_cast<T *>(q);
  return 0;
}

cudaEvent_t take_event(rekf_handle *h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct ProfScope {
  rekf_handle *h;
  int id;
  cudaStream_t stream;
  cudaEvent_t a = nullptr;
  ProfScope(rekf_handle *h_, int id_, cudaStream_t st) : h(h_), id(id_), stream(st) {
    ++h->launches;
    if (h->profiling) {
      a = take_event(h);
      cudaEventRecord(a, stream);
    }
  }
  ~ProfScope() {
    if (a) {
      cudaEvent_t b = take_event(h);
      cudaEventRecord(b, stream);
      h->prof.push_back({id, a, b});
    }
  }
};

// the groups that carry the hot path.  Per-kernel profiling wants one kernel on the GPU at a time and the SAME launch
// shapes as the timed path: rekf_profile_enable re-homes every group onto the handle's main stream meanwhile.
inline std::vector<Group *> active_groups(rekf_handle *h) {
  std::vector<Group *> v;
  if (h->groups.empty()) v.push_back(&h->whole);
  else for (auto &g : h->groups) v.push_back(&g);
  return v;
}

// host-blocking join of every stream of the handle (cold paths: getters, setters, mode switches)
inline cudaError_t join_all(rekf_handle *h) {
  for (auto &g : h->groups) {
    cudaError_t e = cudaStreamSynchronize(g.home_stream);
    if (e != cudaSuccess) return e;
  }
  return cudaStreamSynchronize(h->stream);   // side branches are joined back into these streams inside every step
}

void drain_profile(rekf_handle *h) {
  for (auto &rec : h->prof) {
    cudaEventSynchronize(rec.b);
    float ms = 0;
    cudaEventElapsedTime(&ms, rec.a, rec.b);
    h->prof_us[rec.id] += 1e3 * ms;
    h->prof_calls[rec.id] += 1;
    h->event_pool.push_back(rec.a);
    h->event_pool.push_back(rec.b);
  }
  h->prof.clear();
}

// kind / target lists, float32 + fp64 landmark means, the frame's observations
size_t smem_front(const Layout &L) {
  return sizeof(int) * 2 * (size_t)L.mcap + sizeof(float2) * (size_t)L.Ncap + sizeof(double2) * (size_t)L.Ncap + sizeof(float2) * (size_t)L.mcap + 32;
}
// launch with the programmatic-stream-serialization attribute: the kernel may start while the previous kernel of the stream
// (or graph branch) is still running; it blocks in pdl_wait() until that one has completed (rekf_device.cuh)
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// the same with thread-block clusters of `cluster_x` CTAs along x (grid.x a multiple of it)
template <typename... KArgs, typename... Args>
cudaError_t launch_chain_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, int cluster_x,
                                 Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

size_t smem_chol(const Layout &L) { return sizeof(double) * ((size_t)(L.rcap + 1) * kPS + (size_t)kCholNb * kPS); }
size_t smem_solve(const Layout &L) { return sizeof(double) * (size_t)L.rld * kYS; }

int launch_odometry(rekf_handle *h, Group &grp, const InputRef &in) {
  ProfScope p(h, K_ODOM, grp.stream);
  k_odometry<<<grp.Sg, 1024, 0, grp.stream>>>(grp.L, in);
  CK(cudaGetLastError());
  return 0;
}

// Two pipeline groups are kept in ANTI-phase: a group may start the latency-bound half of a step (association,
// innovation, Cholesky: a handful of CTAs) only once the other group has issued its own and moved on to the
// bandwidth-bound half (TRSM, SYRK: the whole GPU).  Left alone the two streams drift into phase — the timeline
// (scripts/timeline.py) shows both in TRSM, then both in SYRK — and nothing overlaps.
inline bool staggered(rekf_handle *h, Group &grp) {
  return h->stagger_mode != 0 && h->groups.size() == 2 && !h->profiling && &grp != &h->whole;
}
int stagger_wait(rekf_handle *h, Group &grp) {
  if (!staggered(h, grp)) return 0;
  Group &other = h->groups[&grp == &h->groups[0] ? 1 : 0];
  if (other.phase_recorded) CK(cudaStreamWaitEvent(grp.stream, other.phase_ev, 0));
  return 0;
}
int stagger_mark(rekf_handle *h, Group &grp) {
  if (!staggered(h, grp)) return 0;
  CK(cudaEventRecord(grp.phase_ev, grp.stream));
  grp.phase_recorded = true;
  return 0;
}

// HandleObservationMessage, latency-bound half: Predict + ReflectorMatch + measurement rows, S, Cholesky
int launch_obs_narrow(rekf_handle *h, Group &grp, const InputRef &in) {
  const Layout &L = grp.L;
  cudaStream_t stream = grp.stream;
  {
    ProfScope p(h, K_FRONT, stream);
    if (h->front_cluster > 1)
      CK(launch_chain_cluster(k_observation_front, dim3(L.Sg * h->front_cluster), dim3(1024), smem_front(L), stream, h->pdl, h->front_cluster, L, in));
    else
      CK(launch_chain(k_observation_front, dim3(L.Sg), dim3(1024), smem_front(L), stream, h->pdl, L, in));
  }
  {
    ProfScope p(h, K_INNOV, stream);
    const int g = (L.rcap + 15) / 16;
    CK(launch_chain(k_innovation, dim3(g, g, L.Sg), dim3(16, 16), 0, stream, h->pdl, L));
  }
  const bool side_gather = h->solve_w2 && !h->solve_ll;   // k_solve_ll gathers Y = H·Σ itself (solve_ll.cuh)
  if (side_gather) CK(cudaEventRecord(grp.ev_fork, stream));
  {
    ProfScope p(h, K_CHOL, stream);
    if (h->chol_resident) {
      const size_t sm = smem_chol_resident(std::min(L.rcap, kCholResidentMax));
      auto chol = h->chol_newton == 1 ? k_cholesky_smem<1> : k_cholesky_smem<2>;
      CK(launch_chain(chol, dim3(L.Sg), dim3(kCholSmemThreads), sm, stream, h->pdl, L, 0));
      if (L.rcap > kCholResidentMax) {
        // two-level factorisation of frames with more rows than one SM holds (chol_smem.cuh); every stage decides on the
        // device whether the frame is split, so the chain is static
        chol<<<L.Sg, kCholSmemThreads, sm, stream>>>(L, 1);
        k_chol_trsm_rows<<<dim3((L.rcap - 32 + kW3Cols - 1) / kW3Cols, 1, L.Sg), 256, smem_solve_w3(L.rld), stream>>>(L);
        const int nt = (L.rcap + 1 + 31) / 32;
        k_chol_syrk<<<dim3(nt * (nt + 1) / 2, 1, L.Sg), 256, 0, stream>>>(L);
        chol<<<L.Sg, kCholSmemThreads, sm, stream>>>(L, 2);
        k_cholesky<<<L.Sg, 1024, smem_chol(L), stream>>>(L, 1);      // frames too large even for two levels
        h->launches += 5;
      }
    } else {
      k_cholesky<<<L.Sg, 1024, smem_chol(L), stream>>>(L, 0);
    }
  }
  if (side_gather) {
    // fork: Y = H·Σ needs only the match lists and the predicted Σ — a parallel branch (under capture: of the step graph) that
    // runs in the shadow of the Cholesky (one CTA per session, the rest of the GPU idle).  It is forked behind k_innovation
    // and issued after the Cholesky launch: beside k_innovation its ~1000 short blocks tripled that kernel's latency.
    CK(cudaStreamWaitEvent(grp.side_stream, grp.ev_fork, 0));
    {
      ProfScope p(h, K_GATHER, grp.side_stream);
      k_gather_y<<<dim3(L.ld / 128, (L.rcap / 2 + kGYPairs - 1) / kGYPairs, L.Sg), 256, 0, grp.side_stream>>>(L);
    }
    CK(cudaEventRecord(grp.ev_join, grp.side_stream));
  }
  if (side_gather) CK(cudaStreamWaitEvent(stream, grp.ev_join, 0));   // join before the TRSM
  CK(cudaGetLastError());
  return 0;
}

// bandwidth-bound half: W = L⁻¹HΣ and the mean update, Σ −= WᵀW, augmentation
int launch_obs_wide(rekf_handle *h, Group &grp, const InputRef &in) {
  const Layout &L = grp.L;
  cudaStream_t stream = grp.stream;
  {
    ProfScope p(h, K_SOLVE, stream);
    if (h->solve_ll) CK(launch_chain(k_solve_ll, dim3(L.ld / kW3Cols, 1, L.Sg), dim3(kLLThreads), smem_solve_ll(), stream, h->pdl, L));
    else if (h->solve_w2) CK(launch_chain(k_solve_w3, dim3(L.ld / kW3Cols, 1, L.Sg), dim3(256), smem_solve_w3(L.rld), stream, h->pdl && L.rcap <= kCholResidentMax, L));
    else k_solve_w<<<dim3(L.ld / kWCols, 1, L.Sg), 256, smem_solve(L), stream>>>(L);
  }
  if (h->opts.cov_update == REKF_COV_TCGEN05_I8X4) {
    // fp64 side of the hybrid: the whole frame if st.exact_update, else the rows/columns of flagged slots, else nothing
    // (it also rewinds the tile cursor of the persistent kernel that follows)
    ProfScope p(h, K_SYRK_EXACT, stream);
    CK(launch_chain(k_syrk_f64, dim3(148, 1, L.Sg), dim3(256), 0, stream, h->pdl, L));
  }
  {
    ProfScope p(h, K_SYRK, stream);
    if (h->opts.cov_update == REKF_COV_SIMT_F64) {
      k_syrk_f64<<<dim3(592, 1, L.Sg), 256, 0, stream>>>(L);
    } else if (h->opts.cov_update == REKF_COV_TCGEN05_I8X4) {
      SyrkI8P tc8p = h->tc8p;
      if (&grp == &h->whole) tc8p.reserve_sms = 0;                // nothing else is running beside the whole batch
      int rc = syrk_i8p_launch(tc8p, L, stream, h->pdl);
      if (rc != 0) return fail(h, REKF_ERR_CUDA, "tcgen05 int8 SYRK launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
      int rc = syrk_tc_launch(h->tc, L, stream);
      if (rc != 0) return fail(h, REKF_ERR_CUDA, "tcgen05 SYRK launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  {
    ProfScope p(h, K_AUGMENT, stream);
    CK(launch_chain(k_augment, dim3((L.ncap + 255) / 256, 1, L.Sg), dim3(256), 0, stream, h->pdl, L, in));
  }
  CK(cudaGetLastError());
  return 0;
}

// HandleObservationMessage as a fixed launch chain.  The host-message path leaves the two streams free-running: its
// cadence is set by the host's own launches, and the extra event traffic cost more than the stagger gained (measured).
int launch_observation(rekf_handle *h, Group &grp, const InputRef &in) {
  int rc = launch_obs_narrow(h, grp, in);
  if (!rc) rc = launch_obs_wide(h, grp, in);
  return rc;
}

// (re)capture the fused step chain of one group for the inputs `in` (k_observation_front takes the odometry message,
// k_augment advances the replay counter)
constexpr int kMultiSteps = 4;
int capture_step(rekf_handle *h, Group &g, const InputRef &in, Group::StepGraphs &gs, bool want_multi = false) {
  if (gs.step && std::memcmp(&in, &gs.in, sizeof(InputRef)) == 0 && (gs.wide != nullptr) == staggered(h, g)) return 0;
  if (gs.step) { cudaGraphExecDestroy(gs.step); gs.step = nullptr; }
  if (gs.wide) { cudaGraphExecDestroy(gs.wide); gs.wide = nullptr; }
  if (gs.multi) { cudaGraphExecDestroy(gs.multi); gs.multi = nullptr; }
  const bool split = staggered(h, g);          // two graphs, so that the stagger events sit between them
  const int64_t before = h->launches;
  for (int part = 0; part < (split ? 2 : 1); ++part) {
    cudaGraph_t cg = nullptr;
    CK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (part == 0) rc = launch_obs_narrow(h, g, in);
    if (!rc && (part == 1 || !split)) rc = launch_obs_wide(h, g, in);
    cudaError_t ce = cudaStreamEndCapture(g.stream, &cg);
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(h, REKF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    cudaGraphExec_t *exec = part == 0 ? &gs.step : &gs.wide;
    CK(cudaGraphInstantiate(exec, cg, 0));
    cudaGraphDestroy(cg);
    CK(cudaGraphUpload(*exec, g.stream));
  }
  gs.in = in;
  gs.launches = h->launches - before;          // kernel nodes per step
  if (want_multi && !split && h->pdl) {
    cudaGraph_t cg = nullptr;
    CK(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    for (int k = 0; k < kMultiSteps && !rc; ++k) rc = launch_observation(h, g, in);
    cudaError_t ce = cudaStreamEndCapture(g.stream, &cg);
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(h, REKF_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    CK(cudaGraphInstantiate(&gs.multi, cg, 0));
    cudaGraphDestroy(cg);
    CK(cudaGraphUpload(gs.multi, g.stream));
  }
  h->launches = before;                        // the capture pass itself launched nothing
  return 0;
}

int launch_step(rekf_handle *h, Group &g, Group::StepGraphs &gs) {
  int rc = h->stagger_mode == 1 ? stagger_wait(h, g) : 0;
  if (rc) return rc;
  CK(cudaGraphLaunch(gs.step, g.stream));
  if (gs.wide) {
    if ((rc = h->stagger_mode == 1 ? stagger_mark(h, g) : stagger_wait(h, g))) return rc;
    CK(cudaGraphLaunch(gs.wide, g.stream));
    if (h->stagger_mode != 1 && (rc = stagger_mark(h, g))) return rc;
  }
  h->launches += gs.launches;
  return 0;
}

int stage_reserve(rekf_handle *h, size_t elems) {
  if (elems <= h->stage_elems) return 0;
  if (h->stage_dev) cudaFree(h->stage_dev);
  h->stage_dev = nullptr;
  h->stage_elems = 0;
  cudaError_t e = cudaMalloc(&h->stage_dev, elems * sizeof(double));
  if (e != cudaSuccess) return fail(h, REKF_ERR_CUDA, "cudaMalloc(stage %zu) failed: %s", elems * sizeof(double), cudaGetErrorString(e));
  h->stage_elems = elems;
  return 0;
}

int read_state(rekf_handle *h, int s, SessionState *out) {
  if (s < 0 || s >= h->L.S) return fail(h, REKF_ERR_BAD_ARGUMENT, "session %d out of range", s);
  CK(join_all(h));
  CK(cudaMemcpyAsync(out, h->L.st + s, sizeof(SessionState), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// acquire the next pinned staging slot (waits only if the copy issued kSlots messages ago is pending)
char *next_slot(Group &g) {
  g.slot = (g.slot + 1) % Group::kSlots;
  cudaEventSynchronize(g.slot_done[g.slot]);
  return g.mb_host + (size_t)g.slot * g.mb_bytes;
}

// mailbox, staging ring, events and the Layout view of one group (sessions s0 .. s0+Sg-1, counters at `index`)
int init_group(rekf_handle *h, Group &g, int s0, int Sg, int index, cudaStream_t stream, bool own_stream) {
  g.s0 = s0;
  g.Sg = Sg;
  g.stream = stream;
  g.home_stream = stream;
  g.own_stream = own_stream;
  g.L = h->L;
  g.L.s0 = s0;
  g.L.Sg = Sg;
  g.L.step = h->L.step + index;
  g.L.tile_counter = h->L.tile_counter + index;
  g.L.step_ticket = h->L.step_ticket + index;
  const size_t S = (size_t)Sg;
  const Layout &L = h->L;
  g.off_odom = 0;
  g.off_time = g.off_odom + S * 4 * sizeof(double);
  g.off_gps = g.off_time + S * sizeof(double);
  g.off_count = g.off_gps + S * 4 * sizeof(double);
  g.off_xy = g.off_count + round_up((int)(S * sizeof(int)), 16);
  g.mb_bytes = round_up((int)(g.off_xy + S * L.mcap * 2 * sizeof(float)), 256);
  int rc = 0;
  if ((rc = dev_alloc(h, &g.mb_dev, g.mb_bytes))) return rc;
  if ((rc = dev_alloc(h, &g.replay_in_dev, 1))) return rc;
  CK(cudaMallocHost(&g.mb_host, g.mb_bytes * Group::kSlots));
  std::memset(g.mb_host, 0, g.mb_bytes * Group::kSlots);
  for (int i = 0; i < Group::kSlots; ++i) CK(cudaEventCreateWithFlags(&g.slot_done[i], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&g.phase_ev, cudaEventDisableTiming));
  {
    // the side branch (Y = H·Σ: ~1000 short blocks) must not crowd the latency-bound main chain out of the SMs: lowest priority
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&g.side_stream, cudaStreamNonBlocking, least));
  }
  CK(cudaEventCreateWithFlags(&g.ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&g.ev_join, cudaEventDisableTiming));
  CK(cudaMallocHost(&g.pose_host, sizeof(double) * 3 * S * Group::kPoseSlots));
  for (int i = 0; i < Group::kPoseSlots; ++i) CK(cudaEventCreateWithFlags(&g.pose_done[i], cudaEventDisableTiming));
  // kernels index every input by the absolute session: bias the bases by -s0 strides
  InputRef &in = g.host_in;
  in.odom_ss = 4;
  in.time_ss = 1;
  in.xy_ss = (long long)L.mcap * 2;
  in.odom = reinterpret_cast<const double *>(g.mb_dev + g.off_odom) - (long long)s0 * in.odom_ss;
  in.obs_time = reinterpret_cast<const double *>(g.mb_dev + g.off_time) - (long long)s0 * in.time_ss;
  in.gps = reinterpret_cast<const double *>(g.mb_dev + g.off_gps) - (long long)s0 * 4;
  in.obs_count = reinterpret_cast<const int *>(g.mb_dev + g.off_count) - s0;
  in.obs_xy = reinterpret_cast<const float *>(g.mb_dev + g.off_xy) - (long long)s0 * in.xy_ss;
  in.m_stride = L.mcap;
  in.m_fixed = 0;
  in.step = nullptr;
  in.fuse_odom = 0;
  in.pose_out = nullptr;
  in.pose_ss = 0;
  return 0;
}

void destroy_group(Group &g) {
  for (Group::StepGraphs *gs : {&g.replay_gs, &g.host_gs}) {
    if (gs->multi) cudaGraphExecDestroy(gs->multi);
    if (gs->step) cudaGraphExecDestroy(gs->step);
    if (gs->wide) cudaGraphExecDestroy(gs->wide);
  }
  if (g.phase_ev) cudaEventDestroy(g.phase_ev);
  if (g.ev_fork) cudaEventDestroy(g.ev_fork);
  if (g.ev_join) cudaEventDestroy(g.ev_join);
  if (g.side_stream) cudaStreamDestroy(g.side_stream);
  if (g.mb_host) cudaFreeHost(g.mb_host);
  if (g.pose_host) cudaFreeHost(g.pose_host);
  for (auto &e : g.slot_done) if (e) cudaEventDestroy(e);
  for (auto &e : g.pose_done) if (e) cudaEventDestroy(e);
  if (g.done) cudaEventDestroy(g.done);
  if (g.own_stream && g.home_stream) cudaStreamDestroy(g.home_stream);
}

// ---- landmark map text format ---------------------------------------------------------------
// common.cc:5-16 SplitString = std::getline on ',' (empty tokens kept, trailing empty dropped)
std::vector<std::string> split_commas(const std::string &line) {
  std::istringstream ss(line);
  std::string tok;
  std::vector<std::string> out;
  while (std::getline(ss, tok, ',')) out.push_back(tok);
  return out;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

void rekf_default_options(rekf_options *o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->odom_model = REKF_ODOM_DIFF;              // ros_node.cc:285-296 default
  o->linear_velocity_cov = 0.05 * 0.05;        // launch/slam.launch:21, squared at ros_node.cc:207-215
  o->angular_velocity_cov = 0.08 * 0.08;       // launch/slam.launch:22
  o->observation_cov = 0.05 * 0.05;            // launch/slam.launch:23
  o->max_landmarks = 1024;
  o->max_observations = 128;
  o->max_map_landmarks = 1024;
  o->cov_update = REKF_COV_TCGEN05_I8X4;
  o->map_loader = REKF_MAP_LOADER_FIXED;
}

const char *rekf_version(void) { return "rekf-b200 0.1 (sm_100a)"; }

int rekf_create(const rekf_options *opts, rekf_handle **out) { return rekf_create_batch(opts, 1, out); }

static int create_batch_impl(const rekf_options *opts, int sessions, rekf_handle **out);
int rekf_create_batch(const rekf_options *opts, int sessions, rekf_handle **out) {
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  const int rc = create_batch_impl(opts, sessions, out);
  if (prev >= 0) cudaSetDevice(prev);            // the caller's current device is left as it was
  return rc;
}
static int create_batch_impl(const rekf_options *opts, int sessions, rekf_handle **out) {
  if (!opts || !out || sessions < 1) return REKF_ERR_BAD_ARGUMENT;
  *out = nullptr;
  rekf_handle *h = new rekf_handle();
  *out = h;   // returned even on failure so that rekf_last_error() works; caller destroys it
  h->opts = *opts;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(h, REKF_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
  if (opts->device < 0 || opts->device >= ndev) return fail(h, REKF_ERR_BAD_ARGUMENT, "device %d of %d", opts->device, ndev);
  h->device = opts->device;
  CK(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10)
    return fail(h, REKF_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", h->device, prop.major, prop.minor);
  if (opts->stream) {
    h->stream = static_cast<cudaStream_t>(opts->stream);
  } else {
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, greatest));
    h->own_stream = true;
  }
  Layout &L = h->L;
  L.S = sessions;
  L.s0 = 0;
  L.Sg = sessions;
  int G = opts->pipeline_groups > 1 ? std::min(opts->pipeline_groups, sessions) : 1;
  if (const char *e = std::getenv("REKF_STAGGER")) h->stagger_mode = std::atoi(e);
  // With several pipeline groups the blocks of a dependent kernel that sit in pdl_wait() hold SMs the other group's kernels
  // could use (measured: 35.2 k -> 33.4 k steps/s at 8 sessions in 2 groups), alone they only hide launch latency
  // (single session 8.1 k -> 8.4 k): on for one group, off otherwise.
  h->pdl = G == 1;
  if (const char *e = std::getenv("REKF_PDL")) h->pdl = std::atoi(e) != 0;
  L.Ncap = opts->max_landmarks > 0 ? opts->max_landmarks : 1024;
  L.mcap = opts->max_observations > 0 ? opts->max_observations : 128;
  if (L.mcap > 512) return fail(h, REKF_ERR_BAD_ARGUMENT, "max_observations %d > 512", L.mcap);
  L.mapcap = opts->max_map_landmarks > 0 ? opts->max_map_landmarks : 1024;
  L.ncap = internal_dim(L.Ncap);
  L.ld = round_up(L.ncap, kSigmaTile);
  L.rcap = 2 * L.mcap + 4;
  L.rld = round_up(L.rcap, kKBlock);
  L.sld = L.rld + 8;
  L.kq = round_up(L.rcap, 64);
  L.odom_model = opts->odom_model == REKF_ODOM_DIFF ? 0 : 1;   // :13-32: everything but DIFF is the 3x3 model
  L.q_lin = opts->linear_velocity_cov;
  L.q_ang = opts->angular_velocity_cov;
  L.q_obs = opts->observation_cov;
  const size_t S = (size_t)L.S;
  int rc = 0;
  if ((rc = dev_alloc(h, &L.st, S))) return rc;
  if ((rc = dev_alloc(h, &L.mu, S * L.ld))) return rc;
  if ((rc = dev_alloc(h, &L.sigma, S * L.ld * L.ld))) return rc;
  if ((rc = dev_alloc(h, &L.map_xy, (size_t)L.mapcap * 2))) return rc;
  if ((rc = dev_alloc(h, &L.map_cov, (size_t)L.mapcap * 4))) return rc;
  if ((rc = dev_alloc(h, &L.map_count, 1))) return rc;
  if ((rc = dev_alloc(h, &L.state_pairs, S * L.mcap * 2))) return rc;
  if ((rc = dev_alloc(h, &L.map_pairs, S * L.mcap * 2))) return rc;
  if ((rc = dev_alloc(h, &L.new_ids, S * L.mcap))) return rc;
  if ((rc = dev_alloc(h, &L.Hp, S * L.rcap * 4))) return rc;
  if ((rc = dev_alloc(h, &L.Hl, S * L.rcap * 2))) return rc;
  if ((rc = dev_alloc(h, &L.Hslot, S * L.rcap))) return rc;
  if ((rc = dev_alloc(h, &L.innov, S * L.rcap))) return rc;
  if ((rc = dev_alloc(h, &L.Qd, S * L.rcap))) return rc;
  if ((rc = dev_alloc(h, &L.Sbuf, S * L.rld * L.sld))) return rc;
  if ((rc = dev_alloc(h, &L.Dinv, S * (L.rld / kCholNb) * kCholNb * kCholNb))) return rc;
  if ((rc = dev_alloc(h, &L.Ybuf, S * L.rld * L.ld))) return rc;
  if (opts->cov_update == REKF_COV_SIMT_F64) {
    if ((rc = dev_alloc(h, &L.W64, S * L.ld * L.rld))) return rc;
  } else if (opts->cov_update == REKF_COV_TCGEN05_TF32X3) {
    if ((rc = dev_alloc(h, &L.Wt_hi, S * L.ld * L.rld))) return rc;
    if ((rc = dev_alloc(h, &L.Wt_lo, S * L.ld * L.rld))) return rc;
  } else if (opts->cov_update == REKF_COV_TCGEN05_I8X4) {
    if ((rc = dev_alloc(h, &L.Wq, S * 4 * L.ld * L.kq))) return rc;
    if ((rc = dev_alloc(h, &L.W64, S * L.ld * L.rld))) return rc;
    if ((rc = dev_alloc(h, &L.Wexp, S * L.ld))) return rc;
    if ((rc = dev_alloc(h, &L.Wscale, S * L.ld))) return rc;
    if ((rc = dev_alloc(h, &L.Wflag, S * L.ld))) return rc;
    if ((rc = dev_alloc(h, &L.exact_list, S * kMaxExactSlots))) return rc;
  } else {
    return fail(h, REKF_ERR_BAD_ARGUMENT, "unknown cov_update %d", opts->cov_update);
  }
  if (opts->cov_update != REKF_COV_SIMT_F64 && (rc = dev_alloc(h, &L.Wdiag, S * L.ld))) return rc;
  if ((rc = dev_alloc(h, &L.step, G + 1))) return rc;
  if ((rc = dev_alloc(h, &L.tile_counter, G + 1))) return rc;
  if ((rc = dev_alloc(h, &L.step_ticket, G + 1))) return rc;
  if (std::getenv("REKF_TIMELINE") && (rc = dev_alloc(h, &L.tlog, kTimelineCap))) return rc;
  // the TRSM in the shadow of the Cholesky (solve_ll.cuh): every frame fits the single-pass resident factorisation and the fp64
  // W panel exists.  Side by side only along a programmatic-launch chain (one pipeline group); otherwise the same kernel, in order.
  if (const char *e = std::getenv("REKF_FRONT_CLUSTER")) h->front_cluster = std::max(1, std::min(8, std::atoi(e)));
  h->chol_newton = opts->cov_update == REKF_COV_SIMT_F64 ? 2 : 1;
  if (const char *e = std::getenv("REKF_CHOL_NEWTON")) h->chol_newton = std::atoi(e) == 1 ? 1 : 2;
  h->solve_ll = L.rcap <= kCholResidentMax && L.W64 != nullptr && smem_solve_w3(L.rld) <= 227 * 1024;
  if (const char *e = std::getenv("REKF_SOLVE_LL")) h->solve_ll = h->solve_ll && std::atoi(e) != 0;
  if (h->solve_ll) {
    L.sync_n = 8;
    if ((rc = dev_alloc(h, &L.sync, S * L.sync_n))) return rc;
    L.shadow = h->pdl ? 1 : 0;
    if (const char *e = std::getenv("REKF_SHADOW")) L.shadow = (h->pdl && std::atoi(e) != 0) ? 1 : 0;
  }

  // initial state: time, pose (:8-11)
  {
    std::vector<SessionState> st(S);
    std::vector<double> mu0(S * L.ld, 0.0);
    for (size_t s = 0; s < S; ++s) {
      std::memset(&st[s], 0, sizeof(SessionState));
      st[s].time = opts->init_time;
      for (int i = 0; i < 3; ++i) mu0[s * L.ld + i] = opts->init_pose[i];
    }
    CK(cudaMemcpyAsync(L.st, st.data(), sizeof(SessionState) * S, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(L.mu, mu0.data(), sizeof(double) * S * L.ld, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  // groups: mailboxes, staging rings, streams
  if ((rc = init_group(h, h->whole, 0, sessions, 0, h->stream, false))) return rc;
  if (G > 1) {
    h->groups.resize(G);
    for (int g = 0; g < G; ++g) {
      const int lo = (int)((long long)sessions * g / G), hi = (int)((long long)sessions * (g + 1) / G);
      cudaStream_t st = nullptr;
      int least = 0, greatest = 0;
      CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      CK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, greatest));
      if ((rc = init_group(h, h->groups[g], lo, hi - lo, g + 1, st, true))) return rc;
    }
  }
  CK(cudaEventCreate(&h->t0));
  CK(cudaEventCreate(&h->t1));
  // opt in to large dynamic shared memory where needed
  if (smem_front(L) > 200 * 1024) return fail(h, REKF_ERR_CAPACITY, "max_landmarks %d needs %zu B of shared memory in k_observation_front", L.Ncap, smem_front(L));
  if (smem_chol(L) > 227 * 1024 || smem_solve(L) > 227 * 1024)
    return fail(h, REKF_ERR_CAPACITY, "max_observations %d needs %zu / %zu B of shared memory in the Cholesky / TRSM kernels (limit 227 KB: about 400 observations)",
                L.mcap, smem_chol(L), smem_solve(L));
  CK(cudaFuncSetAttribute(k_observation_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_front(L)));
  CK(cudaFuncSetAttribute(k_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chol(L)));
  h->solve_w2 = smem_solve_w3(L.rld) <= 227 * 1024;
  if (h->solve_w2) CK(cudaFuncSetAttribute(k_solve_w3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve_w3(L.rld)));
  // the resident Cholesky (one or two levels) and the DMMA TRSM come as a pair: both need the operands of a frame in one SM
  h->chol_resident = h->solve_w2;
  if (h->chol_resident) {
    CK(cudaFuncSetAttribute(k_cholesky_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chol_resident(std::min(L.rcap, kCholResidentMax))));
    CK(cudaFuncSetAttribute(k_cholesky_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_chol_resident(std::min(L.rcap, kCholResidentMax))));
    CK(cudaFuncSetAttribute(k_chol_trsm_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve_w3(L.rld)));
  }
  CK(cudaFuncSetAttribute(k_solve_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve(L)));
  if (h->solve_ll) {
    CK(cudaFuncSetAttribute(k_solve_ll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_solve_ll()));
    CK(cudaFuncSetAttribute(k_solve_ll, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // four blocks per SM
  }
  if (opts->cov_update == REKF_COV_TCGEN05_TF32X3) {
    const char *why = syrk_tc_init(h->tc, L);
    if (why) return fail(h, REKF_ERR_CUDA, "tcgen05 SYRK setup failed: %s", why);
  }
  if (opts->cov_update == REKF_COV_TCGEN05_I8X4) {
    const char *why = syrk_i8p_init(h->tc8p, L);
    if (why) return fail(h, REKF_ERR_CUDA, "tcgen05 int8 SYRK setup failed: %s", why);
    // with several groups in flight the persistent kernel leaves a few SMs to the other groups' narrow kernels
    h->tc8p.reserve_sms = G > 1 ? (opts->syrk_reserve_sms > 0 ? opts->syrk_reserve_sms : 2 * ((sessions + G - 1) / G) + 4) : 0;
  if (const char *e = std::getenv("REKF_SYRK_RESERVE_SMS")) h->tc8p.reserve_sms = G > 1 ? std::atoi(e) : 0;
  }
  if (opts->map_path && opts->map_path[0]) rekf_load_map_txt(h, opts->map_path);   // :36
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

int rekf_destroy(rekf_handle *h) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_OK;
  if (h->stream) join_all(h);
  drain_profile(h);
  for (auto &g : h->groups) destroy_group(g);
  destroy_group(h->whole);
  syrk_tc_destroy(h->tc);
  for (void *p : h->allocations) cudaFree(p);
  if (h->stage_dev) cudaFree(h->stage_dev);
  if (h->hdr_host) cudaFreeHost(h->hdr_host);
  for (auto &e : h->event_pool) cudaEventDestroy(e);
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->t1) cudaEventDestroy(h->t1);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return REKF_OK;
}

const char *rekf_last_error(const rekf_handle *h) { return h ? h->err.c_str() : "null handle"; }
int rekf_sessions(const rekf_handle *h) { return h ? h->L.S : 0; }
int64_t rekf_launch_count(const rekf_handle *h) { return h ? h->launches : 0; }
void *rekf_stream(rekf_handle *h) { return h ? h->stream : nullptr; }

// ---- hot path -------------------------------------------------------------------------------
int rekf_batch_handle_odometry(rekf_handle *h, const double *odom) {
  DeviceGuard dev_guard(h);
  if (!h || !odom) return REKF_ERR_BAD_ARGUMENT;
  if (h->opts.use_imu) return REKF_OK;   // :213-222: with use_imu the reference does nothing
  for (Group *gp : active_groups(h)) {
    Group &g = *gp;
    char *slot = next_slot(g);
    const size_t bytes = (size_t)g.Sg * 4 * sizeof(double);
    std::memcpy(slot + g.off_odom, odom + (size_t)g.s0 * 4, bytes);
    CK(cudaMemcpyAsync(g.mb_dev + g.off_odom, slot + g.off_odom, bytes, cudaMemcpyHostToDevice, g.stream));
    CK(cudaEventRecord(g.slot_done[g.slot], g.stream));
    int rc = launch_odometry(h, g, g.host_in);
    if (rc) return rc;
  }
  return REKF_OK;
}

int rekf_handle_odometry(rekf_handle *h, double time, double vx, double vy, double wz) {
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  if (h->L.S != 1) return fail(h, REKF_ERR_BAD_ARGUMENT, "rekf_handle_odometry needs a single-session handle");
  const double msg[4] = {time, vx, vy, wz};
  return rekf_batch_handle_odometry(h, msg);
}

static int observation_common(rekf_handle *h, const double *times, const float *xy, const int *counts,
                              int m_stride, const double *gps /*S x 4 or null*/, const double *odom = nullptr /*S x 4: fused step*/) {
  const Layout &L = h->L;
  DeviceGuard dev_guard(h);
  for (int s = 0; s < L.S; ++s) {
    if (counts[s] < 0) return fail(h, REKF_ERR_BAD_ARGUMENT, "negative observation count");
    if (counts[s] > L.mcap) return fail(h, REKF_ERR_CAPACITY, "frame of %d observations exceeds max_observations %d", counts[s], L.mcap);
  }
  for (Group *gp : active_groups(h)) {
    Group &grp = *gp;
    char *slot = next_slot(grp);
    std::memcpy(slot + grp.off_time, times + grp.s0, sizeof(double) * grp.Sg);
    double *g = reinterpret_cast<double *>(slot + grp.off_gps);
    if (gps) std::memcpy(g, gps + (size_t)grp.s0 * 4, sizeof(double) * 4 * grp.Sg);
    else std::memset(g, 0, sizeof(double) * 4 * grp.Sg);
    int *cnt = reinterpret_cast<int *>(slot + grp.off_count);
    float *dst = reinterpret_cast<float *>(slot + grp.off_xy);
    for (int q = 0; q < grp.Sg; ++q) {
      const int m = counts[grp.s0 + q];
      cnt[q] = m;
      if (m > 0) std::memcpy(dst + (size_t)q * L.mcap * 2, xy + (size_t)(grp.s0 + q) * m_stride * 2, sizeof(float) * 2 * m);
    }
    // one copy: [odom |] time | gps | count | xy (only as far as the last session's data reaches)
    const size_t end = grp.off_xy + ((size_t)(grp.Sg - 1) * L.mcap + (size_t)counts[grp.s0 + grp.Sg - 1]) * 2 * sizeof(float);
    const size_t begin = odom ? grp.off_odom : grp.off_time;
    if (odom) std::memcpy(slot + grp.off_odom, odom + (size_t)grp.s0 * 4, sizeof(double) * 4 * grp.Sg);
    CK(cudaMemcpyAsync(grp.mb_dev + begin, slot + begin, end - begin, cudaMemcpyHostToDevice, grp.stream));
    CK(cudaEventRecord(grp.slot_done[grp.slot], grp.stream));
    int rc = 0;
    if (odom) {                                        // both messages of the step: the fused chain, as a graph if allowed
      InputRef in = grp.host_in;
      in.fuse_odom = 1;
      // measured: with several pipeline groups direct launches win on the host path (33.3 k vs 32 k steps/s); with one group,
      // where the step is one programmatic-launch chain, the captured chain does (35.3 k vs 34.4 k)
      const bool host_graphs = std::getenv("REKF_HOST_GRAPHS") ? std::atoi(std::getenv("REKF_HOST_GRAPHS")) != 0 : h->groups.empty();
      if (host_graphs && h->opts.use_graphs && !h->profiling) {
        if (!(rc = capture_step(h, grp, in, grp.host_gs))) rc = launch_step(h, grp, grp.host_gs);
      } else {
        rc = launch_observation(h, grp, in);
      }
    } else {
      rc = launch_observation(h, grp, grp.host_in);
    }
    if (rc) return rc;
  }
  return REKF_OK;
}

int rekf_batch_handle_observation(rekf_handle *h, const double *times, const float *xy, const int *counts, int m_stride) {
  if (!h || !times || !counts || (!xy && m_stride > 0)) return REKF_ERR_BAD_ARGUMENT;
  return observation_common(h, times, xy, counts, m_stride, nullptr);
}

int rekf_batch_handle_step(rekf_handle *h, const double *odom, const double *times, const float *xy, const int *counts, int m_stride) {
  if (!h || !odom || !times || !counts || (!xy && m_stride > 0)) return REKF_ERR_BAD_ARGUMENT;
  if (h->opts.use_imu) return observation_common(h, times, xy, counts, m_stride, nullptr);   // :213-222: odometry ignored
  return observation_common(h, times, xy, counts, m_stride, nullptr, odom);
}

int rekf_handle_observation(rekf_handle *h, double time, const float *xy, int m, const double *gps_pose_or_null) {
  if (!h || m < 0 || (m > 0 && !xy)) return REKF_ERR_BAD_ARGUMENT;
  if (h->L.S != 1) return fail(h, REKF_ERR_BAD_ARGUMENT, "rekf_handle_observation needs a single-session handle");
  double gps[4] = {0, 0, 0, 0};
  if (gps_pose_or_null) { gps[0] = 1.0; gps[1] = gps_pose_or_null[0]; gps[2] = gps_pose_or_null[1]; gps[3] = gps_pose_or_null[2]; }
  return observation_common(h, &time, xy, &m, m, gps_pose_or_null ? gps : nullptr);
}

int rekf_handle_imu(rekf_handle *h, double) { return h ? REKF_OK : REKF_ERR_BAD_ARGUMENT; }   // :224-227 empty

int rekf_replay_device(rekf_handle *h, const void *d_odom, const void *d_obs_time, const void *d_obs_xy, int T, int m, void *d_pose_out) {
  DeviceGuard dev_guard(h);
  if (!h || !d_odom || !d_obs_time || !d_obs_xy || T < 0 || m < 0) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  if (m > L.mcap) return fail(h, REKF_ERR_CAPACITY, "m %d exceeds max_observations %d", m, L.mcap);
  const bool graphs = h->opts.use_graphs && !h->profiling;
  std::vector<Group *> act = active_groups(h);
  std::vector<InputRef> ins(act.size());
  for (size_t gi = 0; gi < act.size(); ++gi) {
    Group &g = *act[gi];
    InputRef &in = ins[gi];
    in = InputRef{};
    in.odom = static_cast<const double *>(d_odom);
    in.obs_time = static_cast<const double *>(d_obs_time);
    in.obs_xy = static_cast<const float *>(d_obs_xy);
    in.obs_count = nullptr;
    in.gps = nullptr;
    in.odom_ss = (long long)T * 4;
    in.time_ss = T;
    in.xy_ss = (long long)T * m * 2;
    in.m_stride = m;
    in.m_fixed = m;
    in.step = g.L.step;
    in.fuse_odom = h->opts.use_imu ? 0 : 1;          // :213-222: with use_imu the reference ignores odometry messages
    in.pose_out = static_cast<double *>(d_pose_out);
    in.pose_ss = (long long)T * 3;
    CK(cudaMemsetAsync(g.L.step, 0, sizeof(int), g.stream));
    if (!graphs) continue;
    // The graphs are captured ONCE per group with a descriptor that only points at replay_in_dev; a new set of input
    // arrays rewrites those 100 bytes on the device, not the graph (no capture / instantiate inside a timed replay).
    CK(cudaMemcpyAsync(g.replay_in_dev, &in, sizeof(InputRef), cudaMemcpyHostToDevice, g.stream));
    InputRef via{};
    via.indirect = g.replay_in_dev;
    int rc = capture_step(h, g, via, g.replay_gs, true);
    if (rc) return rc;
  }
  // groups are issued round-robin; each one's stream orders its own steps, nothing orders groups against each other
  int t_first = 0;
  if (graphs && act.size() == 1 && act[0]->replay_gs.multi) {   // one group: kMultiSteps steps per graph launch
    Group &g = *act[0];
    for (; t_first + kMultiSteps <= T; t_first += kMultiSteps) {
      CK(cudaGraphLaunch(g.replay_gs.multi, g.stream));
      h->launches += g.replay_gs.launches * kMultiSteps;
    }
  }
  for (int t = t_first; t < T; ++t) {
    for (size_t gi = 0; gi < act.size(); ++gi) {
      Group &g = *act[gi];
      if (graphs) {
        int rc = launch_step(h, g, g.replay_gs);
        if (rc) return rc;
        continue;
      }
      int rc = launch_observation(h, g, ins[gi]);
      if (rc) return rc;
    }
  }
  CK(cudaGetLastError());
  return REKF_OK;
}

// ---- accessors --------------------------------------------------------------------------------
int rekf_sync(rekf_handle *h) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  CK(join_all(h));
  std::vector<SessionState> st(h->L.S);
  CK(cudaMemcpy(st.data(), h->L.st, sizeof(SessionState) * h->L.S, cudaMemcpyDeviceToHost));
  for (int s = 0; s < h->L.S; ++s) {
    if (st[s].flags & FLAG_NOT_SPD) return fail(h, REKF_ERR_NOT_SPD, "session %d: innovation matrix not positive definite", s);
    if (st[s].flags & FLAG_TCGEN05_TIMEOUT) return fail(h, REKF_ERR_CUDA, "session %d: tcgen05 SYRK barrier timeout", s);
    if (st[s].flags & FLAG_SYNC_TIMEOUT) return fail(h, REKF_ERR_CUDA, "session %d: TRSM gave up waiting for the Cholesky / gather flags", s);
    if (st[s].flags & (FLAG_LANDMARK_CAPACITY | FLAG_OBS_CAPACITY)) return fail(h, REKF_ERR_CAPACITY, "session %d: capacity exceeded (flags %d)", s, st[s].flags);
  }
  return REKF_OK;
}

int rekf_debug_copy(rekf_handle *h, int session, const char *name, void *out, size_t bytes) {
  DeviceGuard dev_guard(h);
  if (!h || !name || !out) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  if (session < 0 || session >= L.S) return fail(h, REKF_ERR_BAD_ARGUMENT, "session %d out of range", session);
  const std::string n(name);
  const void *src = nullptr;
  size_t size = 0;
  if (n == "sbuf") { src = L.Sbuf + (size_t)session * L.rld * L.sld; size = sizeof(double) * L.rld * L.sld; }
  else if (n == "dinv") { src = L.Dinv + (size_t)session * (L.rld / kCholNb) * kCholNb * kCholNb; size = sizeof(double) * (L.rld / kCholNb) * kCholNb * kCholNb; }
  else if (n == "wdiag" && L.Wdiag) { src = L.Wdiag + (size_t)session * L.ld; size = sizeof(double) * L.ld; }
  else if (n == "innov") { src = L.innov + (size_t)session * L.rcap; size = sizeof(double) * L.rcap; }
  else if (n == "qd") { src = L.Qd + (size_t)session * L.rcap; size = sizeof(double) * L.rcap; }
  else if (n == "tlog" && L.tlog) { src = L.tlog; size = sizeof(unsigned long long) * kTimelineCap; }
  else if (n == "state") { src = L.st + session; size = sizeof(SessionState); }
  else if (n == "mu") { src = L.mu + (size_t)session * L.ld; size = sizeof(double) * L.ld; }
  else if (n == "sigma") { src = L.sigma + (size_t)session * L.ld * L.ld; size = sizeof(double) * L.ld * L.ld; }
  else return fail(h, REKF_ERR_BAD_ARGUMENT, "unknown debug buffer %s", name);
  CK(join_all(h));
  CK(cudaMemcpy(out, src, std::min(bytes, size), cudaMemcpyDeviceToHost));
  return REKF_OK;
}

int rekf_device_error_flags(rekf_handle *h, int session, int *flags_out) {
  DeviceGuard dev_guard(h);
  if (!h || !flags_out) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  *flags_out = st.flags;
  return REKF_OK;
}

int rekf_dim(rekf_handle *h, int session) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  return rc ? rc : 3 + 2 * st.N;
}

int rekf_time(rekf_handle *h, int session, double *time_out) {
  DeviceGuard dev_guard(h);
  if (!h || !time_out) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  *time_out = st.time;
  return REKF_OK;
}

int rekf_get_mu(rekf_handle *h, int session, double *mu, int cap, int *n_out) {
  DeviceGuard dev_guard(h);
  if (!h || (!mu && cap > 0)) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  const int n = 3 + 2 * st.N;
  if (n_out) *n_out = n;
  const int c = std::min(n, cap);
  if (c <= 0) return REKF_OK;
  if ((rc = stage_reserve(h, n))) return rc;
  k_pack_mu<<<(n + 255) / 256, 256, 0, h->stream>>>(h->L, session, h->stage_dev);
  CK(cudaMemcpyAsync(mu, h->stage_dev, sizeof(double) * c, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

int rekf_get_pose(rekf_handle *h, int session, double pose[3], double cov33[9]) {
  DeviceGuard dev_guard(h);
  if (!h || !pose) return REKF_ERR_BAD_ARGUMENT;
  if (session < 0 || session >= h->L.S) return fail(h, REKF_ERR_BAD_ARGUMENT, "session %d out of range", session);
  const Layout &L = h->L;
  CK(join_all(h));
  CK(cudaMemcpyAsync(pose, L.mu + (size_t)session * L.ld, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->stream));
  if (cov33)
    CK(cudaMemcpy2DAsync(cov33, sizeof(double) * 3, L.sigma + (size_t)session * L.ld * L.ld, sizeof(double) * L.ld,
                         sizeof(double) * 3, 3, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (cov33)   // only the upper triangle is stored on the device (row i <= column j sits at cov33[j + 3 i] of the copy)
    for (int i = 0; i < 3; ++i)
      for (int j = i + 1; j < 3; ++j) cov33[i + 3 * j] = cov33[j + 3 * i];
  return REKF_OK;
}

int rekf_batch_get_pose(rekf_handle *h, double *poses) {
  DeviceGuard dev_guard(h);
  if (!h || !poses) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  // every group reads its own sessions on its own stream, then the host waits for all of them
  std::vector<Group *> act = active_groups(h);
  for (Group *g : act)
    CK(cudaMemcpy2DAsync(poses + (size_t)g->s0 * 3, sizeof(double) * 3, L.mu + (size_t)g->s0 * L.ld, sizeof(double) * L.ld,
                         sizeof(double) * 3, g->Sg, cudaMemcpyDeviceToHost, g->stream));
  for (Group *g : act) CK(cudaStreamSynchronize(g->stream));
  return REKF_OK;
}

// Streaming form of the per-step result read: every group copies its sessions' poses into a pinned ring behind its
// own step (stream-ordered, no host wait); the ticket is redeemed later.  At most Group::kPoseSlots tickets may be
// outstanding.
int rekf_batch_request_poses(rekf_handle *h, int64_t *ticket_out) {
  DeviceGuard dev_guard(h);
  if (!h || !ticket_out) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  const int slot = (int)(h->pose_ticket % Group::kPoseSlots);
  for (Group *g : active_groups(h)) {
    CK(cudaMemcpy2DAsync(g->pose_host + (size_t)slot * g->Sg * 3, sizeof(double) * 3, L.mu + (size_t)g->s0 * L.ld,
                         sizeof(double) * L.ld, sizeof(double) * 3, g->Sg, cudaMemcpyDeviceToHost, g->stream));
    CK(cudaEventRecord(g->pose_done[slot], g->stream));
  }
  *ticket_out = h->pose_ticket++;
  return REKF_OK;
}

int rekf_batch_fetch_poses(rekf_handle *h, int64_t ticket, double *poses) {
  DeviceGuard dev_guard(h);
  if (!h || !poses) return REKF_ERR_BAD_ARGUMENT;
  if (ticket < 0 || ticket >= h->pose_ticket || ticket + Group::kPoseSlots <= h->pose_ticket)
    return fail(h, REKF_ERR_BAD_ARGUMENT, "pose ticket %lld is not outstanding", (long long)ticket);
  const int slot = (int)(ticket % Group::kPoseSlots);
  for (Group *g : active_groups(h)) {
    CK(cudaEventSynchronize(g->pose_done[slot]));
    std::memcpy(poses + (size_t)g->s0 * 3, g->pose_host + (size_t)slot * g->Sg * 3, sizeof(double) * 3 * g->Sg);
  }
  return REKF_OK;
}

int rekf_get_landmarks(rekf_handle *h, int session, double *xy, double *cov2x2, int cap, int *count_out) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  if (count_out) *count_out = st.N;
  const int c = std::min(st.N, cap);
  if (c <= 0) return REKF_OK;
  if ((rc = stage_reserve(h, (size_t)st.N * 6))) return rc;
  double *dxy = h->stage_dev, *dcov = h->stage_dev + (size_t)st.N * 2;
  k_pack_landmarks<<<(st.N + 255) / 256, 256, 0, h->stream>>>(h->L, session, dxy, dcov);
  if (xy) CK(cudaMemcpyAsync(xy, dxy, sizeof(double) * 2 * c, cudaMemcpyDeviceToHost, h->stream));
  if (cov2x2) CK(cudaMemcpyAsync(cov2x2, dcov, sizeof(double) * 4 * c, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

int rekf_get_markers(rekf_handle *h, int session, double *markers, int cap, int *count_out) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  if (count_out) *count_out = st.N;
  const int c = std::min(st.N, cap);
  if (c <= 0 || !markers) return REKF_OK;
  if ((rc = stage_reserve(h, (size_t)st.N * 5))) return rc;
  k_pack_markers<<<(st.N + 255) / 256, 256, 0, h->stream>>>(h->L, session, h->stage_dev);
  CK(cudaMemcpyAsync(markers, h->stage_dev, sizeof(double) * 5 * c, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

int rekf_get_sigma(rekf_handle *h, int session, double *sigma, int ld) {
  DeviceGuard dev_guard(h);
  if (!h || !sigma) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  const int n = 3 + 2 * st.N;
  if (ld < n) return fail(h, REKF_ERR_BAD_ARGUMENT, "ld %d < n %d", ld, n);
  if ((rc = stage_reserve(h, (size_t)n * n))) return rc;
  k_pack_sigma<<<dim3((n + 255) / 256, n), 256, 0, h->stream>>>(h->L, session, h->stage_dev, n);
  CK(cudaMemcpy2DAsync(sigma, sizeof(double) * ld, h->stage_dev, sizeof(double) * n, sizeof(double) * n, n,
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

// GetState() (reflector_ekf_slam.h:37-40) in ONE stream synchronisation: time, mu, the full covariance and the sticky
// device flags.  The caller passes the dimension it expects (its mirror's size); when the state has grown the call
// returns REKF_ERR_CAPACITY with *n_out set and nothing copied, and the caller resizes and repeats.
int rekf_get_state(rekf_handle *h, int session, int n_expect, double *time_out, double *mu, double *sigma, int ld, int *n_out, int *flags_out) {
  DeviceGuard dev_guard(h);
  if (!h || n_expect < 3 || !mu || (sigma && ld < n_expect)) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  if (session < 0 || session >= L.S) return fail(h, REKF_ERR_BAD_ARGUMENT, "session %d out of range", session);
  const size_t n = (size_t)n_expect;
  int rc = stage_reserve(h, 8 + n + (sigma ? n * n : 0));
  if (rc) return rc;
  if (!h->hdr_host) CK(cudaMallocHost(&h->hdr_host, 8 * sizeof(double)));
  double *hdr = h->stage_dev, *dmu = h->stage_dev + 8, *dsig = sigma ? dmu + n : nullptr;
  // behind every group's work: the getter's stream waits for the groups instead of the host joining them first
  for (auto &g : h->groups) {
    if (g.stream == h->stream) continue;
    CK(cudaEventRecord(g.done, g.stream));
    CK(cudaStreamWaitEvent(h->stream, g.done, 0));
  }
  k_pack_state<<<dim3((n_expect + 255) / 256, sigma ? n_expect : 1), 256, 0, h->stream>>>(L, session, n_expect, hdr, dmu, dsig);
  CK(cudaMemcpyAsync(h->hdr_host, hdr, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(mu, dmu, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (sigma)
    CK(cudaMemcpy2DAsync(sigma, sizeof(double) * ld, dsig, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int n_dev = (int)h->hdr_host[0];
  if (n_out) *n_out = n_dev;
  if (time_out) *time_out = h->hdr_host[1];
  if (flags_out) *flags_out = (int)h->hdr_host[2];
  if (n_dev != n_expect) return fail(h, REKF_ERR_CAPACITY, "state dimension is %d, caller expected %d", n_dev, n_expect);
  return REKF_OK;
}

// Page-lock a caller-owned host buffer (the adapter's covariance mirror) so that rekf_get_state / rekf_get_sigma copy into
// it at PCIe speed instead of through the driver's pageable staging.  Thin wrappers: no CUDA types cross the ABI.
int rekf_host_register(rekf_handle *h, void *ptr, size_t bytes) {
  DeviceGuard dev_guard(h);
  if (!h || !ptr || !bytes) return REKF_ERR_BAD_ARGUMENT;
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return REKF_OK;
}
int rekf_host_unregister(rekf_handle *h, void *ptr) {
  DeviceGuard dev_guard(h);
  if (!h || !ptr) return REKF_ERR_BAD_ARGUMENT;
  CK(cudaHostUnregister(ptr));
  return REKF_OK;
}

// cumulative counters of one session: {updates, whole frames on the fp64 SYRK, flagged slots done by k_syrk_exact_rows}
int rekf_get_counters(rekf_handle *h, int session, int64_t out[3]) {
  DeviceGuard dev_guard(h);
  if (!h || !out) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  out[0] = st.n_updates; out[1] = st.n_exact_frames; out[2] = st.n_exact_slots;
  return REKF_OK;
}

int rekf_get_match_result(rekf_handle *h, int session, int *state_pairs, int *n_state, int *map_pairs, int *n_map,
                          int *new_ids, int *n_new, int cap) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  const Layout &L = h->L;
  if (n_state) *n_state = st.M;
  if (n_map) *n_map = st.Mmap;
  if (n_new) *n_new = st.N2;
  if (state_pairs && st.M > 0)
    CK(cudaMemcpy(state_pairs, L.state_pairs + (size_t)session * L.mcap * 2, sizeof(int) * 2 * std::min(st.M, cap), cudaMemcpyDeviceToHost));
  if (map_pairs && st.Mmap > 0)
    CK(cudaMemcpy(map_pairs, L.map_pairs + (size_t)session * L.mcap * 2, sizeof(int) * 2 * std::min(st.Mmap, cap), cudaMemcpyDeviceToHost));
  if (new_ids && st.N2 > 0)
    CK(cudaMemcpy(new_ids, L.new_ids + (size_t)session * L.mcap, sizeof(int) * std::min(st.N2, cap), cudaMemcpyDeviceToHost));
  return REKF_OK;
}

int rekf_predict_state(rekf_handle *h, int session, double time, double *mu, int cap, double *sigma, int ld) {
  DeviceGuard dev_guard(h);
  if (!h || !mu) return REKF_ERR_BAD_ARGUMENT;
  SessionState st;
  int rc = read_state(h, session, &st);
  if (rc) return rc;
  const int n = 3 + 2 * st.N;
  if (cap < n || (sigma && ld < n)) return fail(h, REKF_ERR_BAD_ARGUMENT, "buffers too small for n = %d", n);
  if ((rc = stage_reserve(h, (size_t)n * n + n))) return rc;
  double *dmu = h->stage_dev, *dsig = sigma ? h->stage_dev + n : nullptr;
  k_predict_state<<<dim3((n + 255) / 256, sigma ? n : 1), 256, 0, h->stream>>>(h->L, session, time, dmu, dsig, n);
  CK(cudaMemcpyAsync(mu, dmu, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (sigma)
    CK(cudaMemcpy2DAsync(sigma, sizeof(double) * ld, dsig, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

// ---- state injection / persistence ----------------------------------------------------------------
int rekf_set_state(rekf_handle *h, int session, double time, const double vt[3], const double *mu, int n, const double *sigma, int ld) {
  DeviceGuard dev_guard(h);
  if (!h || !mu || !sigma || n < 3 || ((n - 3) & 1) || ld < n) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  if (session < 0 || session >= L.S) return fail(h, REKF_ERR_BAD_ARGUMENT, "session %d out of range", session);
  const int N = (n - 3) / 2;
  if (N > L.Ncap) return fail(h, REKF_ERR_CAPACITY, "%d landmarks exceed max_landmarks %d", N, L.Ncap);
  int rc = stage_reserve(h, (size_t)n * n + n);
  if (rc) return rc;
  CK(join_all(h));
  double *dmu = h->stage_dev, *dsig = h->stage_dev + n;
  CK(cudaMemcpyAsync(dmu, mu, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(dsig, sizeof(double) * n, sigma, sizeof(double) * ld, sizeof(double) * n, n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(L.sigma + (size_t)session * L.ld * L.ld, 0, sizeof(double) * L.ld * L.ld, h->stream));
  CK(cudaMemsetAsync(L.mu + (size_t)session * L.ld, 0, sizeof(double) * L.ld, h->stream));
  k_unpack_state<<<dim3((n + 255) / 256, n), 256, 0, h->stream>>>(L, session, dmu, dsig, n, n);
  SessionState st;
  std::memset(&st, 0, sizeof(st));
  st.time = time;
  if (vt) for (int i = 0; i < 3; ++i) st.vt[i] = vt[i];
  st.N = N;
  CK(cudaMemcpyAsync(L.st + session, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return REKF_OK;
}

int rekf_set_map(rekf_handle *h, const float *xy, const double *cov2x2, int count) {
  DeviceGuard dev_guard(h);
  if (!h || count < 0 || (count > 0 && (!xy || !cov2x2))) return REKF_ERR_BAD_ARGUMENT;
  const Layout &L = h->L;
  if (count > L.mapcap) return fail(h, REKF_ERR_CAPACITY, "%d beacons exceed max_map_landmarks %d", count, L.mapcap);
  h->map_xy.assign(xy, xy + 2 * (size_t)count);
  h->map_cov.assign(cov2x2, cov2x2 + 4 * (size_t)count);
  CK(join_all(h));
  if (count > 0) {
    CK(cudaMemcpy(L.map_xy, xy, sizeof(float) * 2 * count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.map_cov, cov2x2, sizeof(double) * 4 * count, cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(L.map_count, &count, sizeof(int), cudaMemcpyHostToDevice));
  return REKF_OK;
}

int rekf_get_map(rekf_handle *h, float *xy, double *cov2x2, int cap, int *count_out) {
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  const int n = (int)h->map_xy.size() / 2;
  if (count_out) *count_out = n;
  const int c = std::min(n, cap);
  if (xy && c > 0) std::memcpy(xy, h->map_xy.data(), sizeof(float) * 2 * c);
  if (cov2x2 && c > 0) std::memcpy(cov2x2, h->map_cov.data(), sizeof(double) * 4 * c);
  return REKF_OK;
}

// LoadMapFromTxtFile (:43-95).  Missing file / wrong shape / unparsable token: silent no-op.
int rekf_load_map_txt(rekf_handle *h, const char *path) {
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  if (!path || !path[0]) return REKF_OK;                       // :45
  std::ifstream in(path);
  if (!in.good()) return REKF_OK;                              // :45-46, :67-72
  std::vector<std::vector<double>> rows;
  std::string line;
  while (std::getline(in, line)) {                             // :52
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (line.empty()) continue;                                // :55
    std::vector<double> vec;
    for (const auto &tok : split_commas(line)) {               // :58-62
      char *end = nullptr;
      const double v = std::strtod(tok.c_str(), &end);
      if (end == tok.c_str()) return REKF_OK;                  // std::stod would throw; here: no-op
      vec.push_back(v);
    }
    rows.push_back(vec);
  }
  if (rows.size() != 2 || rows[1].size() != 2 * rows[0].size()) return REKF_OK;   // :74
  const int M_ = (int)rows[0].size() / 2;                      // :83
  std::vector<float> xy(2 * (size_t)M_);
  std::vector<double> cov(4 * (size_t)M_, 0.0);
  for (int i = 0; i < M_; ++i) {                               // :85 double → float
    xy[2 * i] = (float)rows[0][2 * i];
    xy[2 * i + 1] = (float)rows[0][2 * i + 1];
  }
  for (int i = 0; i < M_; ++i)
    for (int e = 0; e < 4; ++e) {
      if (h->opts.map_loader == REKF_MAP_LOADER_REFERENCE) {   // :90 reads the positions line
        const size_t idx = 4 * (size_t)i + e;
        cov[4 * i + e] = idx < rows[0].size() ? rows[0][idx] : 0.0;
      } else {
        cov[4 * i + e] = rows[1][4 * (size_t)i + e];
      }
    }
  return rekf_set_map(h, xy.data(), cov.data(), M_);           // :93-94
}

// Node::SaveReflectorResult (ros_node.cc:75-140): same two lines, default ostream formatting.
int rekf_save_map_txt(rekf_handle *h, int session, const char *filebase) {
  if (!h || !filebase) return REKF_ERR_BAD_ARGUMENT;
  int N = 0;
  int rc = rekf_get_landmarks(h, session, nullptr, nullptr, 0, &N);
  if (rc) return rc;
  std::vector<double> xy(2 * (size_t)std::max(N, 1)), cov(4 * (size_t)std::max(N, 1));
  if (N > 0 && (rc = rekf_get_landmarks(h, session, xy.data(), cov.data(), N, nullptr))) return rc;
  const std::string path = std::string(filebase) + ".txt";    // :80
  std::ofstream out(path.c_str(), std::ios::out);
  if (!out.good()) return fail(h, REKF_ERR_IO, "cannot open %s", path.c_str());
  const int Mm = (int)h->map_xy.size() / 2;
  for (int i = 0; i < Mm; ++i) {                               // :87-97
    out << h->map_xy[2 * i] << "," << h->map_xy[2 * i + 1];
    if (i != Mm - 1) out << ",";
  }
  if (N > 0) {                                                 // :98-110 (leading comma unconditional)
    out << ",";
    for (int i = 0; i < N; ++i) {
      out << xy[2 * i] << "," << xy[2 * i + 1];
      if (i != N - 1) out << ",";
    }
  }
  out << std::endl;
  for (int i = 0; i < Mm; ++i) {                               // :112-123
    out << h->map_cov[4 * i] << "," << h->map_cov[4 * i + 1] << "," << h->map_cov[4 * i + 2] << "," << h->map_cov[4 * i + 3];
    if (i != Mm - 1) out << ",";
  }
  if (N > 0) {                                                 // :124-137
    out << ",";
    for (int i = 0; i < N; ++i) {
      out << cov[4 * i] << "," << cov[4 * i + 1] << "," << cov[4 * i + 2] << "," << cov[4 * i + 3];
      if (i != N - 1) out << ",";
    }
  }
  out << std::endl;
  out.close();
  return REKF_OK;
}

// ---- timing ---------------------------------------------------------------------------------------
int rekf_timer_start(rekf_handle *h) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  CK(join_all(h));                                   // the GPU is idle: t0 is stamped now, before any group's work
  CK(cudaEventRecord(h->t0, h->stream));
  return REKF_OK;
}
int rekf_timer_stop(rekf_handle *h, float *ms) {
  DeviceGuard dev_guard(h);
  if (!h || !ms) return REKF_ERR_BAD_ARGUMENT;
  for (auto &g : h->groups) {                        // t1 is stamped after the last group has finished
    CK(cudaEventRecord(g.done, g.stream));
    CK(cudaStreamWaitEvent(h->stream, g.done, 0));
  }
  CK(cudaEventRecord(h->t1, h->stream));
  CK(cudaEventSynchronize(h->t1));
  CK(cudaEventElapsedTime(ms, h->t0, h->t1));
  return REKF_OK;
}
int rekf_profile_enable(rekf_handle *h, int enable) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  CK(join_all(h));
  drain_profile(h);
  h->profiling = enable != 0;
  for (auto &g : h->groups) g.stream = h->profiling ? h->stream : g.home_stream;
  if (enable) {
    std::memset(h->prof_us, 0, sizeof(h->prof_us));
    std::memset(h->prof_calls, 0, sizeof(h->prof_calls));
  }
  return REKF_OK;
}
int rekf_profile_read(rekf_handle *h, const char **names, double *mean_us, int *calls, int cap, int *count_out) {
  DeviceGuard dev_guard(h);
  if (!h) return REKF_ERR_BAD_ARGUMENT;
  CK(join_all(h));
  drain_profile(h);
  int c = 0;
  for (int k = 0; k < K_COUNT && c < cap; ++k) {
    static const char *syrk_names[3] = {"k_syrk_tcgen05", "k_syrk_f64", "k_syrk_tcgen05_i8"};
    if (names) names[c] = (k == K_SYRK) ? syrk_names[h->opts.cov_update] : kKernelNames[k];
    if (mean_us) mean_us[c] = h->prof_calls[k] ? h->prof_us[k] / h->prof_calls[k] : 0.0;
    if (calls) calls[c] = h->prof_calls[k];
    ++c;
  }
  if (count_out) *count_out = c;
  return REKF_OK;
}

}  // extern "C"
