// gsched.cu -- sm_100a discrete-event engine + C ABI (include/gsched.h).
//
// Execution model (B200-first, not a translation of the Python object graph):
//   * one WARP owns one simulation replica ("sim").  The cluster's node table
//     lives in shared memory as (busy-device mask, charged task slots) -- the
//     reference's cpu_used/mem_used always move together in units of 12/60 per
//     task (job.py:105-106, node.py:204-205), so both collapse into one slot
//     counter k with  free_slots = min(cpu/12, mem/60) - k.
//   * lanes stripe over nodes; single-node first fit is a ballot + ffs (argmin
//     over node id), cross-node fill is a warp prefix sum over per-node task
//     capacities with a cut-off -- no per-device objects are ever walked.
//   * all reference per-tick re-scans (pandas filters, _construct_info,
//     pending-time aging, time_processed stepping) are replaced by closed
//     forms and O(1) incremental counters; completions come from a timing
//     wheel keyed by finish tick, appended in start order.
//   * the loop is event stepped: ticks on which nothing arrives, starts or
//     finishes are jumped over; every other tick leaves one 24-byte record of
//     integer aggregates (+ one for the queue while it is non-empty) from which
//     the 64-byte gs_tick_row of EVERY tick is rebuilt on demand; the job table
//     is streamed once from HBM, results are written once (4-byte start tick per job +
//     8-byte gs_cspan per (job,node) + finish order).
//   * thousands of replicas run per launch (one warp each, 148 SMs x 32
//     warps); a single replica is latency bound by construction.
//
// Reference semantics followed (paths relative to the reference root):
//   Scheduler.start            core/scheduling/schedule.py:178-215
//   Scheduler._schedule        core/scheduling/schedule.py:40-60
//   schedule_fifo              core/scheduling/algorithm.py:189-202
//   ms_yarn_placement          core/scheduling/algorithm.py:28-32
//   try_single_node_alloc_ms   core/scheduling/algorithm.py:396-417
//   try_cross_node_alloc_ms    core/scheduling/algorithm.py:301-393
//   Node fit/reserve/release   infra/node.py:71-91,109-127,146-171,200-275
//   Device.can_fit             infra/device.py:67-77
//   gen_jobs / step / finish   core/jobs/jobs_manager.py:65-87,143-148,228-250
//   calculate_network_costs    core/network/network_service.py:3-39
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "gsched.h"

#include "gs_common.cuh"
#include "gs_tick2.cuh"
#include "gs_policy.cuh"
#include "gs_aux.cuh"
#include "gs_switch.cuh"

// ------------------------------------------------------------------ host side

// Per-replica device layout.  The RESULT arrays come first and contiguously (gs_result_layout): statistics records,
// queue records, per-job results, (durations,) finish order, spans -- so that everything a caller reads back from a
// replica is ONE copy, and from all replicas of a handle ONE strided copy (the slabs live side by side in an arena).
struct SimLayout {
  size_t o_ev = 0, o_q = 0, o_ne = 0, o_rec2 = 0, o_dur2 = 0, o_fin = 0, o_spans = 0, out_bytes = 0;
  size_t o_rec = 0, o_rows = 0, o_jst = 0, o_stack = 0, o_wh = 0, o_wm = 0, o_nb = 0, o_nk = 0;
  size_t o_pj = 0, o_run = 0, o_qs = 0, o_end = 0, o_tmp = 0, o_ci = 0, o_ck = 0, o_stale = 0;
  size_t total = 0;
  int64_t rows_cap = 0, qrows_cap = 0;
  int W = 256;
};

struct SimHost {
  gs_cluster cl;
  gs_policy pol;
  bool configured = false, loaded = false, prepared = false;
  int64_t n = 0;
  void *trace_slab = nullptr;
  void *state_slab = nullptr;
  void *git_dev = nullptr;
  long long git_direct_n = 0;       // entries of the direct look-up table behind the two gittins tables in git_dev
  size_t trace_bytes = 0, state_bytes = 0;
  int64_t span_cap = 0, rows_cap = 0, qrows_cap = 0, last_arrive = 0;
  int max_need = 1;
  unsigned char *state_ptr = nullptr;     // the replica's slab: inside the handle's arena or its own allocation (state_slab)
  bool trace_in_arena = false;
  SimDev dev;
  SimLayout layout;
};

struct gs_engine {
  int device = 0, nsims = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEvent_t e_wait = nullptr;   // blocking-sync event: in asynchronous mode the driving thread sleeps while it waits
  std::vector<SimHost> sims;
  SimDev *d_sims = nullptr;
  void *h_stage = nullptr;
  size_t h_stage_bytes = 0;
  void *d_scratch = nullptr;
  size_t d_scratch_bytes = 0;
  std::vector<SimDev> h_back;   // pinned-size-stable host mirror used by gs_run
  // arenas: the replicas' state slabs (result block first) / traces side by side with one stride, so that a whole
  // handle is read back or uploaded with ONE strided copy (gs_fetch_results / gs_load_traces_packed)
  void *arena = nullptr; size_t arena_bytes = 0, arena_stride = 0;
  void *tarena = nullptr; size_t tarena_bytes = 0, tarena_stride = 0;
  std::string err;
  double kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
  long long launches = 0;  // kernels launched by this handle
  int engine_mode = 0;     // event-driven policies: 0 warp-cooperative kernels, 2 one thread per replica
  double span_budget = 0;  // > 0: span pool = min(worst case, budget * n + 4096) records per replica
  int64_t qrows_cap = 0;   // 0: same as rows_cap
  bool async = false;      // gs_set_async
  // sharded single simulation (gs_comm_prepare / gs_comm_init)
  void *comm_buf = nullptr; int64_t comm_cap = 0; int comm_rank = 0, comm_n = 0;
  void *comm_peer[GS_MAX_RANKS] = {nullptr}; bool comm_opened[GS_MAX_RANKS] = {false};
  int comm_min_runnable = 0x7fffffff;      // gs_comm_set_min_runnable: no exchange unless the caller asks for one (include/gsched.h)
  unsigned long long comm_epoch = 0, comm_epoch0 = 0;   // exchange counter: continues across runs / value at the last prepare
  bool dirty = true;       // host mirror of SimDev newer than device copy
};

static std::string g_create_err;

static int fail(gs_handle h, int code, const std::string &msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(h, GS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Wait for the handle's stream.  Asynchronous mode is what many host threads use side by side (one handle each): there the
// waiting thread sleeps on a blocking-sync event instead of spinning, which leaves the cores to the threads that work.
static cudaError_t wait_stream(gs_handle h) {
  if (!h->async) return cudaStreamSynchronize(h->stream);
  cudaError_t e = cudaEventRecord(h->e_wait, h->stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(h->e_wait);
}

extern "C" int gs_abi_version(void) { return GS_ABI_VERSION; }
extern "C" const char *gs_build_tag(void) {
#ifdef __CUDACC__
  return "cuda:sm_100a";
#else
  return "host-emulation";
#endif
}

extern "C" const char *gs_last_error(gs_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int gs_create(int device, int nsims, gs_handle *out) {
  gs_handle h = nullptr;
  if (!out || nsims <= 0) return fail(nullptr, GS_ERR_ARG, "gs_create: bad arguments");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(nullptr, GS_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= count) return fail(nullptr, GS_ERR_ARG, "gs_create: device ordinal out of range");
  CU(cudaSetDevice(device));
  h = new gs_engine();
  h->device = device;
  h->nsims = nsims;
  h->sims.resize((size_t)nsims);
  h->h_back.resize((size_t)nsims);
  cudaError_t e1 = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  cudaError_t e2 = cudaEventCreate(&h->e0);
  cudaError_t e3 = cudaEventCreate(&h->e1);
  cudaError_t e4 = cudaMalloc(&h->d_sims, sizeof(SimDev) * (size_t)nsims);
  if (cudaEventCreateWithFlags(&h->e_wait, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) e4 = cudaErrorUnknown;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
    delete h;
    return fail(nullptr, GS_ERR_CUDA, "gs_create: stream/event/alloc failed");
  }
  *out = h;
  return GS_OK;
}

extern "C" void gs_destroy(gs_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto &s : h->sims) { if (s.trace_slab) cudaFree(s.trace_slab); if (s.state_slab) cudaFree(s.state_slab); if (s.git_dev) cudaFree(s.git_dev); }
  if (h->d_sims) cudaFree(h->d_sims);
  if (h->arena) cudaFree(h->arena);
  if (h->tarena) cudaFree(h->tarena);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->d_scratch) cudaFree(h->d_scratch);
  for (int q = 0; q < GS_MAX_RANKS; ++q) if (h->comm_opened[q] && h->comm_peer[q]) cudaIpcCloseMemHandle(h->comm_peer[q]);
  if (h->comm_buf) cudaFree(h->comm_buf);
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  if (h->e_wait) cudaEventDestroy(h->e_wait);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int check_cluster(gs_handle h, const gs_cluster *c) {
  if (!c) return fail(h, GS_ERR_ARG, "cluster is NULL");
  long long m = (long long)c->num_switch * c->num_node_p_switch;
  if (c->num_switch <= 0 || c->num_node_p_switch <= 0 || m > (1 << 20))
    return fail(h, GS_ERR_ARG, "cluster: num_switch * num_node_p_switch must be in 1..2^20");
  if (c->num_gpu_p_node <= 0 || c->num_gpu_p_node > GS_MAX_GPUS_PER_NODE)
    return fail(h, GS_ERR_ARG, "cluster: num_gpu_p_node must be in 1..64");
  if (c->cpu_per_task <= 0 || c->mem_per_task <= 0 || c->num_cpu_p_node < 0 || c->mem_p_node < 0)
    return fail(h, GS_ERR_ARG, "cluster: cpu/mem per task must be positive");
  if (c->gpu_mem_cap_mib <= 0) return fail(h, GS_ERR_ARG, "cluster: gpu_mem_cap_mib must be positive");
  return GS_OK;
}

extern "C" int gs_config_sim(gs_handle h, int sim, const gs_cluster *cluster, const gs_policy *policy) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_config_sim: sim index out of range");
  int rc = check_cluster(h, cluster);
  if (rc) return rc;
  SimHost &s = h->sims[(size_t)sim];
  if (s.prepared) return fail(h, GS_ERR_STATE, "gs_config_sim: replica already running");
  // validate into locals, commit only on success (a rejected call leaves the replica as it was)
  gs_policy pol;
  if (policy) pol = *policy; else { memset(&pol, 0, sizeof(pol)); pol.num_queue = 1; }
  if (pol.schedule < GS_SCHED_FIFO || pol.schedule > GS_SCHED_GITTINS)
    return fail(h, GS_ERR_ARG, "gs_config_sim: unknown schedule");
  if (pol.scheme != GS_SCHEME_YARN && pol.scheme != GS_SCHEME_COUNT)
    return fail(h, GS_ERR_ARG, "gs_config_sim: unknown scheme");
  if (pol.schedule == GS_SCHED_FIFO && pol.scheme != GS_SCHEME_YARN)
    return fail(h, GS_ERR_ARG, "gs_config_sim: fifo runs with the yarn scheme only");
  if (pol.schedule == GS_SCHED_FIFO &&
      ((long long)cluster->num_switch * cluster->num_node_p_switch * cluster->num_gpu_p_node > 65535 ||
       (long long)cluster->num_cpu_p_node / cluster->cpu_per_task > 32767))
    return fail(h, GS_ERR_ARG, "gs_config_sim: the fifo engine packs its counters for clusters of at most 65535 GPUs");
  if ((pol.schedule == GS_SCHED_DLAS || pol.schedule == GS_SCHED_DLAS_GPU) &&
      (pol.num_queue < 1 || pol.num_queue > GS_MAX_QUEUES))
    return fail(h, GS_ERR_ARG, "gs_config_sim: num_queue must be in 1..8 for dlas");
  void *git_dev = nullptr;
  long long git_direct_n = 0;
  if (pol.schedule == GS_SCHED_GITTINS) {
    if (pol.gittins_n < 1 || !pol.gittins_data || !pol.gittins_index)
      return fail(h, GS_ERR_ARG, "gs_config_sim: gittins needs the (data, index) tables");
    CU(cudaSetDevice(h->device));
    const size_t bytes = 8 * (size_t)pol.gittins_n;
    // direct form of the look-up "index of the first sample above a" for whole-number a (attained service always is one):
    // one entry per integer up to the largest sample, when that is not out of proportion with the table itself
    std::vector<double> direct;
    if (pol.gittins_n >= 2) {
      const double last = pol.gittins_data[pol.gittins_n - 2];
      if (last >= 0.0 && last < 8.0 * (double)pol.gittins_n + 65536.0 && last < 67108864.0) {
        direct.resize((size_t)last + 1);
        size_t idx = 0;
        for (size_t k = 0; k < direct.size(); ++k) {
          while (idx + 1 < (size_t)pol.gittins_n && !(pol.gittins_data[idx] > (double)k)) ++idx;
          direct[k] = pol.gittins_index[idx];
        }
      }
    }
    git_direct_n = (long long)direct.size();
    CU(cudaMalloc(&git_dev, 2 * bytes + 8 * direct.size()));
    cudaError_t e1 = cudaMemcpy(git_dev, pol.gittins_data, bytes, cudaMemcpyHostToDevice);
    cudaError_t e2 = cudaMemcpy((unsigned char *)git_dev + bytes, pol.gittins_index, bytes, cudaMemcpyHostToDevice);
    cudaError_t e3 = direct.empty() ? cudaSuccess : cudaMemcpy((unsigned char *)git_dev + 2 * bytes, direct.data(), 8 * direct.size(), cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cudaFree(git_dev); return fail(h, GS_ERR_CUDA, "gs_config_sim: gittins table upload failed"); }
  }
  if (s.git_dev) cudaFree(s.git_dev);
  s.git_dev = git_dev;
  s.git_direct_n = git_direct_n;
  s.cl = *cluster;
  s.pol = pol;
  s.pol.gittins_data = s.pol.gittins_index = nullptr;   // host pointers are not retained past this call
  s.loaded = false;       // load-time bounds (max_need, span_cap) depend on the cluster: a reconfigured replica needs its trace again
  s.configured = true;
  h->dirty = true;
  return GS_OK;
}

static int ensure_stage(gs_handle h, size_t bytes) {
  if (h->h_stage_bytes >= bytes) return GS_OK;
  if (h->h_stage) { cudaStreamSynchronize(h->stream); cudaFreeHost(h->h_stage); }
  h->h_stage = nullptr; h->h_stage_bytes = 0;
  CU(cudaMallocHost(&h->h_stage, bytes));
  h->h_stage_bytes = bytes;
  return GS_OK;
}

// Validation + load-time bounds over n packed records (read only).
static int scan_trace(gs_handle h, const SimHost &s, int64_t n, const JobIn *ji, bool net, const double *model_mb,
                      const double *iterations, int64_t *span_cap_out, double *max_need_out) {
  const int M = s.cl.num_switch * s.cl.num_node_p_switch;
  const bool netcost = net && s.cl.enable_network_costs;
  int64_t span_cap = 0;
  double max_need = 1.0;
  int prev = 0;
  for (int64_t j = 0; j < n; ++j) {
    const JobIn r = ji[j];
    if (r.arrive < prev) return fail(h, GS_ERR_ARG, "gs_load_trace: arrive_tick must be non-negative and non-decreasing");
    prev = r.arrive;
    if (r.gpc <= 0 || r.gpus < r.gpc || r.gpus % r.gpc != 0 || r.gpus >= (1 << 24) || r.gpc > 255)
      return fail(h, GS_ERR_ARG, "gs_load_trace: gpus must be a positive multiple of gpu_per_task below 2^24 (job.py:96-100)");
    if (r.memb < 0) return fail(h, GS_ERR_ARG, "gs_load_trace: negative mem_bytes");
    if (!(r.dur == r.dur)) return fail(h, GS_ERR_ARG, "gs_load_trace: NaN duration");
    const int64_t tasks = r.gpus / r.gpc;
    span_cap += tasks < M ? tasks : M;
    double d = r.dur;
    if (netcost && r.ps > 1) {
      const double cross = (double)(tasks < M ? tasks : M);
      const double extra = (model_mb[j] / s.cl.bandwidth + cross * s.cl.internode_latency) * (iterations[j] * 2.0);
      if (extra > 0) d += extra;
    }
    if (d > max_need) max_need = d;
  }
  *span_cap_out = span_cap; *max_need_out = max_need;
  return GS_OK;
}

static int load_common(gs_handle h, int sim, int64_t n, const JobIn *packed, const int32_t *arrive_tick,
                       const int32_t *gpus, const int32_t *gpu_per_task, const double *duration,
                       const int64_t *mem_bytes, const double *model_mb, const double *iterations,
                       const int32_t *ps_count) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_load_trace: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.configured) return fail(h, GS_ERR_STATE, "gs_load_trace: call gs_config_sim first");
  if (n < 0 || n >= (1ll << 31) - 64) return fail(h, GS_ERR_ARG, "gs_load_trace: n out of range");
  if (n > 0 && !packed && (!arrive_tick || !gpus || !gpu_per_task || !duration || !mem_bytes))
    return fail(h, GS_ERR_ARG, "gs_load_trace: NULL column");
  const bool net = model_mb && iterations && (packed || ps_count);
  CU(cudaSetDevice(h->device));
  const size_t N = (size_t)(n > 0 ? n : 1);
  const size_t off_model = align_up(sizeof(JobIn) * N), off_iters = align_up(off_model + (net ? 8 * N : 0));
  const size_t total = align_up(off_iters + (net ? 8 * N : 0));
  // asynchronous path: packed records in page-locked caller memory go to the device without staging
  bool direct = false;
  if (h->async && packed && !net && n > 0) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, packed) == cudaSuccess && at.type == cudaMemoryTypeHost) direct = true;
    else (void)cudaGetLastError();
  }
  const JobIn *ji = packed;
  if (!direct) {
    CU(cudaStreamSynchronize(h->stream));         // the staging buffer may still feed an earlier upload
    int rc = ensure_stage(h, total);
    if (rc) return rc;
    unsigned char *st = (unsigned char *)h->h_stage;
    JobIn *dst = (JobIn *)st;
    if (packed) memcpy(dst, packed, sizeof(JobIn) * (size_t)n);
    else
      for (int64_t j = 0; j < n; ++j) {
        JobIn r;
        r.arrive = arrive_tick[j]; r.gpus = gpus[j]; r.gpc = gpu_per_task[j]; r.ps = net ? ps_count[j] : 0;
        r.memb = mem_bytes[j]; r.dur = duration[j];
        dst[j] = r;
      }
    if (net && n > 0) {
      memcpy(st + off_model, model_mb, 8 * (size_t)n);
      memcpy(st + off_iters, iterations, 8 * (size_t)n);
    }
    ji = dst;
  }
  int64_t span_cap = 0;
  double max_need = 1.0;
  int rc = scan_trace(h, s, n, ji, net, model_mb, iterations, &span_cap, &max_need);
  if (rc) return rc;
  if (max_need > (double)(1 << 26)) return fail(h, GS_ERR_ARG, "gs_load_trace: job duration exceeds 2^26 ticks");
  if (s.trace_slab && s.trace_bytes < total) { CU(cudaStreamSynchronize(h->stream)); cudaFree(s.trace_slab); s.trace_slab = nullptr; }
  if (!s.trace_slab) { CU(cudaMalloc(&s.trace_slab, total)); s.trace_bytes = total; }
  if (direct) {
    CU(cudaMemcpyAsync(s.trace_slab, packed, sizeof(JobIn) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  } else {
    CU(cudaEventRecord(h->e0, h->stream));
    CU(cudaMemcpyAsync(s.trace_slab, h->h_stage, total, cudaMemcpyHostToDevice, h->stream));
    CU(cudaEventRecord(h->e1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
    h->h2d_ms += ms;
  }
  unsigned char *d = (unsigned char *)s.trace_slab;
  SimDev &D = s.dev;
  memset(&D, 0, sizeof(D));
  D.jobs = (const JobIn *)d;
  D.model_mb = net ? (const double *)(d + off_model) : nullptr;
  D.iters = net ? (const double *)(d + off_iters) : nullptr;
  if (h->span_budget > 0) {
    const int64_t lim = (int64_t)(h->span_budget * (double)n) + 4096;
    if (span_cap > lim) span_cap = lim;
  }
  s.n = n; s.span_cap = span_cap > 0 ? span_cap : 1;
  s.max_need = (int)max_need + 2;
  s.last_arrive = n > 0 ? ji[n - 1].arrive : 0;
  s.loaded = true;
  s.trace_in_arena = false;
  s.prepared = false;      // (re)loading a trace restarts the replica; slabs are reused when big enough
  h->dirty = true;
  return GS_OK;
}

extern "C" int gs_load_trace(gs_handle h, int sim, int64_t n, const int32_t *arrive_tick, const int32_t *gpus,
                             const int32_t *gpu_per_task, const double *duration, const int64_t *mem_bytes,
                             const double *model_mb, const double *iterations, const int32_t *ps_count) {
  return load_common(h, sim, n, nullptr, arrive_tick, gpus, gpu_per_task, duration, mem_bytes, model_mb, iterations, ps_count);
}

// Same trace, already packed as 32-byte gs_jobin records (saves the column gather on the host).
extern "C" int gs_load_trace_packed(gs_handle h, int sim, int64_t n, const gs_jobin *jobs, const double *model_mb,
                                    const double *iterations) {
  static_assert(sizeof(gs_jobin) == sizeof(JobIn), "gs_jobin layout");
  if (n > 0 && !jobs) return fail(h, GS_ERR_ARG, "gs_load_trace_packed: NULL records");
  return load_common(h, sim, n, reinterpret_cast<const JobIn *>(jobs), nullptr, nullptr, nullptr, nullptr, nullptr,
                     model_mb, iterations, nullptr);
}

static SimLayout layout_sim(gs_handle h, const SimHost &s, int64_t rows_cap) {
  SimLayout L;
  const gs_cluster &c = s.cl;
  const int M = c.num_switch * c.num_node_p_switch;
  const size_t N = (size_t)(s.n > 0 ? s.n : 1);
  int W = 256; while (W < s.max_need + 1) W <<= 1;
  L.W = W;
  if (rows_cap <= 0) rows_cap = s.last_arrive + 2ll * s.max_need + 4096;
  L.rows_cap = rows_cap;
  L.qrows_cap = h->qrows_cap > 0 ? h->qrows_cap : rows_cap;
  const bool evd = s.pol.schedule != GS_SCHED_FIFO;        // event-driven policy: rows + extra scratch
  const bool net = c.enable_network_costs != 0;
  size_t total = 0;
  auto take = [&](size_t bytes) { const size_t o = total; total = align_up(total + bytes); return o; };
  if (evd) {
    L.o_fin = take(4 * N);
    L.o_rec = take(sizeof(gs_job_rec) * N);
    L.o_rows = take(sizeof(gs_tick_row) * (size_t)rows_cap);
    L.out_bytes = total;
    const size_t nql = (size_t)(s.pol.num_queue > 2 ? s.pol.num_queue : 2);
    L.o_pj = take(sizeof(PJob) * N); L.o_run = take(4 * N); L.o_qs = take(4 * N * nql); L.o_end = take(4 * N);
    L.o_tmp = take(4 * N); L.o_ci = take(4 * (size_t)M); L.o_ck = take(4 * (size_t)M); L.o_stale = take(4 * N);
  } else {
    L.o_ev = take(sizeof(gs_evrow) * (size_t)rows_cap);
    L.o_q = take(sizeof(gs_qrow) * (size_t)L.qrows_cap);
    L.o_ne = take(sizeof(gs_nodeev) * (size_t)(M + 2));          // at most one event per node + one per window
    L.o_rec2 = take(4 * N);
    if (net) L.o_dur2 = take(8 * N);
    L.o_fin = take(4 * N);
    L.o_spans = take((c.num_gpu_p_node > 32 ? sizeof(gs_span) : sizeof(gs_cspan)) * (size_t)s.span_cap);
    L.out_bytes = total;
    L.o_jst = take(sizeof(JobState2) * N);
    L.o_stack = take(8 * (N + 1));
    L.o_wh = take(4 * (size_t)W); L.o_wm = take(8 * (size_t)W);
    L.o_nb = take(8 * (size_t)M); L.o_nk = take(4 * (size_t)M);
  }
  L.total = total;
  return L;
}

static int bind_sim(gs_handle h, SimHost &s, const SimLayout &L, unsigned char *d) {
  const gs_cluster &c = s.cl;
  const int M = c.num_switch * c.num_node_p_switch;
  const int W = L.W;
  const int64_t rows_cap = L.rows_cap, qrows_cap = L.qrows_cap;
  const bool evd = s.pol.schedule != GS_SCHED_FIFO;
  const bool net = c.enable_network_costs != 0;
  const size_t o_fin = L.o_fin, o_rec = L.o_rec, o_rows = L.o_rows, o_rec2 = L.o_rec2, o_dur2 = L.o_dur2, o_jst = L.o_jst, o_stack = L.o_stack;
  const size_t o_wh = L.o_wh, o_wm = L.o_wm, o_spans = L.o_spans, o_ev = L.o_ev, o_q = L.o_q, o_nb = L.o_nb, o_nk = L.o_nk;
  const size_t o_pj = L.o_pj, o_run = L.o_run, o_qs = L.o_qs, o_end = L.o_end, o_tmp = L.o_tmp, o_ci = L.o_ci, o_ck = L.o_ck, o_stale = L.o_stale;
  s.layout = L;
  s.state_ptr = d;
  SimDev &D = s.dev;
  D.M = M; D.G = c.num_gpu_p_node;
  int kc = c.num_cpu_p_node / c.cpu_per_task, km = c.mem_p_node / c.mem_per_task;
  D.K = kc < km ? kc : km;
  D.netcost = net ? 1 : 0;
  D.n = (int)s.n; D.wheel_mask = W - 1; D.policy = s.pol.schedule;
  D.cap_bytes = (long long)c.gpu_mem_cap_mib << 20;
  D.fit_limit = D.cap_bytes - ((long long)500 << 20);      // cap - mem > 500 MiB  (device.py:75)
  D.bandwidth = c.bandwidth; D.latency = c.internode_latency;
  D.fin = (int *)(d + o_fin);
  D.rec = nullptr; D.rows = nullptr; D.jstart = nullptr; D.dur2 = nullptr; D.jst2 = nullptr; D.stack = nullptr;
  D.wheel_head = nullptr; D.wheel_mem = nullptr; D.spans = nullptr; D.evrows = nullptr; D.qrows = nullptr; D.nodeev = nullptr; D.nbusy = nullptr; D.nk = nullptr;
  if (evd) {
    D.rec = (gs_job_rec *)(d + o_rec); D.rows = (gs_tick_row *)(d + o_rows);
  } else {
    D.jstart = (int *)(d + o_rec2); D.dur2 = net ? (double *)(d + o_dur2) : nullptr;
    D.jst2 = (JobState2 *)(d + o_jst); D.stack = (int *)(d + o_stack);
    D.wheel_head = (int *)(d + o_wh); D.wheel_mem = (long long *)(d + o_wm);
    D.spans = (void *)(d + o_spans); D.evrows = (gs_evrow *)(d + o_ev); D.qrows = (gs_qrow *)(d + o_q); D.nodeev = (gs_nodeev *)(d + L.o_ne);
    D.nbusy = (unsigned long long *)(d + o_nb); D.nk = (int *)(d + o_nk);
  }
  D.span_cap = s.span_cap; D.rows_cap = rows_cap; D.qrows_cap = qrows_cap;
  s.rows_cap = rows_cap; s.qrows_cap = qrows_cap;
  if (evd) {
    D.pj = (PJob *)(d + o_pj); D.runnable = (int *)(d + o_run); D.queues = (int *)(d + o_qs);
    D.endj = (int *)(d + o_end); D.tmpl = (int *)(d + o_tmp); D.cidle = (int *)(d + o_ci); D.ckfree = (int *)(d + o_ck);
    D.stalej = (int *)(d + o_stale); D.stale_n = 0;
    D.num_queue = s.pol.num_queue > 0 ? s.pol.num_queue : 1;
    for (int q = 0; q < GS_MAX_QUEUES; ++q) { D.queue_limit[q] = s.pol.queue_limit[q]; D.qn[q] = 0; }
    D.gittins_delta = s.pol.gittins_delta; D.next_gittins_unit = s.pol.gittins_delta;
    D.git_n = s.pol.schedule == GS_SCHED_GITTINS ? s.pol.gittins_n : 0;
    D.git_data = (const double *)s.git_dev;
    D.git_index = s.git_dev ? (const double *)((unsigned char *)s.git_dev + 8 * (size_t)s.pol.gittins_n) : nullptr;
    D.git_direct_n = (s.git_dev && D.git_n > 0) ? s.git_direct_n : 0;
    D.git_direct = D.git_direct_n > 0 ? (const double *)((unsigned char *)s.git_dev + 16 * (size_t)s.pol.gittins_n) : nullptr;
    D.rn = 0; D.en = 0; D.end_time = 0x7fffffff; D.next_job_jump = 0x7fffffff;
  }
  D.comm_n = 0; D.comm_rank = 0; D.comm_cap = 0; D.comm_rk_in = nullptr; D.comm_flags = nullptr; D.comm_epoch = 0; D.comm_wait_cycles = 0;
  if (h->comm_n > 1 && s.pol.schedule == GS_SCHED_GITTINS) {
    if ((int64_t)s.n > h->comm_cap) return fail(h, GS_ERR_ARG, "gs_run: the trace has more jobs than gs_comm_prepare sized the exchange buffer for");
    // the event counter never restarts: a new run continues where the last one ended (all ranks execute the same
    // events, so their counters agree), which needs no reset of the flag words and no extra synchronisation
    D.comm_epoch = h->comm_epoch; h->comm_epoch0 = h->comm_epoch;
    D.comm_n = h->comm_n; D.comm_rank = h->comm_rank; D.comm_cap = h->comm_cap; D.comm_min_runnable = h->comm_min_runnable;
    D.comm_flags = (unsigned long long *)h->comm_buf;
    D.comm_rk_in = (double *)((unsigned char *)h->comm_buf + 256);
    for (int q = 0; q < h->comm_n; ++q) {
      D.comm_peer_flags[q] = (unsigned long long *)h->comm_peer[q];
      D.comm_peer_rk[q] = (double *)((unsigned char *)h->comm_peer[q] + 256);
    }
  }
  D.delta = D.p = D.top = D.running = D.finished = D.ever = D.busy_gpus = D.done = D.status = 0;
  D.blocked = D.nev = D.nq = D.nne = 0;
  D.mem_busy = D.sum_arr = D.span_used = D.events = D.evals = D.started = D.ticks = D.row_first = 0;
  D.need_init = 1;
  s.prepared = true;
  return GS_OK;
}

extern "C" int gs_run(gs_handle h, int64_t max_ticks, int64_t rows_cap) {
  if (!h) return GS_ERR_ARG;
  CU(cudaSetDevice(h->device));
  int maxM = 1;
  bool none_prepared = true;
  for (auto &s : h->sims) {
    if (!s.loaded) return fail(h, GS_ERR_STATE, "gs_run: every replica needs gs_config_sim + gs_load_trace");
    none_prepared &= !s.prepared;
    int M = s.cl.num_switch * s.cl.num_node_p_switch;
    if (M > maxM) maxM = M;
  }
  if (none_prepared) {
    // a fresh start of every replica (the common case): the slabs go side by side into one arena with one stride
    std::vector<SimLayout> Ls((size_t)h->nsims);
    size_t stride = 0;
    for (int i = 0; i < h->nsims; ++i) { Ls[(size_t)i] = layout_sim(h, h->sims[(size_t)i], rows_cap); stride = std::max(stride, Ls[(size_t)i].total); }
    stride = align_up(stride, 512);
    const size_t need = stride * (size_t)h->nsims;
    if (h->arena_bytes < need) {
      CU(cudaStreamSynchronize(h->stream));
      if (h->arena) cudaFree(h->arena);
      h->arena = nullptr; h->arena_bytes = 0;
      CU(cudaMalloc(&h->arena, need));
      h->arena_bytes = need;
    }
    h->arena_stride = stride;
    for (int i = 0; i < h->nsims; ++i) {
      int rc = bind_sim(h, h->sims[(size_t)i], Ls[(size_t)i], (unsigned char *)h->arena + stride * (size_t)i);
      if (rc) return rc;
    }
    h->dirty = true;
  } else {
    for (auto &s : h->sims)
      if (!s.prepared) {   // one replica restarts while others keep running: it gets (or keeps) its own allocation
        const SimLayout L = layout_sim(h, s, rows_cap);
        if (s.state_slab && s.state_bytes < L.total) { CU(cudaStreamSynchronize(h->stream)); cudaFree(s.state_slab); s.state_slab = nullptr; }
        if (!s.state_slab) { CU(cudaMalloc(&s.state_slab, L.total)); s.state_bytes = L.total; }
        int rc = bind_sim(h, s, L, (unsigned char *)s.state_slab);
        if (rc) return rc;
        h->dirty = true;
      }
  }
  bool fifo_kind[4] = {false, false, false, false};   // [network costs][more than 32 GPUs per node]
  bool any_evd = false, evd_init = false;
  for (auto &s : h->sims) {
    if (s.pol.schedule == GS_SCHED_FIFO) fifo_kind[(s.cl.enable_network_costs ? 2 : 0) + (s.cl.num_gpu_p_node > 32 ? 1 : 0)] = true;
    else { any_evd = true; evd_init |= s.dev.need_init != 0; }
  }
  if (h->dirty) {
    for (int i = 0; i < h->nsims; ++i) h->h_back[(size_t)i] = h->sims[(size_t)i].dev;
    CU(cudaMemcpyAsync(h->d_sims, h->h_back.data(), sizeof(SimDev) * (size_t)h->nsims, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->dirty = false;
  }
  CU(cudaEventRecord(h->e0, h->stream));
  if (any_evd) {
    if (evd_init) {      // state reset is part of the timed engine work
      gs_init_kernel<<<dim3(16, (unsigned)h->nsims), 256, 0, h->stream>>>(h->d_sims, h->nsims);
      CU(cudaGetLastError());
      h->launches += 1;
    }
    // engine mode 2: one thread per replica (first version); default: one warp per replica
    const int thread_map = h->engine_mode == 2 ? 1 : 0;
    if (thread_map) {
      gs_policy_kernel<<<(unsigned)((h->nsims + 31) / 32), 32, 0, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, 1);
      CU(cudaGetLastError());
      h->launches += 1;
    } else {
      gs_dlas_warp_kernel<<<(unsigned)h->nsims, 32, 0, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks);
      CU(cudaGetLastError());
      const size_t pol_smem = 8 * (size_t)maxM;
      if (pol_smem > 48 * 1024)
        CU(cudaFuncSetAttribute(gs_sortpol_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pol_smem));
      gs_sortpol_warp_kernel<<<(unsigned)h->nsims, 32, pol_smem, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks);
      CU(cudaGetLastError());
      h->launches += 2;
    }
  }
  if (fifo_kind[0] || fifo_kind[1] || fifo_kind[2] || fifo_kind[3]) {
    const int stride = (int)align_up((size_t)maxM * 12 + 8 + SCACHE * 8, 16);
    if (stride > 200 * 1024) return fail(h, GS_ERR_ARG, "gs_run: node table does not fit shared memory (M too large)");
    // one instantiation per (network costs, mask width); each skips the replicas of the other kinds
#define GS_LAUNCH_TICK2(NET_, G64_)                                                                                              \
    do {                                                                                                                         \
      if (stride > 48 * 1024) CU(cudaFuncSetAttribute(gs_tick2_kernel<NET_, G64_>, cudaFuncAttributeMaxDynamicSharedMemorySize, stride)); \
      gs_tick2_kernel<NET_, G64_><<<(unsigned)h->nsims, 32, (size_t)stride, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, stride); \
      CU(cudaGetLastError());                                                                                                    \
      h->launches += 1;                                                                                                          \
    } while (0)
    if (fifo_kind[0]) GS_LAUNCH_TICK2(false, false);
    if (fifo_kind[1]) GS_LAUNCH_TICK2(false, true);
    if (fifo_kind[2]) GS_LAUNCH_TICK2(true, false);
    if (fifo_kind[3]) GS_LAUNCH_TICK2(true, true);
#undef GS_LAUNCH_TICK2
  }
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaMemcpyAsync(h->h_back.data(), h->d_sims, sizeof(SimDev) * (size_t)h->nsims, cudaMemcpyDeviceToHost, h->stream));
  CU(wait_stream(h));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->kernel_ms += ms;
  if (h->comm_n > 1 && h->h_back[0].comm_n > 1) h->comm_epoch = h->h_back[0].comm_epoch;
  int worst = 0;
  for (int i = 0; i < h->nsims; ++i) {
    h->h_back[(size_t)i].need_init = 0;
    h->sims[(size_t)i].dev = h->h_back[(size_t)i];
    if (h->h_back[(size_t)i].status != 0 && worst == 0) worst = h->h_back[(size_t)i].status;
  }
  if (worst != 0) return fail(h, worst, "gs_run: a replica stopped with an in-kernel error (see gs_stats.status)");
  return GS_OK;
}

extern "C" int gs_stats(gs_handle h, int sim, gs_run_stats *out) {
  if (!h || !out) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_stats: sim index out of range");
  const SimDev &D = h->sims[(size_t)sim].dev;
  memset(out, 0, sizeof(*out));
  out->ticks = D.ticks; out->events = D.events; out->finished = D.finished; out->started = D.started;
  out->placement_evals = D.evals; out->done = D.done; out->status = D.status;
  out->kernel_ms = h->kernel_ms; out->h2d_ms = h->h2d_ms; out->d2h_ms = h->d2h_ms;
  return GS_OK;
}

extern "C" int gs_window(gs_handle h, int sim, gs_window_info *out) {
  if (!h || !out) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_window: sim index out of range");
  const SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_window: nothing has run yet");
  const SimDev &D = s.dev;
  out->row_first = D.row_first; out->ticks = D.ticks;
  const bool fifo = s.pol.schedule == GS_SCHED_FIFO;
  out->ev_rows = fifo ? D.nev : 0; out->q_rows = fifo ? D.nq : 0; out->node_events = fifo ? D.nne : 0;
  out->spans_used = D.span_used; out->admitted = D.p; out->finished = D.finished; out->n = s.n;
  return GS_OK;
}

extern "C" int gs_comm_prepare(gs_handle h, int64_t max_jobs, gs_comm_handle *out) {
  if (!h || !out || max_jobs < 1) return fail(h, GS_ERR_ARG, "gs_comm_prepare: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(gs_comm_handle), "IPC handle size");
  if (h->nsims != 1) return fail(h, GS_ERR_ARG, "gs_comm_prepare: a sharded simulation needs a handle with exactly one replica");
  if (h->comm_n > 0) return fail(h, GS_ERR_STATE, "gs_comm_prepare: already initialised");
  CU(cudaSetDevice(h->device));
  if (h->comm_buf) { cudaFree(h->comm_buf); h->comm_buf = nullptr; }
  const size_t bytes = 256 + 2 * 8 * (size_t)max_jobs;
  CU(cudaMalloc(&h->comm_buf, bytes));
  CU(cudaMemset(h->comm_buf, 0, bytes));
  h->comm_cap = max_jobs;
  cudaIpcMemHandle_t ih;
  CU(cudaIpcGetMemHandle(&ih, h->comm_buf));
  memset(out, 0, sizeof(*out));
  memcpy(out->bytes, &ih, sizeof(ih));
  return GS_OK;
}

extern "C" int gs_comm_init(gs_handle h, int rank, int nranks, const gs_comm_handle *all) {
  if (!h || !all || nranks < 1 || nranks > GS_MAX_RANKS || rank < 0 || rank >= nranks)
    return fail(h, GS_ERR_ARG, "gs_comm_init: bad arguments (1 <= nranks <= 8)");
  if (!h->comm_buf) return fail(h, GS_ERR_STATE, "gs_comm_init: call gs_comm_prepare first");
  if (h->comm_n > 0) return fail(h, GS_ERR_STATE, "gs_comm_init: already initialised");
  CU(cudaSetDevice(h->device));
  for (int q = 0; q < nranks; ++q) {
    if (q == rank) { h->comm_peer[q] = h->comm_buf; h->comm_opened[q] = false; continue; }
    cudaIpcMemHandle_t ih;
    memcpy(&ih, all[q].bytes, sizeof(ih));
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int r = 0; r < q; ++r) if (h->comm_opened[r]) { cudaIpcCloseMemHandle(h->comm_peer[r]); h->comm_opened[r] = false; h->comm_peer[r] = nullptr; }
      return fail(h, GS_ERR_COMM, std::string("gs_comm_init: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    h->comm_peer[q] = ptr; h->comm_opened[q] = true;
  }
  h->comm_rank = rank; h->comm_n = nranks;
  for (auto &s : h->sims) s.prepared = false;
  h->dirty = true;
  return GS_OK;
}

// Events whose runnable list has at most `k` entries are evaluated by every rank itself (identical state, identical
// result, nothing to send); longer lists are split and exchanged.  k = 0: exchange on every event.  Must be the same on
// every rank; takes effect for replicas prepared afterwards.
extern "C" int gs_comm_set_min_runnable(gs_handle h, int k) {
  if (!h) return GS_ERR_ARG;
  if (k < 0) return fail(h, GS_ERR_ARG, "gs_comm_set_min_runnable: must be >= 0");
  h->comm_min_runnable = k;
  for (auto &s : h->sims) s.prepared = false;
  h->dirty = true;
  return GS_OK;
}

extern "C" int gs_comm_stats(gs_handle h, int64_t *exchanges, double *mean_us) {
  if (!h) return GS_ERR_ARG;
  const SimDev &D = h->sims[0].dev;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
  const unsigned long long done = D.comm_n > 1 ? D.comm_epoch - h->comm_epoch0 : 0ull;
  if (exchanges) *exchanges = (int64_t)done;
  if (mean_us) *mean_us = (done > 0 && khz > 0) ? (double)D.comm_wait_cycles / (double)done / ((double)khz / 1000.0) : 0.0;
  return GS_OK;
}

extern "C" int gs_sync(gs_handle h) {
  if (!h) return GS_ERR_ARG;
  CU(cudaSetDevice(h->device));
  CU(wait_stream(h));
  return GS_OK;
}

extern "C" int gs_set_async(gs_handle h, int on) {
  if (!h) return GS_ERR_ARG;
  h->async = on != 0;
  return GS_OK;
}

extern "C" int gs_set_queue_rows_cap(gs_handle h, int64_t qrows_cap) {
  if (!h) return GS_ERR_ARG;
  if (qrows_cap < 0) return fail(h, GS_ERR_ARG, "gs_set_queue_rows_cap: must be >= 0");
  h->qrows_cap = qrows_cap;
  return GS_OK;
}

extern "C" int gs_fetch_compact(gs_handle h, int sim, gs_evrow *ev_out, gs_qrow *q_out, gs_nodeev *nodeev_out, gs_job_start *jobs_out,
                                double *duration_out, int32_t *finish_order_out, void *spans_out) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_compact: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_compact: nothing has run yet");
  if (s.pol.schedule != GS_SCHED_FIFO) return fail(h, GS_ERR_ARG, "gs_fetch_compact: the compact records are the fifo engine's output");
  static_assert(sizeof(gs_evrow) == 24 && sizeof(gs_qrow) == 24 && sizeof(gs_nodeev) == 8 && sizeof(gs_job_start) == 4 && sizeof(gs_cspan) == 8,
                "compact record layout");
  const size_t span_bytes = s.cl.num_gpu_p_node > 32 ? sizeof(gs_span) : sizeof(gs_cspan);
  const SimDev &D = s.dev;
  CU(cudaSetDevice(h->device));
  if (ev_out && D.nev > 0) CU(cudaMemcpyAsync(ev_out, D.evrows, sizeof(gs_evrow) * (size_t)D.nev, cudaMemcpyDeviceToHost, h->stream));
  if (q_out && D.nq > 0) CU(cudaMemcpyAsync(q_out, D.qrows, sizeof(gs_qrow) * (size_t)D.nq, cudaMemcpyDeviceToHost, h->stream));
  if (nodeev_out && D.nne > 0) CU(cudaMemcpyAsync(nodeev_out, D.nodeev, sizeof(gs_nodeev) * (size_t)D.nne, cudaMemcpyDeviceToHost, h->stream));
  if (jobs_out && s.n > 0) CU(cudaMemcpyAsync(jobs_out, D.jstart, 4 * (size_t)s.n, cudaMemcpyDeviceToHost, h->stream));
  if (duration_out && D.dur2 && s.n > 0) CU(cudaMemcpyAsync(duration_out, D.dur2, 8 * (size_t)s.n, cudaMemcpyDeviceToHost, h->stream));
  if (finish_order_out && D.finished > 0) CU(cudaMemcpyAsync(finish_order_out, D.fin, 4 * (size_t)D.finished, cudaMemcpyDeviceToHost, h->stream));
  if (spans_out && D.span_used > 0) CU(cudaMemcpyAsync(spans_out, D.spans, span_bytes * (size_t)D.span_used, cudaMemcpyDeviceToHost, h->stream));
  return GS_OK;
}

// Every replica of the handle at once, from ONE host block (record i*pitch_bytes is the trace of replica i): the traces
// go into a device arena with one stride and travel as a single strided copy.  With gs_set_async and a page-locked
// block nothing is staged and the call returns before the copy completes (keep the block until gs_run / gs_sync).
extern "C" int gs_load_traces_packed(gs_handle h, const gs_jobin *jobs, size_t pitch_bytes, const int64_t *n_each) {
  if (!h) return GS_ERR_ARG;
  if (!jobs || !n_each || pitch_bytes % sizeof(JobIn) != 0) return fail(h, GS_ERR_ARG, "gs_load_traces_packed: bad arguments (pitch must be a multiple of 32)");
  CU(cudaSetDevice(h->device));
  int64_t nmax = 1;
  for (int i = 0; i < h->nsims; ++i) {
    const SimHost &s = h->sims[(size_t)i];
    if (!s.configured) return fail(h, GS_ERR_STATE, "gs_load_traces_packed: call gs_config_sim for every replica first");
    if (s.cl.enable_network_costs) return fail(h, GS_ERR_ARG, "gs_load_traces_packed: traces with network columns go through gs_load_trace");
    if (n_each[i] < 0 || n_each[i] >= (1ll << 31) - 64 || (size_t)n_each[i] * sizeof(JobIn) > pitch_bytes)
      return fail(h, GS_ERR_ARG, "gs_load_traces_packed: a trace does not fit the pitch");
    nmax = std::max(nmax, n_each[i]);
  }
  // validate first (read only); nothing changes if a trace is rejected
  std::vector<int64_t> span_cap((size_t)h->nsims); std::vector<double> max_need((size_t)h->nsims);
  for (int i = 0; i < h->nsims; ++i) {
    const JobIn *ji = reinterpret_cast<const JobIn *>(reinterpret_cast<const unsigned char *>(jobs) + pitch_bytes * (size_t)i);
    int rc = scan_trace(h, h->sims[(size_t)i], n_each[i], ji, false, nullptr, nullptr, &span_cap[(size_t)i], &max_need[(size_t)i]);
    if (rc) return rc;
    if (max_need[(size_t)i] > (double)(1 << 26)) return fail(h, GS_ERR_ARG, "gs_load_trace: job duration exceeds 2^26 ticks");
  }
  const size_t stride = align_up(sizeof(JobIn) * (size_t)nmax, 512);
  const size_t need = stride * (size_t)h->nsims;
  CU(wait_stream(h));                                   // earlier work may still read the old traces
  if (h->tarena_bytes < need) {
    if (h->tarena) cudaFree(h->tarena);
    h->tarena = nullptr; h->tarena_bytes = 0;
    CU(cudaMalloc(&h->tarena, need));
    h->tarena_bytes = need;
  }
  h->tarena_stride = stride;
  bool direct = false;
  if (h->async) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, jobs) == cudaSuccess && at.type == cudaMemoryTypeHost) direct = true;
    else (void)cudaGetLastError();
  }
  const void *src = jobs;
  if (!direct) {
    int rc = ensure_stage(h, pitch_bytes * (size_t)h->nsims);
    if (rc) return rc;
    memcpy(h->h_stage, jobs, pitch_bytes * (size_t)h->nsims);
    src = h->h_stage;
  }
  CU(cudaEventRecord(h->e0, h->stream));
  CU(cudaMemcpy2DAsync(h->tarena, stride, src, pitch_bytes, sizeof(JobIn) * (size_t)nmax, (size_t)h->nsims, cudaMemcpyHostToDevice, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  if (!direct) {
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
    h->h2d_ms += ms;
  }
  for (int i = 0; i < h->nsims; ++i) {
    SimHost &s = h->sims[(size_t)i];
    const JobIn *ji = reinterpret_cast<const JobIn *>(reinterpret_cast<const unsigned char *>(jobs) + pitch_bytes * (size_t)i);
    SimDev &D = s.dev;
    memset(&D, 0, sizeof(D));
    D.jobs = (const JobIn *)((unsigned char *)h->tarena + stride * (size_t)i);
    int64_t sc = span_cap[(size_t)i];
    if (h->span_budget > 0) { const int64_t lim = (int64_t)(h->span_budget * (double)n_each[i]) + 4096; if (sc > lim) sc = lim; }
    s.n = n_each[i]; s.span_cap = sc > 0 ? sc : 1;
    s.max_need = (int)max_need[(size_t)i] + 2;
    s.last_arrive = n_each[i] > 0 ? ji[n_each[i] - 1].arrive : 0;
    s.loaded = true; s.prepared = false; s.trace_in_arena = true;
  }
  h->dirty = true;
  return GS_OK;
}

// Where the results of a replica lie inside its result block, and how the blocks of the handle are spaced.
extern "C" int gs_result_layout(gs_handle h, int sim, gs_result_layout_t *out) {
  if (!h || !out) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_result_layout: sim index out of range");
  const SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_result_layout: nothing has run yet");
  if (s.pol.schedule != GS_SCHED_FIFO) return fail(h, GS_ERR_ARG, "gs_result_layout: the compact records are the fifo engine's output");
  const SimLayout &L = s.layout;
  memset(out, 0, sizeof(*out));
  out->block_bytes = (int64_t)L.out_bytes;
  out->off_ev = (int64_t)L.o_ev; out->off_q = (int64_t)L.o_q; out->off_nodeev = (int64_t)L.o_ne; out->off_jobs = (int64_t)L.o_rec2;
  out->cap_nodeev = (int64_t)s.cl.num_switch * s.cl.num_node_p_switch + 2;
  out->off_duration = s.cl.enable_network_costs ? (int64_t)L.o_dur2 : -1;
  out->off_finish_order = (int64_t)L.o_fin; out->off_spans = (int64_t)L.o_spans;
  out->cap_ev = L.rows_cap; out->cap_q = L.qrows_cap; out->cap_spans = s.span_cap; out->n = s.n;
  out->span_bytes = s.cl.num_gpu_p_node > 32 ? (int64_t)sizeof(gs_span) : (int64_t)sizeof(gs_cspan);
  return GS_OK;
}

// The result blocks of replicas [first, first + count) in one strided copy (asynchronous: gs_sync waits).  Row i of
// `out` (out_pitch bytes apart) receives the first block_bytes bytes of replica first + i's block.
extern "C" int gs_fetch_results(gs_handle h, int first, int count, void *out, size_t out_pitch) {
  if (!h) return GS_ERR_ARG;
  if (first < 0 || count < 0 || first + count > h->nsims || (count > 0 && !out)) return fail(h, GS_ERR_ARG, "gs_fetch_results: bad arguments");
  if (count == 0) return GS_OK;
  CU(cudaSetDevice(h->device));
  size_t width = 0;
  bool in_arena = true;
  for (int i = first; i < first + count; ++i) {
    const SimHost &s = h->sims[(size_t)i];
    if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_results: nothing has run yet");
    if (s.pol.schedule != GS_SCHED_FIFO) return fail(h, GS_ERR_ARG, "gs_fetch_results: the compact records are the fifo engine's output");
    width = std::max(width, s.layout.out_bytes);
    in_arena &= s.state_ptr == (unsigned char *)h->arena + h->arena_stride * (size_t)i;
  }
  if (width > out_pitch) return fail(h, GS_ERR_CAPACITY, "gs_fetch_results: out_pitch is smaller than a result block");
  if (in_arena) {
    CU(cudaMemcpy2DAsync(out, out_pitch, (unsigned char *)h->arena + h->arena_stride * (size_t)first, h->arena_stride, width, (size_t)count,
                         cudaMemcpyDeviceToHost, h->stream));
  } else {
    for (int i = first; i < first + count; ++i)
      CU(cudaMemcpyAsync((unsigned char *)out + out_pitch * (size_t)(i - first), h->sims[(size_t)i].state_ptr, h->sims[(size_t)i].layout.out_bytes,
                         cudaMemcpyDeviceToHost, h->stream));
  }
  return GS_OK;
}

static int ensure_scratch(gs_handle h, size_t bytes) {
  if (h->d_scratch_bytes >= bytes) return GS_OK;
  if (h->d_scratch) { cudaStreamSynchronize(h->stream); cudaFree(h->d_scratch); }
  h->d_scratch = nullptr; h->d_scratch_bytes = 0;
  CU(cudaMalloc(&h->d_scratch, bytes));
  h->d_scratch_bytes = bytes;
  return GS_OK;
}

static int timed_d2h(gs_handle h, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return GS_OK;
  CU(cudaEventRecord(h->e0, h->stream));
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->d2h_ms += ms;
  return GS_OK;
}

extern "C" int gs_fetch_rows(gs_handle h, int sim, int64_t first, int64_t count, gs_tick_row *rows_out) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_rows: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_rows: nothing has run yet");
  const SimDev &D = s.dev;
  if (count < 0 || first < D.row_first || first + count > D.ticks)
    return fail(h, GS_ERR_ARG, "gs_fetch_rows: range is outside the rows of the last gs_run window");
  if (count == 0) return GS_OK;
  if (!rows_out) return fail(h, GS_ERR_ARG, "gs_fetch_rows: NULL output");
  CU(cudaSetDevice(h->device));
  if (s.pol.schedule != GS_SCHED_FIFO)
    return timed_d2h(h, rows_out, D.rows + (first - D.row_first), sizeof(gs_tick_row) * (size_t)count);
  // fifo: rebuild the rows of the whole window from its records on the device, copy the requested slice
  const int64_t wrows = D.ticks - D.row_first;
  int rc = ensure_scratch(h, sizeof(gs_tick_row) * (size_t)wrows);
  if (rc) return rc;
  if (D.nev > 0) {
    gs_expand_rows_kernel<<<(unsigned)((D.nev + 127) / 128), 128, 0, h->stream>>>(h->d_sims, sim, D.M, D.G, (gs_tick_row *)h->d_scratch);
    CU(cudaGetLastError());
    h->launches += 1;
  }
  return timed_d2h(h, rows_out, (gs_tick_row *)h->d_scratch + (first - D.row_first), sizeof(gs_tick_row) * (size_t)count);
}

extern "C" int gs_fetch_jobs(gs_handle h, int sim, gs_job_rec *jobs_out, int32_t *finish_order_out) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_jobs: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_jobs: nothing has run yet");
  CU(cudaSetDevice(h->device));
  const SimDev &D = s.dev;
  const gs_job_rec *src = D.rec;
  if (jobs_out && s.n > 0 && s.pol.schedule == GS_SCHED_FIFO) {
    int rc = ensure_scratch(h, sizeof(gs_job_rec) * (size_t)s.n);
    if (rc) return rc;
    gs_expand_jobs_kernel<<<(unsigned)((s.n + 255) / 256), 256, 0, h->stream>>>(h->d_sims, sim, (gs_job_rec *)h->d_scratch);
    CU(cudaGetLastError());
    h->launches += 1;
    src = (const gs_job_rec *)h->d_scratch;
  }
  CU(cudaEventRecord(h->e0, h->stream));
  if (jobs_out && s.n > 0)
    CU(cudaMemcpyAsync(jobs_out, src, sizeof(gs_job_rec) * (size_t)s.n, cudaMemcpyDeviceToHost, h->stream));
  if (finish_order_out && D.finished > 0)
    CU(cudaMemcpyAsync(finish_order_out, D.fin, 4 * (size_t)D.finished, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->d2h_ms += ms;
  return GS_OK;
}

// Spans are pooled in START order while the simulation runs (bit 31 of ntasks marks the first span of a
// job); this entry point hands them out grouped by job (CSR).  Start ticks are unique (one start per
// tick), so the k-th marked span belongs to the job with the k-th smallest start tick.
extern "C" int gs_fetch_spans(gs_handle h, int sim, int64_t *span_off_out, gs_span *spans_out, int64_t spans_cap,
                              int64_t *spans_used) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_spans: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_spans: nothing has run yet");
  CU(cudaSetDevice(h->device));
  const SimDev &D = s.dev;
  const int64_t used = s.pol.schedule == GS_SCHED_FIFO ? D.span_used : 0;
  if (spans_used) *spans_used = used;
  if (!spans_out && !span_off_out) return GS_OK;
  if (spans_out && spans_cap < used) return fail(h, GS_ERR_CAPACITY, "gs_fetch_spans: spans_out too small");
  const int64_t n = s.n;
  if (used == 0) {
    if (span_off_out) for (int64_t j = 0; j <= n; ++j) span_off_out[j] = 0;
    return GS_OK;
  }
  std::vector<gs_span> pool((size_t)used);
  std::vector<int32_t> r2((size_t)n);
  int rc;
  if (s.cl.num_gpu_p_node > 32) {
    rc = timed_d2h(h, pool.data(), D.spans, sizeof(gs_span) * (size_t)used);
    if (rc) return rc;
  } else {                  // compact 8-byte records on the device: widen them here
    std::vector<gs_cspan> cp((size_t)used);
    rc = timed_d2h(h, cp.data(), D.spans, sizeof(gs_cspan) * (size_t)used);
    if (rc) return rc;
    for (int64_t i = 0; i < used; ++i) {
      const uint32_t w = cp[(size_t)i].where;
      gs_span sp; sp.node = (int32_t)GS_CSPAN_NODE(w); sp.devmask = cp[(size_t)i].devmask;
      sp.ntasks = (int32_t)(GS_CSPAN_NTASKS(w) | (GS_CSPAN_FIRST(w) ? GS_SPAN_FIRST : 0u));
      pool[(size_t)i] = sp;
    }
  }
  rc = timed_d2h(h, r2.data(), D.jstart, 4 * (size_t)n);
  if (rc) return rc;
  // jobs in start order: bucket by start tick (unique, < ticks)
  std::vector<int32_t> by_tick((size_t)D.ticks + 1, -1);
  const int64_t admitted = D.p;
  for (int64_t j = 0; j < admitted; ++j) if (r2[(size_t)j] >= 0 && r2[(size_t)j] <= D.ticks) by_tick[(size_t)r2[(size_t)j]] = (int32_t)j;
  std::vector<int32_t> order; order.reserve((size_t)n);
  for (int64_t t = 0; t <= D.ticks; ++t) if (by_tick[(size_t)t] >= 0) order.push_back(by_tick[(size_t)t]);
  std::vector<int64_t> first((size_t)n, 0), cnt((size_t)n, 0);
  int64_t k = -1;
  for (int64_t i = 0; i < used; ++i) {
    if ((uint32_t)pool[(size_t)i].ntasks & GS_SPAN_FIRST) { ++k; if (k < (int64_t)order.size()) first[(size_t)order[(size_t)k]] = i; }
    if (k >= 0 && k < (int64_t)order.size()) cnt[(size_t)order[(size_t)k]] += 1;
  }
  if (k + 1 != (int64_t)order.size()) return fail(h, GS_ERR_STATE, "gs_fetch_spans: span pool and start records disagree");
  int64_t run = 0;
  for (int64_t j = 0; j < n; ++j) {
    if (span_off_out) span_off_out[j] = run;
    if (spans_out)
      for (int64_t i = 0; i < cnt[(size_t)j]; ++i) {
        gs_span sp = pool[(size_t)(first[(size_t)j] + i)];
        sp.ntasks = (int32_t)((uint32_t)sp.ntasks & ~GS_SPAN_FIRST);
        spans_out[run + i] = sp;
      }
    run += cnt[(size_t)j];
  }
  if (span_off_out) span_off_out[n] = run;
  return GS_OK;
}

// Everything a caller needs from one finished (or paused) replica in ONE call, in the row / record
// formats of the reference's LogInfo and job.csv (the compact, asynchronous route is gs_fetch_compact).
extern "C" int gs_fetch_all(gs_handle h, int sim, int64_t first, int64_t count, gs_tick_row *rows_out,
                            gs_job_rec *jobs_out, int32_t *finish_order_out, int64_t *span_off_out,
                            gs_span *spans_out, int64_t spans_cap, int64_t *spans_used) {
  if (!h) return GS_ERR_ARG;
  int rc = GS_OK;
  if (rows_out) rc = gs_fetch_rows(h, sim, first, count, rows_out);
  if (rc == GS_OK && (jobs_out || finish_order_out)) rc = gs_fetch_jobs(h, sim, jobs_out, finish_order_out);
  if (rc == GS_OK) rc = gs_fetch_spans(h, sim, span_off_out, spans_out, spans_cap, spans_used);
  return rc;
}

extern "C" int gs_place_batch(gs_handle h, const gs_cluster *cluster, const gs_node *nodes, int32_t m,
                              const gs_jobreq *jobs, int64_t b, int32_t *first_node, int32_t *nodes_used,
                              const int64_t *task_off, int32_t *task_node, double *kernel_ms) {
  if (!h) return GS_ERR_ARG;
  int rc = check_cluster(h, cluster);
  if (rc) return rc;
  if (!nodes || m <= 0 || m > 16384 || b < 0 || (b > 0 && (!jobs || !first_node)))
    return fail(h, GS_ERR_ARG, "gs_place_batch: bad arguments (1 <= m <= 16384)");
  if ((task_node != nullptr) != (task_off != nullptr))
    return fail(h, GS_ERR_ARG, "gs_place_batch: task_off and task_node go together");
  for (int64_t j = 0; j < b; ++j)
    if (jobs[j].gpu_per_task <= 0 || jobs[j].gpus < jobs[j].gpu_per_task || jobs[j].gpus % jobs[j].gpu_per_task)
      return fail(h, GS_ERR_ARG, "gs_place_batch: gpus must be a positive multiple of gpu_per_task");
  if (b == 0) return GS_OK;
  CU(cudaSetDevice(h->device));
  const int64_t ntask = task_off ? task_off[b] : 0;
  size_t o_nodes = 0, o_jobs = align_up(o_nodes + 16 * (size_t)m), o_first = align_up(o_jobs + 16 * (size_t)b);
  size_t o_used = align_up(o_first + 4 * (size_t)b), o_toff = align_up(o_used + 4 * (size_t)b);
  size_t o_tn = align_up(o_toff + 8 * (size_t)(b + 1)), total = align_up(o_tn + 4 * (size_t)(ntask > 0 ? ntask : 1));
  rc = ensure_scratch(h, total);
  if (rc) return rc;
  unsigned char *d = (unsigned char *)h->d_scratch;
  CU(cudaEventRecord(h->e0, h->stream));
  CU(cudaMemcpyAsync(d + o_nodes, nodes, 16 * (size_t)m, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d + o_jobs, jobs, 16 * (size_t)b, cudaMemcpyHostToDevice, h->stream));
  if (task_off) CU(cudaMemcpyAsync(d + o_toff, task_off, 8 * (size_t)(b + 1), cudaMemcpyHostToDevice, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1); h->h2d_ms += ms;
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device);
  long long want = (b + 255) / 256;
  int grid = (int)(want < (long long)dev_sms * 8 ? want : (long long)dev_sms * 8);
  const long long fit_limit = ((long long)cluster->gpu_mem_cap_mib << 20) - ((long long)500 << 20);
  CU(cudaEventRecord(h->e0, h->stream));
  const size_t place_smem = 12 * (size_t)m + 4 * (GS_MAX_GPUS_PER_NODE + 1);
  if (place_smem > 48 * 1024)
    CU(cudaFuncSetAttribute(gs_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
  gs_place_kernel<<<grid, 256, place_smem, h->stream>>>(
      (const uint4 *)(d + o_nodes), m, cluster->num_gpu_p_node, cluster->num_cpu_p_node, cluster->mem_p_node,
      cluster->cpu_per_task, cluster->mem_per_task, fit_limit, (const uint4 *)(d + o_jobs), (long long)b,
      (int *)(d + o_first), (int *)(d + o_used), task_off ? (const long long *)(d + o_toff) : nullptr,
      task_node ? (int *)(d + o_tn) : nullptr);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->kernel_ms += ms;
  if (kernel_ms) *kernel_ms = ms;
  rc = timed_d2h(h, first_node, d + o_first, 4 * (size_t)b);
  if (rc == GS_OK && nodes_used) rc = timed_d2h(h, nodes_used, d + o_used, 4 * (size_t)b);
  if (rc == GS_OK && task_node && ntask > 0) rc = timed_d2h(h, task_node, d + o_tn, 4 * (size_t)ntask);
  return rc;
}

extern "C" int gs_net_cost(gs_handle h, const gs_cluster *cluster, int64_t b, const int64_t *task_off,
                           const int32_t *task_node, const uint8_t *is_ps, const int32_t *ps_count,
                           const double *model_mb, const double *iterations, double *extra_out) {
  if (!h) return GS_ERR_ARG;
  if (!cluster || b < 0 || (b > 0 && (!task_off || !task_node || !ps_count || !model_mb || !iterations || !extra_out)))
    return fail(h, GS_ERR_ARG, "gs_net_cost: bad arguments");
  if (b == 0) return GS_OK;
  CU(cudaSetDevice(h->device));
  const int64_t nt = task_off[b];
  size_t o_off = 0, o_tn = align_up(o_off + 8 * (size_t)(b + 1)), o_ps = align_up(o_tn + 4 * (size_t)(nt > 0 ? nt : 1));
  size_t o_cnt = align_up(o_ps + (size_t)(nt > 0 ? nt : 1)), o_mm = align_up(o_cnt + 4 * (size_t)b);
  size_t o_it = align_up(o_mm + 8 * (size_t)b), o_out = align_up(o_it + 8 * (size_t)b), total = align_up(o_out + 8 * (size_t)b);
  int rc = ensure_scratch(h, total);
  if (rc) return rc;
  unsigned char *d = (unsigned char *)h->d_scratch;
  CU(cudaMemcpyAsync(d + o_off, task_off, 8 * (size_t)(b + 1), cudaMemcpyHostToDevice, h->stream));
  if (nt > 0) CU(cudaMemcpyAsync(d + o_tn, task_node, 4 * (size_t)nt, cudaMemcpyHostToDevice, h->stream));
  if (is_ps && nt > 0) CU(cudaMemcpyAsync(d + o_ps, is_ps, (size_t)nt, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d + o_cnt, ps_count, 4 * (size_t)b, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d + o_mm, model_mb, 8 * (size_t)b, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d + o_it, iterations, 8 * (size_t)b, cudaMemcpyHostToDevice, h->stream));
  long long want = (b + 3) / 4;
  int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  gs_netcost_kernel<<<grid, 128, 0, h->stream>>>((long long)b, (const long long *)(d + o_off), (const int *)(d + o_tn),
                                                 is_ps ? (const unsigned char *)(d + o_ps) : nullptr,
                                                 (const int *)(d + o_cnt), (const double *)(d + o_mm),
                                                 (const double *)(d + o_it), cluster->bandwidth,
                                                 cluster->internode_latency, (double *)(d + o_out));
  CU(cudaGetLastError());
  h->launches += 1;
  return timed_d2h(h, extra_out, d + o_out, 8 * (size_t)b);
}

extern "C" int gs_switch_yarn(gs_handle h, int32_t ncl, const gs_switch_cluster *clusters, gs_switch_node *nodes, int64_t n_nodes,
                              const gs_switch_job *jobs, int64_t n_jobs, const double *ps_network, int64_t n_ps,
                              double worker_mem, double ps_mem, double p_w_mem,
                              gs_switch_ans *ans, gs_switch_span *spans, int64_t n_spans) {
  if (!h) return GS_ERR_ARG;
  if (ncl < 0 || n_nodes < 0 || n_jobs < 0 || n_ps < 0 || n_spans < 0 ||
      (ncl > 0 && (!clusters || !nodes)) || (n_jobs > 0 && (!jobs || !ans || !spans)) || (n_ps > 0 && !ps_network))
    return fail(h, GS_ERR_ARG, "gs_switch_yarn: bad arguments");
  static_assert(sizeof(gs_switch_cluster) == 40 && sizeof(gs_switch_node) == 24 && sizeof(gs_switch_job) == 32 &&
                sizeof(gs_switch_ans) == 8 && sizeof(gs_switch_span) == 32, "switch record layout");
  for (int32_t c = 0; c < ncl; ++c) {
    const gs_switch_cluster &cl = clusters[c];
    const int64_t m = (int64_t)cl.num_switch * cl.num_node_p_switch;
    if (cl.num_switch <= 0 || cl.num_node_p_switch <= 0 || cl.num_gpu_p_node <= 0 || cl.node_off < 0 || cl.node_off + m > n_nodes ||
        cl.job_off < 0 || cl.job_cnt < 0 || cl.job_off + cl.job_cnt > n_jobs)
      return fail(h, GS_ERR_ARG, "gs_switch_yarn: a cluster record points outside the tables");
    for (int64_t j = cl.job_off; j < cl.job_off + cl.job_cnt; ++j) {
      const gs_switch_job &jb = jobs[j];
      const int64_t slots = jb.num_gpu / cl.num_gpu_p_node + 1;
      if (jb.num_gpu <= 0 || jb.n_ps < 0 || jb.ps_off < 0 || jb.ps_off + jb.n_ps > n_ps || jb.span_off < 0 || jb.span_off + slots > n_spans)
        return fail(h, GS_ERR_ARG, "gs_switch_yarn: a job record points outside the tables (spans need num_gpu / G + 1 slots)");
    }
  }
  if (ncl == 0 || n_jobs == 0) return GS_OK;
  CU(cudaSetDevice(h->device));
  size_t total = 0;
  auto take = [&](size_t bytes) { const size_t o = total; total = align_up(total + (bytes ? bytes : 1)); return o; };
  const size_t o_cl = take(sizeof(gs_switch_cluster) * (size_t)ncl), o_nd = take(sizeof(gs_switch_node) * (size_t)n_nodes);
  const size_t o_jb = take(sizeof(gs_switch_job) * (size_t)n_jobs), o_ps = take(8 * (size_t)n_ps);
  const size_t o_an = take(sizeof(gs_switch_ans) * (size_t)n_jobs), o_sp = take(sizeof(gs_switch_span) * (size_t)n_spans);
  int rc = ensure_scratch(h, total);
  if (rc) return rc;
  unsigned char *d = (unsigned char *)h->d_scratch;
  CU(cudaMemcpyAsync(d + o_cl, clusters, sizeof(gs_switch_cluster) * (size_t)ncl, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d + o_nd, nodes, sizeof(gs_switch_node) * (size_t)n_nodes, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(d + o_jb, jobs, sizeof(gs_switch_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, h->stream));
  if (n_ps > 0) CU(cudaMemcpyAsync(d + o_ps, ps_network, 8 * (size_t)n_ps, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemsetAsync(d + o_sp, 0, sizeof(gs_switch_span) * (size_t)n_spans, h->stream));
  CU(cudaEventRecord(h->e0, h->stream));
  gs_switch_yarn_kernel<<<(unsigned)ncl, 32, 0, h->stream>>>(ncl, (const gs_switch_cluster *)(d + o_cl), (gs_switch_node *)(d + o_nd),
                                                            (const gs_switch_job *)(d + o_jb), (const double *)(d + o_ps),
                                                            worker_mem, ps_mem, p_w_mem, (gs_switch_ans *)(d + o_an), (gs_switch_span *)(d + o_sp));
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaMemcpyAsync(nodes, d + o_nd, sizeof(gs_switch_node) * (size_t)n_nodes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(ans, d + o_an, sizeof(gs_switch_ans) * (size_t)n_jobs, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(spans, d + o_sp, sizeof(gs_switch_span) * (size_t)n_spans, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->kernel_ms += ms;
  return GS_OK;
}

// Restart every replica from tick 0 on the traces already resident in HBM.
extern "C" int gs_reset(gs_handle h) {
  if (!h) return GS_ERR_ARG;
  for (auto &s : h->sims) s.prepared = false;
  h->dirty = true;
  h->kernel_ms = h->h2d_ms = h->d2h_ms = 0;
  return GS_OK;
}

extern "C" int64_t gs_launch_count(gs_handle h) { return h ? h->launches : 0; }

// Span-pool sizing.  Default (0): the worst case sum(min(tasks, nodes)) per replica, which can
// never overflow.  budget > 0: min(worst case, budget * n + 4096) records -- less HBM per replica;
// a replica that would overflow stops with GS_ERR_CAPACITY instead of writing out of bounds.
extern "C" int gs_set_span_budget(gs_handle h, double spans_per_job) {
  if (!h) return GS_ERR_ARG;
  if (!(spans_per_job >= 0)) return fail(h, GS_ERR_ARG, "gs_set_span_budget: must be >= 0");
  h->span_budget = spans_per_job;
  return GS_OK;
}

// event-driven policies: 0 (or 1) = one warp per replica, 2 = one thread per replica; the fifo engine has one mapping
extern "C" int gs_set_engine(gs_handle h, int mode) {
  if (!h) return GS_ERR_ARG;
  if (mode < 0 || mode > 2) return fail(h, GS_ERR_ARG, "gs_set_engine: mode must be 0, 1 or 2");
  h->engine_mode = mode;
  return GS_OK;
}

// Pinned host buffers for callers that want DMA-speed gs_load_trace / gs_fetch_* copies.
extern "C" int gs_host_alloc(size_t bytes, void **out) {
  gs_handle h = nullptr;
  if (!out) return GS_ERR_ARG;
  *out = nullptr;
  CU(cudaMallocHost(out, bytes > 0 ? bytes : 1));
  return GS_OK;
}

extern "C" int gs_host_free(void *p) {
  gs_handle h = nullptr;
  if (p) CU(cudaFreeHost(p));
  return GS_OK;
}
