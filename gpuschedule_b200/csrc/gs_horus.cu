// gs_horus.cu -- C ABI (include/gsched_horus.h) and kernel of the utilisation-aware placement engine.
// One simulation per thread (see gs_horus_core.cuh for the semantics and the reference citations).
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#define GS_HD __host__ __device__ __forceinline__
#include "gs_horus_core.cuh"
#include "gs_horus_host.h"

#ifdef __CUDACC__
// lanes = simulations per warp: 32 (every lane drives one) or 1 (lane 0 only: no divergence inside the warp,
// more warps in flight for the same number of replicas).
#ifndef GS_HORUS_MINBLOCKS
#define GS_HORUS_MINBLOCKS 16      // resident warps per SM the register budget is cut for (one active lane each when lanes == 1):
                                   // 128 registers, no spills; measured 9.5e5 events/s at 9472 replicas (12: 8.7e5, 21: 8.7e5, 25: 8.1e5)
#endif
__global__ void __launch_bounds__(32, GS_HORUS_MINBLOCKS) gs_horus_kernel(HSim *sims, int nsims, long long max_ticks, int lanes) {
  const int i = lanes == 32 ? blockIdx.x * 32 + threadIdx.x : (threadIdx.x == 0 ? (int)blockIdx.x : nsims);
  if (i >= nsims) return;
  HSim s = sims[i];                 // pointers + scalars in registers / local memory
  if (s.n < 0 || s.done || s.status != 0) return;
  h_run(s, max_ticks);
  sims[i] = s;
}

// One simulation per WARP: lane 0 runs the simulation, all lanes score a candidate job's devices together
// (see "Warp-cooperative driver" in gs_horus_core.cuh).  The per-warp copy of the state lives in shared memory.
__global__ void __launch_bounds__(32) gs_horus_coop_kernel(HSim *sims, int nsims, long long max_ticks) {
  __shared__ HSim s;
  __shared__ int req;
  const int b = blockIdx.x, lane = threadIdx.x;
  if (b >= nsims) return;
  if (lane == 0) { s = sims[b]; s.budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL; }
  __syncwarp();
  if (s.n < 0 || s.done || s.status != 0) return;          // uniform: every lane reads the same shared words
  for (;;) {
    if (lane == 0) req = h_coop_advance(s);
    __syncwarp();                                           // lane 0's state (shared and global) is visible to the warp
    const int r = req;
    if (r == H_REQ_DONE) break;
    if (r == H_REQ_PREP) h_coop_prep(s); else if (r == H_REQ_SCORE) h_coop_score(s); else h_coop_stats(s);
    __syncwarp();                                           // the lanes' counts / costs are visible to lane 0
  }
  if (lane == 0) { h_write_records(s); sims[b] = s; }
}

#endif  // __CUDACC__

namespace {
struct WordStream {                 // raw MT19937 words + the per-position sample tables (gs_horus_host.h)
  std::vector<uint32_t> words;
  unsigned char *dev = nullptr; size_t cap = 0; bool dirty = false;
  const unsigned int *d_words = nullptr; const double *d_ret = nullptr, *d_keep = nullptr; const int *d_next = nullptr;
  const int *d_acc = nullptr, *d_rank = nullptr; int cls_off[5] = {0, 0, 0, 0, 0};
};
struct HorusSimHost {
  bool configured = false, loaded = false, prepared = false;
  gs_cluster cl{};
  gs_horus_params par{};
  std::vector<HJob> jobs;
  std::vector<double> stream;
  void *slab = nullptr; size_t slab_bytes = 0;
  double *d_stream = nullptr; size_t stream_cap = 0;
  HSim dev{};                       // host mirror of the device struct
  long long rows_cap = 0;
  bool use_shared = false;          // consume the handle-wide stream (gs_horus_load_stream with sim = -1)
  WordStream ws; int word_mode = 0; // 0: standard-normal values, 1: own words, 2: the handle-wide words
  std::vector<double> mem_avg;
};
}  // namespace

struct gs_horus_handle_s {
  int device = 0;
  std::vector<HorusSimHost> sims;
  HSim *d_sims = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  long long launches = 0;
  std::string err;
  std::vector<double> shared;       // one stream consumed by every replica that did not get its own
  double *d_shared = nullptr; size_t shared_cap = 0; bool shared_dirty = false;
  int lanes = 1;                    // simulations per warp (gs_horus_set_lanes)
  WordStream shared_ws;
};

static std::string g_horus_create_err;
static int hfail(gs_horus_handle h, int code, const std::string &msg) { if (h) h->err = msg; else g_horus_create_err = msg; return code; }
#define HCU(call)                                                                              \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return hfail(h, GS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

static size_t up(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int gs_horus_create(int device, int nsims, gs_horus_handle *out) {
  gs_horus_handle h = nullptr;
  if (!out || nsims <= 0) return hfail(nullptr, GS_ERR_ARG, "gs_horus_create: bad arguments");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return hfail(nullptr, GS_ERR_CUDA, "gs_horus_create: no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= ndev) return hfail(nullptr, GS_ERR_ARG, "gs_horus_create: device out of range");
  h = new gs_horus_handle_s();
  h->device = device;
  h->sims.resize((size_t)nsims);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
      cudaMalloc(&h->d_sims, sizeof(HSim) * (size_t)nsims) != cudaSuccess) {
    delete h;
    return hfail(nullptr, GS_ERR_CUDA, "gs_horus_create: CUDA initialisation failed");
  }
  size_t stack = 0;                 // the scalar kernel keeps ~1.5 KB of per-thread state on its stack frame
  if (cudaDeviceGetLimit(&stack, cudaLimitStackSize) == cudaSuccess && stack < 4096) (void)cudaDeviceSetLimit(cudaLimitStackSize, 4096);
  *out = h;
  return GS_OK;
}

extern "C" int gs_horus_destroy(gs_horus_handle h) {
  if (!h) return GS_ERR_ARG;
  cudaSetDevice(h->device);
  for (auto &s : h->sims) { if (s.slab) cudaFree(s.slab); if (s.d_stream) cudaFree(s.d_stream); if (s.ws.dev) cudaFree(s.ws.dev); }
  if (h->d_shared) cudaFree(h->d_shared);
  if (h->shared_ws.dev) cudaFree(h->shared_ws.dev);
  if (h->d_sims) cudaFree(h->d_sims);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return GS_OK;
}

extern "C" const char *gs_horus_build_tag(void) {
#ifdef __CUDACC__
  return "cuda:sm_100a";
#else
  return "host-emulation";
#endif
}
extern "C" const char *gs_horus_last_error(gs_horus_handle h) { return h ? h->err.c_str() : g_horus_create_err.c_str(); }
extern "C" int64_t gs_horus_launch_count(gs_horus_handle h) { return h ? h->launches : 0; }
extern "C" int gs_horus_set_lanes(gs_horus_handle h, int lanes) {
  if (!h || (lanes != 0 && lanes != 1 && lanes != 32)) return hfail(h, GS_ERR_ARG, "gs_horus_set_lanes: 0 (cooperative warp), 1 or 32");
  h->lanes = lanes;
  return GS_OK;
}

extern "C" int gs_horus_config(gs_horus_handle h, int32_t sim, const gs_cluster *c, const gs_horus_params *p) {
  if (!h || !c || !p || sim < 0 || sim >= (int)h->sims.size()) return hfail(h, GS_ERR_ARG, "gs_horus_config: bad arguments");
  if (c->num_switch <= 0 || c->num_node_p_switch <= 0 || c->num_gpu_p_node <= 0 || c->num_gpu_p_node > 64)
    return hfail(h, GS_ERR_ARG, "gs_horus_config: bad cluster shape");
  if (p->score != GS_HSCORE_HORUS && p->score != GS_HSCORE_GANDIVA) return hfail(h, GS_ERR_ARG, "gs_horus_config: unknown score function");
  if (p->schedule != GS_HSCHED_HORUS && p->schedule != GS_HSCHED_HORUS_PLUS && p->schedule != GS_HSCHED_GANDIVA)
    return hfail(h, GS_ERR_ARG, "gs_horus_config: schedule must be horus, horus+ or gandiva (with fifo the reference raises KeyError in score_fn, algorithm.py:58)");
  if ((p->schedule == GS_HSCHED_GANDIVA) != (p->score == GS_HSCORE_GANDIVA))
    return hfail(h, GS_ERR_ARG, "gs_horus_config: the score function follows the schedule name (gandiva_score <=> schedule gandiva)");
  if (p->schedule == GS_HSCHED_HORUS_PLUS && (p->num_queue < 1 || p->num_queue > H_MAXQ))
    return hfail(h, GS_ERR_ARG, "gs_horus_config: horus+ needs 1..8 queues");
  if (p->placement != GS_HPLACE_HORUS && p->placement != GS_HPLACE_YARN) return hfail(h, GS_ERR_ARG, "gs_horus_config: unknown placement");
  if (c->enable_network_costs) return hfail(h, GS_ERR_ARG, "gs_horus_config: network costs are not part of this path");
  auto &s = h->sims[(size_t)sim];
  s.cl = *c; s.par = *p; s.configured = true; s.prepared = false;
  return GS_OK;
}

extern "C" int gs_horus_load_trace(gs_horus_handle h, int32_t sim, int64_t n, const int32_t *arrive, const int32_t *gpus,
                                   const int32_t *gpc, const double *duration, const int64_t *mem_bytes,
                                   const double *util_avg, const double *util_max, const double *mem_avg_mib) {
  if (!h || sim < 0 || sim >= (int)h->sims.size() || n < 0 || n > 0x3fffffff) return hfail(h, GS_ERR_ARG, "gs_horus_load_trace: bad arguments");
  if (n > 0 && (!arrive || !gpus || !gpc || !duration || !mem_bytes || !util_avg || !util_max)) return hfail(h, GS_ERR_ARG, "gs_horus_load_trace: null column");
  auto &s = h->sims[(size_t)sim];
  std::vector<HJob> jobs((size_t)n);     // validated into a temporary: a rejected trace leaves the replica as it was
  long long first = 0;
  for (int64_t j = 0; j < n; ++j) {
    if (gpc[j] <= 0 || gpus[j] < gpc[j] || gpus[j] % gpc[j] != 0) return hfail(h, GS_ERR_ARG, "gs_horus_load_trace: used_gpus must be a positive multiple of gpu_per_container");
    if (j > 0 && arrive[j] < arrive[j - 1]) return hfail(h, GS_ERR_ARG, "gs_horus_load_trace: rows must be in admission order");
    if (util_max[j] < util_avg[j]) return hfail(h, GS_ERR_ARG, "gs_horus_load_trace: gpu_utilization_max < avg (numpy raises on a negative scale)");
    HJob &o = jobs[(size_t)j];
    o.arrive = arrive[j]; o.gpus = gpus[j]; o.gpc = gpc[j]; o.ntasks = gpus[j] / gpc[j]; o.first_task = (int)first; o.pad = 0;
    o.mem_b = mem_bytes[j]; o.util_avg = util_avg[j]; o.util_max = util_max[j]; o.duration = duration[j];
    o.mem_avg_mib = mem_avg_mib ? mem_avg_mib[j] : 0.0;
    first += o.ntasks;
    if (first > 0x3fffffff) return hfail(h, GS_ERR_ARG, "gs_horus_load_trace: too many tasks");
  }
  s.jobs.swap(jobs);
  s.loaded = true; s.prepared = false;
  return GS_OK;
}

extern "C" int gs_horus_load_stream(gs_horus_handle h, int32_t sim, const double *g, int64_t count) {
  if (!h || sim < -1 || sim >= (int)h->sims.size() || count < 0 || (count > 0 && !g)) return hfail(h, GS_ERR_ARG, "gs_horus_load_stream: bad arguments");
  if (sim == -1) {                  // every replica reads the same samples (each from position 0)
    h->shared.assign(g, g + count); h->shared_dirty = true;
    for (auto &s : h->sims) { s.use_shared = true; s.stream.clear(); s.prepared = false; s.word_mode = 0; }
    return GS_OK;
  }
  auto &s = h->sims[(size_t)sim];
  s.stream.assign(g, g + count);
  s.use_shared = false; s.prepared = false; s.word_mode = 0;
  return GS_OK;
}

extern "C" int gs_horus_load_words(gs_horus_handle h, int32_t sim, const uint32_t *w, int64_t count) {
  if (!h || sim < -1 || sim >= (int)h->sims.size() || count < 0 || count > 0x7ffffff0 || (count > 0 && !w)) return hfail(h, GS_ERR_ARG, "gs_horus_load_words: bad arguments");
  if (sim == -1) {
    h->shared_ws.words.assign(w, w + count); h->shared_ws.dirty = true;
    for (auto &s : h->sims) { s.word_mode = 2; s.prepared = false; }
    return GS_OK;
  }
  auto &s = h->sims[(size_t)sim];
  s.ws.words.assign(w, w + count); s.ws.dirty = true;
  s.word_mode = 1; s.prepared = false;
  return GS_OK;
}

static int upload_words(gs_horus_handle h, WordStream &ws) {
  if (!ws.dirty) return GS_OK;
  const size_t n = ws.words.size(), N = n ? n : 1;
  const size_t o_w = 0, o_ret = up(4 * N), o_keep = up(o_ret + 8 * N), o_next = up(o_keep + 8 * N), o_acc = up(o_next + 4 * N);
  const size_t o_rank = up(o_acc + 4 * N), total = up(o_rank + 4 * N);
  if (ws.dev && ws.cap < total) { cudaFree(ws.dev); ws.dev = nullptr; }
  if (!ws.dev) { HCU(cudaMalloc(&ws.dev, total)); ws.cap = total; }
  std::vector<double> ret(N), keep(N); std::vector<int> next(N);
  std::vector<int> acc(N), rank(N);
  gs_horus_build_gauss_tables(ws.words.data(), (long long)n, ret.data(), keep.data(), next.data());
  gs_horus_build_gauss_index(next.data(), (long long)n, acc.data(), rank.data(), ws.cls_off);
  if (n) {
    HCU(cudaMemcpyAsync(ws.dev + o_acc, acc.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    HCU(cudaMemcpyAsync(ws.dev + o_rank, rank.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    HCU(cudaMemcpyAsync(ws.dev + o_w, ws.words.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
    HCU(cudaMemcpyAsync(ws.dev + o_ret, ret.data(), 8 * n, cudaMemcpyHostToDevice, h->stream));
    HCU(cudaMemcpyAsync(ws.dev + o_keep, keep.data(), 8 * n, cudaMemcpyHostToDevice, h->stream));
    HCU(cudaMemcpyAsync(ws.dev + o_next, next.data(), 4 * n, cudaMemcpyHostToDevice, h->stream));
  }
  HCU(cudaStreamSynchronize(h->stream));
  ws.d_words = (const unsigned int *)(ws.dev + o_w); ws.d_ret = (const double *)(ws.dev + o_ret);
  ws.d_keep = (const double *)(ws.dev + o_keep); ws.d_next = (const int *)(ws.dev + o_next);
  ws.d_acc = (const int *)(ws.dev + o_acc); ws.d_rank = (const int *)(ws.dev + o_rank);
  ws.dirty = false;
  return GS_OK;
}

static int prepare(gs_horus_handle h, HorusSimHost &s, long long rows_cap) {
  const gs_cluster &c = s.cl;
  const int M = c.num_switch * c.num_node_p_switch, G = c.num_gpu_p_node;
  const size_t n = s.jobs.size(), N = n ? n : 1;
  long long ntask = 0; int maxg = 1;
  for (auto &j : s.jobs) { ntask += j.ntasks; maxg = std::max(maxg, j.gpus); }
  const size_t NT = ntask ? (size_t)ntask : 1;
  const int pjw = (M + 63) / 64;
  const int nb = std::max(1, s.par.num_buffer);
  const int nq = s.par.schedule == GS_HSCHED_HORUS_PLUS ? std::max(1, s.par.num_queue) : 1;
  if (s.par.schedule == GS_HSCHED_HORUS_PLUS && s.word_mode == 0)
    return hfail(h, GS_ERR_STATE, "gs_horus_run: horus+ draws integers too: load the raw stream with gs_horus_load_words");
  if (s.word_mode == 1) { int rc = upload_words(h, s.ws); if (rc) return rc; }
  if (rows_cap <= 0) return hfail(h, GS_ERR_ARG, "gs_horus_run: rows_cap must be positive");
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = up(off + bytes); return o; };
  const size_t o_jobs = take(sizeof(HJob) * N), o_js = take(sizeof(HJobState) * N), o_tasks = take(sizeof(HTask) * NT);
  const size_t o_tron = take(4 * NT), o_troo = take(4 * NT), o_nodes = take(sizeof(HNode) * (size_t)M), o_devs = take(sizeof(HDev) * (size_t)M * G);
  const size_t o_pj = take(8 * N * (size_t)pjw), o_q = take(4 * (N + 1) * (size_t)nq), o_run = take(4 * N), o_fin = take(4 * N);
  const size_t o_look = take(4 * (size_t)nb), o_lookq = take(4 * (size_t)nb), o_work = take(4 * N), o_res = take(4 * (size_t)M);
  const size_t o_kall = take(4 * N), o_kas = take(4 * N), o_kold = take(4 * N), o_ksc = take(8 * N);
  const size_t o_sccnt = take(4 * (size_t)M * G), o_scoff = take(4 * (size_t)M * G), o_sccost = take(8 * (size_t)M * G);
  const size_t o_mn = take(4 * (size_t)maxg * maxg), o_mo = take(4 * (size_t)maxg * maxg), o_mc = take(4 * (size_t)maxg);
  const size_t o_ok = take(4 * (size_t)maxg), o_di = take(4 * (size_t)maxg), o_heap = take(sizeof(HCand) * ((size_t)maxg + 2));
  const size_t o_mskip = take(8 * (size_t)maxg);
  const size_t o_rows = take(sizeof(gs_tick_row) * (size_t)rows_cap), o_util = take(8 * (size_t)rows_cap), o_ua = take((size_t)rows_cap);
  const size_t o_recs = take(sizeof(gs_horus_job_rec) * N);
  const size_t total = off;
  if (s.slab && s.slab_bytes < total) { cudaFree(s.slab); s.slab = nullptr; }
  if (!s.slab) { HCU(cudaMalloc(&s.slab, total)); s.slab_bytes = total; }
  HCU(cudaMemsetAsync(s.slab, 0, total, h->stream));
  unsigned char *d = (unsigned char *)s.slab;
  if (n) HCU(cudaMemcpyAsync(d + o_jobs, s.jobs.data(), sizeof(HJob) * n, cudaMemcpyHostToDevice, h->stream));
  std::vector<HTask> tasks(NT);
  std::vector<int> tron(NT, -1);
  gs_horus_init_tasks(s.jobs.data(), (long long)n, (long long)c.gpu_mem_cap_mib << 20, tasks.data());
  HCU(cudaMemcpyAsync(d + o_tasks, tasks.data(), sizeof(HTask) * NT, cudaMemcpyHostToDevice, h->stream));
  HCU(cudaMemcpyAsync(d + o_tron, tron.data(), 4 * NT, cudaMemcpyHostToDevice, h->stream));
  if (s.stream.size() > s.stream_cap) {
    if (s.d_stream) cudaFree(s.d_stream);
    s.d_stream = nullptr;
    HCU(cudaMalloc(&s.d_stream, 8 * s.stream.size()));
    s.stream_cap = s.stream.size();
  }
  if (!s.stream.empty()) HCU(cudaMemcpyAsync(s.d_stream, s.stream.data(), 8 * s.stream.size(), cudaMemcpyHostToDevice, h->stream));
  HCU(cudaStreamSynchronize(h->stream));            // the staging vectors go out of scope below
  HSim &D = s.dev;
  D = HSim{};
  D.M = M; D.G = G; D.S = c.num_switch; D.P = c.num_node_p_switch; D.cpu_cap = c.num_cpu_p_node; D.mem_cap = c.mem_p_node;
  D.placement = s.par.placement;
  D.scheme = s.par.score; D.schedule = s.par.schedule; D.num_buffer = s.par.num_buffer; D.n = (int)n; D.maxg = maxg; D.pjw = pjw;
  D.cap_b = (long long)c.gpu_mem_cap_mib << 20;
  D.jobs = (const HJob *)(d + o_jobs); D.js = (HJobState *)(d + o_js); D.tasks = (HTask *)(d + o_tasks);
  D.tro_node = (int *)(d + o_tron); D.tro_order = (int *)(d + o_troo); D.nodes = (HNode *)(d + o_nodes); D.devs = (HDev *)(d + o_devs);
  D.pj_bits = (unsigned long long *)(d + o_pj); D.queue = (int *)(d + o_q); D.running = (int *)(d + o_run); D.fin = (int *)(d + o_fin);
  D.look = (int *)(d + o_look); D.look_q = (int *)(d + o_lookq); D.work = (int *)(d + o_work); D.res_nodes = (int *)(d + o_res);
  D.km_all = (int *)(d + o_kall); D.km_assign = (int *)(d + o_kas); D.km_old = (int *)(d + o_kold); D.km_score = (double *)(d + o_ksc);
  D.nq = nq;
  D.sc_cnt = (int *)(d + o_sccnt); D.sc_off = (int *)(d + o_scoff); D.sc_cost = (double *)(d + o_sccost);
  if (s.word_mode) {
    const WordStream &ws = s.word_mode == 1 ? s.ws : h->shared_ws;
    D.words = ws.d_words; D.gv_ret = ws.d_ret; D.gv_keep = ws.d_keep; D.gv_next = ws.d_next; D.words_n = (long long)ws.words.size();
    D.gv_acc = ws.d_acc; D.gv_rank = ws.d_rank; for (int c = 0; c < 5; ++c) D.gv_cls_off[c] = ws.cls_off[c];
  }
  D.map_node = (int *)(d + o_mn); D.map_order = (int *)(d + o_mo); D.map_n = (int *)(d + o_mc); D.ok = (int *)(d + o_ok); D.distinct = (int *)(d + o_di);
  D.heap = (HCand *)(d + o_heap); D.map_skip = (long long *)(d + o_mskip);
  D.gauss = s.use_shared ? h->d_shared : s.d_stream;
  D.gauss_n = (long long)(s.use_shared ? h->shared.size() : s.stream.size()); D.gauss_pos = 0;
  D.rows = (gs_tick_row *)(d + o_rows); D.util = (double *)(d + o_util); D.util_arr = d + o_ua; D.recs = (gs_horus_job_rec *)(d + o_recs);
  D.rows_cap = rows_cap;
  D.current_remaining = (long long)n; D.running_jobs = 0;
  s.rows_cap = rows_cap;
  s.prepared = true;
  return GS_OK;
}

extern "C" int gs_horus_run(gs_horus_handle h, int64_t max_ticks, int64_t rows_cap) {
  if (!h) return GS_ERR_ARG;
  HCU(cudaSetDevice(h->device));
  const int nsims = (int)h->sims.size();
  std::vector<HSim> host((size_t)nsims);
  if (h->shared_dirty) {
    if (h->shared.size() > h->shared_cap) {
      if (h->d_shared) cudaFree(h->d_shared);
      h->d_shared = nullptr;
      HCU(cudaMalloc(&h->d_shared, 8 * h->shared.size()));
      h->shared_cap = h->shared.size();
    }
    if (!h->shared.empty()) HCU(cudaMemcpyAsync(h->d_shared, h->shared.data(), 8 * h->shared.size(), cudaMemcpyHostToDevice, h->stream));
    HCU(cudaStreamSynchronize(h->stream));
    h->shared_dirty = false;
    for (auto &s : h->sims) if (s.use_shared) s.prepared = false;          // the buffer may have moved
  }
  if (h->shared_ws.dirty) {
    int rc = upload_words(h, h->shared_ws); if (rc) return rc;
    for (auto &s : h->sims) if (s.word_mode == 2) s.prepared = false;
  }
  for (int i = 0; i < nsims; ++i) {
    auto &s = h->sims[(size_t)i];
    if (!s.configured || !s.loaded) return hfail(h, GS_ERR_STATE, "gs_horus_run: every replica needs gs_horus_config + gs_horus_load_trace");
    if (!s.prepared) { int rc = prepare(h, s, rows_cap); if (rc) return rc; }
    host[(size_t)i] = s.dev;
  }
  HCU(cudaMemcpyAsync(h->d_sims, host.data(), sizeof(HSim) * (size_t)nsims, cudaMemcpyHostToDevice, h->stream));
  HCU(cudaEventRecord(h->ev0, h->stream));
#ifdef __CUDACC__
  if (h->lanes == 32) gs_horus_kernel<<<(nsims + 31) / 32, 32, 0, h->stream>>>(h->d_sims, nsims, (long long)max_ticks, 32);
  else if (h->lanes == 1) gs_horus_kernel<<<nsims, 32, 0, h->stream>>>(h->d_sims, nsims, (long long)max_ticks, 1);
  else gs_horus_coop_kernel<<<nsims, 32, 0, h->stream>>>(h->d_sims, nsims, (long long)max_ticks);
#else   // host build for tests/emu (fake_cuda/cuda_runtime.h): what the kernels do, one simulation after the other
  for (int i = 0; i < nsims; ++i) {
    HSim &sm = h->d_sims[i];
    if (sm.n < 0 || sm.done || sm.status != 0) continue;
    if (h->lanes == 0) h_run_coop(sm, (long long)max_ticks); else h_run(sm, (long long)max_ticks);
  }
#endif
  h->launches += 1;
  HCU(cudaGetLastError());
  HCU(cudaEventRecord(h->ev1, h->stream));
  HCU(cudaMemcpyAsync(host.data(), h->d_sims, sizeof(HSim) * (size_t)nsims, cudaMemcpyDeviceToHost, h->stream));
  HCU(cudaStreamSynchronize(h->stream));
  HCU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  int worst = GS_OK;
  for (int i = 0; i < nsims; ++i) {
    h->sims[(size_t)i].dev = host[(size_t)i];
    if (host[(size_t)i].status != 0 && worst == GS_OK) worst = host[(size_t)i].status;
  }
  if (worst == GS_ERR_CAPACITY) return hfail(h, worst, "gs_horus_run: a replica ran out of rows or of random samples (raise rows_cap / load a longer stream)");
  if (worst != GS_OK) return hfail(h, worst, "gs_horus_run: a replica reached an inconsistent state");
  return GS_OK;
}

extern "C" int gs_horus_stats(gs_horus_handle h, int32_t sim, gs_horus_run_stats *out) {
  if (!h || !out || sim < 0 || sim >= (int)h->sims.size()) return hfail(h, GS_ERR_ARG, "gs_horus_stats: bad arguments");
  const HSim &D = h->sims[(size_t)sim].dev;
  out->ticks = D.ticks; out->events = D.events; out->draws = D.draws;
  int queued = 0;
  for (int q = 0; q < D.nq; ++q) queued += D.qn[q];
  out->finished = D.nfin; out->queued = queued; out->running = D.nrun; out->done = D.done; out->status = D.status; out->reserved = 0;
  out->kernel_ms = h->last_ms; out->reserved2 = 0.f;
  return GS_OK;
}

extern "C" int gs_horus_fetch(gs_horus_handle h, int32_t sim, gs_tick_row *rows, double *util, uint8_t *util_is_array,
                              int64_t rows_cap, gs_horus_job_rec *recs, int32_t *finish_order, int64_t *n_rows, int64_t *n_finished) {
  if (!h || sim < 0 || sim >= (int)h->sims.size()) return hfail(h, GS_ERR_ARG, "gs_horus_fetch: bad arguments");
  auto &s = h->sims[(size_t)sim];
  if (!s.prepared) return hfail(h, GS_ERR_STATE, "gs_horus_fetch: nothing has run");
  HCU(cudaSetDevice(h->device));
  const HSim &D = s.dev;
  if (D.ticks > rows_cap) return hfail(h, GS_ERR_CAPACITY, "gs_horus_fetch: row buffer too small");
  if (rows && D.ticks) HCU(cudaMemcpyAsync(rows, D.rows, sizeof(gs_tick_row) * (size_t)D.ticks, cudaMemcpyDeviceToHost, h->stream));
  if (util && D.ticks) HCU(cudaMemcpyAsync(util, D.util, 8 * (size_t)D.ticks, cudaMemcpyDeviceToHost, h->stream));
  if (util_is_array && D.ticks) HCU(cudaMemcpyAsync(util_is_array, D.util_arr, (size_t)D.ticks, cudaMemcpyDeviceToHost, h->stream));
  if (recs && D.n) HCU(cudaMemcpyAsync(recs, D.recs, sizeof(gs_horus_job_rec) * (size_t)D.n, cudaMemcpyDeviceToHost, h->stream));
  if (finish_order && D.nfin) HCU(cudaMemcpyAsync(finish_order, D.fin, 4 * (size_t)D.nfin, cudaMemcpyDeviceToHost, h->stream));
  HCU(cudaStreamSynchronize(h->stream));
  if (n_rows) *n_rows = D.ticks;
  if (n_finished) *n_finished = D.nfin;
  return GS_OK;
}
