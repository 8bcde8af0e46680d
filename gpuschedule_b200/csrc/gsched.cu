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
//   * every tick emits one 64-byte gs_tick_row of integer aggregates; the job
//     table is streamed once from HBM (arrival-ordered SoA), results are
//     written once (24-byte gs_job_rec + 16-byte gs_span per (job,node)).
//   * thousands of replicas run per launch (one warp each, 148 SMs x many
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

#include <string>
#include <vector>

#include "gsched.h"

#ifndef GS_TICK_MINBLOCKS
#define GS_TICK_MINBLOCKS 24
#endif
#define FULL 0xffffffffu
// ballot over the lanes of one replica group, bit 0 = the group's first lane (needs GM, gbase, SUB in scope)
#define GBALLOT(pred) ((SUB == 32) ? __ballot_sync(GM, (pred)) : ((__ballot_sync(GM, (pred)) >> gbase) & ((1u << SUB) - 1u)))

// ------------------------------------------------------------------ device state

struct __align__(16) JobState {   // 32 B, written at start, read once at completion
  int next;                  // next job in the same finish-tick bucket (start order)
  int node0;                 // span_cnt == 1: the node;  span_cnt > 1: first index in the span pool
  unsigned long long mask0;  // span_cnt == 1: devices held on node0
  long long memc;            // gpus * min(device capacity, memory_max): the job's share of the memory column
  int gpus;
  int cnt_gpc;               // span_cnt (bits 0-23) | gpu_per_task (bits 24-31)
};
#define JS_CNT(x) ((x) & 0xffffff)
#define JS_GPC(x) ((int)((unsigned)(x) >> 24))
// JobState.gpus: gpus (bits 0-23) | tasks of a single-span job (bits 24-31, <= 64 because gpus <= G <= 64 there)
#define JS_GPUS(x) ((x) & 0xffffff)
#define JS_NT0(x) ((int)((unsigned)(x) >> 24))

struct __align__(32) JobIn {   // 32 B = one DRAM sector per job, read once in admission order
  int arrive;       // first tick with normalized_time <= tick
  int gpus;
  int gpc;          // gpu_per_container
  int ps;           // ps_count (0 when the trace has no network columns)
  long long memb;   // memory_max, bytes
  double dur;       // minutes * 0.5
};

struct PJob {     // 32 B: the fields of the legacy job dict the policies touch (run_sim.py:208-230,730-779)
  int last_check, total_exec, exec, pending, last_pending, start, resume;
  unsigned char status, q_id, pad0, pad1;
};
enum { PST_NONE = 0, PST_PENDING = 1, PST_RUNNING = 2, PST_END = 3 };

struct SimDev {
  // ---- configuration
  int M, G, K;          // nodes, gpus/node, task slots/node = min(cpu/cpu_pt, mem/mem_pt)
  int netcost, n, wheel_mask, policy, pad0;
  long long cap_bytes;  // Device.memory in bytes
  long long fit_limit;  // a task fits an empty device iff mem_bytes < fit_limit
  double bandwidth, latency;
  // ---- trace (read-only)
  const JobIn *jobs;
  const double *model_mb, *iters;
  // ---- results / scratch
  gs_job_rec *rec;
  JobState *jst;
  int2 *sref;                 // per job: {first index in the span pool, span count}
  int *stack, *fin, *wheel_head, *wheel_tail;
  gs_span *spans;
  gs_tick_row *rows;
  unsigned long long *nbusy;  // persisted node table (between launches)
  int *nk;                    // bit31 = node ever hosted a placement (node.py:93-97, never cleared)
  long long span_cap, rows_cap;
  // ---- event-driven policies (sjf / dlas / dlas-gpu / gittins): scratch + parameters
  struct PJob *pj;            // per-job dynamic state
  int *runnable, *queues, *endj, *tmpl, *cidle, *ckfree;   // queues: num_queue lists of n entries
  const double *git_data, *git_index;                      // device copies of the gittins tables
  double queue_limit[GS_MAX_QUEUES];
  double gittins_delta, next_gittins_unit;
  int num_queue, git_n, rn, en, end_time, next_job_jump, qn[GS_MAX_QUEUES];
  // ---- loop state (persisted)
  int delta, p, top, running, finished, ever, busy_gpus, done, status, need_init;
  long long mem_busy, sum_arr, span_used, events, evals, started, ticks, row_first;
};

#define EVER_BIT 0x80000000u

__device__ __forceinline__ int meta_cap(unsigned mt, int gpc) {
  // tasks a node can still take: min(idle devices / gpus per task, free task slots)
  const int idle = (int)(mt & 0xffu), kfree = (int)(mt >> 16);
  return min(gpc == 1 ? idle : idle / gpc, kfree);
}

// Shared-memory geometry shared by both tick kernels
#define LW 128          // finish-tick buckets kept in shared memory (ticks, power of two)
#define SCACHE 4        // cached top-of-queue entries (power of two)
#define WARP_EXTRA_BYTES (LW * 8 + SCACHE * 8)

__device__ __forceinline__ unsigned long long take_lowest(unsigned long long idle, int cnt, int G) {
  // the `cnt` lowest set bits of `idle` (devices are claimed in index order, node.py:208-216)
  if (cnt == 1) return idle & (~idle + 1ull);
  if (G <= 32) {
    unsigned m = (unsigned)idle;
    if (cnt >= __popc(m)) return idle;
    for (int i = 0; i < cnt; ++i) m &= m - 1u;
    return (unsigned long long)((unsigned)idle ^ m);
  }
  unsigned long long m = idle;
  if (cnt >= __popcll(m)) return idle;
  for (int i = 0; i < cnt; ++i) m &= m - 1ull;
  return idle ^ m;
}

// One warp advances one replica.  Lanes stripe over the node table for placement; all other
// state is warp-uniform.  Per warp in shared memory: node table (busy mask u64 + slot counter
// i32 per node), a LW-tick window of the finish wheel (head, tail) and the top SCACHE queue
// entries.  Every global load on the per-tick path is issued one tick before its value is
// needed (pending wheel bucket, release record of the next tick's first finisher) or comes
// from a register-resident 32-record window of the trace, so the loop body has no dependent
// DRAM/L2 round trip in the common case.
template <int SUB>
__global__ void __launch_bounds__(32, GS_TICK_MINBLOCKS) gs_tick_kernel(SimDev *sims, int nsims, long long max_ticks, int smem_stride) {
  // SUB lanes own one replica: 32 = a whole warp, 16 = two replicas per warp sharing the
  // (mostly warp-uniform) instruction stream.  Every collective uses the group's own lane
  // mask, so the groups of a warp may diverge freely.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wl = threadIdx.x & 31;              // lane within the warp
  const int lane = wl & (SUB - 1);              // lane within the replica's group
  const int gbase = wl & ~(SUB - 1);            // first warp lane of the group
  const unsigned GM = (SUB == 32) ? 0xffffffffu : (((1u << SUB) - 1u) << gbase);
  const int grp = wl / SUB;
  const int sim = blockIdx.x * (32 / SUB) + grp;
  if (sim >= nsims) return;
  SimDev &S = sims[sim];
  if (S.done || S.status != 0 || S.policy != GS_SCHED_FIFO) return;

  const int M = S.M, G = S.G, K = S.K, n = S.n;
  unsigned long long *busy = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)grp * smem_stride);
  int *kk = reinterpret_cast<int *>(busy + M);
  int2 *sstk = reinterpret_cast<int2 *>(kk + M + (M & 1));   // 8-byte aligned

  const JobIn *__restrict__ jobs = S.jobs;
  gs_job_rec *rec = S.rec;
  JobState *jst = S.jst;
  int2 *stack = reinterpret_cast<int2 *>(S.stack);
  int *fin = S.fin, *gwh = S.wheel_head, *gwt = S.wheel_tail;
  gs_span *spans = S.spans;
  const int wmask = S.wheel_mask;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int netcost = S.netcost;
  const long long span_cap = S.span_cap;      // SimDev lives in global memory: keep loop-invariant fields in registers
  int2 *sref = S.sref;
  const unsigned long long gmask = (G >= 64) ? ~0ull : ((1ull << G) - 1ull);

  int delta = S.delta, p = S.p, top = S.top, running = S.running, finished = S.finished;
  int ever = S.ever, busy_gpus = S.busy_gpus, status = 0;
  long long mem_busy = S.mem_busy, sum_arr = S.sum_arr, span_used = S.span_used;
  long long evals = S.evals, started = S.started, ticks = S.ticks;
  const long long row_first = ticks;
  gs_tick_row *rows = S.rows;
  const long long rows_cap = S.rows_cap;
  long long budget_ll = max_ticks > 0 ? max_ticks : 0x7fffffffLL;
  if (budget_ll > rows_cap) budget_ll = rows_cap;     // the row window bounds a launch anyway
  int budget = (int)(budget_ll > 0x7fffffffLL ? 0x7fffffffLL : budget_ll);
  int4 *rowp = reinterpret_cast<int4 *>(rows);
  int tick_i = 0;                                     // ticks done in this launch

  // ---- stage persistent state: node table, wheel window, queue top
  for (int i = lane; i < M; i += SUB) {
    const unsigned long long bz = S.nbusy[i];
    const unsigned kv = (unsigned)S.nk[i];
    busy[i] = bz;
    kk[i] = (int)((unsigned)(G - __popcll(bz)) | ((kv & EVER_BIT) ? 0x100u : 0u) | ((unsigned)(K - (int)(kv & ~EVER_BIT)) << 16));
  }
  for (int i = max(top - SCACHE, 0) + lane; i < top; i += SUB) sstk[i & (SCACHE - 1)] = stack[i];
  int cache_lo = max(top - SCACHE, 0);           // queue entries [cache_lo, top) are cached
  __syncwarp(GM);

  // trace window: lane l holds record wbase + l
  int wbase = p & ~(SUB - 1);
  JobIn wj;
  wj.arrive = 0x7fffffff; wj.gpus = 1; wj.gpc = 1; wj.ps = 0; wj.memb = 0; wj.dur = 0.0;
  if (wbase + lane < n) wj = jobs[wbase + lane];
  // arrival tick of the next trace record: most ticks admit nothing and only compare this scalar
  int next_arr = 0x7fffffff;
  if (p < n) next_arr = __shfl_sync(GM, wj.arrive, gbase + (p - wbase));

  // queue head (cached while it stays the head)
  int head = -1, hg = 1, hgpc = 1, htasks = 1, hps = 0, harr = 0;
  long long hmemb = 0;
  double hdur = 0.0;
  bool head_valid = false;
  int bottom_arr = (top > 0) ? stack[0].y : 0;


  bool done = (n == 0);

  while (!done && tick_i < budget && status == 0) {
    // ---------------- A. admit arrivals (gen_jobs + head insert)
    if (next_arr <= delta) {
      int cnt = 0, q = p;
      while (q < n) {
        const int idx = wbase + lane;
        const unsigned b = GBALLOT(idx >= q && idx < n && wj.arrive <= delta);
        const int c = __popc(b);
        if (c > 0 && cnt == 0) {
          // the batch's first job becomes the queue head: take its record out of the window now
          const int src = p - wbase;
          hg = __shfl_sync(GM, wj.gpus, gbase + src); hgpc = __shfl_sync(GM, wj.gpc, gbase + src);
          harr = delta;
          if (netcost) hps = __shfl_sync(GM, wj.ps, gbase + src);
          hmemb = __shfl_sync(GM, wj.memb, gbase + src);
          hdur = __longlong_as_double(__shfl_sync(GM, __double_as_longlong(wj.dur), gbase + src));
        }
        cnt += c; q += c;
        if (q < wbase + SUB || q >= n) break;
        wbase += SUB;
        wj.arrive = 0x7fffffff;
        if (wbase + lane < n) wj = jobs[wbase + lane];
      }
      if (cnt > 0) {
        // batch [p, p+cnt) lands AHEAD of the queue, first of the batch on top (quirk Q2)
        if (cnt == 1) {
          if (lane == 0) { const int2 e = make_int2(p, delta); stack[top] = e; sstk[top & (SCACHE - 1)] = e; }
        } else {
          for (int i = lane; i < cnt; i += SUB) {
            const int2 e = make_int2(p + cnt - 1 - i, delta);
            stack[top + i] = e;
            if (i >= cnt - SCACHE) sstk[(top + i) & (SCACHE - 1)] = e;
          }
        }
        if (top == 0) bottom_arr = delta;
        head = p; head_valid = true; htasks = hgpc == 1 ? hg : hg / hgpc;
        top += cnt; p += cnt;
        if (top - cache_lo > SCACHE) cache_lo = top - SCACHE;
        sum_arr += (long long)cnt * delta;
        __syncwarp(GM);
      }
      next_arr = 0x7fffffff;
      if (p < n) next_arr = __shfl_sync(GM, wj.arrive, gbase + (p - wbase));
    }
    // ---------------- B. one scheduling attempt on the queue head (quirks Q1, Q3)
    if (top > 0) {
      if (!head_valid) {
        if (top - 1 >= cache_lo) head = sstk[(top - 1) & (SCACHE - 1)].x;
        else { head = stack[top - 1].x; cache_lo = top; }
        if (head >= wbase && head < wbase + SUB) {
          const int src = head - wbase;
          hg = __shfl_sync(GM, wj.gpus, gbase + src); hgpc = __shfl_sync(GM, wj.gpc, gbase + src);
          hps = __shfl_sync(GM, wj.ps, gbase + src); harr = __shfl_sync(GM, wj.arrive, gbase + src);
          hmemb = __shfl_sync(GM, wj.memb, gbase + src);
          hdur = __longlong_as_double(__shfl_sync(GM, __double_as_longlong(wj.dur), gbase + src));
        } else {
          const JobIn jr = jobs[head];
          hg = jr.gpus; hgpc = jr.gpc; hmemb = jr.memb; hdur = jr.dur; hps = jr.ps; harr = jr.arrive;
        }
        htasks = hgpc == 1 ? hg : hg / hgpc;
        head_valid = true;
      }
      const bool placeable = hmemb < fit_limit;   // Device.can_fit on an empty device
      bool ok = false;
      int first_node = -1, nspans = 0;
      const int span_first = (int)span_used;
      unsigned long long mask0 = 0;
      if (hg <= G) {
        // try_single_node_alloc_ms: first node (id order) that fits the whole job
        int found = -1;
        for (int base = 0; base < M; base += SUB) {
          const int nd = base + lane;
          bool fit = false;
          if (nd < M) {
            const unsigned mt = (unsigned)kk[nd];
            fit = (int)(mt & 0xffu) >= hg && (int)(mt >> 16) >= htasks;
          }
          const unsigned b = GBALLOT(fit);
          if (!placeable) {          // quirk Q21: cpu/mem charged for every task, never refunded
            if (fit) kk[nd] -= htasks << 16;
            continue;
          }
          if (b) { found = base + __ffs(b) - 1; break; }
        }
        if (found >= 0 && span_used + 1 > span_cap) { status = GS_ERR_CAPACITY; found = -1; }
        if (found >= 0) {
          ok = true; first_node = found; nspans = 1;
          bool fresh = false;
          if (lane == (found & (SUB - 1))) {
            const unsigned long long take = take_lowest(~busy[found] & gmask, hg, G);
            busy[found] |= take;
            const unsigned kv = (unsigned)kk[found];
            kk[found] = (int)((kv - (unsigned)hg - ((unsigned)htasks << 16)) | 0x100u);
            mask0 = take;
            fresh = !(kv & 0x100u);
            gs_span sp; sp.node = found; sp.ntasks = htasks; sp.devmask = take;
            spans[span_first] = sp;
          }
          mask0 = __shfl_sync(GM, mask0, gbase + (found & (SUB - 1)));
          ever += __popc(GBALLOT(fresh));
          evals += found + 1;
        } else {
          evals += M;
        }
      } else {
        // try_cross_node_alloc_ms: walk nodes in id order, each takes what it can hold
        int cum = 0, last_base = -1;
        if (placeable) {
          for (int base = 0; base < M; base += SUB) {
            const int nd = base + lane;
            const int c = (nd < M) ? meta_cap((unsigned)kk[nd], hgpc) : 0;
            cum += __reduce_add_sync(GM, c);
            if (cum >= htasks) { last_base = base; break; }
          }
        } else {
          for (int base = 0; base < M; base += SUB) {   // quirk Q21, cross-node flavour: one task charged per node
            const int nd = base + lane;
            if (nd < M && meta_cap((unsigned)kk[nd], hgpc) > 0) kk[nd] -= 1 << 16;
          }
        }
        if (last_base >= 0 && span_used + min(htasks, M) > span_cap) { status = GS_ERR_CAPACITY; last_base = -1; }
        if (last_base >= 0) {
          // pass 1 proved the job fits: commit (a failed walk is rolled back exactly by the
          // reference, algorithm.py:378-387, so no state changes in that case)
          ok = true;
          int rem = htasks, last_node = 0;
          for (int base = 0; base <= last_base; base += SUB) {
            const int nd = base + lane;
            const int c = (nd < M) ? meta_cap((unsigned)kk[nd], hgpc) : 0;
            int incl = c;
            #pragma unroll
            for (int o = 1; o < SUB; o <<= 1) { int v = __shfl_up_sync(GM, incl, o, SUB); if (lane >= o) incl += v; }
            const int take = min(c, max(rem - (incl - c), 0));
            const unsigned tb = GBALLOT(take > 0);
            bool fresh = false;
            if (take > 0) {
              const unsigned long long tk = take_lowest(~busy[nd] & gmask, take * hgpc, G);
              busy[nd] |= tk;
              const unsigned kv = (unsigned)kk[nd];
              kk[nd] = (int)((kv - (unsigned)(take * hgpc) - ((unsigned)take << 16)) | 0x100u);
              fresh = !(kv & 0x100u);
              const int slot = nspans + __popc(tb & ((1u << lane) - 1u));
              gs_span sp; sp.node = nd; sp.ntasks = take; sp.devmask = tk;
              spans[span_first + slot] = sp;
              if (slot == 0) { mask0 = tk; first_node = nd; }
            }
            ever += __popc(GBALLOT(fresh));
            if (tb) last_node = base + 31 - __clz(tb);
            nspans += __popc(tb);
            const int tot = __shfl_sync(GM, incl, gbase + SUB - 1);
            rem -= min(rem, tot);
          }
          {  // first-span fields live in whichever lane owned slot 0
            const int src = __ffs(GBALLOT(first_node >= 0)) - 1;
            mask0 = __shfl_sync(GM, mask0, gbase + src);
            first_node = __shfl_sync(GM, first_node, gbase + src);
          }
          evals += last_node + 1;
        } else {
          evals += M;
        }
      }
      __syncwarp(GM);
      if (ok) {
        // ---- commit: pop, network cost, start (algorithm.py:198-200, schedule.py:49-54,164-167)
        const int j = head;
        double dur2 = hdur;
        if (netcost && hps > 1) {
          // (model_size/bandwidth + cross*latency) * (iterations*2.0), network_service.py:34-37
          const double mps = __ddiv_rn(S.model_mb[j], S.bandwidth);
          const double nis = __dmul_rn((double)nspans, S.latency);
          const double rt = __dmul_rn(S.iters[j], 2.0);
          dur2 = __dadd_rn(hdur, __dmul_rn(__dadd_rn(mps, nis), rt));
        }
        const double eff = dur2 > hdur ? dur2 : hdur;               // Job.get_duration (job.py:206-210)
        const double cl = ceil(eff);
        int need = cl < 1.0 ? 1 : (cl > 1.0e9 ? 0x7fffffff : (int)cl);   // quirk Q11
        if (need > wmask) { status = GS_ERR_ARG; need = wmask; }
        const int endt = delta + need;
        span_used += nspans;
        const long long memc = (long long)hg * (hmemb < cap_bytes ? hmemb : cap_bytes);
        JobState js; js.next = -1; js.node0 = nspans == 1 ? first_node : span_first; js.mask0 = mask0;
        js.memc = memc; js.gpus = hg | ((nspans == 1 ? htasks : 0) << 24); js.cnt_gpc = nspans | (hgpc << 24);
        // append to the finish-tick bucket of the timing wheel (start order)
        const int gs_ = endt & wmask;
        const int tl = gwt[gs_];
        if (lane == 0) {
          if (tl < 0) gwh[gs_] = j;
          gwt[gs_] = j;
          gs_job_rec r; r.start = delta; r.end = endt; r.jct = need; r.preempt = 1; r.duration = dur2;
          rec[j] = r;
          jst[j] = js;
          sref[j] = make_int2(span_first, nspans);
          if (tl >= 0) jst[tl].next = j;
        }
        top -= 1;
        sum_arr -= harr;
        running += 1; started += 1;
        busy_gpus += hg;
        mem_busy += memc;
        head = -1; head_valid = false;
        __syncwarp(GM);
      }
    }
    // ---------------- D/E. time advances; release jobs whose finish tick is now
    const int now = delta + 1;
    {
      const int sl = now & wmask;
      int h = gwh[sl];
      if (h >= 0) {
        __syncwarp(GM);
        if (lane == 0) { gwh[sl] = -1; gwt[sl] = -1; }
        while (h >= 0) {
          const JobState js = jst[h];
          const int scnt = JS_CNT(js.cnt_gpc), sgpc = JS_GPC(js.cnt_gpc);
          if (scnt == 1) {
            if (lane == 0) { busy[js.node0] &= ~js.mask0; kk[js.node0] += JS_GPUS(js.gpus) + (JS_NT0(js.gpus) << 16); }
          } else {
            for (int i = lane; i < scnt; i += SUB) {
              const gs_span sp = spans[js.node0 + i];
              busy[sp.node] &= ~sp.devmask; kk[sp.node] += sp.ntasks * sgpc + (sp.ntasks << 16);
            }
          }
          if (lane == 0) fin[finished] = h;
          finished += 1; running -= 1;
          busy_gpus -= JS_GPUS(js.gpus);
          mem_busy -= js.memc;
          h = js.next;
        }
        __syncwarp(GM);
      }
    }
    // ---------------- H. statistics row (schedule.py:95-133) from O(1) counters
    {
      int pmax = 0, mlo = 0, mhi = 0;
      if (top > 0) {
        // queue is a stack with non-decreasing arrival ticks bottom->top, so the sorted
        // pending list is the stack read top->bottom: median/max are index look-ups
        const int ilo = top - 1 - ((top - 1) >> 1), ihi = top - 1 - (top >> 1);
        const int a_lo = ilo >= cache_lo ? sstk[ilo & (SCACHE - 1)].y : stack[ilo].y;
        const int a_hi = ihi >= cache_lo ? sstk[ihi & (SCACHE - 1)].y : stack[ihi].y;
        pmax = now - bottom_arr; mlo = now - a_lo; mhi = now - a_hi;
      }
      if (lane == 0) {
        int4 *dst = rowp;
        const int tg = M * G;
        const long long ps = top > 0 ? (long long)top * now - sum_arr : 0;
        dst[0] = make_int4(now, M - ever, ever, busy_gpus);
        dst[1] = make_int4(tg - busy_gpus, running, top, finished);
        dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(ps & 0xffffffffLL), (int)(ps >> 32));
        dst[3] = make_int4(pmax, mlo, mhi, 0);
      }
    }
    tick_i += 1; rowp += 4;
    delta = now;
    done = (n - p) + running == 0;      // schedule.py:185 -- the queue is NOT counted (quirk Q4)
  }

  // ---------------- persist: node table, wheel window and pending bucket go back to global memory
  __syncwarp(GM);
  for (int i = lane; i < M; i += SUB) {
    const unsigned mt = (unsigned)kk[i];
    S.nbusy[i] = busy[i];
    S.nk[i] = (int)((unsigned)(K - (int)(mt >> 16)) | ((mt & 0x100u) ? EVER_BIT : 0u));
  }
  if (lane == 0) {
    S.delta = delta; S.p = p; S.top = top; S.running = running; S.finished = finished;
    S.ever = ever; S.busy_gpus = busy_gpus; S.mem_busy = mem_busy; S.sum_arr = sum_arr;
    S.span_used = span_used; S.events = (long long)p + started + finished; S.evals = evals; S.started = started;
    S.ticks = ticks + tick_i; S.row_first = row_first; S.done = done ? 1 : 0; S.status = status;
  }
}

// ------------------------------------------------------------------ lane engine
// One THREAD owns one replica; a warp carries up to 32 unrelated replicas.  This is the
// throughput kernel: a replica's tick is almost all scalar bookkeeping, so a whole warp
// per replica wastes 31/32 of the issue slots, and the number of replicas in flight is
// capped by HBM capacity (~18 MB per 100k-job replica), so per-tick LATENCY decides
// throughput.  Everything the common path touches therefore lives in shared memory or
// registers, and every global load is issued one iteration before its value is needed:
//   * node table        meta word per node: idle devices (0-7) | ever (8) | free slots (16-31),
//                       plus the busy-device bitmap (32 or 64 bit)
//   * wheel window      finish-tick buckets (head, tail) for the next LW ticks; far buckets
//                       stay in the global wheel and are pulled in LW ticks ahead
//   * job ring          the next few 32-byte trace records, refilled one per tick
//   * stack cache       the top 4 queue entries (job, arrival tick)
//   * release record    JobState of the job finishing next tick, prefetched into registers
// Shared memory is laid out [word][lane] so lane l always hits bank l: conflict-free no
// matter which node / slot each lane is looking at.  First fit is a serial scan from `lo`,
// the lowest node with an idle device (first fit packs low ids, so the scan is short).
// Lanes never share data: no warp collectives except the per-tick reconvergence barrier.
#define META_IDLE(m) ((int)((m) & 0xffu))
#define META_EVER 0x100u
#define META_KFREE(m) ((int)((m) >> 16))
#define RING 8          // job-record ring, entries (power of two)
#define LANE_EXTRA_WORDS (2 * LW + RING * 8 + SCACHE * 2)

template <typename MaskT>
__global__ void __launch_bounds__(32) gs_lane_kernel(SimDev *sims, int nsims, long long max_ticks, int Mmax, int L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int sim = blockIdx.x * L + lane;
  const bool in_range = lane < L && sim < nsims;
  SimDev &S = sims[in_range ? sim : 0];
  const bool alive = in_range && !S.done && S.status == 0 && S.policy == GS_SCHED_FIFO;

  const int M = S.M, G = S.G, K = S.K, n = S.n;
  const int MW = (sizeof(MaskT) == 8) ? 3 : 2;
  uint32_t *meta = reinterpret_cast<uint32_t *>(smem_raw) + (lane < L ? lane : 0);   // [nd * L]
  uint32_t *mlo = meta + (size_t)Mmax * L;
  uint32_t *mhi = meta + (size_t)2 * Mmax * L;                 // only when MaskT is 64 bit
  uint32_t *swh = meta + (size_t)MW * Mmax * L;                // [slot * L]  bucket head
  uint32_t *swt = swh + (size_t)LW * L;                        //             bucket tail
  uint32_t *ring = swt + (size_t)LW * L;                       // [(slot * 8 + word) * L]
  uint32_t *sstk = ring + (size_t)RING * 8 * L;                // [(slot * 2 + {job,arrive}) * L]
  const MaskT gmask = (G >= (int)(8 * sizeof(MaskT))) ? (MaskT)~(MaskT)0 : (MaskT)(((MaskT)1 << G) - 1);

  const JobIn *__restrict__ jobs = S.jobs;
  gs_job_rec *rec = S.rec;
  JobState *jst = S.jst;
  int2 *sref = S.sref;
  int2 *stack = reinterpret_cast<int2 *>(S.stack);
  int *fin = S.fin, *gwh = S.wheel_head, *gwt = S.wheel_tail;
  gs_span *spans = S.spans;
  const int wmask = S.wheel_mask;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int netcost = S.netcost;

  int delta = S.delta, p = S.p, top = S.top, running = S.running, finished = S.finished;
  int ever = S.ever, busy_gpus = S.busy_gpus, status = 0;
  long long mem_busy = S.mem_busy, sum_arr = S.sum_arr, span_used = S.span_used;
  long long evals = S.evals, started = S.started, ticks = S.ticks;
  const long long row_first = ticks;
  gs_tick_row *rows = S.rows;
  const long long rows_cap = S.rows_cap;
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;

  // ---- stage the persistent state into shared memory
  int lo = M;
  int pf = p;                        // ring holds trace records [max(ring_lo, pf - RING), pf)
  int ring_lo = p;
  int pend_h = -1, pend_t = -1;      // bucket of tick delta + LW, loaded but not yet in the window
  if (alive) {
    for (int nd = 0; nd < M; ++nd) {
      unsigned long long bz = S.nbusy[nd];
      unsigned kv = (unsigned)S.nk[nd];
      int idle = G - __popcll(bz);
      meta[nd * L] = (uint32_t)idle | ((kv & EVER_BIT) ? META_EVER : 0u) | ((uint32_t)(K - (int)(kv & ~EVER_BIT)) << 16);
      mlo[nd * L] = (uint32_t)bz;
      if (sizeof(MaskT) == 8) mhi[nd * L] = (uint32_t)(bz >> 32);
      if (idle > 0 && nd < lo) lo = nd;
    }
    for (int t = delta + 1; t <= delta + LW - 1; ++t) {      // window ticks move from the global wheel
      int gs_ = t & wmask;
      swh[(t & (LW - 1)) * L] = (uint32_t)gwh[gs_]; swt[(t & (LW - 1)) * L] = (uint32_t)gwt[gs_];
      gwh[gs_] = -1; gwt[gs_] = -1;
    }
    { int gs_ = (delta + LW) & wmask; pend_h = gwh[gs_]; pend_t = gwt[gs_]; gwh[gs_] = -1; gwt[gs_] = -1; }
    for (; pf < n && pf < p + RING - 2; ++pf) {
      const uint4 *src = reinterpret_cast<const uint4 *>(&jobs[pf]);
      uint4 a0 = src[0], a1 = src[1];
      uint32_t *r = ring + (size_t)((pf & (RING - 1)) * 8) * L;
      r[0] = a0.x; r[L] = a0.y; r[2 * L] = a0.z; r[3 * L] = a0.w;
      r[4 * L] = a1.x; r[5 * L] = a1.y; r[6 * L] = a1.z; r[7 * L] = a1.w;
    }
    for (int i = max(top - SCACHE, 0); i < top; ++i) {
      int2 e = stack[i];
      sstk[((i & (SCACHE - 1)) * 2) * L] = (uint32_t)e.x; sstk[((i & (SCACHE - 1)) * 2 + 1) * L] = (uint32_t)e.y;
    }
  }
  int cache_lo = max(top - SCACHE, 0);     // stack entries [cache_lo, top) are in the cache

  // trace record q -> registers (ring if resident, else global)
  auto load_job = [&](int q) -> JobIn {
    JobIn r;
    if (q < pf && q >= pf - RING && q >= ring_lo) {
      const uint32_t *w = ring + (size_t)((q & (RING - 1)) * 8) * L;
      r.arrive = (int)w[0]; r.gpus = (int)w[L]; r.gpc = (int)w[2 * L]; r.ps = (int)w[3 * L];
      r.memb = (long long)(((unsigned long long)w[5 * L] << 32) | w[4 * L]);
      r.dur = __longlong_as_double((long long)(((unsigned long long)w[7 * L] << 32) | w[6 * L]));
    } else {
      r = jobs[q];
    }
    return r;
  };
  auto arrive_of = [&](int q) -> int {
    if (q >= n) return 0x7fffffff;
    if (q < pf && q >= pf - RING && q >= ring_lo) return (int)ring[(size_t)((q & (RING - 1)) * 8) * L];
    return jobs[q].arrive;
  };

  int next_arrive = alive ? arrive_of(p) : 0x7fffffff;
  int head = -1, htasks = 1;
  JobIn hj;
  hj.arrive = 0; hj.gpus = 1; hj.gpc = 1; hj.ps = 0; hj.memb = 0; hj.dur = 0.0;
  int bottom_arr = (alive && top > 0) ? stack[0].y : 0;
  // pipelined loads: issued at the end of iteration d, consumed in iteration d + 1
  bool rp_valid = false; uint4 rp0 = make_uint4(0, 0, 0, 0), rp1 = make_uint4(0, 0, 0, 0);   // trace record pf
  int pre_h = -1; JobState pre_js;                                                          // release record
  pre_js.next = -1; pre_js.node0 = 0; pre_js.mask0 = 0; pre_js.memc = 0; pre_js.gpus = 0; pre_js.cnt_gpc = 1;
  int com_j = -1; JobState com_js = pre_js;                                                 // last commit
  bool done = !alive || (n == 0);

  // Every tick starts with the whole warp reconverged (no break/return inside the body,
  // explicit barrier): otherwise independent thread scheduling lets the replicas drift
  // apart and run one at a time.
  while (true) {
    const bool go = !done && status == 0 && budget > 0 && (ticks - row_first) < rows_cap;
    if (!__any_sync(FULL, go)) break;
    if (go) {
      // ---------------- A. admit arrivals: the batch lands ahead of the queue, first job on top (Q2)
      if (next_arrive <= delta) {
        const int a = p;
        int b = p, na;
        do { ++b; na = arrive_of(b); } while (na <= delta);
        if (top == 0) bottom_arr = delta;
        for (int i = b - 1; i >= a; --i) {
          stack[top] = make_int2(i, delta);
          sstk[((top & (SCACHE - 1)) * 2) * L] = (uint32_t)i; sstk[((top & (SCACHE - 1)) * 2 + 1) * L] = (uint32_t)delta;
          ++top;
        }
        if (top - cache_lo > SCACHE) cache_lo = top - SCACHE;
        sum_arr += (long long)(b - a) * delta;
        head = a; hj = load_job(a); htasks = hj.gpc == 1 ? hj.gpus : hj.gpus / hj.gpc;
        p = b; next_arrive = na;
      }
      // ---------------- B. one attempt on the queue head (Q1, Q3)
      com_j = -1;
      if (top > 0) {
        if (head < 0) {
          if (top - 1 >= cache_lo) head = (int)sstk[(((top - 1) & (SCACHE - 1)) * 2) * L];
          else { head = stack[top - 1].x; cache_lo = top; }      // cache exhausted: deeper entries are global only
          hj = load_job(head); htasks = hj.gpc == 1 ? hj.gpus : hj.gpus / hj.gpc;
        }
        const int hg = hj.gpus, hgpc = hj.gpc;
        const bool placeable = hj.memb < fit_limit;
        bool ok = false;
        int first_node = -1, nspans = 0;
        const int span_first = (int)span_used;
        MaskT mask0 = 0;
        if (hg <= G) {
          int found = -1;
          for (int nd = lo; nd < M; ++nd) {
            uint32_t mt = meta[nd * L];
            if (META_IDLE(mt) >= hg && META_KFREE(mt) >= htasks) {
              if (!placeable) { meta[nd * L] = mt - ((uint32_t)htasks << 16); continue; }   // Q21 leak
              found = nd; break;
            }
          }
          if (found >= 0 && span_used + 1 > S.span_cap) { status = GS_ERR_CAPACITY; found = -1; }
          if (found >= 0) {
            uint32_t mt = meta[found * L];
            MaskT bz = (MaskT)mlo[found * L];
            if (sizeof(MaskT) == 8) bz |= (MaskT)((unsigned long long)mhi[found * L] << 32);
            MaskT m = (MaskT)(~bz & gmask), take = 0;
            for (int i = 0; i < hg; ++i) { MaskT bit = (MaskT)(m & (MaskT)(~m + 1)); take |= bit; m ^= bit; }
            bz |= take;
            mlo[found * L] = (uint32_t)bz;
            if (sizeof(MaskT) == 8) mhi[found * L] = (uint32_t)((unsigned long long)bz >> 32);
            if (!(mt & META_EVER)) ever += 1;
            meta[found * L] = (mt - (uint32_t)hg - ((uint32_t)htasks << 16)) | META_EVER;
            gs_span sp; sp.node = found; sp.ntasks = htasks; sp.devmask = (unsigned long long)take;
            spans[span_first] = sp;
            ok = true; first_node = found; nspans = 1; mask0 = take;
            evals += found + 1;
          } else {
            evals += M;
          }
        } else {
          int cum = 0, last = -1;
          for (int nd = lo; nd < M; ++nd) {
            uint32_t mt = meta[nd * L];
            int idle = META_IDLE(mt);
            int c = min(hgpc == 1 ? idle : idle / hgpc, META_KFREE(mt));
            if (c <= 0) continue;
            if (!placeable) { meta[nd * L] = mt - (1u << 16); continue; }     // Q21 leak, one task per node
            cum += c;
            if (cum >= htasks) { last = nd; break; }
          }
          if (last >= 0 && span_used + min(htasks, M) > S.span_cap) { status = GS_ERR_CAPACITY; last = -1; }
          if (last >= 0) {
            int rem = htasks;
            for (int nd = lo; nd <= last; ++nd) {
              uint32_t mt = meta[nd * L];
              int idle = META_IDLE(mt);
              int c = min(hgpc == 1 ? idle : idle / hgpc, META_KFREE(mt));
              if (c <= 0) continue;
              int take_n = min(c, rem);
              MaskT bz = (MaskT)mlo[nd * L];
              if (sizeof(MaskT) == 8) bz |= (MaskT)((unsigned long long)mhi[nd * L] << 32);
              MaskT m = (MaskT)(~bz & gmask), take = 0;
              for (int i = 0; i < take_n * hgpc; ++i) { MaskT bit = (MaskT)(m & (MaskT)(~m + 1)); take |= bit; m ^= bit; }
              bz |= take;
              mlo[nd * L] = (uint32_t)bz;
              if (sizeof(MaskT) == 8) mhi[nd * L] = (uint32_t)((unsigned long long)bz >> 32);
              if (!(mt & META_EVER)) ever += 1;
              meta[nd * L] = (mt - (uint32_t)(take_n * hgpc) - ((uint32_t)take_n << 16)) | META_EVER;
              gs_span sp; sp.node = nd; sp.ntasks = take_n; sp.devmask = (unsigned long long)take;
              spans[span_first + nspans] = sp;
              if (nspans == 0) { first_node = nd; mask0 = take; }
              ++nspans;
              rem -= take_n;
            }
            ok = true;
            evals += last + 1;
          } else {
            evals += M;
          }
        }
        if (ok) {
          // ---- commit: pop, network cost, start (algorithm.py:198-200, schedule.py:49-54,164-167)
          while (lo < M && META_IDLE(meta[lo * L]) == 0) ++lo;
          const int j = head;
          double dur2 = hj.dur;
          if (netcost && hj.ps > 1) {
            double mps = __ddiv_rn(S.model_mb[j], S.bandwidth);
            double nis = __dmul_rn((double)nspans, S.latency);
            double rt = __dmul_rn(S.iters[j], 2.0);
            dur2 = __dadd_rn(hj.dur, __dmul_rn(__dadd_rn(mps, nis), rt));
          }
          double eff = dur2 > hj.dur ? dur2 : hj.dur;
          double cl = ceil(eff);
          int need = cl < 1.0 ? 1 : (cl > 1.0e9 ? 0x7fffffff : (int)cl);
          if (need > wmask) { status = GS_ERR_ARG; need = wmask; }
          const int endt = delta + need;
          span_used += nspans;
          gs_job_rec r; r.start = delta; r.end = endt; r.jct = need; r.preempt = 1; r.duration = dur2;
          rec[j] = r;
          sref[j] = make_int2(span_first, nspans);
          const long long memc = (long long)hg * (hj.memb < cap_bytes ? hj.memb : cap_bytes);
          JobState js; js.next = -1; js.node0 = nspans == 1 ? first_node : span_first;
          js.mask0 = (unsigned long long)mask0; js.memc = memc; js.gpus = hg | ((nspans == 1 ? htasks : 0) << 24); js.cnt_gpc = nspans | (hgpc << 24);
          jst[j] = js;
          com_j = j; com_js = js;
          // append to the finish-tick bucket (start order): window / pending register / global wheel
          int tl;
          if (need <= LW - 1) {
            const int sl = (endt & (LW - 1)) * L;
            tl = (int)swt[sl];
            if (tl < 0) swh[sl] = (uint32_t)j;
            swt[sl] = (uint32_t)j;
          } else if (need == LW) {
            tl = pend_t;
            if (tl < 0) pend_h = j;
            pend_t = j;
          } else {
            const int gs_ = endt & wmask;
            tl = gwt[gs_];
            if (tl < 0) gwh[gs_] = j;
            gwt[gs_] = j;
          }
          if (tl >= 0) {
            jst[tl].next = j;
            if (tl == pre_h) pre_js.next = j;
          }
          top -= 1;
          sum_arr -= hj.arrive;
          running += 1; started += 1;
          busy_gpus += hg;
          mem_busy += memc;
          head = -1;
        }
      }
      // ---------------- D/E. release jobs whose finish tick is now
      const int now = delta + 1;
      {
        const int sl = (now & (LW - 1)) * L;
        int h = (int)swh[sl];
        if (h >= 0) {
          swh[sl] = 0xffffffffu; swt[sl] = 0xffffffffu;
          do {
            JobState js;
            if (h == pre_h) js = pre_js;
            else if (h == com_j) js = com_js;
            else js = jst[h];
            const int scnt = JS_CNT(js.cnt_gpc), sgpc = JS_GPC(js.cnt_gpc);
            if (scnt == 1) {
              const int nd = js.node0;
              mlo[nd * L] &= ~(uint32_t)js.mask0;
              if (sizeof(MaskT) == 8) mhi[nd * L] &= ~(uint32_t)(js.mask0 >> 32);
              meta[nd * L] += (uint32_t)JS_GPUS(js.gpus) + ((uint32_t)JS_NT0(js.gpus) << 16);
              if (nd < lo) lo = nd;
            } else {
              for (int i = 0; i < scnt; ++i) {
                gs_span sp = spans[js.node0 + i];
                mlo[sp.node * L] &= ~(uint32_t)sp.devmask;
                if (sizeof(MaskT) == 8) mhi[sp.node * L] &= ~(uint32_t)(sp.devmask >> 32);
                meta[sp.node * L] += (uint32_t)(sp.ntasks * sgpc) + ((uint32_t)sp.ntasks << 16);
                if (sp.node < lo) lo = sp.node;
              }
            }
            fin[finished] = h;
            finished += 1; running -= 1;
            busy_gpus -= JS_GPUS(js.gpus);
            mem_busy -= js.memc;
            h = js.next;
          } while (h >= 0);
        }
      }
      // ---------------- H. statistics row (schedule.py:95-133) from O(1) counters
      {
        int pmax = 0, mlo_p = 0, mhi_p = 0;
        if (top > 0) {
          const int ilo = top - 1 - (top - 1) / 2, ihi = top - 1 - top / 2;
          const int alo = ilo >= cache_lo ? (int)sstk[((ilo & (SCACHE - 1)) * 2 + 1) * L] : stack[ilo].y;
          const int ahi = ihi >= cache_lo ? (int)sstk[((ihi & (SCACHE - 1)) * 2 + 1) * L] : stack[ihi].y;
          pmax = now - bottom_arr; mlo_p = now - alo; mhi_p = now - ahi;
        }
        int4 *dst = reinterpret_cast<int4 *>(&rows[ticks - row_first]);
        const int tg = M * G;
        const long long ps = top > 0 ? (long long)top * now - sum_arr : 0;
        dst[0] = make_int4(now, M - ever, ever, busy_gpus);
        dst[1] = make_int4(tg - busy_gpus, running, top, finished);
        dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(ps & 0xffffffffLL), (int)(ps >> 32));
        dst[3] = make_int4(pmax, mlo_p, mhi_p, 0);
      }
      // ---------------- pipeline stage: retire last iteration's loads, issue the next ones
      {
        // bucket of tick delta + LW enters the window (its slot held tick delta, consumed last iteration)
        const int sl = (delta & (LW - 1)) * L;
        swh[sl] = (uint32_t)pend_h; swt[sl] = (uint32_t)pend_t;
        const int gs_ = (delta + 1 + LW) & wmask;
        pend_h = gwh[gs_]; pend_t = gwt[gs_];
        gwh[gs_] = -1; gwt[gs_] = -1;
        // trace record ring: one record per tick
        if (rp_valid) {
          uint32_t *r = ring + (size_t)((pf & (RING - 1)) * 8) * L;
          r[0] = rp0.x; r[L] = rp0.y; r[2 * L] = rp0.z; r[3 * L] = rp0.w;
          r[4 * L] = rp1.x; r[5 * L] = rp1.y; r[6 * L] = rp1.z; r[7 * L] = rp1.w;
          ++pf;
        }
        // never run further ahead than RING - 2 past p (the 2 slots behind p keep the records of the
        // jobs admitted last, which are the ones popped next), never fall behind p
        rp_valid = false;
        if (pf < p) { pf = p; ring_lo = p; }         // a burst outran the ring: restart it at p
        if (pf < n && pf < p + RING - 2) {
          const uint4 *src = reinterpret_cast<const uint4 *>(&jobs[pf]);
          rp0 = src[0]; rp1 = src[1]; rp_valid = true;
        }
        // release record of the job that heads the bucket of tick now + 1
        pre_h = (int)swh[((now + 1) & (LW - 1)) * L];
        if (pre_h >= 0) pre_js = jst[pre_h];
      }
      ticks += 1; budget -= 1;
      delta = now;
      done = (n - p) + running == 0;      // schedule.py:185 -- the queue is NOT counted (Q4)
    }   // go
    __syncwarp();
  }

  if (!alive) return;
  // ---- persist: node table, wheel window and pending bucket go back to global memory
  for (int nd = 0; nd < M; ++nd) {
    uint32_t mt = meta[nd * L];
    unsigned long long bz = mlo[nd * L];
    if (sizeof(MaskT) == 8) bz |= (unsigned long long)mhi[nd * L] << 32;
    S.nbusy[nd] = bz;
    S.nk[nd] = (int)((uint32_t)(K - META_KFREE(mt)) | ((mt & META_EVER) ? EVER_BIT : 0u));
  }
  for (int t = delta + 1; t <= delta + LW - 1; ++t) {
    gwh[t & wmask] = (int)swh[(t & (LW - 1)) * L]; gwt[t & wmask] = (int)swt[(t & (LW - 1)) * L];
  }
  gwh[(delta + LW) & wmask] = pend_h; gwt[(delta + LW) & wmask] = pend_t;
  S.delta = delta; S.p = p; S.top = top; S.running = running; S.finished = finished;
  S.ever = ever; S.busy_gpus = busy_gpus; S.mem_busy = mem_busy; S.sum_arr = sum_arr;
  S.span_used = span_used; S.events = (long long)p + started + finished; S.evals = evals; S.started = started;
  S.ticks = ticks; S.row_first = row_first; S.done = done ? 1 : 0; S.status = status;
}

// ------------------------------------------------------------------ event-driven policies
// sjf / dlas / dlas-gpu / gittins: restated from the reference's dead Tiresias-style loops
// (run_sim.py:162-287, 664-947, 949-1203); the decisions taken where
// that code is undefined are listed in oracle/policy_oracle.c, which this kernel matches
// bit for bit.  First version: ONE THREAD per replica (a warp carries 32 replicas); every
// event re-evaluates all runnable jobs (counter update, ordering, emptied-cluster greedy
// re-admission), exactly as the specification does.  Lists live in global memory.
__device__ __forceinline__ double git_lookup(const SimDev &S, double a) {
  const int n = S.git_n;
  if (n < 2 || a > S.git_data[n - 2]) return 0.0;
  int lo = 0, hi = n - 1;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (S.git_data[mid] > a) hi = mid; else lo = mid + 1; }
  return S.git_index[lo];
}

__device__ __forceinline__ void plist_remove(int *v, int &n, int x) {
  int w = 0;
  for (int i = 0; i < n; ++i) { int e = v[i]; if (e != x) v[w++] = e; }
  n = w;
}

__device__ bool pol_yarn_place(const SimDev &S, int gpus, int gpc, bool placeable) {
  if (!placeable) return false;
  const int M = S.M, G = S.G, tasks = gpus / gpc;
  int *idle = S.cidle, *kfree = S.ckfree;
  if (gpus <= G) {
    for (int nd = 0; nd < M; ++nd)
      if (idle[nd] >= gpus && kfree[nd] >= tasks) { idle[nd] -= gpus; kfree[nd] -= tasks; return true; }
    return false;
  }
  int cum = 0, last = -1;
  for (int nd = 0; nd < M; ++nd) {
    int cap = min(idle[nd] / gpc, kfree[nd]);
    if (cap <= 0) continue;
    cum += cap;
    if (cum >= tasks) { last = nd; break; }
  }
  if (last < 0) return false;
  int rem = tasks;
  for (int nd = 0; nd <= last; ++nd) {
    int cap = min(idle[nd] / gpc, kfree[nd]);
    if (cap <= 0) continue;
    int take = min(cap, rem);
    idle[nd] -= take * gpc; kfree[nd] -= take; rem -= take;
  }
  return true;
}

__global__ void __launch_bounds__(32) gs_policy_kernel(SimDev *sims, int nsims, long long max_ticks, int take_dlas) {
  const int sim = blockIdx.x * blockDim.x + threadIdx.x;
  if (sim >= nsims) return;
  SimDev &S = sims[sim];
  if (S.policy == GS_SCHED_FIFO || S.done || S.status != 0) return;
  if (!take_dlas) return;   // every event-driven policy has a warp-cooperative kernel; this one is the fallback (engine mode 2)
  const int policy = S.policy, n = S.n, M = S.M, G = S.G, K = S.K;
  const bool is_dlas = policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU;
  const bool gputime = policy == GS_SCHED_DLAS_GPU || policy == GS_SCHED_GITTINS;
  const int nq = is_dlas ? S.num_queue : 1;
  const JobIn *__restrict__ jobs = S.jobs;
  PJob *pj = S.pj;
  int *runnable = S.runnable, *endj = S.endj, *tmpl = S.tmpl;
  gs_job_rec *rec = S.rec;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int total_gpus = M * G;
  int p = S.p, rn = S.rn, en = S.en, end_time = S.end_time, next_job_jump = S.next_job_jump, nfin = S.finished;
  double next_git = S.next_gittins_unit;
  long long events = S.events, ticks = S.ticks;
  const long long row_first = ticks;
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;
  int status = 0;
  bool done = false;

  while (budget > 0 && (ticks - row_first) < S.rows_cap) {
    if (!((n - p) + rn > 0)) { done = true; break; }
    if (p >= n && end_time == 0x7fffffff) { done = true; break; }     // "cluster is not large enough"
    const int start_time = p < n ? jobs[p].arrive : 0x7fffffff;
    int event_time; bool has_start = false, has_end = false;
    if (end_time < start_time) { event_time = end_time; has_end = true; }
    else if (end_time > start_time) { event_time = start_time; has_start = true; }
    else { event_time = start_time; has_start = has_end = true; }
    if (is_dlas && event_time > next_job_jump) { event_time = next_job_jump; has_start = has_end = false; }
    if (policy == GS_SCHED_GITTINS && (double)event_time > next_git) { event_time = (int)next_git; has_start = has_end = false; }
    if (has_end) {
      for (int i = 0; i < en; ++i) {
        const int j = endj[i];
        PJob &r = pj[j];
        r.status = PST_END;
        gs_job_rec o; o.start = r.start; o.end = event_time;
        double cl = ceil(jobs[j].dur); o.jct = cl < 1.0 ? 1 : (int)cl; o.preempt = r.resume; o.duration = jobs[j].dur;
        rec[j] = o;
        S.fin[nfin++] = j; ++events;
        plist_remove(runnable, rn, j);
        plist_remove(S.queues + (size_t)r.q_id * n, S.qn[r.q_id], j);
      }
    }
    if (has_start) {
      while (p < n && jobs[p].arrive == event_time) {
        const int j = p++;
        PJob r; r.last_check = event_time; r.total_exec = 0; r.exec = 0; r.pending = 0; r.last_pending = 0; r.start = -1;
        r.resume = 0; r.status = PST_PENDING; r.q_id = 0; r.pad0 = 0; r.pad1 = 0;
        pj[j] = r;
        runnable[rn++] = j; S.queues[S.qn[0]++] = j; ++events;
      }
    }
    for (int i = 0; i < rn; ++i) {
      const int j = runnable[i];
      PJob &r = pj[j];
      const int dt = event_time - r.last_check;
      r.last_check = event_time;
      if (r.status == PST_RUNNING) {
        r.total_exec += dt; r.exec += dt;
        if (is_dlas) {
          const double j_gt = gputime ? (double)r.exec * jobs[j].gpus : (double)r.exec;
          if (r.q_id < nq - 1 && j_gt >= S.queue_limit[r.q_id]) {
            plist_remove(S.queues + (size_t)r.q_id * n, S.qn[r.q_id], j);
            r.q_id += 1;
            S.queues[(size_t)r.q_id * n + S.qn[r.q_id]++] = j;
          }
        }
      } else {
        r.pending += dt;
        if (r.exec > 0) r.last_pending += dt;
      }
    }
    // ---- order, empty the cluster, greedy re-admission
    int nrun = 0, npre = 0, busy = 0;
    long long mem_busy = 0;
    int *run_jobs = tmpl, *pre_jobs = tmpl + (n > 0 ? n - 1 : 0);
    if (policy == GS_SCHED_SJF) {
      for (int i = 1; i < rn; ++i) {          // stable insertion sort by num_gpu (list is nearly sorted)
        const int x = runnable[i]; const int kx = jobs[x].gpus; int k = i;
        while (k > 0 && jobs[runnable[k - 1]].gpus > kx) { runnable[k] = runnable[k - 1]; --k; }
        runnable[k] = x;
      }
      for (int nd = 0; nd < M; ++nd) { S.cidle[nd] = G; S.ckfree[nd] = K; }
      for (int i = 0; i < rn; ++i) {
        const int j = runnable[i];
        const JobIn jr = jobs[j];
        PJob &r = pj[j];
        if (pol_yarn_place(S, jr.gpus, jr.gpc, jr.memb < fit_limit)) {
          if (r.start < 0) r.start = event_time;
          if (r.status == PST_PENDING) run_jobs[nrun++] = j;
          busy += jr.gpus; mem_busy += (long long)jr.gpus * (jr.memb < cap_bytes ? jr.memb : cap_bytes);
        } else if (r.status == PST_RUNNING) { pre_jobs[-(npre++)] = j; }
      }
    } else {
      if (policy == GS_SCHED_GITTINS) {       // stable insertion sort by rank, ascending
        double *rk = reinterpret_cast<double *>(S.queues);     // gittins has no queues: reuse as rank scratch
        for (int i = 0; i < rn; ++i) {
          const int j = runnable[i]; const PJob &r = pj[j];
          rk[i] = git_lookup(S, r.status == PST_RUNNING ? (double)r.exec * jobs[j].gpus : (double)r.exec);
        }
        for (int i = 1; i < rn; ++i) {
          const int x = runnable[i]; const double kx = rk[i]; int k = i;
          while (k > 0 && rk[k - 1] > kx) { runnable[k] = runnable[k - 1]; rk[k] = rk[k - 1]; --k; }
          runnable[k] = x; rk[k] = kx;
        }
      }
      int free_gpu = total_gpus;
      const int nlists = policy == GS_SCHED_GITTINS ? 1 : nq;
      for (int q = 0; q < nlists; ++q) {
        const int *lst = policy == GS_SCHED_GITTINS ? runnable : S.queues + (size_t)q * n;
        const int ln = policy == GS_SCHED_GITTINS ? rn : S.qn[q];
        for (int i = 0; i < ln; ++i) {
          const int j = lst[i];
          const JobIn jr = jobs[j];
          PJob &r = pj[j];
          if (free_gpu >= jr.gpus) {
            if (r.status == PST_PENDING) run_jobs[nrun++] = j;
            free_gpu -= jr.gpus;
            busy += jr.gpus; mem_busy += (long long)jr.gpus * (jr.memb < cap_bytes ? jr.memb : cap_bytes);
          } else if (r.status == PST_RUNNING) { pre_jobs[-(npre++)] = j; }
        }
      }
    }
    for (int i = 0; i < npre; ++i) { pj[pre_jobs[-i]].status = PST_PENDING; ++events; }
    for (int i = 0; i < nrun; ++i) {
      PJob &r = pj[run_jobs[i]];
      r.status = PST_RUNNING; r.resume += 1; ++events;
      if (r.start < 0) r.start = event_time;
    }
    if (is_dlas) {
      for (int q = 0; q < nq; ++q) {
        int *qv = S.queues + (size_t)q * n;
        int w = 0, pn = 0;
        for (int i = 0; i < S.qn[q]; ++i) { const int j = qv[i]; if (pj[j].status == PST_PENDING) tmpl[pn++] = j; else qv[w++] = j; }
        for (int i = 0; i < pn; ++i) qv[w++] = tmpl[i];
      }
    }
    end_time = 0x7fffffff; en = 0;
    next_job_jump = 0x7fffffff;
    int running = 0, queued = 0, pmax = 0;
    long long psum = 0;
    for (int i = 0; i < rn; ++i) {
      const int j = runnable[i];
      const PJob r = pj[j];
      if (r.status != PST_RUNNING) { ++queued; psum += r.pending; pmax = max(pmax, r.pending); continue; }
      ++running;
      const JobIn jr = jobs[j];
      double cl = ceil(jr.dur);
      const int D = cl < 1.0 ? 1 : (int)cl;
      const int e = event_time + (D - r.total_exec);
      if (e < end_time) { end_time = e; en = 0; endj[en++] = j; }
      else if (e == end_time) endj[en++] = j;
      if (is_dlas && r.q_id < nq - 1) {
        const double lim = S.queue_limit[r.q_id];
        const double jt = gputime ? ceil((lim - (double)r.exec) / (double)jr.gpus) + event_time : lim - (double)r.exec + event_time;
        int jti = jt > 2.0e9 ? 0x7fffffff : (int)jt;
        if (jti <= event_time) jti = event_time + 1;
        next_job_jump = min(next_job_jump, jti);
      }
    }
    if (policy == GS_SCHED_GITTINS) next_git += (double)event_time;
    {
      int busy_nodes = 0;
      if (policy == GS_SCHED_SJF) for (int nd = 0; nd < M; ++nd) busy_nodes += (S.cidle[nd] < G);
      int4 *dst = reinterpret_cast<int4 *>(&S.rows[ticks - row_first]);
      dst[0] = make_int4(event_time, M - busy_nodes, busy_nodes, busy);
      dst[1] = make_int4(total_gpus - busy, running, queued, nfin);
      dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(psum & 0xffffffffLL), (int)(psum >> 32));
      dst[3] = make_int4(pmax, 0, 0, 0);
    }
    ticks += 1; budget -= 1;
  }
  if (!done && !((n - p) + rn > 0)) done = true;
  if (!done && p >= n && end_time == 0x7fffffff) done = true;
  if (done) {   // jobs that started but never completed keep their start and restart count
    for (int j = 0; j < n; ++j) { const PJob r = pj[j]; if (r.status != PST_END && r.status != PST_NONE && r.start >= 0) { rec[j].start = r.start; rec[j].preempt = r.resume; } }
  }
  S.p = p; S.rn = rn; S.en = en; S.end_time = end_time; S.next_job_jump = next_job_jump; S.finished = nfin;
  S.next_gittins_unit = next_git; S.events = events; S.ticks = ticks; S.row_first = row_first;
  S.done = done ? 1 : 0; S.status = status; S.running = 0; S.top = 0; S.started = 0;
}

// ------------------------------------------------------------------ event-driven policies, warp cooperative
// dlas / dlas-gpu (MLFQ with GPU counting), one WARP per replica.  Same semantics as
// gs_policy_kernel / oracle/policy_oracle.c, but every O(runnable) loop of an event runs 32
// entries at a time: counter update + END compaction (ballot prefix), demotion list in runnable
// order, greedy admission as a warp prefix sum with skip, RUNNING-before-PENDING stable partition
// of each queue, min-reduction for the next completion / queue jump.
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
  return v;
}

__global__ void __launch_bounds__(32) gs_dlas_warp_kernel(SimDev *sims, int nsims, long long max_ticks) {
  const int sim = blockIdx.x;
  const int lane = threadIdx.x;
  if (sim >= nsims) return;
  SimDev &S = sims[sim];
  const int policy = S.policy;
  if (!(policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU) || S.done || S.status != 0) return;
  const int n = S.n, M = S.M, G = S.G;
  const bool gputime = policy == GS_SCHED_DLAS_GPU;
  const int nq = S.num_queue;
  const JobIn *__restrict__ jobs = S.jobs;
  PJob *pj = S.pj;
  int *runnable = S.runnable, *endj = S.endj, *tmpl = S.tmpl;
  gs_job_rec *rec = S.rec;
  int *fin = S.fin, *queues = S.queues;
  gs_tick_row *rows = S.rows;
  const long long rows_cap = S.rows_cap;
  const long long cap_bytes = S.cap_bytes;
  const int total_gpus = M * G;
  const unsigned lt = (1u << lane) - 1u;
  int p = S.p, rn = S.rn, en = S.en, end_time = S.end_time, next_job_jump = S.next_job_jump, nfin = S.finished;
  int qn[GS_MAX_QUEUES];
  #pragma unroll
  for (int q = 0; q < GS_MAX_QUEUES; ++q) qn[q] = S.qn[q];
  double qlim[GS_MAX_QUEUES];
  #pragma unroll
  for (int q = 0; q < GS_MAX_QUEUES; ++q) qlim[q] = S.queue_limit[q];
  long long events = S.events, ticks = S.ticks;
  const long long row_first = ticks;
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;
  bool done = false;

  while (budget > 0 && (ticks - row_first) < rows_cap) {
    if (!((n - p) + rn > 0)) { done = true; break; }
    if (p >= n && end_time == 0x7fffffff) { done = true; break; }
    const int start_time = p < n ? jobs[p].arrive : 0x7fffffff;
    int event_time; bool has_start = false, has_end = false;
    if (end_time < start_time) { event_time = end_time; has_end = true; }
    else if (end_time > start_time) { event_time = start_time; has_start = true; }
    else { event_time = start_time; has_start = has_end = true; }
    if (event_time > next_job_jump) { event_time = next_job_jump; has_start = has_end = false; }
    // ---- completions (end_jobs is in runnable order)
    if (has_end) {
      for (int i = lane; i < en; i += 32) {
        const int j = endj[i];
        PJob r = pj[j];
        r.status = PST_END;
        pj[j] = r;
        const double dur = jobs[j].dur;
        const double cl = ceil(dur);
        gs_job_rec o; o.start = r.start; o.end = event_time; o.jct = cl < 1.0 ? 1 : (int)cl; o.preempt = r.resume; o.duration = dur;
        rec[j] = o;
        fin[nfin + i] = j;
      }
      nfin += en; events += en;
    }
    // ---- arrivals: appended to runnable and to queue 0 in trace order
    if (has_start) {
      int cnt = 0;
      while (true) {
        const int idx = p + cnt + lane;
        const unsigned b = __ballot_sync(FULL, idx < n && jobs[idx].arrive == event_time);
        const int c = (b == FULL) ? 32 : __ffs(~b) - 1;       // run of arrivals from the front
        cnt += c;
        if (c < 32) break;
      }
      for (int i = lane; i < cnt; i += 32) {
        const int j = p + i;
        PJob r; r.last_check = event_time; r.total_exec = 0; r.exec = 0; r.pending = 0; r.last_pending = 0; r.start = -1;
        r.resume = 0; r.status = PST_PENDING; r.q_id = 0; r.pad0 = 0; r.pad1 = 0;
        pj[j] = r;
        runnable[rn + i] = j;
        queues[qn[0] + i] = j;
      }
      rn += cnt; qn[0] += cnt; events += cnt; p += cnt;
    }
    __syncwarp();
    // ---- pass 1 over runnable: drop END, age counters, detect demotions (kept in runnable order)
    int nd = 0;
    {
      int w = 0;
      for (int base = 0; base < rn; base += 32) {
        const int idx = base + lane;
        const bool valid = idx < rn;
        const int j = valid ? runnable[idx] : 0;
        PJob r;
        if (valid) r = pj[j]; else { r.status = PST_END; r.q_id = 0; r.last_check = 0; r.total_exec = 0; r.exec = 0; r.pending = 0; r.last_pending = 0; r.start = -1; r.resume = 0; }
        const bool keep = valid && r.status != PST_END;
        bool demote = false;
        if (keep) {
          const int dt = event_time - r.last_check;
          r.last_check = event_time;
          if (r.status == PST_RUNNING) {
            r.total_exec += dt; r.exec += dt;
            const double j_gt = gputime ? (double)r.exec * jobs[j].gpus : (double)r.exec;
            if (r.q_id < nq - 1 && j_gt >= qlim[r.q_id]) { demote = true; r.q_id += 1; }
          } else {
            r.pending += dt;
            if (r.exec > 0) r.last_pending += dt;
          }
          pj[j] = r;
        }
        const unsigned kb = __ballot_sync(FULL, keep), db = __ballot_sync(FULL, demote);
        if (keep) runnable[w + __popc(kb & lt)] = j;
        if (demote) tmpl[nd + __popc(db & lt)] = j;
        w += __popc(kb); nd += __popc(db);
      }
      rn = w;
    }
    __syncwarp();
    // ---- queues: drop END / demoted-away entries, then append this event's demotions
    for (int q = 0; q < nq; ++q) {
      int *qv = queues + (size_t)q * n;
      int w = 0;
      for (int base = 0; base < qn[q]; base += 32) {
        const int idx = base + lane;
        const bool valid = idx < qn[q];
        const int j = valid ? qv[idx] : 0;
        bool keep = false;
        if (valid) { const PJob r = pj[j]; keep = r.status != PST_END && r.q_id == q; }
        const unsigned kb = __ballot_sync(FULL, keep);
        if (keep) qv[w + __popc(kb & lt)] = j;
        w += __popc(kb);
      }
      qn[q] = w;
      __syncwarp();
      if (q > 0) {          // jobs demoted into q, in runnable order
        for (int base = 0; base < nd; base += 32) {
          const int idx = base + lane;
          const int j = idx < nd ? tmpl[idx] : 0;
          const bool mine = idx < nd && pj[j].q_id == q;
          const unsigned mb = __ballot_sync(FULL, mine);
          if (mine) qv[qn[q] + __popc(mb & lt)] = j;
          qn[q] += __popc(mb);
        }
      }
      __syncwarp();
    }
    // ---- greedy re-admission on the emptied cluster (GPU counting), queue by queue, and the
    //      RUNNING-before-PENDING stable partition of each queue
    int free_gpu = total_gpus, busy = 0;
    long long mem_busy = 0;
    for (int q = 0; q < nq; ++q) {
      int *qv = queues + (size_t)q * n;
      int w = 0, pn = 0;   // RUNNING entries written so far / PENDING entries parked in tmpl
      for (int base = 0; base < qn[q]; base += 32) {
        const int idx = base + lane;
        const bool valid = idx < qn[q];
        const int j = valid ? qv[idx] : 0;
        PJob r; JobIn jr;
        int g = 0;
        if (valid) { r = pj[j]; jr = jobs[j]; g = jr.gpus; } else { r.status = PST_NONE; r.start = -1; r.resume = 0; jr.memb = 0; }
        // sequential greedy over the 32 entries: admit while the prefix fits, skip the first that does not
        bool admitted = false, decided = !valid;
        while (true) {
          const unsigned ub = __ballot_sync(FULL, !decided);
          if (ub == 0) break;
          if (free_gpu == 0) { decided = true; continue; }
          const int inc = warp_incl_scan(decided ? 0 : g, lane);
          const bool fits = !decided && inc <= free_gpu;
          const unsigned fb = __ballot_sync(FULL, !decided && !fits);     // undecided entries that do not fit
          const int first_fail = fb ? __ffs(fb) - 1 : 32;
          if (!decided && lane < first_fail) { admitted = true; decided = true; }
          if (!decided && lane == first_fail) decided = true;            // rejected
          const int last_ok = first_fail - 1;
          const int used_now = last_ok >= 0 ? __shfl_sync(FULL, inc, last_ok < 0 ? 0 : last_ok) : 0;
          free_gpu -= used_now;
        }
        // status transitions (each one is an event): PENDING->RUNNING = resume, RUNNING->PENDING = preempt
        const bool flip_run = valid && admitted && r.status == PST_PENDING;
        const bool flip_pre = valid && !admitted && r.status == PST_RUNNING;
        if (flip_run) { r.status = PST_RUNNING; r.resume += 1; if (r.start < 0) r.start = event_time; pj[j] = r; }
        if (flip_pre) { r.status = PST_PENDING; pj[j] = r; }
        events += __popc(__ballot_sync(FULL, flip_run)) + __popc(__ballot_sync(FULL, flip_pre));
        busy += __reduce_add_sync(FULL, admitted ? g : 0);
        {
          long long mc = admitted ? (long long)g * (jr.memb < cap_bytes ? jr.memb : cap_bytes) : 0;
          #pragma unroll
          for (int o = 16; o > 0; o >>= 1) mc += __shfl_xor_sync(FULL, mc, o);
          mem_busy += mc;
        }
        // stable partition: RUNNING entries stay in place order, PENDING go behind
        const bool is_run = valid && admitted;
        const bool is_pen = valid && !admitted;
        const unsigned rb = __ballot_sync(FULL, is_run), pb = __ballot_sync(FULL, is_pen);
        if (is_run) qv[w + __popc(rb & lt)] = j;
        if (is_pen) tmpl[pn + __popc(pb & lt)] = j;
        w += __popc(rb); pn += __popc(pb);
      }
      __syncwarp();
      for (int i = lane; i < pn; i += 32) qv[w + i] = tmpl[i];
      __syncwarp();
    }
    // ---- final pass over runnable: transitions are counted, next completion / jump, statistics
    end_time = 0x7fffffff; en = 0; next_job_jump = 0x7fffffff;
    int running = 0, queued = 0, pmax = 0;
    long long psum = 0;
    for (int base = 0; base < rn; base += 32) {
      const int idx = base + lane;
      const bool valid = idx < rn;
      const int j = valid ? runnable[idx] : 0;
      int e = 0x7fffffff, jt = 0x7fffffff, pend = 0;
      bool isrun = false;
      if (valid) {
        const PJob r = pj[j];
        isrun = r.status == PST_RUNNING;
        if (isrun) {
          const JobIn jr = jobs[j];
          const double cl = ceil(jr.dur);
          const int D = cl < 1.0 ? 1 : (int)cl;
          e = event_time + (D - r.total_exec);
          if (r.q_id < nq - 1) {
            const double lim = qlim[r.q_id];
            const double t = gputime ? ceil((lim - (double)r.exec) / (double)jr.gpus) + event_time : lim - (double)r.exec + event_time;
            jt = t > 2.0e9 ? 0x7fffffff : (int)t;
            if (jt <= event_time) jt = event_time + 1;
          }
        } else pend = r.pending;
      }
      const int cmin = __reduce_min_sync(FULL, e);
      if (cmin < end_time) { end_time = cmin; en = 0; }
      const unsigned eb = __ballot_sync(FULL, valid && isrun && e == end_time);
      if (valid && isrun && e == end_time) endj[en + __popc(eb & lt)] = j;
      en += __popc(eb);
      next_job_jump = min(next_job_jump, __reduce_min_sync(FULL, jt));
      running += __popc(__ballot_sync(FULL, valid && isrun));
      queued += __popc(__ballot_sync(FULL, valid && !isrun));
      pmax = max(pmax, __reduce_max_sync(FULL, pend));
      psum += (long long)__reduce_add_sync(FULL, pend);
    }
    __syncwarp();
    if (lane == 0) {
      int4 *dst = reinterpret_cast<int4 *>(&rows[ticks - row_first]);
      dst[0] = make_int4(event_time, M, 0, busy);
      dst[1] = make_int4(total_gpus - busy, running, queued, nfin);
      dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(psum & 0xffffffffLL), (int)(psum >> 32));
      dst[3] = make_int4(pmax, 0, 0, 0);
    }
    ticks += 1; budget -= 1;
  }
  if (!done && !((n - p) + rn > 0)) done = true;
  if (!done && p >= n && end_time == 0x7fffffff) done = true;
  __syncwarp();
  if (done) {
    for (int j = lane; j < n; j += 32) { const PJob r = pj[j]; if (r.status != PST_END && r.status != PST_NONE && r.start >= 0) { rec[j].start = r.start; rec[j].preempt = r.resume; } }
  }
  if (lane == 0) {
    S.p = p; S.rn = rn; S.en = en; S.end_time = end_time; S.next_job_jump = next_job_jump; S.finished = nfin;
    #pragma unroll
    for (int q = 0; q < GS_MAX_QUEUES; ++q) S.qn[q] = qn[q];
    S.events = events; S.ticks = ticks; S.row_first = row_first;
    S.done = done ? 1 : 0; S.running = 0; S.top = 0; S.started = 0;
  }
}

// sjf (stable order by num_gpu + live-yarn placement on the emptied cluster) and gittins (stable
// order by gittins rank + GPU counting), one WARP per replica.  Same semantics as
// gs_policy_kernel / oracle/policy_oracle.c.  The runnable list stays sorted between events for
// sjf (keys never change), so new arrivals are INSERTED (count of keys <= k, warp-parallel shift);
// gittins ranks move a little every event, so the list is repaired with stable odd-even
// transposition rounds (adjacent swaps only when strictly greater == the unique stable order).
__global__ void __launch_bounds__(32) gs_sortpol_warp_kernel(SimDev *sims, int nsims, long long max_ticks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int sim = blockIdx.x;
  const int lane = threadIdx.x;
  if (sim >= nsims) return;
  SimDev &S = sims[sim];
  const int policy = S.policy;
  if (!(policy == GS_SCHED_SJF || policy == GS_SCHED_GITTINS) || S.done || S.status != 0) return;
  const bool sjf = policy == GS_SCHED_SJF;
  const int n = S.n, M = S.M, G = S.G, K = S.K;
  int *nidle = reinterpret_cast<int *>(smem_raw);        // sjf: (idle devices, free slots) per node
  int *nkfree = nidle + M;
  const JobIn *__restrict__ jobs = S.jobs;
  PJob *pj = S.pj;
  int *runnable = S.runnable, *endj = S.endj;
  double *rk = reinterpret_cast<double *>(S.queues);     // gittins: rank of runnable[i] (no queues in these policies)
  gs_job_rec *rec = S.rec;
  int *fin = S.fin;
  gs_tick_row *rows = S.rows;
  const long long rows_cap = S.rows_cap;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int total_gpus = M * G;
  const unsigned lt = (1u << lane) - 1u;
  int p = S.p, rn = S.rn, en = S.en, end_time = S.end_time, nfin = S.finished;
  double next_git = S.next_gittins_unit;
  long long events = S.events, ticks = S.ticks;
  const long long row_first = ticks;
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;
  bool done = false;

  while (budget > 0 && (ticks - row_first) < rows_cap) {
    if (!((n - p) + rn > 0)) { done = true; break; }
    if (p >= n && end_time == 0x7fffffff) { done = true; break; }
    const int start_time = p < n ? jobs[p].arrive : 0x7fffffff;
    int event_time; bool has_start = false, has_end = false;
    if (end_time < start_time) { event_time = end_time; has_end = true; }
    else if (end_time > start_time) { event_time = start_time; has_start = true; }
    else { event_time = start_time; has_start = has_end = true; }
    if (!sjf && (double)event_time > next_git) { event_time = (int)next_git; has_start = has_end = false; }
    // ---- completions
    if (has_end) {
      for (int i = lane; i < en; i += 32) {
        const int j = endj[i];
        PJob r = pj[j];
        r.status = PST_END;
        pj[j] = r;
        const double dur = jobs[j].dur;
        const double cl = ceil(dur);
        gs_job_rec o; o.start = r.start; o.end = event_time; o.jct = cl < 1.0 ? 1 : (int)cl; o.preempt = r.resume; o.duration = dur;
        rec[j] = o;
        fin[nfin + i] = j;
      }
      nfin += en; events += en;
    }
    __syncwarp();
    // ---- pass 1: drop END, age counters, (gittins) rank of every survivor at its new position
    {
      int w = 0;
      for (int base = 0; base < rn; base += 32) {
        const int idx = base + lane;
        const bool valid = idx < rn;
        const int j = valid ? runnable[idx] : 0;
        PJob r;
        r.status = PST_END; r.q_id = 0; r.last_check = 0; r.total_exec = 0; r.exec = 0; r.pending = 0; r.last_pending = 0; r.start = -1; r.resume = 0;
        if (valid) r = pj[j];
        const bool keep = valid && r.status != PST_END;
        double rank = 0.0;
        if (keep) {
          const int dt = event_time - r.last_check;
          r.last_check = event_time;
          if (r.status == PST_RUNNING) { r.total_exec += dt; r.exec += dt; }
          else { r.pending += dt; if (r.exec > 0) r.last_pending += dt; }
          pj[j] = r;
          if (!sjf) rank = git_lookup(S, r.status == PST_RUNNING ? (double)r.exec * jobs[j].gpus : (double)r.exec);
        }
        const unsigned kb = __ballot_sync(FULL, keep);
        if (keep) { const int pos = w + __popc(kb & lt); runnable[pos] = j; if (!sjf) rk[pos] = rank; }
        w += __popc(kb);
      }
      rn = w;
    }
    __syncwarp();
    // ---- arrivals (after the survivors, like the list append of the specification)
    int cnt = 0;
    if (has_start) {
      while (true) {
        const int idx = p + cnt + lane;
        const unsigned b = __ballot_sync(FULL, idx < n && jobs[idx].arrive == event_time);
        const int c = (b == FULL) ? 32 : __ffs(~b) - 1;
        cnt += c;
        if (c < 32) break;
      }
      for (int i = lane; i < cnt; i += 32) {
        const int j = p + i;
        PJob r; r.last_check = event_time; r.total_exec = 0; r.exec = 0; r.pending = 0; r.last_pending = 0; r.start = -1;
        r.resume = 0; r.status = PST_PENDING; r.q_id = 0; r.pad0 = 0; r.pad1 = 0;
        pj[j] = r;
      }
      events += cnt;
      if (!sjf) {
        const double r0 = git_lookup(S, 0.0);              // a new job: executed_time == 0
        for (int i = lane; i < cnt; i += 32) { runnable[rn + i] = p + i; rk[rn + i] = r0; }
        rn += cnt;
      }
    }
    __syncwarp();
    if (sjf) {
      // stable insertion of each new job: position = number of runnable entries with num_gpu <= its own
      for (int i = 0; i < cnt; ++i) {
        const int j = p + i;
        const int kx = jobs[j].gpus;
        int pos = 0;
        for (int base = 0; base < rn; base += 32) {
          const int idx = base + lane;
          const bool le = idx < rn && jobs[runnable[idx]].gpus <= kx;
          pos += __popc(__ballot_sync(FULL, le));
        }
        for (int hi = rn; hi > pos; hi -= 32) {              // shift [pos, rn) right by one, from the tail
          const int idx = hi - 1 - lane;
          const int v = idx >= pos ? runnable[idx] : 0;
          __syncwarp();
          if (idx >= pos) runnable[idx + 1] = v;
          __syncwarp();
        }
        if (lane == 0) runnable[pos] = j;
        rn += 1;
        __syncwarp();
      }
    } else {
      // stable odd-even transposition until a full round makes no swap
      bool again = rn > 1;
      while (again) {
        unsigned any = 0;
        for (int phase = 0; phase < 2; ++phase) {
          for (int base = phase; base + 1 < rn; base += 64) {
            const int a = base + 2 * lane;
            bool sw = false;
            if (a + 1 < rn) {
              const double ka = rk[a], kb2 = rk[a + 1];
              if (ka > kb2) { const int ja = runnable[a], jb = runnable[a + 1]; runnable[a] = jb; runnable[a + 1] = ja; rk[a] = kb2; rk[a + 1] = ka; sw = true; }
            }
            any |= __ballot_sync(FULL, sw);
          }
          __syncwarp();
        }
        again = any != 0;
      }
    }
    p += cnt;
    __syncwarp();
    // ---- greedy re-admission on the emptied cluster, in list order
    int busy = 0;
    long long mem_busy = 0;
    if (sjf) {
      for (int nd = lane; nd < M; nd += 32) { nidle[nd] = G; nkfree[nd] = K; }
      __syncwarp();
      for (int i = 0; i < rn; ++i) {
        const int j = runnable[i];
        const JobIn jr = jobs[j];
        PJob r = pj[j];
        const int hg = jr.gpus, hc = jr.gpc, tasks = hc == 1 ? hg : hg / hc;
        bool ok = false;
        if (jr.memb < fit_limit) {
          if (hg <= G) {
            int found = -1;
            for (int base = 0; base < M && found < 0; base += 32) {
              const int nd = base + lane;
              const bool fit = nd < M && nidle[nd] >= hg && nkfree[nd] >= tasks;
              const unsigned b = __ballot_sync(FULL, fit);
              if (b) found = base + __ffs(b) - 1;
            }
            if (found >= 0) { ok = true; if (lane == 0) { nidle[found] -= hg; nkfree[found] -= tasks; } }
          } else {
            int cum = 0, last_base = -1;
            for (int base = 0; base < M; base += 32) {
              const int nd = base + lane;
              const int c = nd < M ? max(min(nidle[nd] / hc, nkfree[nd]), 0) : 0;
              cum += __reduce_add_sync(FULL, c);
              if (cum >= tasks) { last_base = base; break; }
            }
            if (last_base >= 0) {
              ok = true;
              int rem = tasks;
              for (int base = 0; base <= last_base; base += 32) {
                const int nd = base + lane;
                const int c = nd < M ? max(min(nidle[nd] / hc, nkfree[nd]), 0) : 0;
                const int incl = warp_incl_scan(c, lane);
                const int take = min(c, max(rem - (incl - c), 0));
                if (take > 0) { nidle[nd] -= take * hc; nkfree[nd] -= take; }
                rem -= min(rem, __shfl_sync(FULL, incl, 31));
              }
            }
          }
          __syncwarp();
        }
        if (ok) {
          busy += hg;
          mem_busy += (long long)hg * (jr.memb < cap_bytes ? jr.memb : cap_bytes);
          if (r.status == PST_PENDING) {
            r.status = PST_RUNNING; r.resume += 1; if (r.start < 0) r.start = event_time;
            if (lane == 0) pj[j] = r;
            events += 1;
          } else if (r.start < 0) { r.start = event_time; if (lane == 0) pj[j] = r; }
        } else if (r.status == PST_RUNNING) {
          r.status = PST_PENDING;
          if (lane == 0) pj[j] = r;
          events += 1;
        }
        __syncwarp();
      }
    } else {
      int free_gpu = total_gpus;
      for (int base = 0; base < rn; base += 32) {
        const int idx = base + lane;
        const bool valid = idx < rn;
        const int j = valid ? runnable[idx] : 0;
        PJob r; JobIn jr;
        r.status = PST_NONE; r.start = -1; r.resume = 0; jr.memb = 0; jr.gpus = 0;
        int g = 0;
        if (valid) { r = pj[j]; jr = jobs[j]; g = jr.gpus; }
        bool admitted = false, decided = !valid;
        while (true) {
          const unsigned ub = __ballot_sync(FULL, !decided);
          if (ub == 0) break;
          if (free_gpu == 0) { decided = true; continue; }
          const int inc = warp_incl_scan(decided ? 0 : g, lane);
          const bool fits = !decided && inc <= free_gpu;
          const unsigned fb = __ballot_sync(FULL, !decided && !fits);
          const int first_fail = fb ? __ffs(fb) - 1 : 32;
          if (!decided && lane < first_fail) { admitted = true; decided = true; }
          if (!decided && lane == first_fail) decided = true;
          const int used_now = first_fail > 0 ? __shfl_sync(FULL, inc, first_fail - 1) : 0;
          free_gpu -= used_now;
        }
        const bool flip_run = valid && admitted && r.status == PST_PENDING;
        const bool flip_pre = valid && !admitted && r.status == PST_RUNNING;
        if (flip_run) { r.status = PST_RUNNING; r.resume += 1; if (r.start < 0) r.start = event_time; pj[j] = r; }
        if (flip_pre) { r.status = PST_PENDING; pj[j] = r; }
        events += __popc(__ballot_sync(FULL, flip_run)) + __popc(__ballot_sync(FULL, flip_pre));
        busy += __reduce_add_sync(FULL, admitted ? g : 0);
        long long mc = admitted ? (long long)g * (jr.memb < cap_bytes ? jr.memb : cap_bytes) : 0;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mc += __shfl_xor_sync(FULL, mc, o);
        mem_busy += mc;
      }
    }
    __syncwarp();
    // ---- final pass: next completion (ties in list order) and statistics
    end_time = 0x7fffffff; en = 0;
    int running = 0, queued = 0, pmax = 0;
    long long psum = 0;
    for (int base = 0; base < rn; base += 32) {
      const int idx = base + lane;
      const bool valid = idx < rn;
      const int j = valid ? runnable[idx] : 0;
      int e = 0x7fffffff, pend = 0;
      bool isrun = false;
      if (valid) {
        const PJob r = pj[j];
        isrun = r.status == PST_RUNNING;
        if (isrun) {
          const double cl = ceil(jobs[j].dur);
          const int D = cl < 1.0 ? 1 : (int)cl;
          e = event_time + (D - r.total_exec);
        } else pend = r.pending;
      }
      const int cmin = __reduce_min_sync(FULL, e);
      if (cmin < end_time) { end_time = cmin; en = 0; }
      const unsigned eb = __ballot_sync(FULL, valid && isrun && e == end_time);
      if (valid && isrun && e == end_time) endj[en + __popc(eb & lt)] = j;
      en += __popc(eb);
      running += __popc(__ballot_sync(FULL, valid && isrun));
      queued += __popc(__ballot_sync(FULL, valid && !isrun));
      pmax = max(pmax, __reduce_max_sync(FULL, pend));
      psum += (long long)__reduce_add_sync(FULL, pend);
    }
    if (!sjf) next_git += (double)event_time;
    int busy_nodes = 0;
    if (sjf) for (int base = 0; base < M; base += 32) { const int nd = base + lane; busy_nodes += __popc(__ballot_sync(FULL, nd < M && nidle[nd] < G)); }
    __syncwarp();
    if (lane == 0) {
      int4 *dst = reinterpret_cast<int4 *>(&rows[ticks - row_first]);
      dst[0] = make_int4(event_time, M - busy_nodes, busy_nodes, busy);
      dst[1] = make_int4(total_gpus - busy, running, queued, nfin);
      dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(psum & 0xffffffffLL), (int)(psum >> 32));
      dst[3] = make_int4(pmax, 0, 0, 0);
    }
    ticks += 1; budget -= 1;
  }
  if (!done && !((n - p) + rn > 0)) done = true;
  if (!done && p >= n && end_time == 0x7fffffff) done = true;
  __syncwarp();
  if (done) {
    for (int j = lane; j < n; j += 32) { const PJob r = pj[j]; if (r.status != PST_END && r.status != PST_NONE && r.start >= 0) { rec[j].start = r.start; rec[j].preempt = r.resume; } }
  }
  if (lane == 0) {
    S.p = p; S.rn = rn; S.en = en; S.end_time = end_time; S.finished = nfin; S.next_gittins_unit = next_git;
    S.events = events; S.ticks = ticks; S.row_first = row_first;
    S.done = done ? 1 : 0; S.running = 0; S.top = 0; S.started = 0;
  }
}

// One launch (re)initialises every replica flagged need_init: job records (never-started jobs
// report start=end=-1, jct=preempt=0 and their input duration), empty wheel, idle node table.
__global__ void gs_init_kernel(SimDev *sims, int nsims) {
  const int sim = blockIdx.y;
  if (sim >= nsims) return;
  const SimDev &S = sims[sim];
  if (!S.need_init) return;
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < S.n; i += stride) {
    gs_job_rec r; r.start = -1; r.end = -1; r.jct = 0; r.preempt = 0; r.duration = S.jobs[i].dur;
    S.rec[i] = r;
  }
  for (int i = t0; i <= S.wheel_mask; i += stride) { S.wheel_head[i] = -1; S.wheel_tail[i] = -1; }
  for (int i = t0; i < S.M; i += stride) { S.nbusy[i] = 0ull; S.nk[i] = 0; }
  for (int i = t0; i < S.n; i += stride) S.sref[i] = make_int2(0, 0);
  if (S.policy != GS_SCHED_FIFO)
    for (int i = t0; i < S.n; i += stride) { PJob z; memset(&z, 0, sizeof(z)); z.start = -1; S.pj[i] = z; }
}

// ------------------------------------------------------------------ result regrouping
// Spans are pooled in START order while the simulation runs; callers want them grouped by
// job (CSR).  One block scans the per-job span counts, a second kernel gathers.
__global__ void __launch_bounds__(1024) gs_span_scan_kernel(const gs_job_rec *__restrict__ rec, const int2 *__restrict__ sref,
                                                            int n, long long *__restrict__ off) {
  // one block: every thread sums a contiguous chunk, the block scans the 1024 chunk sums,
  // every thread rewrites its chunk as an exclusive prefix
  __shared__ long long warp_sum[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int chunk = (n + 1023) / 1024;
  const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
  long long v = 0;
  for (int j = lo; j < hi; ++j) v += (rec[j].start >= 0) ? sref[j].y : 0;
  long long incl = v;
  #pragma unroll
  for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_sum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const long long w = warp_sum[lane];
    long long wi = w;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += t; }
    warp_sum[lane] = wi - w;
  }
  __syncthreads();
  long long run = warp_sum[wid] + incl - v;
  for (int j = lo; j < hi; ++j) { off[j] = run; run += (rec[j].start >= 0) ? sref[j].y : 0; }
  if (tid == 1023) off[n] = run;
}

__global__ void gs_span_gather_kernel(const gs_job_rec *__restrict__ rec, const int2 *__restrict__ sref,
                                      const gs_span *__restrict__ pool, const long long *__restrict__ off, int n,
                                      gs_span *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || rec[j].start < 0) return;
  const int2 sr = sref[j];
  const long long o = off[j];
  for (int i = 0; i < sr.y; ++i) out[o + i] = pool[sr.x + i];
}

// ------------------------------------------------------------------ stateless candidate scoring
// gs_place_batch: b independent jobs scored against ONE cluster state (nothing is modified).
// The block first turns the node table (one 16-byte load per node) into a small capacity
// index in shared memory:
//    cap[nd]  = min(idle devices, free task slots)          tasks of a 1-GPU-per-task job the node can take
//    ff[t]    = first node with cap >= t                    -> single-node first fit is ONE look-up
//    P[nd]    = inclusive prefix sum of cap,  Q[nd] = inclusive count of nodes with cap > 0
//                                                           -> cross-node fill is a binary search on P
// and then streams the job requests through it, ONE THREAD PER JOB: 16 bytes in, 8 bytes out,
// a handful of instructions -- the kernel is bound by HBM bandwidth, not by the node scan.
// Jobs with gpu_per_task != 1 (or a requested per-task node list) take the general per-node walk.
// general walk over the (idle, slots) table: any gpu_per_task, optional per-task node list
__device__ void place_general(const short2 *tab, int M, int G, int gpus, int gpc, int *tn, int &fn, int &used) {
  const int tasks = gpus / gpc;
  if (gpus <= G) {
    for (int nd = 0; nd < M; ++nd) {
      const short2 t = tab[nd];
      if (t.x >= gpus && t.y >= tasks) { fn = nd; used = 1; break; }
    }
    if (fn >= 0 && tn) for (int t = 0; t < tasks; ++t) tn[t] = fn;
    return;
  }
  int cum = 0, last = -1;
  for (int nd = 0; nd < M; ++nd) {
    const short2 t = tab[nd];
    cum += max(min((int)t.x / gpc, (int)t.y), 0);
    if (cum >= tasks) { last = nd; break; }
  }
  if (last < 0) return;
  int done_tasks = 0;
  for (int nd = 0; nd <= last; ++nd) {
    const short2 t = tab[nd];
    const int c = max(min((int)t.x / gpc, (int)t.y), 0);
    const int take = min(c, tasks - done_tasks);
    if (take > 0) {
      if (fn < 0) fn = nd;
      ++used;
      if (tn) for (int q = 0; q < take; ++q) tn[done_tasks + q] = nd;
      done_tasks += take;
    }
  }
}

__global__ void __launch_bounds__(256) gs_place_kernel(const uint4 *__restrict__ nodes, int M, int G, int cpu_cnt,
                                                       int mem_sz, int cpu_pt, int mem_pt, long long fit_limit,
                                                       const uint4 *__restrict__ jobs, long long b,
                                                       int *__restrict__ first_node, int *__restrict__ nodes_used,
                                                       const long long *__restrict__ task_off, int *__restrict__ task_node) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  short2 *tab = reinterpret_cast<short2 *>(smem_raw);          // (idle, slots) per node
  int *P = reinterpret_cast<int *>(tab + M);                   // prefix of cap (gpc == 1)
  int *Q = P + M;                                              // prefix count of cap > 0
  int *ff = Q + M;                                             // [GS_MAX_GPUS_PER_NODE + 1]
  __shared__ int first_pos_s;
  for (int i = threadIdx.x; i <= GS_MAX_GPUS_PER_NODE; i += blockDim.x) ff[i] = 0x7fffffff;
  if (threadIdx.x == 0) first_pos_s = 0x7fffffff;
  __syncthreads();
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const uint4 v = nodes[i];                                  // {busy_lo, busy_hi, cpu_used, mem_used}
    unsigned long long bm = ((unsigned long long)v.y << 32) | v.x;
    if (G < 64) bm &= (1ull << G) - 1ull;
    const int idle = G - __popcll(bm);
    const int cf = cpu_cnt - (int)v.z, mf = mem_sz - (int)v.w;
    const int slots = min(min(cf > 0 ? cf / cpu_pt : 0, mf > 0 ? mf / mem_pt : 0), 32767);
    tab[i] = make_short2((short)idle, (short)slots);
    const int cap = min(idle, slots);
    P[i] = cap;
    Q[i] = cap > 0 ? 1 : 0;
    for (int t = 1; t <= cap; ++t) atomicMin(&ff[t], i);
    if (cap > 0) atomicMin(&first_pos_s, i);
  }
  __syncthreads();
  if (threadIdx.x < 32) {                                      // warp 0: inclusive scans of P and Q
    const int lane = threadIdx.x;
    int cp = 0, cq = 0;
    for (int base = 0; base < M; base += 32) {
      const int i = base + lane;
      int vp = i < M ? P[i] : 0, vq = i < M ? Q[i] : 0;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tp = __shfl_up_sync(FULL, vp, o), tq = __shfl_up_sync(FULL, vq, o);
        if (lane >= o) { vp += tp; vq += tq; }
      }
      if (i < M) { P[i] = cp + vp; Q[i] = cq + vq; }
      cp += __shfl_sync(FULL, vp, 31); cq += __shfl_sync(FULL, vq, 31);
    }
  }
  __syncthreads();
  const int first_pos = first_pos_s;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // fast path: four requests per thread per iteration, all four 16-byte loads in flight together
  long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (task_node == nullptr) {
    for (; j0 + 3 * stride < b; j0 += 4 * stride) {
      uint4 jr[4];
      #pragma unroll
      for (int u = 0; u < 4; ++u) jr[u] = __ldcs(&jobs[j0 + u * stride]);
      #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gpus = (int)jr[u].x, gpc = (int)jr[u].y;
        const long long memb = (long long)(((unsigned long long)jr[u].w << 32) | jr[u].z);
        int fn = -1, used = 0;
        if (memb < fit_limit) {
          if (gpc == 1) {
            if (gpus <= G) {
              const int f = ff[gpus];
              if (f != 0x7fffffff) { fn = f; used = 1; }
            } else if (P[M - 1] >= gpus) {
              int lo = 0, hi = M - 1;
              while (lo < hi) { const int mid = (lo + hi) >> 1; if (P[mid] >= gpus) hi = mid; else lo = mid + 1; }
              fn = first_pos; used = Q[lo];
            }
          } else {
            place_general(tab, M, G, gpus, gpc, nullptr, fn, used);
          }
        }
        __stcs(&first_node[j0 + u * stride], fn);
        if (nodes_used) __stcs(&nodes_used[j0 + u * stride], used);
      }
    }
  }
  for (long long j = j0; j < b; j += stride) {
    const uint4 jr = jobs[j];                                  // {gpus, gpc, mem_lo, mem_hi}
    const int gpus = (int)jr.x, gpc = (int)jr.y;
    const long long memb = (long long)(((unsigned long long)jr.w << 32) | jr.z);
    const int tasks = gpc == 1 ? gpus : gpus / gpc;
    int *tn = task_node ? task_node + task_off[j] : nullptr;
    int fn = -1, used = 0;
    if (memb < fit_limit) {
      if (gpc == 1 && tn == nullptr) {
        if (gpus <= G) {
          const int f = ff[gpus];
          if (f != 0x7fffffff) { fn = f; used = 1; }
        } else if (P[M - 1] >= tasks) {
          int lo = 0, hi = M - 1;                              // smallest nd with P[nd] >= tasks
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (P[mid] >= tasks) hi = mid; else lo = mid + 1; }
          fn = first_pos; used = Q[lo];
        }
      } else {
        place_general(tab, M, G, gpus, gpc, tn, fn, used);
      }
    }
    if (fn < 0 && tn) for (int t = 0; t < tasks; ++t) tn[t] = -1;
    first_node[j] = fn;
    if (nodes_used) nodes_used[j] = used;
  }
}

// gs_net_cost: one warp per job.  cross = |ps_nodes symmetric-difference wk_nodes|
// (network_service.py:16-24); extra = (model/bw + cross*lat) * (iters*2.0) with the
// reference's association and no FMA contraction (:34-37).
__global__ void gs_netcost_kernel(long long b, const long long *__restrict__ task_off,
                                  const int *__restrict__ task_node, const unsigned char *__restrict__ is_ps,
                                  const int *__restrict__ ps_count, const double *__restrict__ model_mb,
                                  const double *__restrict__ iters, double bandwidth, double latency,
                                  double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long j = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < b; j += warps) {
    const long long a = task_off[j], e = task_off[j + 1];
    int cross = 0;
    if (ps_count[j] > 1) {
      for (long long t = a + lane; t < e; t += 32) {
        int nd = task_node[t];
        bool first = true;
        for (long long u = a; u < t && first; ++u) first = task_node[u] != nd;
        if (!first) continue;
        bool in_ps = false, in_wk = false;
        for (long long u = a; u < e; ++u)
          if (task_node[u] == nd) { if (is_ps && is_ps[u]) in_ps = true; else in_wk = true; }
        cross += (in_ps != in_wk);
      }
      cross = __reduce_add_sync(FULL, cross);
    }
    if (lane == 0) {
      double extra = 0.0;
      if (cross > 0) {
        double mps = __ddiv_rn(model_mb[j], bandwidth);
        double nis = __dmul_rn((double)cross, latency);
        double rt = __dmul_rn(iters[j], 2.0);
        extra = __dmul_rn(__dadd_rn(mps, nis), rt);
      }
      out[j] = extra;
    }
  }
}

// ------------------------------------------------------------------ host side

struct SimHost {
  gs_cluster cl;
  gs_policy pol;
  bool configured = false, loaded = false, prepared = false;
  int64_t n = 0;
  void *trace_slab = nullptr;
  void *state_slab = nullptr;
  void *git_dev = nullptr;
  size_t trace_bytes = 0, state_bytes = 0;
  int64_t span_cap = 0, rows_cap = 0, last_arrive = 0;
  int max_need = 1;
  SimDev dev;
};

struct gs_engine {
  int device = 0, nsims = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void *d_scratch2 = nullptr;
  size_t d_scratch2_bytes = 0;
  std::vector<SimHost> sims;
  SimDev *d_sims = nullptr;
  void *h_stage = nullptr;
  size_t h_stage_bytes = 0;
  void *d_scratch = nullptr;
  size_t d_scratch_bytes = 0;
  std::string err;
  double kernel_ms = 0, h2d_ms = 0, d2h_ms = 0;
  long long launches = 0;  // kernels launched by this handle
  int engine_mode = 0;     // 0 auto, 1 warp-per-replica, 2 lane-per-replica, 3 half-warp-per-replica
  double span_budget = 0;  // > 0: span pool = min(worst case, budget * n + 4096) records per replica
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

extern "C" int gs_abi_version(void) { return GS_ABI_VERSION; }

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
  cudaError_t e1 = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  cudaError_t e1b = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
  (void)e1b;
  cudaError_t e2 = cudaEventCreate(&h->e0);
  cudaError_t e3 = cudaEventCreate(&h->e1);
  cudaError_t e4 = cudaMalloc(&h->d_sims, sizeof(SimDev) * (size_t)nsims);
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
  for (auto &s : h->sims) { if (s.trace_slab) cudaFree(s.trace_slab); if (s.state_slab) cudaFree(s.state_slab); if (s.git_dev) cudaFree(s.git_dev); }
  if (h->d_sims) cudaFree(h->d_sims);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->d_scratch) cudaFree(h->d_scratch);
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->stream2) cudaStreamDestroy(h->stream2);
  if (h->d_scratch2) cudaFree(h->d_scratch2);
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
  s.cl = *cluster;
  if (policy) s.pol = *policy; else { memset(&s.pol, 0, sizeof(s.pol)); s.pol.num_queue = 1; }
  if (s.pol.schedule < GS_SCHED_FIFO || s.pol.schedule > GS_SCHED_GITTINS)
    return fail(h, GS_ERR_ARG, "gs_config_sim: unknown schedule");
  if (s.pol.scheme != GS_SCHEME_YARN && s.pol.scheme != GS_SCHEME_COUNT)
    return fail(h, GS_ERR_ARG, "gs_config_sim: unknown scheme");
  if (s.pol.schedule == GS_SCHED_FIFO && s.pol.scheme != GS_SCHEME_YARN)
    return fail(h, GS_ERR_ARG, "gs_config_sim: fifo runs with the yarn scheme only");
  if ((s.pol.schedule == GS_SCHED_DLAS || s.pol.schedule == GS_SCHED_DLAS_GPU) &&
      (s.pol.num_queue < 1 || s.pol.num_queue > GS_MAX_QUEUES))
    return fail(h, GS_ERR_ARG, "gs_config_sim: num_queue must be in 1..8 for dlas");
  if (s.git_dev) { cudaFree(s.git_dev); s.git_dev = nullptr; }
  if (s.pol.schedule == GS_SCHED_GITTINS) {
    if (s.pol.gittins_n < 1 || !s.pol.gittins_data || !s.pol.gittins_index)
      return fail(h, GS_ERR_ARG, "gs_config_sim: gittins needs the (data, index) tables");
    CU(cudaSetDevice(h->device));
    const size_t bytes = 8 * (size_t)s.pol.gittins_n;
    CU(cudaMalloc(&s.git_dev, 2 * bytes));
    CU(cudaMemcpy(s.git_dev, s.pol.gittins_data, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy((unsigned char *)s.git_dev + bytes, s.pol.gittins_index, bytes, cudaMemcpyHostToDevice));
  }
  s.configured = true;
  return GS_OK;
}

static int ensure_stage(gs_handle h, size_t bytes) {
  if (h->h_stage_bytes >= bytes) return GS_OK;
  if (h->h_stage) cudaFreeHost(h->h_stage);
  h->h_stage = nullptr; h->h_stage_bytes = 0;
  CU(cudaMallocHost(&h->h_stage, bytes));
  h->h_stage_bytes = bytes;
  return GS_OK;
}

// Shared tail of the two loaders: `ji` already holds n validated JobIn records in the pinned
// staging buffer (plus the optional network columns); bounds are known.
static int finish_load(gs_handle h, SimHost &s, int64_t n, bool net, size_t off_model, size_t off_iters, size_t total,
                       int64_t span_cap, double max_need, int64_t last_arrive) {
  if (max_need > (double)(1 << 26)) return fail(h, GS_ERR_ARG, "gs_load_trace: job duration exceeds 2^26 ticks");
  unsigned char *st = (unsigned char *)h->h_stage;
  if (s.trace_slab && s.trace_bytes < total) { cudaFree(s.trace_slab); s.trace_slab = nullptr; }
  if (!s.trace_slab) { CU(cudaMalloc(&s.trace_slab, total)); s.trace_bytes = total; }
  CU(cudaEventRecord(h->e0, h->stream));
  CU(cudaMemcpyAsync(s.trace_slab, st, total, cudaMemcpyHostToDevice, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->h2d_ms += ms;
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
  s.last_arrive = last_arrive;
  s.loaded = true;
  s.prepared = false;      // (re)loading a trace restarts the replica; slabs are reused when big enough
  h->dirty = true;
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
  const int M = s.cl.num_switch * s.cl.num_node_p_switch;
  CU(cudaSetDevice(h->device));
  const size_t N = (size_t)(n > 0 ? n : 1);
  const size_t off_model = align_up(sizeof(JobIn) * N), off_iters = align_up(off_model + (net ? 8 * N : 0));
  const size_t total = align_up(off_iters + (net ? 8 * N : 0));
  int rc = ensure_stage(h, total);
  if (rc) return rc;
  unsigned char *st = (unsigned char *)h->h_stage;
  JobIn *ji = (JobIn *)st;
  if (packed) memcpy(ji, packed, sizeof(JobIn) * (size_t)n);
  // one pass: (pack,) validate, bounds
  int64_t span_cap = 0;
  double max_need = 1.0;
  int prev = 0;
  const bool netcost = net && s.cl.enable_network_costs;
  for (int64_t j = 0; j < n; ++j) {
    JobIn r;
    if (packed) r = ji[j];
    else {
      r.arrive = arrive_tick[j]; r.gpus = gpus[j]; r.gpc = gpu_per_task[j]; r.ps = net ? ps_count[j] : 0;
      r.memb = mem_bytes[j]; r.dur = duration[j];
      ji[j] = r;
    }
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
  if (net && n > 0) {
    memcpy(st + off_model, model_mb, 8 * (size_t)n);
    memcpy(st + off_iters, iterations, 8 * (size_t)n);
  }
  return finish_load(h, s, n, net, off_model, off_iters, total, span_cap, max_need, n > 0 ? ji[n - 1].arrive : 0);
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

static int prepare_sim(gs_handle h, SimHost &s, int64_t rows_cap) {
  const gs_cluster &c = s.cl;
  const int M = c.num_switch * c.num_node_p_switch;
  const size_t N = (size_t)(s.n > 0 ? s.n : 1);
  int W = 256; while (W < s.max_need + 1) W <<= 1;    // >= 2x the lane engine's shared-memory window
  if (rows_cap <= 0) rows_cap = s.last_arrive + 2ll * s.max_need + 4096;
  size_t o_rec = 0, o_jst = align_up(o_rec + sizeof(gs_job_rec) * N), o_stack = align_up(o_jst + sizeof(JobState) * N);
  size_t o_fin = align_up(o_stack + 8 * N), o_sref = align_up(o_fin + 4 * N), o_wh = align_up(o_sref + 8 * N), o_wt = align_up(o_wh + 4 * (size_t)W);
  size_t o_spans = align_up(o_wt + 4 * (size_t)W), o_rows = align_up(o_spans + sizeof(gs_span) * (size_t)s.span_cap);
  size_t o_nb = align_up(o_rows + sizeof(gs_tick_row) * (size_t)rows_cap), o_nk = align_up(o_nb + 8 * (size_t)M);
  size_t total = align_up(o_nk + 4 * (size_t)M);
  const bool evd = s.pol.schedule != GS_SCHED_FIFO;        // event-driven policy: extra scratch
  const size_t nql = (size_t)(s.pol.num_queue > 2 ? s.pol.num_queue : 2);
  size_t o_pj = total, o_run = 0, o_q = 0, o_end = 0, o_tmp = 0, o_ci = 0, o_ck = 0;
  if (evd) {
    o_run = align_up(o_pj + sizeof(PJob) * N); o_q = align_up(o_run + 4 * N); o_end = align_up(o_q + 4 * N * nql);
    o_tmp = align_up(o_end + 4 * N); o_ci = align_up(o_tmp + 4 * N); o_ck = align_up(o_ci + 4 * (size_t)M);
    total = align_up(o_ck + 4 * (size_t)M);
  }
  if (s.state_slab && s.state_bytes < total) { cudaFree(s.state_slab); s.state_slab = nullptr; }
  if (!s.state_slab) { CU(cudaMalloc(&s.state_slab, total)); s.state_bytes = total; }
  unsigned char *d = (unsigned char *)s.state_slab;
  SimDev &D = s.dev;
  D.M = M; D.G = c.num_gpu_p_node;
  int kc = c.num_cpu_p_node / c.cpu_per_task, km = c.mem_p_node / c.mem_per_task;
  D.K = kc < km ? kc : km;
  D.netcost = c.enable_network_costs ? 1 : 0;
  D.n = (int)s.n; D.wheel_mask = W - 1; D.policy = s.pol.schedule;
  D.cap_bytes = (long long)c.gpu_mem_cap_mib << 20;
  D.fit_limit = D.cap_bytes - ((long long)500 << 20);      // cap - mem > 500 MiB  (device.py:75)
  D.bandwidth = c.bandwidth; D.latency = c.internode_latency;
  D.rec = (gs_job_rec *)(d + o_rec); D.jst = (JobState *)(d + o_jst);
  D.stack = (int *)(d + o_stack); D.fin = (int *)(d + o_fin); D.sref = (int2 *)(d + o_sref);
  D.wheel_head = (int *)(d + o_wh); D.wheel_tail = (int *)(d + o_wt);
  D.spans = (gs_span *)(d + o_spans); D.rows = (gs_tick_row *)(d + o_rows);
  D.nbusy = (unsigned long long *)(d + o_nb); D.nk = (int *)(d + o_nk);
  D.span_cap = s.span_cap; D.rows_cap = rows_cap;
  s.rows_cap = rows_cap;
  if (evd) {
    D.pj = (PJob *)(d + o_pj); D.runnable = (int *)(d + o_run); D.queues = (int *)(d + o_q);
    D.endj = (int *)(d + o_end); D.tmpl = (int *)(d + o_tmp); D.cidle = (int *)(d + o_ci); D.ckfree = (int *)(d + o_ck);
    D.num_queue = s.pol.num_queue > 0 ? s.pol.num_queue : 1;
    for (int q = 0; q < GS_MAX_QUEUES; ++q) { D.queue_limit[q] = s.pol.queue_limit[q]; D.qn[q] = 0; }
    D.gittins_delta = s.pol.gittins_delta; D.next_gittins_unit = s.pol.gittins_delta;
    D.git_n = s.pol.schedule == GS_SCHED_GITTINS ? s.pol.gittins_n : 0;
    D.git_data = (const double *)s.git_dev;
    D.git_index = s.git_dev ? (const double *)((unsigned char *)s.git_dev + 8 * (size_t)s.pol.gittins_n) : nullptr;
    D.rn = 0; D.en = 0; D.end_time = 0x7fffffff; D.next_job_jump = 0x7fffffff;
  }
  D.delta = D.p = D.top = D.running = D.finished = D.ever = D.busy_gpus = D.done = D.status = 0;
  D.mem_busy = D.sum_arr = D.span_used = D.events = D.evals = D.started = D.ticks = D.row_first = 0;
  D.need_init = 1;
  s.prepared = true;
  return GS_OK;
}

extern "C" int gs_run(gs_handle h, int64_t max_ticks, int64_t rows_cap) {
  if (!h) return GS_ERR_ARG;
  CU(cudaSetDevice(h->device));
  int maxM = 1;
  for (auto &s : h->sims) {
    if (!s.loaded) return fail(h, GS_ERR_STATE, "gs_run: every replica needs gs_config_sim + gs_load_trace");
    if (!s.prepared) { int rc = prepare_sim(h, s, rows_cap); if (rc) return rc; h->dirty = true; }
    int M = s.cl.num_switch * s.cl.num_node_p_switch;
    if (M > maxM) maxM = M;
  }
  bool any_init = false;
  if (h->dirty) {
    std::vector<SimDev> tmp((size_t)h->nsims);
    for (int i = 0; i < h->nsims; ++i) { tmp[(size_t)i] = h->sims[(size_t)i].dev; any_init |= tmp[(size_t)i].need_init != 0; }
    CU(cudaMemcpyAsync(h->d_sims, tmp.data(), sizeof(SimDev) * (size_t)h->nsims, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->dirty = false;
  }
  int maxG = 1;
  bool any_fifo = false, any_evd = false;
  for (auto &s : h->sims) {
    if (s.cl.num_gpu_p_node > maxG) maxG = s.cl.num_gpu_p_node;
    if (s.pol.schedule == GS_SCHED_FIFO) any_fifo = true; else any_evd = true;
  }
  const size_t lane_words = (size_t)maxM * (maxG > 32 ? 3 : 2) + LANE_EXTRA_WORDS;
  int L = 32;
  if (const char *e = getenv("GSCHED_LANES")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) L = v; }
  while (L > 1 && lane_words * (size_t)L * 4 > 100 * 1024) L >>= 1;
  const size_t lane_smem = lane_words * (size_t)L * 4;
  bool use_lane = h->engine_mode == 2;   // auto == warp mapping (measured faster at every replica count that fits HBM)
  if (use_lane && lane_smem > 200 * 1024) {
    if (h->engine_mode == 2) return fail(h, GS_ERR_ARG, "gs_run: node table too large for the lane engine");
    use_lane = false;
  }
  CU(cudaEventRecord(h->e0, h->stream));
  if (any_init) {      // state reset is part of the timed engine work
    gs_init_kernel<<<dim3(16, (unsigned)h->nsims), 256, 0, h->stream>>>(h->d_sims, h->nsims);
    CU(cudaGetLastError());
    h->launches += 1;
  }
  if (any_evd) {
    // dlas / dlas-gpu: one warp per replica; sjf / gittins (and everything with engine mode 2): one thread per replica
    const int thread_dlas = h->engine_mode == 2 ? 1 : 0;
    gs_policy_kernel<<<(unsigned)((h->nsims + 31) / 32), 32, 0, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, thread_dlas);
    CU(cudaGetLastError());
    h->launches += 1;
    if (!thread_dlas) {
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
  if (!any_fifo) {
    // nothing for the tick kernels
  } else if (use_lane) {
    const unsigned grid = (unsigned)((h->nsims + L - 1) / L);
    if (maxG > 32) {
      if (lane_smem > 48 * 1024)
        CU(cudaFuncSetAttribute(gs_lane_kernel<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lane_smem));
      gs_lane_kernel<unsigned long long><<<grid, 32, lane_smem, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, maxM, L);
    } else {
      if (lane_smem > 48 * 1024)
        CU(cudaFuncSetAttribute(gs_lane_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lane_smem));
      gs_lane_kernel<uint32_t><<<grid, 32, lane_smem, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, maxM, L);
    }
  } else {
    const int stride = (int)align_up((size_t)maxM * 12 + 8 + SCACHE * 8, 16);
    if (stride > 200 * 1024) return fail(h, GS_ERR_ARG, "gs_run: node table does not fit shared memory (M too large)");
    if (2 * stride > 48 * 1024)
    {
      CU(cudaFuncSetAttribute(gs_tick_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, stride));
      CU(cudaFuncSetAttribute(gs_tick_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * stride));
    }
    int sub = 32;
    if (const char *e = getenv("GSCHED_SUB")) { if (atoi(e) == 16) sub = 16; }
    if (h->engine_mode == 3) sub = 16;
    if (sub == 16)
      gs_tick_kernel<16><<<(unsigned)((h->nsims + 1) / 2), 32, (size_t)stride * 2, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, stride);
    else
      gs_tick_kernel<32><<<(unsigned)h->nsims, 32, (size_t)stride, h->stream>>>(h->d_sims, h->nsims, (long long)max_ticks, stride);
  }
  CU(cudaGetLastError());
  if (any_fifo) h->launches += 1;
  CU(cudaEventRecord(h->e1, h->stream));
  std::vector<SimDev> back((size_t)h->nsims);
  CU(cudaMemcpyAsync(back.data(), h->d_sims, sizeof(SimDev) * (size_t)h->nsims, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->kernel_ms += ms;
  int worst = 0;
  for (int i = 0; i < h->nsims; ++i) {
    back[(size_t)i].need_init = 0;
    h->sims[(size_t)i].dev = back[(size_t)i];
    if (back[(size_t)i].status != 0 && worst == 0) worst = back[(size_t)i].status;
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
  return timed_d2h(h, rows_out, D.rows + (first - D.row_first), sizeof(gs_tick_row) * (size_t)count);
}

extern "C" int gs_fetch_jobs(gs_handle h, int sim, gs_job_rec *jobs_out, int32_t *finish_order_out) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_jobs: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_jobs: nothing has run yet");
  CU(cudaSetDevice(h->device));
  CU(cudaEventRecord(h->e0, h->stream));
  if (jobs_out && s.n > 0)
    CU(cudaMemcpyAsync(jobs_out, s.dev.rec, sizeof(gs_job_rec) * (size_t)s.n, cudaMemcpyDeviceToHost, h->stream));
  if (finish_order_out && s.dev.finished > 0)
    CU(cudaMemcpyAsync(finish_order_out, s.dev.fin, 4 * (size_t)s.dev.finished, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->d2h_ms += ms;
  return GS_OK;
}

static int ensure_scratch(gs_handle h, size_t bytes) {
  if (h->d_scratch_bytes >= bytes) return GS_OK;
  if (h->d_scratch) cudaFree(h->d_scratch);
  h->d_scratch = nullptr; h->d_scratch_bytes = 0;
  CU(cudaMalloc(&h->d_scratch, bytes));
  h->d_scratch_bytes = bytes;
  return GS_OK;
}

extern "C" int gs_fetch_spans(gs_handle h, int sim, int64_t *span_off_out, gs_span *spans_out, int64_t spans_cap,
                              int64_t *spans_used) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_spans: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_spans: nothing has run yet");
  CU(cudaSetDevice(h->device));
  const int64_t used = s.dev.span_used;
  if (spans_used) *spans_used = used;
  if (!spans_out && !span_off_out) return GS_OK;
  if (spans_out && spans_cap < used) return fail(h, GS_ERR_CAPACITY, "gs_fetch_spans: spans_out too small");
  const size_t N = (size_t)(s.n > 0 ? s.n : 1);
  const size_t o_off = 0, o_sp = align_up(8 * (N + 1)), total = align_up(o_sp + sizeof(gs_span) * (size_t)(used > 0 ? used : 1));
  int rc = ensure_scratch(h, total);
  if (rc) return rc;
  unsigned char *d = (unsigned char *)h->d_scratch;
  long long *d_off = (long long *)(d + o_off);
  gs_span *d_sp = (gs_span *)(d + o_sp);
  CU(cudaEventRecord(h->e0, h->stream));
  gs_span_scan_kernel<<<1, 1024, 0, h->stream>>>(s.dev.rec, s.dev.sref, (int)s.n, d_off);
  if (s.n > 0 && used > 0 && spans_out)
    gs_span_gather_kernel<<<(unsigned)((s.n + 255) / 256), 256, 0, h->stream>>>(s.dev.rec, s.dev.sref, s.dev.spans, d_off,
                                                                              (int)s.n, d_sp);
  CU(cudaGetLastError());
  h->launches += 2;
  if (span_off_out) CU(cudaMemcpyAsync(span_off_out, d_off, 8 * (size_t)(s.n + 1), cudaMemcpyDeviceToHost, h->stream));
  if (spans_out && used > 0) CU(cudaMemcpyAsync(spans_out, d_sp, sizeof(gs_span) * (size_t)used, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->d2h_ms += ms;
  return GS_OK;
}

// Everything a caller needs from one finished (or paused) replica in ONE call: the rows of the last
// window, job records, finish order and the spans grouped by job.  The big row copy runs on the
// main stream while the regroup kernels and the small copies run on a second stream.
extern "C" int gs_fetch_all(gs_handle h, int sim, int64_t first, int64_t count, gs_tick_row *rows_out,
                            gs_job_rec *jobs_out, int32_t *finish_order_out, int64_t *span_off_out,
                            gs_span *spans_out, int64_t spans_cap, int64_t *spans_used) {
  if (!h) return GS_ERR_ARG;
  if (sim < 0 || sim >= h->nsims) return fail(h, GS_ERR_ARG, "gs_fetch_all: sim index out of range");
  SimHost &s = h->sims[(size_t)sim];
  if (!s.prepared) return fail(h, GS_ERR_STATE, "gs_fetch_all: nothing has run yet");
  const SimDev &D = s.dev;
  if (rows_out && (count < 0 || first < D.row_first || first + count > D.ticks))
    return fail(h, GS_ERR_ARG, "gs_fetch_all: row range is outside the last gs_run window");
  const int64_t used = D.span_used;
  if (spans_used) *spans_used = used;
  if (spans_out && spans_cap < used) return fail(h, GS_ERR_CAPACITY, "gs_fetch_all: spans_out too small");
  CU(cudaSetDevice(h->device));
  const size_t N = (size_t)(s.n > 0 ? s.n : 1);
  const size_t o_sp = align_up(8 * (N + 1)), total = align_up(o_sp + sizeof(gs_span) * (size_t)(used > 0 ? used : 1));
  if (h->d_scratch2_bytes < total) {
    if (h->d_scratch2) cudaFree(h->d_scratch2);
    h->d_scratch2 = nullptr; h->d_scratch2_bytes = 0;
    CU(cudaMalloc(&h->d_scratch2, total));
    h->d_scratch2_bytes = total;
  }
  unsigned char *d = (unsigned char *)h->d_scratch2;
  long long *d_off = (long long *)d;
  gs_span *d_sp = (gs_span *)(d + o_sp);
  CU(cudaEventRecord(h->e0, h->stream));
  if (rows_out && count > 0)
    CU(cudaMemcpyAsync(rows_out, D.rows + (first - D.row_first), sizeof(gs_tick_row) * (size_t)count, cudaMemcpyDeviceToHost, h->stream));
  if (span_off_out || spans_out) {
    gs_span_scan_kernel<<<1, 1024, 0, h->stream2>>>(D.rec, D.sref, (int)s.n, d_off);
    if (s.n > 0 && used > 0 && spans_out)
      gs_span_gather_kernel<<<(unsigned)((s.n + 255) / 256), 256, 0, h->stream2>>>(D.rec, D.sref, D.spans, d_off, (int)s.n, d_sp);
    CU(cudaGetLastError());
    h->launches += 2;
    if (span_off_out) CU(cudaMemcpyAsync(span_off_out, d_off, 8 * (size_t)(s.n + 1), cudaMemcpyDeviceToHost, h->stream2));
    if (spans_out && used > 0) CU(cudaMemcpyAsync(spans_out, d_sp, sizeof(gs_span) * (size_t)used, cudaMemcpyDeviceToHost, h->stream2));
  }
  if (jobs_out && s.n > 0) CU(cudaMemcpyAsync(jobs_out, D.rec, sizeof(gs_job_rec) * (size_t)s.n, cudaMemcpyDeviceToHost, h->stream2));
  if (finish_order_out && D.finished > 0) CU(cudaMemcpyAsync(finish_order_out, D.fin, 4 * (size_t)D.finished, cudaMemcpyDeviceToHost, h->stream2));
  CU(cudaStreamSynchronize(h->stream2));
  CU(cudaEventRecord(h->e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0; cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->d2h_ms += ms;
  return GS_OK;
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

// 0 = auto (lane engine from 32 replicas up), 1 = one warp per replica, 2 = one lane per replica
extern "C" int gs_set_engine(gs_handle h, int mode) {
  if (!h) return GS_ERR_ARG;
  if (mode < 0 || mode > 3) return fail(h, GS_ERR_ARG, "gs_set_engine: mode must be 0..3");
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
