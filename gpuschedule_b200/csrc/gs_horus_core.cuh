// Utilisation-aware placement engine (horus / gandiva), one simulation per thread -- first correct
// device path for SURVEY 8(f) rank 1; the per-replica state below is laid out so that the scoring,
// statistics and aging loops can later be spread over the lanes of a warp (the draw index of every
// (node, device, task) sample is a prefix sum of known counts, so the stream order survives).
//
// Reference semantics (file:line in /root/reference):
//   horus_placement / score functions   core/scheduling/algorithm.py:34-180, core/scheduling/horus.py:6-56
//   schedule_horus / schedule_fifo       core/scheduling/algorithm.py:189-240
//   time_slice_check + preempt (gandiva) core/scheduling/algorithm.py:420-440, core/jobs/jobs_manager.py:150-201
//   Device / Node packing rules          infra/device.py:20-76, infra/node.py:57-232
//   tick loop, completion, statistics    core/scheduling/schedule.py:39-213
// The reference samples numpy's global legacy stream inside these decisions; the engine consumes the
// same stream as an array of standard-normal values g[k] supplied by the host
// (numpy.random.standard_normal continues the stream numpy.random.normal(loc, scale, size=1) uses:
// the k-th sample is loc + scale * g[k]), so results are bit-identical to a seeded reference run.
//
// Everything in this header is plain scalar C++ (GS_HD = __host__ __device__), which lets
// tests/emu/ compile the very same functions with g++ and check the logic on a CPU-only box;
// the product only ever runs them inside gs_horus_kernel.
#pragma once
#include <stdint.h>

#include "gsched.h"
#include "gsched_horus.h"

#ifndef GS_HD
#define GS_HD __host__ __device__ __forceinline__
#endif

// numpy evaluates these expressions as separate IEEE multiplications and additions; nvcc would contract
// a * b + c into one FMA (different rounding), so the device build spells them with the _rn intrinsics.
#ifdef __CUDA_ARCH__
#define H_MUL(a, b) __dmul_rn((a), (b))
#define H_ADD(a, b) __dadd_rn((a), (b))
#else
#define H_MUL(a, b) ((a) * (b))
#define H_ADD(a, b) ((a) + (b))
#endif

#define H_TASK_CPU 12
#define H_TASK_MEM 60
#define H_DEV_SLOTS 4
#define H_MAXQ 8

struct HJob {            // one trace row (read-only), 64 bytes
  int arrive, gpus, gpc, ntasks;
  int first_task, pad;
  long long mem_b;
  double util_avg, util_max, duration;
  double mem_avg_mib;    // Job.gpu_mem_avg: only a k-means feature (core/jobs/utils.py:4-22)
};
struct HJobState {
  int pending, start, end, migration, tasks_finished, tro_n;
  unsigned char running, finished, in_running, pad;
};
struct HTask {           // 72 bytes; the constants are copied from the job so that the sampling loops touch one record
  double duration, original;           // Task.duration / original_duration (job.py:33-34)
  double util_avg, half_spread;        // loc and scale of the utilisation sample: avg, (max - avg) / 2 (device.py:52)
  double quarter_spread;               // scale of the interference sample: (max - avg) / 4 (device.py:31)
  long long mem_clamped;               // min(device memory, gpu_memory_max) in bytes (device.py:59)
  int job, time_processed;
  int placed_node, run_node;           // membership in Node.placed_tasks / Node.running_tasks
  unsigned char interfered, running, finished, pad;
  int pad2;
};
// Device.running_tasks in insertion order + the sum of the clamped task memories (get_current_memory is
// min(capacity, that sum): clamping after every term or once gives the same value for non-negative terms)
struct HDev { int nt; int t[H_DEV_SLOTS]; int pad; long long mem_sum; };
struct HNode { int cpu_used, mem_used, n_running, n_placed_tasks, n_placed_jobs; };
struct HCand { double min_score; int node, pad; };

struct HSim {
  // ---- configuration
  int M, G, S, P, cpu_cap, mem_cap, scheme, schedule, num_buffer, n, maxg, pjw;
  int placement, pad_cfg;               // GS_HPLACE_*: horus_placement or ms_yarn_placement
  long long cap_b;
  // ---- trace and simulation state
  const HJob *jobs; HJobState *js; HTask *tasks;
  int *tro_node, *tro_order;            // Job.tasks_running_on: value per task / key insertion order
  HNode *nodes; HDev *devs;
  unsigned long long *pj_bits;          // Node.placed_jobs membership, bit (job, node)
  int *queue;                           // nq queues of n + 1 slots each (heap order for horus / horus+)
  int *running, *fin;
  int qn[H_MAXQ];
  double credits[H_MAXQ];               // JobQueueManager.queue_credits
  int nq, nrun, nfin, pad0;
  // ---- scratch
  int *look, *look_q, *work, *res_nodes, *map_node, *map_order, *map_n, *ok, *distinct;
  long long *map_skip;                  // samples a candidate's trial consumed
  int *km_all, *km_assign, *km_old;     // horus+: jobs being re-clustered, their assignment, the previous one
  double *km_score;
  HCand *heap;
  // ---- the sampled stream: either standard-normal values (horus / gandiva) ...
  const double *gauss; long long gauss_n, gauss_pos;
  // ... or raw MT19937 words with per-position tables of the polar-method outcome (any schedule; needed by
  // horus+, whose integer draws shift where the next normal sample starts).  See gs_horus_host.h.
  const unsigned int *words; const double *gv_ret, *gv_keep; const int *gv_next;
  const int *gv_acc, *gv_rank; int gv_cls_off[5], pad_idx;   // accepted start positions per residue class + rank (gs_horus_host.h)
  long long words_n, words_pos, draws;
  double gauss_kept; int has_gauss, pad1;
  // ---- results
  gs_tick_row *rows; double *util; unsigned char *util_arr; gs_horus_job_rec *recs; long long rows_cap;
  // ---- scheduling step in flight (candidates of this tick) and the cooperative scoring request
  int la_n, la_i, la_pos, la_nres;
  int phase, req, sc_job, sc_k, sc_total, sc_hn;
  long long sc_base, budget;
  int *sc_cnt, *sc_off; double *sc_cost;   // per device: samples it consumes, their offset, its cost (cooperative scoring)
  // ---- loop state (persisted between launches)
  int delta, p, status, done;
  long long ticks, events, current_remaining, running_jobs;
};

GS_HD double h_gauss(HSim &s) {
  s.draws += 1;
  if (s.words) {                                            // legacy_gauss over the word stream: cached second value first
    if (s.has_gauss) { s.has_gauss = 0; return s.gauss_kept; }
    if (s.words_pos >= s.words_n || s.gv_next[s.words_pos] < 0) { s.status = GS_ERR_CAPACITY; return 0.0; }
    const long long p = s.words_pos;
    s.gauss_kept = s.gv_keep[p]; s.has_gauss = 1; s.words_pos = s.gv_next[p];
    return s.gv_ret[p];
  }
  if (s.gauss_pos >= s.gauss_n) { s.status = GS_ERR_CAPACITY; return 0.0; }      // host supplies a longer stream and re-runs
  return s.gauss[s.gauss_pos++];
}
// numpy.random.randint(n) / choice(n) of the legacy generator: masked rejection over 32-bit words
GS_HD long long h_below(HSim &s, long long n) {
  const unsigned long long mx = (unsigned long long)(n - 1);
  if (mx == 0) return 0;
  unsigned long long mask = mx;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
  for (;;) {
    if (!s.words || s.words_pos >= s.words_n) { s.status = GS_ERR_CAPACITY; return 0; }
    const unsigned long long v = (unsigned long long)s.words[s.words_pos++] & mask;     // n < 2^32 here (n <= jobs)
    if (v <= mx) return (long long)v;
  }
}
// Sample number idx (0-based) counted from the CURRENT state of the stream, without consuming anything.
// Values form: a plain offset.  Word form: the kept second value first (if any), then pairs along one residue class.
GS_HD double h_sample_at(const HSim &s, long long idx) {
  if (!s.words) return s.gauss[s.gauss_pos + idx];
  if (s.has_gauss) { if (idx == 0) return s.gauss_kept; idx -= 1; }
  const long long p = s.words_pos;
  const int at = s.gv_acc[s.gv_cls_off[p & 3] + s.gv_rank[p] + (int)(idx >> 1)];
  return (idx & 1) ? s.gv_keep[at] : s.gv_ret[at];
}
// are `count` more samples available from the current state?
GS_HD bool h_samples_available(const HSim &s, long long count) {
  if (!s.words) return s.gauss_pos + count <= s.gauss_n;
  if (s.has_gauss) count -= 1;
  if (count <= 0) return true;
  const long long p = s.words_pos;
  if (p >= s.words_n) return false;
  const long long pairs = (count + 1) >> 1, c = p & 3;
  return s.gv_rank[p] + pairs <= (long long)(s.gv_cls_off[c + 1] - s.gv_cls_off[c]);
}
// consume `count` samples (what `count` calls of h_gauss would have done to the state)
GS_HD void h_skip_samples(HSim &s, long long count) {
  s.draws += count;
  if (!s.words) { s.gauss_pos += count; return; }
  if (count <= 0) return;
  if (s.has_gauss) { s.has_gauss = 0; count -= 1; if (count == 0) return; }
  const long long p = s.words_pos, pairs = (count + 1) >> 1;
  const int last = s.gv_acc[s.gv_cls_off[p & 3] + s.gv_rank[p] + (int)(pairs - 1)];
  s.words_pos = (long long)last + 4;
  if (count & 1) { s.has_gauss = 1; s.gauss_kept = s.gv_keep[last]; }
}
GS_HD double h_normal(HSim &s, double loc, double scale) { return H_ADD(loc, H_MUL(scale, h_gauss(s))); }
GS_HD HDev &h_dev(HSim &s, int nd, int d) { return s.devs[(long long)nd * s.G + d]; }
GS_HD long long h_task_mem(const HSim &s, int t) { return s.jobs[s.tasks[t].job].mem_b; }
GS_HD bool h_node_is_free(const HSim &s, int nd) { return s.cpu_cap - s.nodes[nd].cpu_used > 0 || s.mem_cap - s.nodes[nd].mem_used > 0; }   // node.py:57-58

// Device.get_current_memory (device.py:56-62) in bytes: every term is an exact binary fraction of a MiB
GS_HD long long h_dev_mem(const HSim &s, const HDev &d) { return d.mem_sum > s.cap_b ? s.cap_b : d.mem_sum; }
// Device.can_fit (device.py:67-76): at most 4 tasks, 500 MiB margin
GS_HD bool h_dev_can_fit(const HSim &s, const HDev &d, int t) {
  const long long cur = h_dev_mem(s, d);
  if (d.nt >= 4) return false;
  return s.cap_b - (cur + h_task_mem(s, t)) > (500LL << 20);
}
// Device.get_current_utilization (device.py:48-54); is_arr: the Python value is a 1-element numpy array
GS_HD double h_dev_util(HSim &s, const HDev &d, int *is_arr) {
  double u = 0.0; int arr = 0;
  for (int i = 0; i < d.nt; ++i) {
    const HTask &o = s.tasks[d.t[i]];
    const double x = h_normal(s, o.util_avg, o.half_spread);
    if (x < 100.0) { u = H_ADD(u, x); arr = 1; } else { u = H_ADD(u, 100.0); }    // min(100, sample)
    if (100.0 < u) { u = 100.0; arr = 0; }                                        // min(util, 100)
  }
  if (is_arr) *is_arr = arr;
  return u;
}
// Device.add_task (device.py:20-43)
GS_HD bool h_dev_add_task(HSim &s, HDev &d, int t, bool pack) {
  if (!h_dev_can_fit(s, d, t)) return false;
  if (!pack && d.nt > 0) return false;
  HTask &tk = s.tasks[t];
  if (d.nt >= 2) {
    // one interference sample per resident task is drawn, but the slowed duration they feed is only logged
    // (device.py:28-37): the values are never used, so the stream is simply advanced
    if (h_samples_available(s, d.nt)) h_skip_samples(s, d.nt); else s.status = GS_ERR_CAPACITY;
    tk.interfered = 1;
  } else { tk.interfered = 0; tk.duration = tk.original; }
  for (int i = 0; i < d.nt; ++i) if (d.t[i] == t) return true;                   // key already present: position kept
  if (d.nt >= H_DEV_SLOTS) { s.status = GS_ERR_STATE; return true; }
  d.t[d.nt++] = t; d.mem_sum += tk.mem_clamped;
  return true;
}
// Node.can_fit (node.py:136-162)
GS_HD bool h_node_can_fit(const HSim &s, int nd, int t, bool pack) {
  if (s.cpu_cap - s.nodes[nd].cpu_used - H_TASK_CPU < 0 || s.mem_cap - s.nodes[nd].mem_used - H_TASK_MEM < 0) return false;
  const HDev *dv = s.devs + (long long)nd * s.G;
  if (!pack) {
    int idle = 0;
    for (int d = 0; d < s.G; ++d) idle += (dv[d].nt == 0);
    return idle - s.jobs[s.tasks[t].job].gpc >= 0;
  }
  for (int d = 0; d < s.G; ++d) if (h_dev_can_fit(s, dv[d], t)) return true;
  return false;
}
// Node.try_reserve_and_placed_task (node.py:190-211); a partial placement keeps what it took
GS_HD bool h_node_reserve_fitting(HSim &s, int nd, int t, bool pack);
GS_HD bool h_node_reserve_task(HSim &s, int nd, int t, bool pack) {
  if (!h_node_can_fit(s, nd, t, pack)) return false;
  return h_node_reserve_fitting(s, nd, t, pack);
}
// the part after Node.can_fit said yes
GS_HD bool h_node_reserve_fitting(HSim &s, int nd, int t, bool pack) {
  s.nodes[nd].cpu_used += H_TASK_CPU; s.nodes[nd].mem_used += H_TASK_MEM;
  int need = s.jobs[s.tasks[t].job].gpc;
  for (int d = 0; d < s.G; ++d) {
    if (need <= 0) break;
    if (h_dev_add_task(s, h_dev(s, nd, d), t, pack)) need -= 1;
  }
  if (need == 0 && s.tasks[t].placed_node != nd) { s.tasks[t].placed_node = nd; s.nodes[nd].n_placed_tasks += 1; }
  return need == 0;
}
GS_HD void h_node_place_job(HSim &s, int nd, int j) {     // placed_jobs[job_id] = job (node.py:213-232)
  unsigned long long &w = s.pj_bits[(long long)j * s.pjw + (nd >> 6)]; const unsigned long long b = 1ull << (nd & 63);
  if (!(w & b)) { w |= b; s.nodes[nd].n_placed_jobs += 1; }
}
GS_HD void h_node_pop_job(HSim &s, int nd, int j) {       // placed_jobs.pop(job_id)
  unsigned long long &w = s.pj_bits[(long long)j * s.pjw + (nd >> 6)]; const unsigned long long b = 1ull << (nd & 63);
  if (w & b) { w &= ~b; s.nodes[nd].n_placed_jobs -= 1; }
}
// JobsManager.reset_interference for one task (jobs_manager.py:189-201)
GS_HD void h_reset_task(HSim &s, int t) {
  HTask &tk = s.tasks[t];
  if (!s.js[tk.job].in_running || !tk.interfered) return;
  tk.interfered = 0;
  const double diff = tk.duration - tk.original;
  const long long half = (long long)(diff / 2);                                  // int(diff / 2)
  tk.duration = tk.original + (double)(half > 5 ? half : 5);
}
// Node.release_allocated_resources (node.py:64-84).  With `lift`, the tasks left alone on a device get
// their interference mark lifted right away (the caller's reset_interference on the returned set).
GS_HD void h_node_release(HSim &s, int nd, int t, bool lift) {
  s.nodes[nd].cpu_used -= H_TASK_CPU; s.nodes[nd].mem_used -= H_TASK_MEM;
  for (int d = 0; d < s.G; ++d) {
    HDev &dv = h_dev(s, nd, d);
    for (int i = 0; i < dv.nt; ++i) if (dv.t[i] == t) { for (int k = i; k + 1 < dv.nt; ++k) dv.t[k] = dv.t[k + 1]; dv.nt--; dv.mem_sum -= s.tasks[t].mem_clamped; break; }
  }
  if (!lift) return;
  for (int d = 0; d < s.G; ++d) {          // the set is built over ALL devices first, then applied: same outcome, marks are per task
    const HDev &dv = h_dev(s, nd, d);
    if (dv.nt <= 1) for (int i = 0; i < dv.nt; ++i) if (s.tasks[dv.t[i]].interfered) h_reset_task(s, dv.t[i]);
  }
}

GS_HD double h_polyval(double x) { double y = 0.0; y = H_ADD(H_MUL(y, x), 4E-5); y = H_ADD(H_MUL(y, x), -0.00302); y = H_ADD(H_MUL(y, x), 1.16664); return y; }   // np.polyval(NV_2080_COEF, x)

// score_fn[scheme](node, task) -> min_cost (horus.py:6-56)
GS_HD double h_score_node(HSim &s, int nd, int t) {
  const HJob &jb = s.jobs[s.tasks[t].job];
  const double cap_mib = (double)(s.cap_b >> 20), tm = (double)jb.mem_b / 1048576.0;
  double min_cost = 999.0;
  for (int d = 0; d < s.G; ++d) {
    const HDev &dv = h_dev(s, nd, d);
    if (!h_dev_can_fit(s, dv, t)) continue;
    const double cur = (double)h_dev_mem(s, dv) / 1048576.0;
    double cost;
    if (s.scheme == GS_HSCORE_HORUS) {
      const double mem_cost = (cur + tm) / cap_mib;
      const double val = H_ADD(h_dev_util(s, dv, nullptr), jb.util_avg);
      const double util_cost = h_polyval(val);
      cost = H_ADD(H_ADD(H_MUL(mem_cost, 0.5), H_MUL(util_cost, 0.5)), (double)dv.nt);
    } else {
      const double mem_cost = cur + tm / cap_mib;          // precedence as written, horus.py:16
      const double util_cost = h_dev_util(s, dv, nullptr);
      cost = H_ADD(H_ADD(H_MUL(mem_cost, 0.5), util_cost / 100), (double)dv.nt);
    }
    if (cost < min_cost) min_cost = cost;
  }
  return min_cost;
}

// heapq over NodeDeviceInfo.__lt__ (algorithm.py:25-26): a < b  <=>  a.min_score > b.min_score
GS_HD void h_heap_push(HCand *h, int &n, HCand x) {
  int pos = n++;
  while (pos > 0) { const int parent = (pos - 1) >> 1; if (x.min_score > h[parent].min_score) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = x;
}
GS_HD void h_heap_pop(HCand *h, int &n) {
  const HCand last = h[--n];
  if (n == 0) return;
  int pos = 0, child = 1;
  while (child < n) {
    const int right = child + 1;
    if (right < n && !(h[child].min_score > h[right].min_score)) child = right;
    h[pos] = h[child]; pos = child; child = 2 * pos + 1;
  }
  while (pos > 0) { const int parent = (pos - 1) >> 1; if (last.min_score > h[parent].min_score) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = last;
}

// horus_placement (algorithm.py:34-180): true on success, res_nodes = distinct nodes in first-use order
// second half of horus_placement: the candidates are in s.heap[0..hn) (heap order); sort, try, commit
GS_HD bool h_placement_finish(HSim &s, int j, int hn, int &n_res) {
  const HJob &jb = s.jobs[j];
  HJobState &st = s.js[j];
  const int T = jb.ntasks, t0 = jb.first_task;
  HCand *heap = s.heap;
  for (int i = 1; i < hn; ++i) {                          // sorted(nodes_stack, key=min_score): stable
    const HCand x = heap[i]; int k = i - 1;
    while (k >= 0 && heap[k].min_score > x.min_score) { heap[k + 1] = heap[k]; --k; }
    heap[k + 1] = x;
  }
  const int C = hn;
  // Two shortcuts that leave every outcome as it is.  (a) The tasks of a job are alike, and a Node.can_fit refusal
  // changes nothing: once a node refuses one task it refuses the rest, so the task loop stops there.  (b) A node
  // sits in the candidate list once per task; with one device per task (no partial placements) a trial is undone
  // completely (but for the placed_jobs marks, which are re-applied), so trying the same node again gives the same
  // plan and advances the stream by the same number of (unused) interference samples -- the plan is copied and the
  // stream skipped instead.
  const bool replay = jb.gpc == 1;
  for (int i = 0; i < C; ++i) {
    int *mn = s.map_node + (long long)i * T, *mo = s.map_order + (long long)i * T;
    int mapped = 0;
    const int cand = heap[i].node;
    if (replay) {
      int same = -1;
      for (int e = 0; e < i; ++e) if (heap[e].node == cand) { same = e; break; }
      if (same >= 0) {
        const int *pn = s.map_node + (long long)same * T, *po = s.map_order + (long long)same * T;
        for (int k = 0; k < T; ++k) { mn[k] = pn[k]; mo[k] = po[k]; }
        s.map_n[i] = s.map_n[same]; s.ok[i] = s.ok[same]; s.distinct[i] = s.distinct[same]; s.map_skip[i] = s.map_skip[same];
        if (h_samples_available(s, s.map_skip[i])) h_skip_samples(s, s.map_skip[i]); else { s.status = GS_ERR_CAPACITY; return false; }
        // what a trial leaves behind even after its undo: placed_jobs[job] on every node it used except the first
        // (algorithm.py:127-137) -- another candidate's undo may have popped one of them in between
        for (int q = 0; q < s.map_n[i]; ++q) h_node_place_job(s, mn[mo[q]], j);
        if (s.map_n[i] > 0) h_node_pop_job(s, mn[mo[0]], j);
        continue;
      }
    }
    const long long draws_before = s.draws;
    for (int k = 0; k < T; ++k) mn[k] = -1;
    for (int k = 0; k < T; ++k) {
      if (!h_node_can_fit(s, cand, t0 + k, true)) break;                       // (a)
      if (h_node_reserve_fitting(s, cand, t0 + k, true)) { h_node_place_job(s, cand, j); mn[k] = cand; mo[mapped++] = k; }
    }
    const int home = cand / s.P;
    bool stop = false;
    for (int dist = 0; dist < s.S && !stop; ++dist)      // get_racks_by_dist: stable sort of the racks by |rack - home|
      for (int r = 0; r < s.S && !stop; ++r) {
        const int dr = r > home ? r - home : home - r;
        if (dr != dist) continue;
        if (mapped >= T) { stop = true; break; }
        for (int q = 0; q < s.P; ++q) {
          const int nd = r * s.P + q;
          if (mapped >= T) break;
          for (int k = 0; k < T; ++k) {
            if (mn[k] >= 0) continue;
            if (!h_node_can_fit(s, nd, t0 + k, true)) break;                   // (a)
            if (h_node_reserve_fitting(s, nd, t0 + k, true)) { h_node_place_job(s, nd, j); mn[k] = nd; mo[mapped++] = k; }
            if (mapped >= T) break;
          }
        }
      }
    for (int q = 0; q < mapped; ++q) {                    // undo the trial (:127-137)
      const int k = mo[q], nd = mn[k];
      if (q == 0) h_node_pop_job(s, nd, j);
      if (s.tasks[t0 + k].placed_node == nd) {
        s.tasks[t0 + k].placed_node = -1; s.nodes[nd].n_placed_tasks -= 1;
        h_node_release(s, nd, t0 + k, false);
      }
    }
    s.map_n[i] = mapped; s.ok[i] = mapped >= T; s.map_skip[i] = s.draws - draws_before;
    int dn = 0;
    for (int q = 0; q < mapped; ++q) { bool seen = false; for (int p2 = 0; p2 < q; ++p2) seen |= (mn[mo[p2]] == mn[mo[q]]); dn += !seen; }
    s.distinct[i] = dn;
  }
  int best = -1;
  for (int i = 0; i < C; ++i) if (s.ok[i] && (best < 0 || s.distinct[i] < s.distinct[best])) best = i;    // stable sort by len(nodes)
  if (best < 0) return false;
  const int *mn = s.map_node + (long long)best * T, *mo = s.map_order + (long long)best * T;
  n_res = 0;
  for (int q = 0; q < T; ++q) {
    const int k = mo[q], nd = mn[k];
    if (!h_node_reserve_task(s, nd, t0 + k, true)) { s.status = GS_ERR_STATE; return false; }   // the reference asserts here
    bool seen = false;
    for (int p2 = 0; p2 < n_res; ++p2) seen |= (s.res_nodes[p2] == nd);
    if (!seen) s.res_nodes[n_res++] = nd;
    if (s.tro_node[t0 + k] < 0) s.tro_order[t0 + st.tro_n++] = k;               // dict: a known key keeps its position
    s.tro_node[t0 + k] = nd;
    h_node_place_job(s, nd, j);
  }
  return true;
}
GS_HD bool h_placement(HSim &s, int j, int &n_res) {
  const HJob &jb = s.jobs[j];
  const int T = jb.ntasks, demand = jb.gpus, t0 = jb.first_task;
  HCand *heap = s.heap;
  int hn = 0;
  for (int k = 0; k < T; ++k)
    for (int nd = 0; nd < s.M; ++nd) {
      if (!h_node_is_free(s, nd) || !h_node_can_fit(s, nd, t0 + k, true)) continue;
      HCand x; x.node = nd; x.pad = 0; x.min_score = h_score_node(s, nd, t0 + k);
      h_heap_push(heap, hn, x);
      if (hn > demand) h_heap_pop(heap, hn);
    }
  return h_placement_finish(s, j, hn, n_res);
}

// ---- --scheme yarn under these schedulers: ms_yarn_placement (algorithm.py:28-32,301-417), no packing
GS_HD int h_node_idle_devices(const HSim &s, int nd) { int c = 0; const HDev *dv = s.devs + (long long)nd * s.G; for (int d = 0; d < s.G; ++d) c += (dv[d].nt == 0); return c; }
GS_HD int h_can_fit_num_task(const HSim &s, int nd, int gpc, int remaining) {              // node.py:109-127
  const int g = h_node_idle_devices(s, nd) / gpc - remaining;
  const int c = (s.cpu_cap - s.nodes[nd].cpu_used) / H_TASK_CPU - remaining, m = (s.mem_cap - s.nodes[nd].mem_used) / H_TASK_MEM - remaining;
  const int ng = g >= 0 ? remaining : remaining + g, nc = c >= 0 ? remaining : remaining + c, nm = m >= 0 ? remaining : remaining + m;
  const int r = nc < nm ? nc : nm;
  return r < ng ? r : ng;
}
GS_HD void h_tro_set(HSim &s, int j, int k, int nd) {          // job.tasks_running_on[task] = node (a known key keeps its position)
  const int t0 = s.jobs[j].first_task; HJobState &st = s.js[j];
  if (s.tro_node[t0 + k] < 0) {
    bool known = false;
    for (int q = 0; q < st.tro_n; ++q) known |= (s.tro_order[t0 + q] == k);
    if (!known) s.tro_order[t0 + st.tro_n++] = k;
  }
  s.tro_node[t0 + k] = nd;
}
GS_HD bool h_yarn_placement(HSim &s, int j, int &n_res) {
  const HJob &jb = s.jobs[j];
  const int T = jb.ntasks, t0 = jb.first_task, gpc = jb.gpc;
  n_res = 0;
  if (jb.gpus <= s.G) {                                         // try_single_node_alloc_ms (algorithm.py:396-417)
    for (int nd = 0; nd < s.M; ++nd) {
      if (!h_node_is_free(s, nd)) continue;
      if (!(h_node_idle_devices(s, nd) >= jb.gpus && s.cpu_cap - s.nodes[nd].cpu_used >= H_TASK_CPU * T && s.mem_cap - s.nodes[nd].mem_used >= H_TASK_MEM * T)) continue;
      if (h_can_fit_num_task(s, nd, gpc, T) < T) continue;      // Node.try_alloc_job (node.py:234-263)
      int placed = 0;
      for (int k = 0; k < T; ++k) if (h_node_reserve_task(s, nd, t0 + k, false)) { h_tro_set(s, j, k, nd); ++placed; }
      if (placed == 0) continue;                                // devices refused every task: what was charged stays (Q21)
      h_node_place_job(s, nd, j);
      s.res_nodes[n_res++] = nd;
      return true;
    }
    return false;
  }
  int assigned = 0; const int least = (jb.gpus + s.G - 1) / s.G;   // try_cross_node_alloc_ms (algorithm.py:301-393)
  for (int nd = 0; nd < s.M; ++nd) {
    if (!h_node_is_free(s, nd)) continue;
    if (assigned == T) break;
    const int can = h_can_fit_num_task(s, nd, gpc, T - assigned);
    if (can == 0) continue;
    int worker_count = 0; bool check_next = false;
    for (int k = assigned; k < T; ++k) {
      if (!(worker_count <= can)) continue;                     // the `<=` over-try (:341)
      ++worker_count;
      if (!h_node_reserve_task(s, nd, t0 + k, false)) { --worker_count; check_next = true; break; }
      h_tro_set(s, j, k, nd);
    }
    if (worker_count > 0) { assigned += worker_count; h_node_place_job(s, nd, j); s.res_nodes[n_res++] = nd; }
    if (check_next) continue;
    if (n_res >= least && assigned == T) break;
  }
  if (assigned == T && n_res >= least) return true;
  for (int a = 0; a < n_res; ++a) {                             // not enough: clear everything (:378-387)
    const int nd = s.res_nodes[a];
    h_node_pop_job(s, nd, j);
    for (int k = 0; k < T; ++k)
      if (s.tasks[t0 + k].placed_node == nd) { s.tasks[t0 + k].placed_node = -1; s.nodes[nd].n_placed_tasks -= 1; h_node_release(s, nd, t0 + k, false); }
  }
  n_res = 0;
  return false;
}
GS_HD bool h_place(HSim &s, int j, int &n_res) {               // placement_algorithms[--scheme] (algorithm.py:182-187)
  return s.placement == GS_HPLACE_YARN ? h_yarn_placement(s, j, n_res) : h_placement(s, j, n_res);
}

// ---- queues: heapq over Job.__lt__ (base_factory.py:7-11) for horus / horus+, plain list for fifo / gandiva
GS_HD bool h_job_lt(const HSim &s, int a, int b) { return s.jobs[a].util_avg != 0.0 ? s.jobs[a].util_avg < s.jobs[b].util_avg : false; }
GS_HD bool h_is_pq(const HSim &s) { return s.schedule == GS_HSCHED_HORUS || s.schedule == GS_HSCHED_HORUS_PLUS; }
GS_HD int h_queued(const HSim &s) { int t = 0; for (int q = 0; q < s.nq; ++q) t += s.qn[q]; return t; }
GS_HD void h_queue_insert(HSim &s, int q, int j, int pos) {   // JobQueueManager.insert (job_queue_manager.py:146-154)
  int *h = s.queue + (long long)q * (s.n + 1);
  s.credits[q] = s.credits[q] + 1;
  if (h_is_pq(s)) {
    int at = s.qn[q]++;
    while (at > 0) { const int parent = (at - 1) >> 1; if (h_job_lt(s, j, h[parent])) { h[at] = h[parent]; at = parent; } else break; }
    h[at] = j;
    return;
  }
  if (pos > s.qn[q]) pos = s.qn[q];
  for (int i = s.qn[q]; i > pos; --i) h[i] = h[i - 1];
  h[pos] = j; s.qn[q]++;
}
GS_HD int h_queue_pop(HSim &s, int q) {                       // job_queue_manager.py:129-135
  int *h = s.queue + (long long)q * (s.n + 1);
  int &n = s.qn[q];
  if (!h_is_pq(s)) { const int j = h[0]; --n; for (int i = 0; i < n; ++i) h[i] = h[i + 1]; return j; }
  const int last = h[--n];
  if (n == 0) return last;
  const int ret = h[0];
  int pos = 0, child = 1;
  while (child < n) {
    const int right = child + 1;
    if (right < n && !h_job_lt(s, h[child], h[right])) child = right;
    h[pos] = h[child]; pos = child; child = 2 * pos + 1;
  }
  while (pos > 0) { const int parent = (pos - 1) >> 1; if (h_job_lt(s, last, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = last;
  return ret;
}
GS_HD void h_sort_ints(int *a, int n) {                    // heapsort (pending times for the medians)
  for (int i = n / 2 - 1; i >= 0; --i) { int r = i, v = a[r]; for (;;) { int c = 2 * r + 1; if (c >= n) break; if (c + 1 < n && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; } a[r] = v; }
  for (int e = n - 1; e > 0; --e) { int v = a[e]; a[e] = a[0]; int r = 0; for (;;) { int c = 2 * r + 1; if (c >= e) break; if (c + 1 < e && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; } a[r] = v; }
}
// JobQueueManager.update_credits (job_queue_manager.py:115-127): median pending time x queue length
GS_HD void h_update_credits(HSim &s) {
  for (int q = 0; q < s.nq; ++q) {
    const int n = s.qn[q];
    if (n == 0) { s.credits[q] = 0; continue; }
    const int *h = s.queue + (long long)q * (s.n + 1);
    for (int i = 0; i < n; ++i) s.work[i] = s.js[h[i]].pending;
    h_sort_ints(s.work, n);
    double mp = (n & 1) ? (double)s.work[n / 2] : ((double)s.work[n / 2 - 1] + (double)s.work[n / 2]) / 2.0;   // np.median
    if (mp < 0) mp = 0;
    s.credits[q] = mp < 1 ? (double)n : H_MUL(mp, (double)n);
  }
}

// ---- horus+: every insert without explicit queue positions re-clusters ALL queued jobs (jobs_manager.py:93-139)
GS_HD double h_job_score(const HSim &s, int j) {            // transform_to_dist (core/jobs/utils.py:14-22)
  const HJob &x = s.jobs[j];
  double sc = (double)x.ntasks;
  sc = H_ADD(sc, x.util_avg); sc = H_ADD(sc, (double)x.gpc); sc = H_ADD(sc, (double)x.gpus); sc = H_ADD(sc, x.util_max);
  sc = H_ADD(sc, x.mem_avg_mib); sc = H_ADD(sc, (double)x.mem_b / 1048576.0);
  return sc;
}
GS_HD double h_fabs(double v) { return v < 0 ? -v : v; }
GS_HD double h_job_dist(const HSim &s, int a, int b) {      // job_dist (core/jobs/utils.py:4-12)
  const HJob &x = s.jobs[a], &y = s.jobs[b];
  const int dt = x.ntasks - y.ntasks, dg = x.gpc - y.gpc, du = x.gpus - y.gpus;
  double sc = (double)(dt < 0 ? -dt : dt);
  sc = H_ADD(sc, h_fabs(H_ADD(x.util_avg, -y.util_avg))); sc = H_ADD(sc, (double)(dg < 0 ? -dg : dg)); sc = H_ADD(sc, (double)(du < 0 ? -du : du));
  sc = H_ADD(sc, h_fabs(H_ADD(x.util_max, -y.util_max))); sc = H_ADD(sc, h_fabs(H_ADD(x.mem_avg_mib, -y.mem_avg_mib)));
  sc = H_ADD(sc, h_fabs(H_ADD((double)x.mem_b / 1048576.0, -((double)y.mem_b / 1048576.0))));
  return sc;
}
GS_HD double h_block_sum(const double *a, int n) {          // numpy's add.reduce kernel for n <= 128
  if (n < 8) { double r = 0.; for (int i = 0; i < n; ++i) r = H_ADD(r, a[i]); return r; }
  double r[8]; int i;
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] = H_ADD(r[k], a[i + k]);
  double res = H_ADD(H_ADD(H_ADD(r[0], r[1]), H_ADD(r[2], r[3])), H_ADD(H_ADD(r[4], r[5]), H_ADD(r[6], r[7])));
  for (; i < n; ++i) res = H_ADD(res, a[i]);
  return res;
}
GS_HD double h_pairwise_sum(const double *a, int n) {       // numpy pairwise summation, recursion unrolled with a small stack
  if (n <= 128) return h_block_sum(a, n);
  // the recursion splits [lo, lo+len) at n2 = (len / 2) rounded down to a multiple of 8 and adds left + right
  int lo_st[40], len_st[40], state[40]; double acc[40];
  int sp = 0; lo_st[0] = 0; len_st[0] = n; state[0] = 0; acc[0] = 0.0;
  double ret = 0.0;
  for (;;) {
    if (len_st[sp] <= 128) { ret = h_block_sum(a + lo_st[sp], len_st[sp]); if (sp == 0) return ret; --sp; }
    else if (state[sp] == 0) { int n2 = len_st[sp] / 2; n2 -= n2 % 8; state[sp] = 1; lo_st[sp + 1] = lo_st[sp]; len_st[sp + 1] = n2; state[sp + 1] = 0; ++sp; continue; }
    if (state[sp] == 1) { acc[sp] = ret; int n2 = len_st[sp] / 2; n2 -= n2 % 8; state[sp] = 2; lo_st[sp + 1] = lo_st[sp] + n2; len_st[sp + 1] = len_st[sp] - n2; state[sp + 1] = 0; ++sp; continue; }
    if (state[sp] == 2) { ret = H_ADD(acc[sp], ret); if (sp == 0) return ret; --sp; }
  }
}
// clusterize (core/jobs/utils.py:36-67) over km_all[0..m): fills km_assign
GS_HD void h_clusterize(HSim &s, int m) {
  const int k = s.nq;
  int cent[H_MAXQ];
  for (int c = 0; c < k; ++c) cent[c] = s.km_all[h_below(s, m)];                   // np.random.randint(len(jobs), size=k)
  for (int i = 0; i < m; ++i) { s.km_assign[i] = -1; s.km_old[i] = -1; }
  int iter = 0;
  while (iter < 1000 && s.status == 0) {
    if (iter != 0) { bool same = true; for (int i = 0; i < m; ++i) same &= (s.km_assign[i] == s.km_old[i]); if (same) break; }
    for (int i = 0; i < m; ++i) s.km_old[i] = s.km_assign[i];
    iter += 1;
    for (int i = 0; i < m; ++i) {
      int bi = 0; double bd = 0;
      for (int c = 0; c < k; ++c) { const double d = h_job_dist(s, s.km_all[i], cent[c]); if (c == 0 || d < bd) { bd = d; bi = c; } }   // np.argmin
      s.km_assign[i] = bi;
    }
    for (int c = 0; c < k; ++c) {
      int cnt = 0;
      for (int i = 0; i < m; ++i) if (s.km_assign[i] == c) s.km_score[cnt++] = h_job_score(s, s.km_all[i]);
      if (cnt > 0) {
        const double mean = h_pairwise_sum(s.km_score, cnt) / (double)cnt;
        const double target = (double)(long long)mean;                               // .astype(int)
        int best = -1; double bscore = 99999999999.0;
        for (int i = 0; i < m; ++i) if (s.km_assign[i] == c) { const double t = h_fabs(H_ADD(h_job_score(s, s.km_all[i]), -target)); if (t < bscore) { best = s.km_all[i]; bscore = t; } }
        cent[c] = best;
      } else cent[c] = s.km_all[h_below(s, m)];                                      // np.random.choice(len(jobs))
    }
  }
}
// JobsManager.insert without queue positions under horus+ (jobs_manager.py:114-139): ALL queued jobs are popped
// (queue by queue, heap order), the new jobs [first_new, first_new + n_new) go behind them, k-means assigns queues
GS_HD void h_insert_reclustered(HSim &s, int first_new, int n_new) {
  int m = 0;
  for (int q = 0; q < s.nq; ++q) { const int cnt = s.qn[q]; for (int i = 0; i < cnt; ++i) s.km_all[m++] = h_queue_pop(s, q); }
  for (int i = 0; i < n_new; ++i) s.km_all[m++] = first_new + i;
  if (m == 0) return;
  h_clusterize(s, m);
  for (int i = 0; i < m; ++i) h_queue_insert(s, s.km_assign[i], s.km_all[i], i);
}

GS_HD int h_time_processed(const HSim &s, int j) { const HJob &jb = s.jobs[j]; int m = 0; for (int k = 0; k < jb.ntasks; ++k) { const int v = s.tasks[jb.first_task + k].time_processed; if (v > m) m = v; } return m; }
GS_HD double h_get_duration(const HSim &s, int j) { const HJob &jb = s.jobs[j]; double m = 0; for (int k = 0; k < jb.ntasks; ++k) { const double v = s.tasks[jb.first_task + k].duration; if (v > m) m = v; } return m > jb.duration ? m : jb.duration; }
GS_HD void h_running_remove(HSim &s, int j) {
  int w = 0;
  for (int i = 0; i < s.nrun; ++i) if (s.running[i] != j) s.running[w++] = s.running[i];
  s.nrun = w; s.js[j].in_running = 0;
}
// Scheduler.add_to_running -> Node.execute_job -> Job.try_execute (schedule.py:159-162, node.py:164-188, job.py:153-169)
GS_HD void h_start_job(HSim &s, int j, int nn) {
  const HJob &jb = s.jobs[j]; HJobState &st = s.js[j];
  for (int a = 0; a < nn; ++a) {
    const int nd = s.res_nodes[a];
    for (int q = 0; q < st.tro_n; ++q) {
      const int t = jb.first_task + s.tro_order[jb.first_task + q];
      if (s.tro_node[t] != nd) continue;
      HTask &tk = s.tasks[t];
      if (tk.placed_node == nd) { tk.placed_node = -1; s.nodes[nd].n_placed_tasks -= 1; }
      tk.running = 1;
      if (tk.run_node != nd) { tk.run_node = nd; s.nodes[nd].n_running += 1; }
    }
    int cnt = 0;
    for (int k = 0; k < jb.ntasks; ++k) cnt += (s.tasks[jb.first_task + k].running && !s.tasks[jb.first_task + k].finished);
    if (cnt == jb.ntasks) {
      st.start = s.delta; st.migration += 1; st.running = 1;
      if (!st.in_running) { st.in_running = 1; s.running[s.nrun++] = j; }
    }
  }
}
// JobsManager.preempt (jobs_manager.py:150-187)
GS_HD void h_preempt(HSim &s, int j) {
  const HJob &jb = s.jobs[j]; HJobState &st = s.js[j];
  h_running_remove(s, j);
  for (int q = 0; q < st.tro_n; ++q) {
    const int t = jb.first_task + s.tro_order[jb.first_task + q], nd = s.tro_node[t];
    bool first = true;
    for (int p2 = 0; p2 < q; ++p2) first &= (s.tro_node[jb.first_task + s.tro_order[jb.first_task + p2]] != nd);
    if (first) h_node_pop_job(s, nd, j);
    if (s.tasks[t].run_node == nd) { s.tasks[t].run_node = -1; s.nodes[nd].n_running -= 1; h_node_release(s, nd, t, true); }
  }
  st.running = 0; st.pending = 0;
  for (int k = 0; k < jb.ntasks; ++k) s.tasks[jb.first_task + k].running = 0;
  s.events += 1;
  h_queue_insert(s, 0, j, 0);          // gandiva only: one plain list
}

// ---- Scheduler.start (schedule.py:178-213) in pieces, shared by the scalar driver (h_run) and the warp-cooperative one
// gen_jobs; false = the loop is over (done, or out of rows)
GS_HD bool h_tick_begin(HSim &s) {
  if (!(s.current_remaining + s.running_jobs > 0)) { s.done = 1; return false; }
  if (s.ticks >= s.rows_cap) { s.status = GS_ERR_CAPACITY; return false; }
  const int first_new = s.p;                               // rows with normalized_time <= delta (jobs_manager.py:228-241)
  while (s.p < s.n && s.jobs[s.p].arrive <= s.delta) { ++s.p; ++s.events; }
  if (s.schedule == GS_HSCHED_HORUS_PLUS) h_insert_reclustered(s, first_new, s.p - first_new);   // every tick, even without arrivals
  else for (int j = first_new; j < s.p; ++j) h_queue_insert(s, 0, j, j - first_new);
  return true;
}
// _schedule (schedule.py:39-58) up to the placement attempts: the candidate jobs are s.look[0..la_n)
GS_HD bool h_sched_setup(HSim &s) {
  s.la_n = 0; s.la_i = 0; s.la_pos = -1; s.la_nres = 0;
  if (h_queued(s) <= 0) return false;
  int free_nodes = 0;
  for (int nd = 0; nd < s.M; ++nd) free_nodes += h_node_is_free(s, nd);
  if (free_nodes < 1) return false;
  if (s.schedule == GS_HSCHED_HORUS || s.schedule == GS_HSCHED_HORUS_PLUS) {       // schedule_horus / schedule_horus_plus (algorithm.py:204-290)
    const bool plus = s.schedule == GS_HSCHED_HORUS_PLUS;
    const int qd = h_queued(s);
    int min_k = s.num_buffer < qd ? s.num_buffer : qd;
    if (min_k < 0) min_k = 0;
    for (int i = 0; i < min_k; ++i) {
      int qi = 0;
      if (plus) {                                           // the queue with the most credit (leaky bucket)
        h_update_credits(s);
        for (int q = 1; q < s.nq; ++q) if (s.credits[q] > s.credits[qi]) qi = q;              // np.argmax: first maximum
      }
      s.look[i] = h_queue_pop(s, qi); s.look_q[i] = qi;
    }
    s.la_n = min_k;
  } else { s.look[0] = s.queue[0]; s.look_q[0] = 0; s.la_n = 1; }                   // schedule_fifo: the head, popped on success
  return true;
}
// after the attempts: la_pos = index of the job that was placed (or -1), la_nres = its node count
GS_HD void h_sched_finish(HSim &s) {
  int placed = -1;
  if (s.schedule == GS_HSCHED_HORUS || s.schedule == GS_HSCHED_HORUS_PLUS) {
    int min_k = s.la_n;
    if (s.la_pos >= 0) { placed = s.look[s.la_pos]; for (int i = s.la_pos; i + 1 < min_k; ++i) { s.look[i] = s.look[i + 1]; s.look_q[i] = s.look_q[i + 1]; } min_k -= 1; }
    for (int i = 0; i < min_k; ++i) h_queue_insert(s, s.look_q[i], s.look[i], i);               // back to the queue they came from
  } else if (s.la_pos >= 0) { placed = s.look[0]; (void)h_queue_pop(s, 0); }
  if (placed >= 0) { h_start_job(s, placed, s.la_nres); s.events += 1; }
}
// the rest of the tick before the statistics row: aging, completions, time slicing
GS_HD void h_tick_mid(HSim &s) {
    s.current_remaining = s.n - s.p;
    s.delta += 1;
    // JobsManager.step (jobs_manager.py:141-148)
    for (int q = 0; q < s.nq; ++q) { const int *h = s.queue + (long long)q * (s.n + 1); for (int i = 0; i < s.qn[q]; ++i) s.js[h[i]].pending += 1; }
    for (int i = 0; i < s.nrun; ++i) {
      const int j = s.running[i];
      if (!s.js[j].running) continue;
      const HJob &jb = s.jobs[j];
      for (int k = 0; k < jb.ntasks; ++k) if (s.tasks[jb.first_task + k].running) s.tasks[jb.first_task + k].time_processed += 1;
    }
    if (s.schedule == GS_HSCHED_HORUS_PLUS) h_update_credits(s);
    // release_finished_jobs (schedule.py:136-157)
    int nf = 0;
    for (int i = 0; i < s.nrun; ++i) { const int j = s.running[i]; if (!((double)h_time_processed(s, j) < h_get_duration(s, j))) s.work[nf++] = j; }
    for (int f = 0; f < nf; ++f) {
      const int j = s.work[f]; const HJob &jb = s.jobs[j]; HJobState &st = s.js[j];
      for (int q = 0; q < st.tro_n; ++q) {
        const int t = jb.first_task + s.tro_order[jb.first_task + q], nd = s.tro_node[t];
        HTask &tk = s.tasks[t];
        if (tk.run_node == nd) { tk.run_node = -1; s.nodes[nd].n_running -= 1; }
        if (!tk.finished) { tk.finished = 1; st.tasks_finished += 1; }
        h_node_release(s, nd, t, true);
        if (!st.finished && st.tasks_finished == jb.ntasks) {
          st.running = 0; st.finished = 1; st.end = s.delta;
          h_running_remove(s, j);
          s.fin[s.nfin++] = j; s.events += 1;
        }
      }
    }
    s.running_jobs = s.nrun;
    // plugin: gandiva time slicing (algorithm.py:420-440)
    if (s.schedule == GS_HSCHED_GANDIVA && s.qn[0] > 0) {
      int nt = 0;
      for (int i = 0; i < s.nrun; ++i) { const int tp = h_time_processed(s, s.running[i]); if (tp > 1 && tp % 100 == 0) s.work[nt++] = s.running[i]; }
      for (int i = 0; i < nt; ++i) h_preempt(s, s.work[i]);
    }
}
// _construct_info (schedule.py:95-133).  from_arrays: the per-device utilisation (sc_cost) and "is a numpy array" flag
// (sc_off) were computed by the lanes (h_coop_stats); otherwise they are sampled here, device by device.
GS_HD void h_stats_row(HSim &s, bool from_arrays) {
    gs_tick_row row;
    row.now = s.delta; row.idle_nodes = 0; row.busy_nodes = 0; row.busy_gpus = 0; row.idle_gpus = 0;
    row.pend_sum = 0; row.pend_max = 0; row.pend_med_lo = 0; row.pend_med_hi = 0; row.reserved = 0;
    double usum = 0.0; int uarr = 0; long long msum = 0;
    for (int nd = 0; nd < s.M; ++nd) {
      const HNode &nn = s.nodes[nd];
      if (nn.n_running + nn.n_placed_tasks + nn.n_placed_jobs == 0) row.idle_nodes += 1; else row.busy_nodes += 1;
      for (int d = 0; d < s.G; ++d) {
        const HDev &dv = h_dev(s, nd, d);
        if (dv.nt == 0) { row.idle_gpus += 1; continue; }
        row.busy_gpus += 1;
        int a = 0; double u;
        if (from_arrays) { u = s.sc_cost[nd * s.G + d]; a = s.sc_off[nd * s.G + d]; } else u = h_dev_util(s, dv, &a);
        usum = H_ADD(usum, u); uarr |= a;                    // summed in device order: the order is part of the result
        msum += h_dev_mem(s, dv);
      }
    }
    row.mem_busy_bytes = msum;
    const int queued = h_queued(s);
    row.running = s.nrun; row.queued = queued; row.finished = s.nfin;
    {
      int m = 0;
      for (int q = 0; q < s.nq; ++q) { const int *h = s.queue + (long long)q * (s.n + 1); for (int i = 0; i < s.qn[q]; ++i) { const int pd = s.js[h[i]].pending; s.work[m++] = pd; row.pend_sum += pd; if (pd > row.pend_max) row.pend_max = pd; } }
      if (m > 0) { h_sort_ints(s.work, m); row.pend_med_lo = s.work[(m - 1) / 2]; row.pend_med_hi = s.work[m / 2]; }
    }
    s.rows[s.ticks] = row;
    s.util[s.ticks] = usum / (double)(row.idle_gpus + row.busy_gpus); s.util_arr[s.ticks] = (unsigned char)uarr;
    s.ticks += 1;
    if (!(s.current_remaining + s.running_jobs > 0)) s.done = 1;
}
GS_HD void h_tick_end(HSim &s) { h_tick_mid(s); h_stats_row(s, false); }
GS_HD void h_write_records(HSim &s) {
  if (s.done && s.status == 0)
    for (int j = 0; j < s.n; ++j) {
      gs_horus_job_rec r; const HJobState &st = s.js[j];
      r.start = st.start; r.end = st.end; r.jct = h_time_processed(s, j); r.preempt = st.migration;
      r.original = s.jobs[j].duration; r.actual = h_get_duration(s, j);
      s.recs[j] = r;
    }
}

// scalar driver: one simulation per thread
GS_HD void h_run(HSim &s, long long max_ticks) {
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;
  while (!s.done && s.status == 0 && budget > 0) {
    if (!h_tick_begin(s)) break;
    if (h_sched_setup(s)) {
      for (int i = 0; i < s.la_n; ++i) { int nres = 0; if (h_place(s, s.look[i], nres)) { s.la_pos = i; s.la_nres = nres; break; } }
      h_sched_finish(s);
    }
    h_tick_end(s);
    budget -= 1;
  }
  h_write_records(s);
}

// ======================================================================================================================
// Warp-cooperative driver (one simulation per WARP).  Over 90 % of the samples are drawn while scoring
// (task x node x device x tasks-on-device, for up to num_buffer candidate jobs every tick), so that loop is spread
// over the lanes: lane 0 runs the scalar simulation (h_coop_advance) until a candidate job needs scoring, posts a
// request, and all 32 lanes execute it.  Nothing moves while a job is scored, so what each device consumes is known
// up front: PREP computes per device whether it fits and how many samples it draws; lane 0 turns the counts into
// offsets (one prefix sum) and, per task k, SCORE lets every lane read ITS devices' samples at
// base + k * total + offset -- the stream order of the sequential code -- and compute their cost.  Lane 0 then
// reduces per node and feeds the heap exactly as h_placement does.  Both stream forms are indexable (h_sample_at): the
// standard-normal values by offset, the raw words through the per-class index of accepted positions.
//
// The phases are written with H_FOR_LANE_ITEMS / H_SYNC so that the host build (tests/emu) executes the same
// statements with the lane loop run sequentially: the index arithmetic is checked on the CPU against the oracle.
#ifdef __CUDA_ARCH__
#define H_FOR_LANE_ITEMS(i, n) for (int i = (int)(threadIdx.x & 31); i < (n); i += 32)
#define H_SYNC() __syncwarp()
#else
#define H_FOR_LANE_ITEMS(i, n) for (int i = 0; i < (n); ++i)
#define H_SYNC()
#endif
enum { H_REQ_DONE = 0, H_REQ_PREP = 1, H_REQ_SCORE = 2, H_REQ_STATS = 3 };
enum { H_PH_TICK = 0, H_PH_JOB = 1, H_PH_PREPPED = 2, H_PH_SCORED = 3, H_PH_STATS = 4 };

// PREP: per device, does the task fit and how many samples does scoring it draw (its running tasks)
GS_HD void h_coop_prep(HSim &s) {
  const int t = s.jobs[s.sc_job].first_task;               // all tasks of a job are alike for these predicates
  H_FOR_LANE_ITEMS(i, s.M * s.G) {
    const int nd = i / s.G;
    const bool node_ok = h_node_is_free(s, nd) && s.cpu_cap - s.nodes[nd].cpu_used - H_TASK_CPU >= 0 && s.mem_cap - s.nodes[nd].mem_used - H_TASK_MEM >= 0;
    const HDev &dv = s.devs[i];
    s.sc_cnt[i] = (node_ok && h_dev_can_fit(s, dv, t)) ? dv.nt : -1;          // -1: not scored
  }
}
// Device.get_current_utilization with the samples at a known index of the stream
GS_HD double h_dev_util_at(const HSim &s, const HDev &d, long long pos) {
  double u = 0.0;
  for (int i = 0; i < d.nt; ++i) {
    const HTask &o = s.tasks[d.t[i]];
    const double x = H_ADD(o.util_avg, H_MUL(o.half_spread, h_sample_at(s, pos + i)));
    if (x < 100.0) u = H_ADD(u, x); else u = H_ADD(u, 100.0);
    if (100.0 < u) u = 100.0;
  }
  return u;
}
// SCORE: the cost of every scored device for task sc_k (horus.py:6-56)
GS_HD void h_coop_score(HSim &s) {
  const HJob &jb = s.jobs[s.sc_job];
  const double cap_mib = (double)(s.cap_b >> 20), tm = (double)jb.mem_b / 1048576.0;
  const long long base = (long long)s.sc_k * s.sc_total;        // sample index relative to the (unchanged) stream state
  H_FOR_LANE_ITEMS(i, s.M * s.G) {
    if (s.sc_cnt[i] < 0) continue;
    const HDev &dv = s.devs[i];
    const double cur = (double)h_dev_mem(s, dv) / 1048576.0;
    const double util = h_dev_util_at(s, dv, base + s.sc_off[i]);
    double cost;
    if (s.scheme == GS_HSCORE_HORUS) {
      const double mem_cost = (cur + tm) / cap_mib;
      const double util_cost = h_polyval(H_ADD(util, jb.util_avg));
      cost = H_ADD(H_ADD(H_MUL(mem_cost, 0.5), H_MUL(util_cost, 0.5)), (double)dv.nt);
    } else {
      const double mem_cost = cur + tm / cap_mib;
      cost = H_ADD(H_ADD(H_MUL(mem_cost, 0.5), util / 100), (double)dv.nt);
    }
    s.sc_cost[i] = cost;
  }
}
// STATS: Device.get_current_utilization of every busy device for the statistics row (offsets in sc_off, from lane 0);
// leaves the value in sc_cost and the "is a numpy array" flag in sc_off
GS_HD void h_coop_stats(HSim &s) {
  H_FOR_LANE_ITEMS(i, s.M * s.G) {
    const HDev &d = s.devs[i];
    if (d.nt == 0) continue;
    const long long pos = s.sc_off[i];
    double u = 0.0; int arr = 0;
    for (int t = 0; t < d.nt; ++t) {
      const HTask &o = s.tasks[d.t[t]];
      const double x = H_ADD(o.util_avg, H_MUL(o.half_spread, h_sample_at(s, pos + t)));
      if (x < 100.0) { u = H_ADD(u, x); arr = 1; } else { u = H_ADD(u, 100.0); }
      if (100.0 < u) { u = 100.0; arr = 0; }
    }
    s.sc_cost[i] = u; s.sc_off[i] = arr;
  }
}
// lane 0, end of a tick in the cooperative driver: everything before the row, then the sample offsets of the row
GS_HD int h_coop_tick_end(HSim &s) {
  h_tick_mid(s);
  int total = 0;
  for (int i = 0; i < s.M * s.G; ++i) { s.sc_off[i] = total; total += s.devs[i].nt; }
  s.sc_total = total;
  if (!h_samples_available(s, total)) { s.status = GS_ERR_CAPACITY; s.phase = H_PH_TICK; return H_REQ_DONE; }
  s.phase = H_PH_STATS;
  return H_REQ_STATS;
}
// lane 0: run the simulation up to the next cooperative request (or the end of this launch)
GS_HD int h_coop_advance(HSim &s) {
  for (;;) {
    if (s.phase == H_PH_TICK) {
      if (s.done || s.status != 0 || s.budget <= 0) return H_REQ_DONE;
      if (!h_tick_begin(s)) return H_REQ_DONE;
      if (h_sched_setup(s)) { s.phase = H_PH_JOB; continue; }
      return h_coop_tick_end(s);
    }
    if (s.phase == H_PH_STATS) {                             // the lanes have sampled the busy devices
      h_stats_row(s, true);
      h_skip_samples(s, s.sc_total);
      s.budget -= 1; s.phase = H_PH_TICK;
      continue;
    }
    if (s.phase == H_PH_JOB) {
      if (s.la_pos >= 0 || s.la_i >= s.la_n) { h_sched_finish(s); return h_coop_tick_end(s); }
      const int j = s.look[s.la_i];
      if (s.placement == GS_HPLACE_YARN) {                   // nothing to score: the scalar placement
        int nres = 0;
        if (h_place(s, j, nres)) { s.la_pos = s.la_i; s.la_nres = nres; } else s.la_i += 1;
        continue;
      }
      s.sc_job = j; s.sc_k = 0; s.sc_hn = 0; s.phase = H_PH_PREPPED;
      return H_REQ_PREP;
    }
    if (s.phase == H_PH_PREPPED) {                           // counts -> offsets; the samples of one task span sc_total
      int total = 0;
      for (int i = 0; i < s.M * s.G; ++i) { s.sc_off[i] = total; if (s.sc_cnt[i] > 0) total += s.sc_cnt[i]; }
      s.sc_total = total;
      if (!h_samples_available(s, (long long)total * s.jobs[s.sc_job].ntasks)) { s.status = GS_ERR_CAPACITY; s.phase = H_PH_TICK; return H_REQ_DONE; }
      s.phase = H_PH_SCORED;
      return H_REQ_SCORE;
    }
    // H_PH_SCORED: per node the cheapest scored device -> heap (algorithm.py:50-66), then the next task or the trials
    const HJob &jb = s.jobs[s.sc_job];
    for (int nd = 0; nd < s.M; ++nd) {
      double min_cost = 999.0; bool any = false;
      for (int d = 0; d < s.G; ++d) { const int i = nd * s.G + d; if (s.sc_cnt[i] < 0) continue; any = true; if (s.sc_cost[i] < min_cost) min_cost = s.sc_cost[i]; }
      if (!any) continue;                                    // node.can_fit(t, pack=True) is false
      HCand x; x.node = nd; x.pad = 0; x.min_score = min_cost;
      h_heap_push(s.heap, s.sc_hn, x);
      if (s.sc_hn > jb.gpus) h_heap_pop(s.heap, s.sc_hn);
    }
    s.sc_k += 1;
    if (s.sc_k < jb.ntasks) return H_REQ_SCORE;
    h_skip_samples(s, (long long)s.sc_total * jb.ntasks);
    int nres = 0;
    if (h_placement_finish(s, s.sc_job, s.sc_hn, nres)) { s.la_pos = s.la_i; s.la_nres = nres; } else s.la_i += 1;
    s.phase = H_PH_JOB;
  }
}
// host form of the cooperative driver (the kernel interleaves the same calls with __syncwarp)
GS_HD void h_run_coop(HSim &s, long long max_ticks) {
  s.budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;
  for (;;) {
    const int req = h_coop_advance(s);
    if (req == H_REQ_DONE) break;
    if (req == H_REQ_PREP) h_coop_prep(s); else if (req == H_REQ_SCORE) h_coop_score(s); else h_coop_stats(s);
  }
  h_write_records(s);
}
