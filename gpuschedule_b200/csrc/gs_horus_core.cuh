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

struct HJob {            // one trace row (read-only), 56 bytes
  int arrive, gpus, gpc, ntasks;
  int first_task, pad;
  long long mem_b;
  double util_avg, util_max, duration;
};
struct HJobState {
  int pending, start, end, migration, tasks_finished, tro_n;
  unsigned char running, finished, in_running, pad;
};
struct HTask {
  double duration, original;           // Task.duration / original_duration (job.py:33-34)
  int job, time_processed;
  int placed_node, run_node;           // membership in Node.placed_tasks / Node.running_tasks
  unsigned char interfered, running, finished, pad;
};
struct HDev { int nt; int t[H_DEV_SLOTS]; };        // Device.running_tasks in insertion order
struct HNode { int cpu_used, mem_used, n_running, n_placed_tasks, n_placed_jobs; };
struct HCand { double min_score; int node, pad; };

struct HSim {
  // ---- configuration
  int M, G, S, P, cpu_cap, mem_cap, scheme, schedule, num_buffer, n, maxg, pjw;
  long long cap_b;
  // ---- trace and simulation state
  const HJob *jobs; HJobState *js; HTask *tasks;
  int *tro_node, *tro_order;            // Job.tasks_running_on: value per task / key insertion order
  HNode *nodes; HDev *devs;
  unsigned long long *pj_bits;          // Node.placed_jobs membership, bit (job, node)
  int *queue, *running, *fin;
  int qn, nrun, nfin, pad0;
  // ---- scratch
  int *look, *work, *res_nodes, *map_node, *map_order, *map_n, *ok, *distinct;
  HCand *heap;
  // ---- the sampled stream
  const double *gauss; long long gauss_n, gauss_pos;
  // ---- results
  gs_tick_row *rows; double *util; unsigned char *util_arr; gs_horus_job_rec *recs; long long rows_cap;
  // ---- loop state (persisted between launches)
  int delta, p, status, done;
  long long ticks, events, current_remaining, running_jobs;
};

GS_HD double h_gauss(HSim &s) {
  if (s.gauss_pos >= s.gauss_n) { s.status = GS_ERR_CAPACITY; return 0.0; }      // host supplies a longer stream and re-runs
  return s.gauss[s.gauss_pos++];
}
GS_HD double h_normal(HSim &s, double loc, double scale) { return H_ADD(loc, H_MUL(scale, h_gauss(s))); }
GS_HD HDev &h_dev(HSim &s, int nd, int d) { return s.devs[(long long)nd * s.G + d]; }
GS_HD long long h_task_mem(const HSim &s, int t) { return s.jobs[s.tasks[t].job].mem_b; }
GS_HD bool h_node_is_free(const HSim &s, int nd) { return s.cpu_cap - s.nodes[nd].cpu_used > 0 || s.mem_cap - s.nodes[nd].mem_used > 0; }   // node.py:57-58

// Device.get_current_memory (device.py:56-62) in bytes: every term is an exact binary fraction of a MiB
GS_HD long long h_dev_mem(const HSim &s, const HDev &d) {
  long long m = 0;
  for (int i = 0; i < d.nt; ++i) { long long x = h_task_mem(s, d.t[i]); if (x > s.cap_b) x = s.cap_b; m += x; if (m > s.cap_b) m = s.cap_b; }
  return m;
}
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
    const HJob &j = s.jobs[s.tasks[d.t[i]].job];
    const double x = h_normal(s, j.util_avg, (j.util_max - j.util_avg) / 2);
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
    for (int i = 0; i < d.nt; ++i) {      // interference samples are drawn; the slowed duration is only logged (:35-37)
      const HJob &j = s.jobs[s.tasks[d.t[i]].job];
      (void)h_normal(s, j.util_avg, (j.util_max - j.util_avg) / 4);
    }
    tk.interfered = 1;
  } else { tk.interfered = 0; tk.duration = tk.original; }
  for (int i = 0; i < d.nt; ++i) if (d.t[i] == t) return true;                   // key already present: position kept
  if (d.nt >= H_DEV_SLOTS) { s.status = GS_ERR_STATE; return true; }
  d.t[d.nt++] = t;
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
GS_HD bool h_node_reserve_task(HSim &s, int nd, int t, bool pack) {
  if (!h_node_can_fit(s, nd, t, pack)) return false;
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
    for (int i = 0; i < dv.nt; ++i) if (dv.t[i] == t) { for (int k = i; k + 1 < dv.nt; ++k) dv.t[k] = dv.t[k + 1]; dv.nt--; break; }
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
GS_HD bool h_placement(HSim &s, int j, int &n_res) {
  const HJob &jb = s.jobs[j];
  HJobState &st = s.js[j];
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
  for (int i = 1; i < hn; ++i) {                          // sorted(nodes_stack, key=min_score): stable
    const HCand x = heap[i]; int k = i - 1;
    while (k >= 0 && heap[k].min_score > x.min_score) { heap[k + 1] = heap[k]; --k; }
    heap[k + 1] = x;
  }
  const int C = hn;
  for (int i = 0; i < C; ++i) {
    int *mn = s.map_node + (long long)i * T, *mo = s.map_order + (long long)i * T;
    int mapped = 0;
    for (int k = 0; k < T; ++k) mn[k] = -1;
    const int cand = heap[i].node;
    for (int k = 0; k < T; ++k)
      if (h_node_reserve_task(s, cand, t0 + k, true)) { h_node_place_job(s, cand, j); mn[k] = cand; mo[mapped++] = k; }
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
            if (h_node_reserve_task(s, nd, t0 + k, true)) { h_node_place_job(s, nd, j); mn[k] = nd; mo[mapped++] = k; }
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
    s.map_n[i] = mapped; s.ok[i] = mapped >= T;
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

// ---- queue: heapq over Job.__lt__ (base_factory.py:7-11) for horus, plain list for fifo / gandiva
GS_HD bool h_job_lt(const HSim &s, int a, int b) { return s.jobs[a].util_avg != 0.0 ? s.jobs[a].util_avg < s.jobs[b].util_avg : false; }
GS_HD void h_queue_insert(HSim &s, int j, int pos) {       // JobQueueManager.insert (job_queue_manager.py:146-154)
  int *h = s.queue;
  if (s.schedule == GS_HSCHED_HORUS) {
    int at = s.qn++;
    while (at > 0) { const int parent = (at - 1) >> 1; if (h_job_lt(s, j, h[parent])) { h[at] = h[parent]; at = parent; } else break; }
    h[at] = j;
    return;
  }
  if (pos > s.qn) pos = s.qn;
  for (int i = s.qn; i > pos; --i) h[i] = h[i - 1];
  h[pos] = j; s.qn++;
}
GS_HD int h_queue_pop(HSim &s) {                           // job_queue_manager.py:129-135
  int *h = s.queue;
  if (s.schedule != GS_HSCHED_HORUS) { const int j = h[0]; --s.qn; for (int i = 0; i < s.qn; ++i) h[i] = h[i + 1]; return j; }
  const int last = h[--s.qn];
  if (s.qn == 0) return last;
  const int ret = h[0];
  int pos = 0, child = 1;
  while (child < s.qn) {
    const int right = child + 1;
    if (right < s.qn && !h_job_lt(s, h[child], h[right])) child = right;
    h[pos] = h[child]; pos = child; child = 2 * pos + 1;
  }
  while (pos > 0) { const int parent = (pos - 1) >> 1; if (h_job_lt(s, last, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = last;
  return ret;
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
  h_queue_insert(s, j, 0);
}
GS_HD void h_sort_ints(int *a, int n) {                    // heapsort (pending times for the median)
  for (int i = n / 2 - 1; i >= 0; --i) { int r = i, v = a[r]; for (;;) { int c = 2 * r + 1; if (c >= n) break; if (c + 1 < n && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; } a[r] = v; }
  for (int e = n - 1; e > 0; --e) { int v = a[e]; a[e] = a[0]; int r = 0; for (;;) { int c = 2 * r + 1; if (c >= e) break; if (c + 1 < e && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; } a[r] = v; }
}

// Scheduler.start (schedule.py:178-213): runs until done, max_ticks or the row buffer is full
GS_HD void h_run(HSim &s, long long max_ticks) {
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;
  while (!s.done && s.status == 0 && budget > 0) {
    if (!(s.current_remaining + s.running_jobs > 0)) { s.done = 1; break; }
    if (s.ticks >= s.rows_cap) { s.status = GS_ERR_CAPACITY; break; }
    // gen_jobs: rows with normalized_time <= delta in trace order (jobs_manager.py:228-241)
    { int pos = 0; while (s.p < s.n && s.jobs[s.p].arrive <= s.delta) { h_queue_insert(s, s.p, pos++); ++s.p; ++s.events; } }
    // _schedule (schedule.py:39-58)
    if (s.qn > 0) {
      int free_nodes = 0;
      for (int nd = 0; nd < s.M; ++nd) free_nodes += h_node_is_free(s, nd);
      if (free_nodes >= 1) {
        int placed = -1, nres = 0;
        if (s.schedule == GS_HSCHED_HORUS) {                                       // schedule_horus (algorithm.py:204-240)
          int min_k = s.num_buffer < s.qn ? s.num_buffer : s.qn;
          if (min_k < 0) min_k = 0;
          for (int i = 0; i < min_k; ++i) s.look[i] = h_queue_pop(s);
          int pos = -1;
          for (int i = 0; i < min_k; ++i) if (h_placement(s, s.look[i], nres)) { pos = i; break; }
          if (pos >= 0) { placed = s.look[pos]; for (int i = pos; i + 1 < min_k; ++i) s.look[i] = s.look[i + 1]; min_k -= 1; }
          for (int i = 0; i < min_k; ++i) h_queue_insert(s, s.look[i], i);
        } else {                                                                   // schedule_fifo (algorithm.py:189-202)
          const int j = s.queue[0];
          if (h_placement(s, j, nres)) { (void)h_queue_pop(s); placed = j; }
        }
        if (placed >= 0) { h_start_job(s, placed, nres); s.events += 1; }
      }
    }
    s.current_remaining = s.n - s.p;
    s.delta += 1;
    // JobsManager.step (jobs_manager.py:141-148)
    for (int i = 0; i < s.qn; ++i) s.js[s.queue[i]].pending += 1;
    for (int i = 0; i < s.nrun; ++i) {
      const int j = s.running[i];
      if (!s.js[j].running) continue;
      const HJob &jb = s.jobs[j];
      for (int k = 0; k < jb.ntasks; ++k) if (s.tasks[jb.first_task + k].running) s.tasks[jb.first_task + k].time_processed += 1;
    }
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
    if (s.schedule == GS_HSCHED_GANDIVA && s.qn > 0) {
      int nt = 0;
      for (int i = 0; i < s.nrun; ++i) { const int tp = h_time_processed(s, s.running[i]); if (tp > 1 && tp % 100 == 0) s.work[nt++] = s.running[i]; }
      for (int i = 0; i < nt; ++i) h_preempt(s, s.work[i]);
    }
    // _construct_info (schedule.py:95-133)
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
        int a = 0; const double u = h_dev_util(s, dv, &a);
        usum = H_ADD(usum, u); uarr |= a;
        msum += h_dev_mem(s, dv);
      }
    }
    row.mem_busy_bytes = msum;
    row.running = s.nrun; row.queued = s.qn; row.finished = s.nfin;
    for (int i = 0; i < s.qn; ++i) { const int pd = s.js[s.queue[i]].pending; s.work[i] = pd; row.pend_sum += pd; if (pd > row.pend_max) row.pend_max = pd; }
    if (s.qn > 0) { h_sort_ints(s.work, s.qn); row.pend_med_lo = s.work[(s.qn - 1) / 2]; row.pend_med_hi = s.work[s.qn / 2]; }
    s.rows[s.ticks] = row;
    s.util[s.ticks] = usum / (double)(row.idle_gpus + row.busy_gpus); s.util_arr[s.ticks] = (unsigned char)uarr;
    s.ticks += 1; budget -= 1;
    if (!(s.current_remaining + s.running_jobs > 0)) s.done = 1;
  }
  if (s.done && s.status == 0)
    for (int j = 0; j < s.n; ++j) {
      gs_horus_job_rec r; const HJobState &st = s.js[j];
      r.start = st.start; r.end = st.end; r.jct = h_time_processed(s, j); r.preempt = st.migration;
      r.original = s.jobs[j].duration; r.actual = h_get_duration(s, j);
      s.recs[j] = r;
    }
}
