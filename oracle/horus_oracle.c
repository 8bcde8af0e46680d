/*
 * horus_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's LIVE utilisation-aware paths:
 *     placement  horus_placement + horus_score / gandiva_score   core/scheduling/algorithm.py:34-180,
 *                                                                 core/scheduling/horus.py:6-56
 *     schedulers schedule_horus (look-ahead of num_buffer jobs)   algorithm.py:204-240
 *                schedule_horus_plus (credit queues + k-means)    algorithm.py:242-290,
 *                                                                 core/jobs/job_queue_manager.py:103-127,
 *                                                                 core/jobs/jobs_manager.py:93-139, core/jobs/utils.py:36-67
 *                schedule_fifo + time_slice_check (gandiva)       algorithm.py:189-202,420-440, jobs_manager.py:150-187
 *     device / node packing rules (<= 4 tasks per device, 500 MiB margin, interference marks)
 *                                                                 infra/device.py:20-76, infra/node.py:64-221
 *     tick loop, completion, statistics row                       core/scheduling/schedule.py:39-213
 *
 * These paths consume numpy's GLOBAL legacy MT19937 stream (np.random.normal in
 * Device.get_current_utilization / add_task, np.random.randint / choice in clusterize), so the
 * restatement carries its own MT19937 + legacy polar gauss + masked-rejection integers and replays the
 * stream draw for draw: results are compared BYTE FOR BYTE with runs of the unmodified reference under
 * numpy.random.seed(S) (tests/golden/make_horus_golden.py -> tests/golden/horus_*).
 *
 * Parity status: PINNED by those fixtures (tests/test_horus_oracle.py) and by tests/golden/fuzz_horus_reference.py
 * (random clusters / traces / score x scheduler combinations against the unmodified reference).  It checks
 * gs_horus_kernel (include/gsched_horus.h) and the host build of the same device functions (tests/emu).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gsched.h"

/* ------------------------------------------------------------------ numpy legacy random stream */
typedef struct { uint32_t mt[624]; int pos; int has_gauss; double gauss; uint64_t draws; } rng_t;

static void rng_seed(rng_t *r, uint32_t seed) {                 /* RandomState.seed(int): init_genrand */
  for (int i = 0; i < 624; i++) { r->mt[i] = seed; seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u; }
  r->pos = 624; r->has_gauss = 0; r->gauss = 0.0; r->draws = 0;
}
static void rng_refill(rng_t *r) {
  uint32_t *mt = r->mt, y; int k;
  for (k = 0; k < 624 - 397; k++) { y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu); mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
  for (; k < 623; k++) { y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu); mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u); }
  y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  r->pos = 0;
}
static uint32_t rng_u32(rng_t *r) {
  if (r->pos == 624) rng_refill(r);
  uint32_t y = r->mt[r->pos++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static double rng_double(rng_t *r) { int32_t a = (int32_t)(rng_u32(r) >> 5), b = (int32_t)(rng_u32(r) >> 6); return (a * 67108864.0 + b) / 9007199254740992.0; }
static double rng_gauss(rng_t *r) {                              /* legacy_gauss: polar method, second value cached */
  r->draws++;
  if (r->has_gauss) { double t = r->gauss; r->gauss = 0.0; r->has_gauss = 0; return t; }
  double f, x1, x2, r2;
  do { x1 = 2.0 * rng_double(r) - 1.0; x2 = 2.0 * rng_double(r) - 1.0; r2 = x1 * x1 + x2 * x2; } while (r2 >= 1.0 || r2 == 0.0);
  f = sqrt(-2.0 * log(r2) / r2);
  r->gauss = f * x1; r->has_gauss = 1;
  return f * x2;
}
static double rng_normal(rng_t *r, double loc, double scale) { return loc + scale * rng_gauss(r); }
static int64_t rng_below(rng_t *r, int64_t n) {                  /* randint(n) / choice(n): masked rejection, 32-bit draws */
  uint64_t max = (uint64_t)(n - 1), mask = max, v;
  if (max == 0) return 0;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
  if (max <= 0xffffffffULL) { while ((v = (rng_u32(r) & mask)) > max) {} }
  else { while ((v = ((((uint64_t)rng_u32(r)) << 32) | rng_u32(r)) & mask) > max) {} }
  return (int64_t)v;
}

/* ------------------------------------------------------------------ state */
enum { HS_SCHEME_HORUS = 0, HS_SCHEME_GANDIVA = 1 };
enum { HS_SCHED_FIFO = 0, HS_SCHED_HORUS = 1, HS_SCHED_HORUS_PLUS = 2, HS_SCHED_GANDIVA = 3 };
#define DEV_MAXT 8
#define TASK_CPU 12
#define TASK_MEM 60
#define MAXQ 16

typedef struct {
  int32_t job;
  double duration, original;                 /* Task.duration / original_duration (job.py:33-34) */
  uint8_t interfered, running, finished;
  int32_t time_processed;
  int32_t placed_node, run_node;             /* membership in Node.placed_tasks / Node.running_tasks */
} htask_t;
typedef struct { int nt; int32_t t[DEV_MAXT]; } hdev_t;         /* Device.running_tasks (ordered) */
typedef struct { int cpu_used, mem_used, n_running, n_placed_tasks, n_placed_jobs, rack; } hnode_t;
typedef struct {
  int32_t gpus, gpc, ntasks, first_task;
  double util_avg, util_max, mem_avg_mib;
  int64_t mem_b;
  double duration;                            /* Job.duration */
  int32_t pending, start, end, migration, tasks_finished, tro_n;
  uint8_t running, finished, in_running;
} hjob_t;

typedef struct {
  int M, G, S, P, cpu_cap, mem_cap, scheme, schedule, num_buffer, nq, yarn;
  int64_t cap_b, n;
  hnode_t *nodes; hdev_t *devs; hjob_t *jobs; htask_t *tasks;
  int32_t *tro_node, *tro_order;              /* Job.tasks_running_on: value per task, key insertion order */
  uint64_t *pj_bits; int pjw;                 /* Node.placed_jobs membership: bit (job, node) */
  int32_t *queue[MAXQ]; int64_t qn[MAXQ]; double credits[MAXQ];
  int32_t *running; int64_t nrun;             /* JobsManager.running_jobs (insertion ordered dict) */
  rng_t rng;
  int64_t events;
} hsim_t;

static inline hdev_t *dev_of(hsim_t *s, int node, int d) { return &s->devs[(size_t)node * s->G + d]; }
static inline int64_t task_mem(const hsim_t *s, int t) { return s->jobs[s->tasks[t].job].mem_b; }
static inline int cpu_free(const hsim_t *s, int nd) { return s->cpu_cap - s->nodes[nd].cpu_used; }
static inline int mem_free(const hsim_t *s, int nd) { return s->mem_cap - s->nodes[nd].mem_used; }
static inline int node_is_free(const hsim_t *s, int nd) { return cpu_free(s, nd) > 0 || mem_free(s, nd) > 0; }   /* node.py:57-58 */

/* Device.get_current_memory  device.py:56-62 (bytes; every term is an exact binary fraction of a MiB) */
static int64_t dev_mem(const hsim_t *s, const hdev_t *d) {
  int64_t m = 0;
  for (int i = 0; i < d->nt; ++i) { int64_t x = task_mem(s, d->t[i]); if (x > s->cap_b) x = s->cap_b; m += x; if (m > s->cap_b) m = s->cap_b; }
  return m;
}
/* Device.can_fit  device.py:67-76 */
static int dev_can_fit(const hsim_t *s, const hdev_t *d, int t) {
  int64_t cur = dev_mem(s, d);
  if (d->nt >= 4) return 0;
  return s->cap_b - (cur + task_mem(s, t)) > ((int64_t)500 << 20);
}
/* Device.get_current_utilization  device.py:48-54; *is_arr = the Python value is a 1-element numpy array */
static double dev_util(hsim_t *s, const hdev_t *d, int *is_arr) {
  double u = 0.0; int arr = 0;
  for (int i = 0; i < d->nt; ++i) {
    const hjob_t *j = &s->jobs[s->tasks[d->t[i]].job];
    double x = rng_normal(&s->rng, j->util_avg, (j->util_max - j->util_avg) / 2);
    if (x < 100.0) { u = u + x; arr = 1; } else { u = u + 100.0; }          /* min(100, array) */
    if (100.0 < u) { u = 100.0; arr = 0; }                                  /* min(util, 100) */
  }
  if (is_arr) *is_arr = arr;
  return u;
}
/* Device.add_task  device.py:20-43 */
static int dev_add_task(hsim_t *s, hdev_t *d, int t, int pack) {
  if (!dev_can_fit(s, d, t)) return 0;
  if (!pack && d->nt > 0) return 0;
  htask_t *tk = &s->tasks[t];
  if (d->nt >= 2) {
    for (int i = 0; i < d->nt; ++i) {          /* the draws are consumed; the slowed duration is only logged (:35-37) */
      const hjob_t *j = &s->jobs[s->tasks[d->t[i]].job];
      (void)rng_normal(&s->rng, j->util_avg, (j->util_max - j->util_avg) / 4);
    }
    tk->interfered = 1;
  } else { tk->interfered = 0; tk->duration = tk->original; }
  for (int i = 0; i < d->nt; ++i) if (d->t[i] == t) return 1;               /* dict key already present: position kept */
  if (d->nt >= DEV_MAXT) abort();
  d->t[d->nt++] = t;
  return 1;
}
/* Node.can_fit  node.py:136-162 */
static int node_can_fit(const hsim_t *s, int nd, int t, int pack) {
  if (cpu_free(s, nd) - TASK_CPU < 0 || mem_free(s, nd) - TASK_MEM < 0) return 0;
  if (!pack) {
    int idle = 0;
    for (int d = 0; d < s->G; ++d) idle += (s->devs[(size_t)nd * s->G + d].nt == 0);
    return idle - s->jobs[s->tasks[t].job].gpc >= 0;
  }
  for (int d = 0; d < s->G; ++d) if (dev_can_fit(s, &s->devs[(size_t)nd * s->G + d], t)) return 1;
  return 0;
}
/* Node.try_reserve_and_placed_task  node.py:190-211 (a partial placement keeps what it took, like the original) */
static int node_reserve_task(hsim_t *s, int nd, int t, int pack) {
  if (!node_can_fit(s, nd, t, pack)) return 0;
  s->nodes[nd].cpu_used += TASK_CPU; s->nodes[nd].mem_used += TASK_MEM;
  int need = s->jobs[s->tasks[t].job].gpc;
  for (int d = 0; d < s->G; ++d) {
    if (need <= 0) break;
    if (dev_add_task(s, dev_of(s, nd, d), t, pack)) need -= 1;
  }
  if (need == 0 && s->tasks[t].placed_node != nd) {
    if (s->tasks[t].placed_node >= 0) s->nodes[s->tasks[t].placed_node].n_placed_tasks -= 0;   /* other node keeps its own entry */
    s->tasks[t].placed_node = nd; s->nodes[nd].n_placed_tasks += 1;
  }
  return need == 0;
}
/* Node.try_reserve_and_placed_job(job, False)  node.py:213-232: placed_jobs[job_id] = job */
static void node_place_job(hsim_t *s, int nd, int j) {
  uint64_t *w = &s->pj_bits[(size_t)j * s->pjw + (nd >> 6)], b = 1ull << (nd & 63);
  if (!(*w & b)) { *w |= b; s->nodes[nd].n_placed_jobs += 1; }
}
static void node_pop_job(hsim_t *s, int nd, int j) {              /* placed_jobs.pop(job_id) */
  uint64_t *w = &s->pj_bits[(size_t)j * s->pjw + (nd >> 6)], b = 1ull << (nd & 63);
  if (*w & b) { *w &= ~b; s->nodes[nd].n_placed_jobs -= 1; }
}
/* Node.release_allocated_resources  node.py:64-84; fills `set` with the tasks whose interference may be lifted */
static int node_release(hsim_t *s, int nd, int t, int32_t *set, int set_cap) {
  int ns = 0;
  s->nodes[nd].cpu_used -= TASK_CPU; s->nodes[nd].mem_used -= TASK_MEM;
  for (int d = 0; d < s->G; ++d) {
    hdev_t *dv = dev_of(s, nd, d);
    for (int i = 0; i < dv->nt; ++i) if (dv->t[i] == t) { for (int k = i; k + 1 < dv->nt; ++k) dv->t[k] = dv->t[k + 1]; dv->nt--; break; }
    if (dv->nt <= 1)
      for (int i = 0; i < dv->nt; ++i) if (s->tasks[dv->t[i]].interfered) {
        int dup = 0;
        for (int k = 0; k < ns; ++k) dup |= (set[k] == dv->t[i]);
        if (!dup && set && ns < set_cap) set[ns++] = dv->t[i];
      }
  }
  return ns;
}
/* JobsManager.reset_interference  jobs_manager.py:189-201 */
static void reset_interference(hsim_t *s, const int32_t *set, int ns) {
  for (int i = 0; i < ns; ++i) {
    htask_t *tk = &s->tasks[set[i]];
    if (!s->jobs[tk->job].in_running) continue;
    if (tk->interfered) {
      tk->interfered = 0;
      double diff = tk->duration - tk->original;
      long half = (long)(diff / 2);                                       /* int(diff/2) */
      tk->duration = tk->original + (double)(half > 5 ? half : 5);
    }
  }
}

/* ------------------------------------------------------------------ scoring + placement */
static double polyval_nv2080(double x) { double y = 0.0; y = y * x + 4E-5; y = y * x + -0.00302; y = y * x + 1.16664; return y; }   /* np.polyval, fitted_fn.py:1 */

static double score_node(hsim_t *s, int nd, int t) {              /* horus.py:6-56: returns min_cost */
  const hjob_t *jb = &s->jobs[s->tasks[t].job];
  const double cap_mib = (double)(s->cap_b >> 20), tm = (double)jb->mem_b / 1048576.0;
  double min_cost = 999.0;
  for (int d = 0; d < s->G; ++d) {
    hdev_t *dv = dev_of(s, nd, d);
    if (!dev_can_fit(s, dv, t)) continue;
    double cur = (double)dev_mem(s, dv) / 1048576.0, cost;
    if (s->scheme == HS_SCHEME_HORUS) {
      double mem_cost = (cur + tm) / cap_mib;
      double val = dev_util(s, dv, NULL) + jb->util_avg;
      double util_cost = polyval_nv2080(val);
      cost = (mem_cost * 0.5) + (util_cost * 0.5) + (double)dv->nt;
    } else {
      double mem_cost = cur + tm / cap_mib;                       /* precedence as written, horus.py:16 */
      double util_cost = dev_util(s, dv, NULL);
      cost = (mem_cost * 0.5) + (util_cost / 100) + (double)dv->nt;
    }
    if (cost < min_cost) min_cost = cost;
  }
  return min_cost;
}

typedef struct { int node; double min_score; } ninfo_t;
static inline int ninfo_lt(const ninfo_t *a, const ninfo_t *b) { return a->min_score > b->min_score; }   /* algorithm.py:25-26 */
static void heap_push(ninfo_t *h, int *n, ninfo_t x) {            /* heapq.heappush: _siftdown(heap, 0, len-1) */
  int pos = (*n)++;
  while (pos > 0) { int parent = (pos - 1) >> 1; if (ninfo_lt(&x, &h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = x;
}
static ninfo_t heap_pop(ninfo_t *h, int *n) {                     /* heapq.heappop: _siftup(heap, 0) */
  ninfo_t last = h[--(*n)];
  if (*n == 0) return last;
  ninfo_t ret = h[0];
  int end = *n, pos = 0, child = 1;
  while (child < end) {
    int right = child + 1;
    if (right < end && !ninfo_lt(&h[child], &h[right])) child = right;
    h[pos] = h[child]; pos = child; child = 2 * pos + 1;
  }
  while (pos > 0) { int parent = (pos - 1) >> 1; if (ninfo_lt(&last, &h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = last;
  return ret;
}

/* horus_placement  algorithm.py:34-180.  Returns 1 on success and fills res_nodes (distinct nodes, first-use order). */
static int horus_placement(hsim_t *s, int j, int32_t *res_nodes, int *n_res) {
  hjob_t *jb = &s->jobs[j];
  const int T = jb->ntasks, demand = jb->gpus, t0 = jb->first_task;
  ninfo_t *heap = (ninfo_t *)malloc(sizeof(ninfo_t) * (size_t)(demand + 2));
  int hn = 0;
  for (int k = 0; k < T; ++k)
    for (int nd = 0; nd < s->M; ++nd) {
      if (!node_is_free(s, nd)) continue;
      if (!node_can_fit(s, nd, t0 + k, 1)) continue;
      ninfo_t x; x.node = nd; x.min_score = score_node(s, nd, t0 + k);
      heap_push(heap, &hn, x);
      if (hn > demand) (void)heap_pop(heap, &hn);
    }
  for (int i = 1; i < hn; ++i) {                                  /* sorted(nodes_stack, key=min_score): stable */
    ninfo_t x = heap[i]; int k = i - 1;
    while (k >= 0 && heap[k].min_score > x.min_score) { heap[k + 1] = heap[k]; --k; }
    heap[k + 1] = x;
  }
  const int C = hn;
  int32_t *map_node = (int32_t *)malloc(sizeof(int32_t) * (size_t)(C > 0 ? C : 1) * (size_t)T);
  int32_t *map_order = (int32_t *)malloc(sizeof(int32_t) * (size_t)(C > 0 ? C : 1) * (size_t)T);
  int *map_n = (int *)calloc((size_t)(C > 0 ? C : 1), sizeof(int)), *ok = (int *)calloc((size_t)(C > 0 ? C : 1), sizeof(int));
  int *distinct = (int *)calloc((size_t)(C > 0 ? C : 1), sizeof(int));
  int any = 0;
  for (int i = 0; i < C; ++i) {
    int32_t *mn = map_node + (size_t)i * T, *mo = map_order + (size_t)i * T;
    for (int k = 0; k < T; ++k) mn[k] = -1;
    const int cand = heap[i].node;
    for (int k = 0; k < T; ++k)
      if (node_reserve_task(s, cand, t0 + k, 1)) { node_place_job(s, cand, j); mn[k] = cand; mo[map_n[i]++] = k; }
    const int home = s->nodes[cand].rack;
    for (int dist = 0; dist < s->S; ++dist)                       /* get_racks_by_dist: stable sort by |rack - home| */
      for (int r = 0; r < s->S; ++r) {
        if (abs(r - home) != dist) continue;
        if (map_n[i] >= T) { any = 1; goto racks_done; }
        for (int q = 0; q < s->P; ++q) {
          const int nd = r * s->P + q;
          if (map_n[i] >= T) break;
          for (int k = 0; k < T; ++k) {
            if (mn[k] >= 0) continue;
            if (node_reserve_task(s, nd, t0 + k, 1)) { node_place_job(s, nd, j); mn[k] = nd; mo[map_n[i]++] = k; }
            if (map_n[i] >= T) break;
          }
        }
      }
racks_done:
    for (int q = 0; q < map_n[i]; ++q) {                          /* undo the trial  (:127-137) */
      const int k = mo[q], nd = mn[k];
      if (q == 0) node_pop_job(s, nd, j);
      if (s->tasks[t0 + k].placed_node == nd) {
        s->tasks[t0 + k].placed_node = -1; s->nodes[nd].n_placed_tasks -= 1;
        (void)node_release(s, nd, t0 + k, NULL, 0);
      }
    }
    if (map_n[i] >= T) { any = 1; ok[i] = 1; }
    for (int q = 0; q < map_n[i]; ++q) { int seen = 0; for (int p = 0; p < q; ++p) seen |= (mn[mo[p]] == mn[mo[q]]); distinct[i] += !seen; }
  }
  int best = -1;
  if (any) for (int i = 0; i < C; ++i) if (ok[i] && (best < 0 || distinct[i] < distinct[best])) best = i;   /* stable sort by len(nodes) */
  int success = 0;
  if (best >= 0) {
    const int32_t *mn = map_node + (size_t)best * T, *mo = map_order + (size_t)best * T;
    *n_res = 0;
    for (int q = 0; q < T; ++q) {
      const int k = mo[q], nd = mn[k];
      if (!node_reserve_task(s, nd, t0 + k, 1)) abort();           /* the reference asserts cnt == len(tasks) */
      int seen = 0;
      for (int p = 0; p < *n_res; ++p) seen |= (res_nodes[p] == nd);
      if (!seen) res_nodes[(*n_res)++] = nd;
      if (s->tro_node[t0 + k] < 0) s->tro_order[t0 + jb->tro_n++] = k;      /* dict: a known key keeps its position */
      s->tro_node[t0 + k] = nd;
      node_place_job(s, nd, j);
    }
    success = 1;
  }
  free(heap); free(map_node); free(map_order); free(map_n); free(ok); free(distinct);
  return success;
}

/* ------------------------------------------------------------------ yarn placement under these schedulers
 * --scheme yarn with --schedule horus | horus+ | gandiva: ms_yarn_placement (algorithm.py:28-32,301-417) on the same
 * node / device state.  No packing (pack=False), so every device holds at most one task. */
static int node_idle_devices(const hsim_t *s, int nd) { int c = 0; for (int d = 0; d < s->G; ++d) c += (s->devs[(size_t)nd * s->G + d].nt == 0); return c; }
static int node_can_fit_num_task(const hsim_t *s, int nd, int gpc, int remaining) {          /* node.py:109-127 */
  int g = node_idle_devices(s, nd) / gpc - remaining, c = cpu_free(s, nd) / TASK_CPU - remaining, m = mem_free(s, nd) / TASK_MEM - remaining;
  int ng = g >= 0 ? remaining : remaining + g, nc = c >= 0 ? remaining : remaining + c, nm = m >= 0 ? remaining : remaining + m;
  int r = nc < nm ? nc : nm;
  return r < ng ? r : ng;
}
static void tro_set(hsim_t *s, int j, int k, int nd) {            /* job.tasks_running_on[task] = node (dict semantics) */
  hjob_t *jb = &s->jobs[j];
  if (s->tro_node[jb->first_task + k] < 0) {
    int known = 0;
    for (int q = 0; q < jb->tro_n; ++q) known |= (s->tro_order[jb->first_task + q] == k);
    if (!known) s->tro_order[jb->first_task + jb->tro_n++] = k;
  }
  s->tro_node[jb->first_task + k] = nd;
}
static int yarn_placement(hsim_t *s, int j, int32_t *res_nodes, int *n_res) {
  hjob_t *jb = &s->jobs[j];
  const int T = jb->ntasks, t0 = jb->first_task, gpc = jb->gpc;
  *n_res = 0;
  if (jb->gpus <= s->G) {                                         /* try_single_node_alloc_ms  algorithm.py:396-417 */
    for (int nd = 0; nd < s->M; ++nd) {
      if (!node_is_free(s, nd)) continue;
      if (!(node_idle_devices(s, nd) >= jb->gpus && cpu_free(s, nd) >= TASK_CPU * T && mem_free(s, nd) >= TASK_MEM * T)) continue;
      if (node_can_fit_num_task(s, nd, gpc, T) < T) continue;     /* Node.try_alloc_job  node.py:234-263 */
      int placed = 0;
      for (int k = 0; k < T; ++k) if (node_reserve_task(s, nd, t0 + k, 0)) { tro_set(s, j, k, nd); ++placed; }
      if (placed == 0) continue;                                  /* devices refused every task: what was charged stays */
      node_place_job(s, nd, j);
      res_nodes[(*n_res)++] = nd;
      return 1;
    }
    return 0;
  }
  int assigned = 0, least = (jb->gpus + s->G - 1) / s->G;         /* try_cross_node_alloc_ms  algorithm.py:301-393 */
  for (int nd = 0; nd < s->M; ++nd) {
    if (!node_is_free(s, nd)) continue;
    if (assigned == T) break;
    int can = node_can_fit_num_task(s, nd, gpc, T - assigned);
    if (can == 0) continue;
    int worker_count = 0, check_next = 0;
    for (int k = assigned; k < T; ++k) {
      if (!(worker_count <= can)) continue;                       /* the `<=` over-try (:341) */
      ++worker_count;
      if (!node_reserve_task(s, nd, t0 + k, 0)) { --worker_count; check_next = 1; break; }
      tro_set(s, j, k, nd);
    }
    if (worker_count > 0) { assigned += worker_count; node_place_job(s, nd, j); res_nodes[(*n_res)++] = nd; }
    if (check_next) continue;
    if (*n_res >= least && assigned == T) break;
  }
  if (assigned == T && *n_res >= least) return 1;
  for (int a = 0; a < *n_res; ++a) {                              /* not enough: clear everything (:378-387) */
    const int nd = res_nodes[a];
    node_pop_job(s, nd, j);
    for (int k = 0; k < T; ++k)
      if (s->tasks[t0 + k].placed_node == nd) { s->tasks[t0 + k].placed_node = -1; s->nodes[nd].n_placed_tasks -= 1; (void)node_release(s, nd, t0 + k, NULL, 0); }
  }
  *n_res = 0;
  return 0;
}
static int place_job(hsim_t *s, int j, int32_t *res_nodes, int *n_res);

/* ------------------------------------------------------------------ queues (heapq over Job.__lt__) */
static inline int job_lt(const hsim_t *s, int a, int b) {         /* base_factory.py:7-11 CompareAbleByUtilization */
  if (s->jobs[a].util_avg != 0.0) return s->jobs[a].util_avg < s->jobs[b].util_avg;
  return 0;
}
static void q_heappush(hsim_t *s, int q, int j) {
  int32_t *h = s->queue[q]; int64_t pos = s->qn[q]++;
  while (pos > 0) { int64_t parent = (pos - 1) >> 1; if (job_lt(s, j, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = j;
}
static int q_heappop(hsim_t *s, int q) {
  int32_t *h = s->queue[q];
  int last = h[--s->qn[q]];
  if (s->qn[q] == 0) return last;
  int ret = h[0];
  int64_t end = s->qn[q], pos = 0, child = 1;
  while (child < end) {
    int64_t right = child + 1;
    if (right < end && !job_lt(s, h[child], h[right])) child = right;
    h[pos] = h[child]; pos = child; child = 2 * pos + 1;
  }
  while (pos > 0) { int64_t parent = (pos - 1) >> 1; if (job_lt(s, last, h[parent])) { h[pos] = h[parent]; pos = parent; } else break; }
  h[pos] = last;
  return ret;
}
static int is_pq(const hsim_t *s) { return s->schedule == HS_SCHED_HORUS || s->schedule == HS_SCHED_HORUS_PLUS; }
/* JobQueueManager.insert  job_queue_manager.py:146-154 */
static void q_insert(hsim_t *s, int j, int q, int64_t pos) {
  s->credits[q] = s->credits[q] + 1;
  if (is_pq(s)) { q_heappush(s, q, j); return; }
  int32_t *h = s->queue[q];
  if (pos > s->qn[q]) pos = s->qn[q];
  memmove(h + pos + 1, h + pos, sizeof(int32_t) * (size_t)(s->qn[q] - pos));
  h[pos] = j; s->qn[q]++;
}
static int q_pop(hsim_t *s, int q) {                               /* job_queue_manager.py:129-135 */
  if (is_pq(s)) return q_heappop(s, q);
  int j = s->queue[q][0];
  memmove(s->queue[q], s->queue[q] + 1, sizeof(int32_t) * (size_t)(--s->qn[q]));
  return j;
}
static int64_t q_total(const hsim_t *s) { int64_t t = 0; for (int q = 0; q < s->nq; ++q) t += s->qn[q]; return t; }

static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y); }
static double median_pending(const hsim_t *s, const int32_t *jobs, int64_t n, int32_t *scratch) {   /* np.median of ints */
  for (int64_t i = 0; i < n; ++i) scratch[i] = s->jobs[jobs[i]].pending;
  qsort(scratch, (size_t)n, sizeof(int32_t), cmp_i32);
  if (n & 1) return (double)scratch[n / 2];
  return ((double)scratch[n / 2 - 1] + (double)scratch[n / 2]) / 2.0;          /* mean of the two middle values */
}
/* JobQueueManager.update_credits  job_queue_manager.py:115-127 */
static void update_credits(hsim_t *s, int32_t *scratch) {
  for (int q = 0; q < s->nq; ++q) {
    if (s->qn[q] > 0) {
      double mp = median_pending(s, s->queue[q], s->qn[q], scratch);
      if (mp < 0) mp = 0;
      s->credits[q] = mp < 1 ? (double)s->qn[q] : mp * (double)s->qn[q];
    } else s->credits[q] = 0;
  }
}

/* ------------------------------------------------------------------ k-means queue assignment (horus+) */
static double job_score(const hsim_t *s, int j) {                  /* transform_to_dist  utils.py:14-22 */
  const hjob_t *x = &s->jobs[j];
  double sc = (double)x->ntasks;
  sc += x->util_avg; sc += (double)x->gpc; sc += (double)x->gpus; sc += x->util_max; sc += x->mem_avg_mib; sc += (double)x->mem_b / 1048576.0;
  return sc;
}
static double job_dist(const hsim_t *s, int a, int b) {            /* utils.py:4-12 */
  const hjob_t *x = &s->jobs[a], *y = &s->jobs[b];
  double sc = (double)abs(x->ntasks - y->ntasks);
  sc += fabs(x->util_avg - y->util_avg); sc += (double)abs(x->gpc - y->gpc); sc += (double)abs(x->gpus - y->gpus);
  sc += fabs(x->util_max - y->util_max); sc += fabs(x->mem_avg_mib - y->mem_avg_mib);
  sc += fabs((double)x->mem_b / 1048576.0 - (double)y->mem_b / 1048576.0);
  return sc;
}
static double pairwise_sum(const double *a, int64_t n) {           /* numpy's add.reduce over a contiguous float64 vector */
  if (n < 8) { double r = 0.; for (int64_t i = 0; i < n; ++i) r += a[i]; return r; }
  if (n <= 128) {
    double r[8]; int64_t i;
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int64_t n2 = n / 2; n2 -= n2 % 8;
  return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}
/* clusterize  utils.py:36-67: fills assign[i] for jobs[i] */
static void clusterize(hsim_t *s, const int32_t *jobs, int64_t n, int k, int32_t *assign, double *dscratch) {
  int32_t cent[MAXQ];
  for (int c = 0; c < k; ++c) cent[c] = jobs[rng_below(&s->rng, n)];
  int32_t *old = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) { assign[i] = -1; old[i] = -1; }
  int iter = 0;
  while (iter < 1000) {
    if (iter != 0) { int same = 1; for (int64_t i = 0; i < n; ++i) same &= (assign[i] == old[i]); if (same) break; }
    memcpy(old, assign, sizeof(int32_t) * (size_t)n);
    iter += 1;
    for (int64_t i = 0; i < n; ++i) {
      int bi = 0; double bd = 0;
      for (int c = 0; c < k; ++c) { double d = job_dist(s, jobs[i], cent[c]); if (c == 0 || d < bd) { bd = d; bi = c; } }   /* np.argmin: first minimum */
      assign[i] = bi;
    }
    for (int c = 0; c < k; ++c) {
      int64_t m = 0;
      for (int64_t i = 0; i < n; ++i) if (assign[i] == c) dscratch[m++] = job_score(s, jobs[i]);
      if (m > 0) {
        double mean = pairwise_sum(dscratch, m) / (double)m;
        double target = (double)(long long)mean;                             /* .astype(int) */
        int best = -1; double bscore = 99999999999.0;
        for (int64_t i = 0; i < n; ++i) if (assign[i] == c) { double t = fabs(job_score(s, jobs[i]) - target); if (t < bscore) { best = jobs[i]; bscore = t; } }
        cent[c] = best;
      } else cent[c] = jobs[rng_below(&s->rng, n)];                           /* np.random.choice(len(jobs)) */
    }
  }
  free(old);
}

/* JobsManager.insert  jobs_manager.py:114-139 */
static void jm_insert(hsim_t *s, const int32_t *jobs_in, int64_t n_in, const int32_t *qpos, int32_t *scratch, double *dscratch) {
  if (s->schedule == HS_SCHED_HORUS_PLUS && qpos == NULL) {
    int64_t total = q_total(s) + n_in, m = 0;
    int32_t *all = (int32_t *)malloc(sizeof(int32_t) * (size_t)(total > 0 ? total : 1));
    for (int q = 0; q < s->nq; ++q) { int64_t cnt = s->qn[q]; for (int64_t i = 0; i < cnt; ++i) all[m++] = q_pop(s, q); }   /* pop_all_queuing_jobs */
    for (int64_t i = 0; i < n_in; ++i) all[m++] = jobs_in[i];
    if (m > 0) {
      int32_t *assign = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
      clusterize(s, all, m, s->nq, assign, dscratch);
      for (int64_t i = 0; i < m; ++i) q_insert(s, all[i], assign[i], i);
      free(assign);
    }
    free(all);
    (void)scratch;
    return;
  }
  for (int64_t i = 0; i < n_in; ++i) q_insert(s, jobs_in[i], qpos ? qpos[i] : 0, i);
}

static int place_job(hsim_t *s, int j, int32_t *res_nodes, int *n_res) {   /* placement_algorithms[--scheme]  algorithm.py:182-187 */
  return s->yarn ? yarn_placement(s, j, res_nodes, n_res) : horus_placement(s, j, res_nodes, n_res);
}

/* ------------------------------------------------------------------ start / finish / preempt */
static void start_job(hsim_t *s, int j, const int32_t *nodes, int nn, int delta) {   /* schedule.py:159-162, node.py:164-188, job.py:153-169 */
  hjob_t *jb = &s->jobs[j];
  for (int a = 0; a < nn; ++a) {
    const int nd = nodes[a];
    for (int q = 0; q < jb->tro_n; ++q) {
      const int k = s->tro_order[jb->first_task + q], t = jb->first_task + k;
      if (s->tro_node[t] != nd) continue;
      htask_t *tk = &s->tasks[t];
      if (tk->placed_node == nd) { tk->placed_node = -1; s->nodes[nd].n_placed_tasks -= 1; }
      tk->running = 1;
      if (tk->run_node != nd) { tk->run_node = nd; s->nodes[nd].n_running += 1; }
    }
    int cnt = 0;
    for (int k = 0; k < jb->ntasks; ++k) cnt += (s->tasks[jb->first_task + k].running && !s->tasks[jb->first_task + k].finished);
    if (cnt == jb->ntasks) {
      jb->start = delta; jb->migration += 1; jb->running = 1;
      if (!jb->in_running) { jb->in_running = 1; s->running[s->nrun++] = j; }
    }
  }
}
static int job_time_processed(const hsim_t *s, const hjob_t *jb) { int m = 0; for (int k = 0; k < jb->ntasks; ++k) if (s->tasks[jb->first_task + k].time_processed > m) m = s->tasks[jb->first_task + k].time_processed; return m; }
static double job_get_duration(const hsim_t *s, const hjob_t *jb) { double m = 0; for (int k = 0; k < jb->ntasks; ++k) if (s->tasks[jb->first_task + k].duration > m) m = s->tasks[jb->first_task + k].duration; return m > jb->duration ? m : jb->duration; }
static void running_remove(hsim_t *s, int j) {
  int64_t w = 0;
  for (int64_t i = 0; i < s->nrun; ++i) if (s->running[i] != j) s->running[w++] = s->running[i];
  s->nrun = w; s->jobs[j].in_running = 0;
}
/* JobsManager.preempt  jobs_manager.py:150-187 */
static void preempt_job(hsim_t *s, int j, int32_t *scratch, double *dscratch) {
  hjob_t *jb = &s->jobs[j];
  running_remove(s, j);
  int32_t set[64];
  for (int q = 0; q < jb->tro_n; ++q) {
    const int k = s->tro_order[jb->first_task + q], t = jb->first_task + k, nd = s->tro_node[t];
    int first = 1;
    for (int p = 0; p < q; ++p) first &= (s->tro_node[jb->first_task + s->tro_order[jb->first_task + p]] != nd);
    if (first) node_pop_job(s, nd, j);
    if (s->tasks[t].run_node == nd) {
      s->tasks[t].run_node = -1; s->nodes[nd].n_running -= 1;
      int ns = node_release(s, nd, t, set, 64);
      if (ns > 0) reset_interference(s, set, ns);
    }
  }
  jb->running = 0; jb->pending = 0;
  for (int k = 0; k < jb->ntasks; ++k) s->tasks[jb->first_task + k].running = 0;
  s->events += 1;
  int32_t one = j;
  jm_insert(s, &one, 1, NULL, scratch, dscratch);
}

typedef struct { int32_t start, end, jct, preempt; double original, actual; } horus_job_rec;

/* `scheme` = which score function (0 horus_score, 1 gandiva_score).  The reference picks it by the SCHEDULE name
 * (schedule.py:47 -> algorithm.py:196 -> :58 score_fn[scheme] with scheme == self.schedule); the Python wrapper maps. */
int64_t oracle_run_horus(const gs_cluster *c, int32_t scheme, int32_t schedule, int32_t num_buffer, int32_t num_queue,
                         uint32_t seed, int64_t n, const int32_t *arrive, const int32_t *gpus, const int32_t *gpc,
                         const double *duration, const int64_t *mem_bytes, const double *mem_avg_mib,
                         const double *util_avg, const double *util_max,
                         gs_tick_row *rows_out, double *util_out, uint8_t *util_is_array, int64_t rows_cap,
                         horus_job_rec *recs_out, int32_t *finish_order_out, int64_t *n_finished_out,
                         int64_t *events_out, uint64_t *draws_out) {
  hsim_t S; memset(&S, 0, sizeof(S));
  hsim_t *s = &S;
  s->S = c->num_switch; s->P = c->num_node_p_switch; s->M = s->S * s->P; s->G = c->num_gpu_p_node;
  s->cpu_cap = c->num_cpu_p_node; s->mem_cap = c->mem_p_node; s->cap_b = (int64_t)c->gpu_mem_cap_mib << 20;
  s->yarn = (scheme >> 8) & 1;                /* bit 8 of `scheme`: --scheme yarn (ms_yarn_placement) */
  scheme &= 0xff;
  s->scheme = scheme; s->schedule = schedule; s->num_buffer = num_buffer; s->n = n;
  s->nq = num_queue > 0 ? num_queue : 1;
  if (s->nq > MAXQ || s->G > 64) return GS_ERR_ARG;
  rng_seed(&s->rng, seed);
  int64_t ntask = 0;
  for (int64_t j = 0; j < n; ++j) ntask += gpus[j] / gpc[j];
  const size_t N = (size_t)(n > 0 ? n : 1), NT = (size_t)(ntask > 0 ? ntask : 1);
  s->nodes = (hnode_t *)calloc((size_t)s->M, sizeof(hnode_t));
  s->devs = (hdev_t *)calloc((size_t)s->M * (size_t)s->G, sizeof(hdev_t));
  s->jobs = (hjob_t *)calloc(N, sizeof(hjob_t));
  s->tasks = (htask_t *)calloc(NT, sizeof(htask_t));
  s->tro_node = (int32_t *)malloc(NT * sizeof(int32_t)); s->tro_order = (int32_t *)malloc(NT * sizeof(int32_t));
  s->pjw = (s->M + 63) / 64; s->pj_bits = (uint64_t *)calloc(N * (size_t)s->pjw, sizeof(uint64_t));
  for (int q = 0; q < s->nq; ++q) s->queue[q] = (int32_t *)malloc(N * sizeof(int32_t));
  s->running = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *scratch = (int32_t *)malloc(N * sizeof(int32_t)), *look = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *look_q = (int32_t *)malloc(N * sizeof(int32_t)), *to_finish = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *res_nodes = (int32_t *)malloc(sizeof(int32_t) * (size_t)s->M);
  double *dscratch = (double *)malloc(N * sizeof(double));
  for (int nd = 0; nd < s->M; ++nd) s->nodes[nd].rack = nd / s->P;
  int64_t tcur = 0;
  for (int64_t j = 0; j < n; ++j) {
    hjob_t *jb = &s->jobs[j];
    jb->gpus = gpus[j]; jb->gpc = gpc[j]; jb->ntasks = gpus[j] / gpc[j]; jb->first_task = (int32_t)tcur;
    jb->util_avg = util_avg[j]; jb->util_max = util_max[j]; jb->mem_b = mem_bytes[j]; jb->mem_avg_mib = mem_avg_mib ? mem_avg_mib[j] : 0.0;
    jb->duration = duration[j]; jb->start = 0; jb->end = 0;
    for (int k = 0; k < jb->ntasks; ++k) {
      htask_t *tk = &s->tasks[tcur + k];
      tk->job = (int32_t)j; tk->duration = tk->original = duration[j]; tk->placed_node = -1; tk->run_node = -1;
      s->tro_node[tcur + k] = -1;
    }
    tcur += jb->ntasks;
  }
  int64_t p = 0, nfin = 0, ticks = 0, rc = 0;
  int delta = 0;
  int64_t current_remaining = n, running_jobs = 0;
  while (current_remaining + running_jobs > 0) {                  /* schedule.py:185 */
    /* gen_jobs: rows with normalized_time <= delta, in trace order  (jobs_manager.py:228-241) */
    int64_t b = p;
    while (b < n && arrive[b] <= delta) ++b;
    for (int64_t i = p; i < b; ++i) look[i - p] = (int32_t)i;
    s->events += b - p;
    jm_insert(s, look, b - p, NULL, scratch, dscratch);
    p = b;
    /* _schedule  schedule.py:39-58 */
    if (q_total(s) > 0) {
      int free_nodes = 0;
      for (int nd = 0; nd < s->M; ++nd) free_nodes += node_is_free(s, nd);
      if (free_nodes >= 1) {
        int placed = -1, nres = 0;
        if (s->schedule == HS_SCHED_FIFO || s->schedule == HS_SCHED_GANDIVA) {          /* schedule_fifo */
          int j = s->queue[0][0];
          if (place_job(s, j, res_nodes, &nres)) { (void)q_pop(s, 0); placed = j; }
        } else if (s->schedule == HS_SCHED_HORUS) {                                      /* schedule_horus */
          int64_t qd = q_total(s), min_k = num_buffer < qd ? num_buffer : qd;
          if (min_k < 0) min_k = 0;
          for (int64_t i = 0; i < min_k; ++i) look[i] = q_pop(s, 0);
          int pos = -1;
          for (int64_t i = 0; i < min_k; ++i) if (place_job(s, look[i], res_nodes, &nres)) { pos = (int)i; break; }
          if (pos >= 0) { placed = look[pos]; for (int64_t i = pos; i + 1 < min_k; ++i) look[i] = look[i + 1]; min_k -= 1; }
          jm_insert(s, look, min_k, NULL, scratch, dscratch);
        } else {                                                                          /* schedule_horus_plus */
          int64_t qd = q_total(s), min_k = num_buffer < qd ? num_buffer : qd;
          if (min_k > 0) {
            for (int64_t i = 0; i < min_k; ++i) {
              update_credits(s, scratch);
              int qi = 0;
              for (int q = 1; q < s->nq; ++q) if (s->credits[q] > s->credits[qi]) qi = q;     /* np.argmax: first maximum */
              look[i] = q_pop(s, qi); look_q[i] = qi;
            }
            int pos = -1;
            for (int64_t i = 0; i < min_k; ++i) if (place_job(s, look[i], res_nodes, &nres)) { pos = (int)i; break; }
            if (pos >= 0) { placed = look[pos]; for (int64_t i = pos; i + 1 < min_k; ++i) { look[i] = look[i + 1]; look_q[i] = look_q[i + 1]; } min_k -= 1; }
            jm_insert(s, look, min_k, look_q, scratch, dscratch);
          }
        }
        if (placed >= 0) { start_job(s, placed, res_nodes, nres, delta); s->events += 1; }
      }
    }
    current_remaining = n - p;
    delta += 1;
    /* JobsManager.step  jobs_manager.py:141-148 */
    for (int q = 0; q < s->nq; ++q) for (int64_t i = 0; i < s->qn[q]; ++i) s->jobs[s->queue[q][i]].pending += 1;
    for (int64_t i = 0; i < s->nrun; ++i) {
      hjob_t *jb = &s->jobs[s->running[i]];
      if (!jb->running) continue;
      for (int k = 0; k < jb->ntasks; ++k) if (s->tasks[jb->first_task + k].running) s->tasks[jb->first_task + k].time_processed += 1;
    }
    if (s->schedule == HS_SCHED_HORUS_PLUS) update_credits(s, scratch);
    /* release_finished_jobs  schedule.py:136-157 */
    int64_t nf = 0;
    for (int64_t i = 0; i < s->nrun; ++i) { hjob_t *jb = &s->jobs[s->running[i]]; if (!((double)job_time_processed(s, jb) < job_get_duration(s, jb))) to_finish[nf++] = s->running[i]; }
    for (int64_t f = 0; f < nf; ++f) {
      const int j = to_finish[f]; hjob_t *jb = &s->jobs[j];
      int32_t set[64];
      for (int q = 0; q < jb->tro_n; ++q) {
        const int k = s->tro_order[jb->first_task + q], t = jb->first_task + k, nd = s->tro_node[t];
        htask_t *tk = &s->tasks[t];
        if (tk->run_node == nd) { tk->run_node = -1; s->nodes[nd].n_running -= 1; }
        if (!tk->finished) { tk->finished = 1; jb->tasks_finished += 1; }
        int ns = node_release(s, nd, t, set, 64);
        if (ns > 0) reset_interference(s, set, ns);
        if (!jb->finished && jb->tasks_finished == jb->ntasks) {
          jb->running = 0; jb->finished = 1; jb->end = delta;
          running_remove(s, j);
          finish_order_out[nfin++] = j; s->events += 1;
        }
      }
    }
    running_jobs = s->nrun;
    /* plugin: gandiva time slicing  algorithm.py:420-440 */
    if (s->schedule == HS_SCHED_GANDIVA && q_total(s) > 0) {
      int64_t nt = 0;
      for (int64_t i = 0; i < s->nrun; ++i) { int tp = job_time_processed(s, &s->jobs[s->running[i]]); if (tp > 1 && tp % 100 == 0) to_finish[nt++] = s->running[i]; }
      for (int64_t i = 0; i < nt; ++i) preempt_job(s, to_finish[i], scratch, dscratch);
    }
    /* _construct_info  schedule.py:95-133 */
    if (ticks >= rows_cap) { rc = GS_ERR_CAPACITY; break; }
    gs_tick_row *row = &rows_out[ticks];
    memset(row, 0, sizeof(*row));
    row->now = delta;
    double usum = 0.0; int uarr = 0; int64_t msum = 0;
    for (int nd = 0; nd < s->M; ++nd) {
      const hnode_t *nn = &s->nodes[nd];
      if (nn->n_running + nn->n_placed_tasks + nn->n_placed_jobs == 0) row->idle_nodes += 1; else row->busy_nodes += 1;
      for (int d = 0; d < s->G; ++d) {
        hdev_t *dv = dev_of(s, nd, d);
        if (dv->nt == 0) { row->idle_gpus += 1; continue; }
        row->busy_gpus += 1;
        int a = 0; double u = dev_util(s, dv, &a);
        usum = usum + u; uarr |= a;
        msum += dev_mem(s, dv);
      }
    }
    util_out[ticks] = usum / (double)(row->idle_gpus + row->busy_gpus); util_is_array[ticks] = (uint8_t)uarr;
    row->mem_busy_bytes = msum;
    row->running = (int32_t)s->nrun; row->queued = (int32_t)q_total(s); row->finished = (int32_t)nfin;
    {
      int64_t m = 0;
      for (int q = 0; q < s->nq; ++q) for (int64_t i = 0; i < s->qn[q]; ++i) { int pd = s->jobs[s->queue[q][i]].pending; scratch[m++] = pd; row->pend_sum += pd; if (pd > row->pend_max) row->pend_max = pd; }
      if (m > 0) { qsort(scratch, (size_t)m, sizeof(int32_t), cmp_i32); row->pend_med_lo = scratch[(m - 1) / 2]; row->pend_med_hi = scratch[m / 2]; }
    }
    ticks += 1;
  }
  for (int64_t j = 0; j < n; ++j) {
    const hjob_t *jb = &s->jobs[j];
    recs_out[j].start = jb->start; recs_out[j].end = jb->end; recs_out[j].jct = job_time_processed(s, jb);
    recs_out[j].preempt = jb->migration; recs_out[j].original = jb->duration; recs_out[j].actual = job_get_duration(s, jb);
  }
  *n_finished_out = nfin;
  if (events_out) *events_out = s->events;
  if (draws_out) *draws_out = s->rng.draws;
  free(s->nodes); free(s->devs); free(s->jobs); free(s->tasks); free(s->tro_node); free(s->tro_order); free(s->pj_bits);
  for (int q = 0; q < s->nq; ++q) free(s->queue[q]);
  free(s->running); free(scratch); free(look); free(look_q); free(to_finish); free(res_nodes); free(dscratch);
  return rc < 0 ? rc : ticks;
}
