/*
 * policy_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's event-driven policy simulators, which exist only as
 * dead code (they read globals JOBS / CLUSTER / LOG / scheduler that are never defined,
 * SURVEY section 0):
 *     sjf        smallest_first_sim_jobs     /root/reference/run_sim.py:162-287
 *     dlas(-gpu) dlas_sim_jobs(gputime)      /root/reference/run_sim.py:664-947
 *     gittins    gittins_sim_jobs            /root/reference/run_sim.py:956-1203
 *                get_gittins_index           /root/reference/run_sim.py:949-954
 *
 * PARITY PINNED TO THE REFERENCE'S LOOP CODE (not to a reference run: run_sim.py cannot enter
 * these loops).  tests/golden/make_policy_golden.py extracts the four functions above plus
 * cal_r_gittins_index / parse_job_dist from /root/reference/run_sim.py with `ast`, exec()s them
 * UNCHANGED in a namespace whose JOBS / CLUSTER / LOG / scheduler are the minimal stubs listed
 * below, and records every LOG.job_complete and every LOG.checkpoint; tests/golden/policy_* hold
 * 9 such fixtures and tests/golden/fuzz_policy_reference.py compared 460 random cases (all
 * identical).  What stays an assumption is therefore ONLY the stub completion, not the loop logic.
 * Completion of what the dead code leaves undefined (kept deliberately minimal):
 *   - a job dict = {job_idx, num_gpu, submit_time = admission tick, duration D =
 *     max(1, ceil(minutes * 0.5)) ticks}; move_to_runnable() = status PENDING,
 *     last_check_time = event time, all counters 0, start_time = "maxsize".
 *   - job_events = one start event per distinct submit tick, jobs in trace order.
 *   - sjf placement "try_get_job_res" = the LIVE yarn placement
 *     (core/scheduling/algorithm.py:28-32) on the emptied cluster; dlas / gittins use GPU
 *     counting only, exactly as the dead code does (run_sim.py:812-817, 1107-1113).
 *   - solve_starvation = 0 (its default): no promotion.
 * Defects of the spec kept verbatim and flagged: the gputime jump subtracts the un-scaled
 * executed_time (:929); pending jobs rank by executed_time without the GPU factor (:1071);
 * next_gittins_unit += event_time (:1202); a queue jump may be computed AT or BEFORE the current
 * event time (:929) -- the loop then handles a zero / negative-dt jump event, exactly as the code does
 * (each such event demotes at least one job one more level, so the loop still terminates);
 * Q25: on an end/start tie the START event dict receives the key 'end_jobs' (:708-710, :994-996); if
 * a jump event is served first, the key survives on the dict and those jobs are completed when the
 * start event is finally consumed -- even if they were preempted meanwhile and have work left.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gsched.h"

#define T_INF 0x7fffffff
enum { ST_NONE = 0, ST_PENDING = 1, ST_RUNNING = 2, ST_END = 3 };

typedef struct {
  int status, q_id, last_check, total_exec, exec, pending, last_pending, start, preempt, resume, remaining;
  double rank;
} pjob_t;

typedef struct {
  int M, G, K;
  int *idle, *kfree;     /* emptied every event; sjf only */
} pcluster_t;

/* live yarn placement on the (re)filled cluster: same rules as the tick oracle, condensed
 * to (idle devices, free task slots) per node since devices are anonymous here.            */
static int yarn_place(pcluster_t *c, int gpus, int gpc, int placeable) {
  int tasks = gpus / gpc;
  if (!placeable) return 0;
  if (gpus <= c->G) {
    for (int nd = 0; nd < c->M; ++nd)
      if (c->idle[nd] >= gpus && c->kfree[nd] >= tasks) { c->idle[nd] -= gpus; c->kfree[nd] -= tasks; return 1; }
    return 0;
  }
  int cum = 0, last = -1;
  for (int nd = 0; nd < c->M; ++nd) {
    int cap = c->idle[nd] / gpc; if (c->kfree[nd] < cap) cap = c->kfree[nd];
    if (cap <= 0) continue;
    cum += cap;
    if (cum >= tasks) { last = nd; break; }
  }
  if (last < 0) return 0;
  int rem = tasks;
  for (int nd = 0; nd <= last; ++nd) {
    int cap = c->idle[nd] / gpc; if (c->kfree[nd] < cap) cap = c->kfree[nd];
    if (cap <= 0) continue;
    int take = cap < rem ? cap : rem;
    c->idle[nd] -= take * gpc; c->kfree[nd] -= take; rem -= take;
  }
  return 1;
}

/* get_gittins_index                                             run_sim.py:949-954 */
static double gittins_index(const gs_policy *p, double a) {
  int n = p->gittins_n;                       /* table length including the sentinel */
  if (n < 2 || a > p->gittins_data[n - 2]) return 0.0;
  int lo = 0, hi = n - 1;                     /* first i with data[i] > a */
  while (lo < hi) { int mid = (lo + hi) / 2; if (p->gittins_data[mid] > a) hi = mid; else lo = mid + 1; }
  return p->gittins_index[lo];
}

/* stable insertion sort of an index list by an integer / double key (Python list.sort is stable) */
static void stable_sort_int(int32_t *v, int64_t n, const int32_t *key) {
  for (int64_t i = 1; i < n; ++i) { int32_t x = v[i]; int64_t j = i; while (j > 0 && key[v[j - 1]] > key[x]) { v[j] = v[j - 1]; --j; } v[j] = x; }
}
static void stable_sort_rank(int32_t *v, int64_t n, const pjob_t *jb) {
  for (int64_t i = 1; i < n; ++i) { int32_t x = v[i]; int64_t j = i; while (j > 0 && jb[v[j - 1]].rank > jb[x].rank) { v[j] = v[j - 1]; --j; } v[j] = x; }
}
static void list_remove(int32_t *v, int64_t *n, int32_t x) {
  int64_t w = 0; for (int64_t i = 0; i < *n; ++i) if (v[i] != x) v[w++] = v[i]; *n = w;
}

int64_t oracle_run_policy(const gs_cluster *c, const gs_policy *pol, int64_t n, const int32_t *arrive,
                          const int32_t *gpus, const int32_t *gpc, const double *duration,
                          const int64_t *mem_bytes, gs_tick_row *rows_out, int64_t rows_cap,
                          gs_job_rec *jobs_out, int32_t *finish_order_out, int64_t *n_finished_out,
                          int64_t *events_out) {
  const int policy = pol->schedule;
  const int gputime = (policy == GS_SCHED_DLAS_GPU || policy == GS_SCHED_GITTINS);
  const int nq = (policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU) ? (pol->num_queue > 0 ? pol->num_queue : 1) : 1;
  if (nq > GS_MAX_QUEUES) return GS_ERR_ARG;
  pcluster_t cl;
  cl.M = c->num_switch * c->num_node_p_switch; cl.G = c->num_gpu_p_node;
  { int kc = c->num_cpu_p_node / c->cpu_per_task, km = c->mem_p_node / c->mem_per_task; cl.K = kc < km ? kc : km; }
  cl.idle = (int *)malloc(sizeof(int) * (size_t)cl.M); cl.kfree = (int *)malloc(sizeof(int) * (size_t)cl.M);
  const int total_gpus = cl.M * cl.G;
  const int64_t cap_bytes = (int64_t)c->gpu_mem_cap_mib << 20, fit_limit = cap_bytes - ((int64_t)500 << 20);
  size_t N = (size_t)(n > 0 ? n : 1);
  pjob_t *jb = (pjob_t *)calloc(N, sizeof(pjob_t));
  int32_t *D = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *runnable = (int32_t *)malloc(N * sizeof(int32_t)); int64_t rn = 0;
  int32_t *queue[GS_MAX_QUEUES]; int64_t qn[GS_MAX_QUEUES];
  for (int q = 0; q < GS_MAX_QUEUES; ++q) { queue[q] = (int32_t *)malloc(N * sizeof(int32_t)); qn[q] = 0; }
  int32_t *end_jobs = (int32_t *)malloc(N * sizeof(int32_t)); int64_t en = 0; int end_time = T_INF;
  int32_t *tmp = (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *stale_jobs = (int32_t *)malloc(N * sizeof(int32_t)); int64_t stale_n = 0;   /* Q25, see the event selection */
  for (int64_t j = 0; j < n; ++j) {
    double cl_ = ceil(duration[j]); D[j] = cl_ < 1.0 ? 1 : (int32_t)cl_;
    jb[j].start = -1;
    jobs_out[j].start = -1; jobs_out[j].end = -1; jobs_out[j].jct = 0; jobs_out[j].preempt = 0; jobs_out[j].duration = duration[j];
  }
  int64_t next_row = 0, nfin = 0, events = 0, ticks = 0, rc = 0;
  int next_job_jump = T_INF;
  double next_gittins_unit = pol->gittins_delta;

  while ((n - next_row) + rn > 0) {                               /* run_sim.py:168 / :678 / :964 */
    if (next_row >= n && end_time == T_INF) break;               /* "cluster is not large enough" */
    int start_time = next_row < n ? arrive[next_row] : T_INF;
    int event_time, has_start = 0, has_end = 0;
    const int32_t *elist = end_jobs; int64_t ecount = en;
    if (end_time < start_time) { event_time = end_time; has_end = 1; }
    else if (end_time > start_time) { event_time = start_time; has_start = 1; }
    else {                                                        /* tie: the START event dict gets the key */
      event_time = start_time; has_start = 1; has_end = 1;        /* 'end_jobs' = this end list (:708-710)  */
      memcpy(stale_jobs, end_jobs, (size_t)en * sizeof(int32_t)); stale_n = en;
    }
    int jumped = 0;
    if ((policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU) && event_time > next_job_jump) {
      event_time = next_job_jump; jumped = 1;                    /* :715-717  event = dict() */
    }
    if (policy == GS_SCHED_GITTINS && (double)event_time > next_gittins_unit) {
      event_time = (int)next_gittins_unit; jumped = 1;           /* :1006-1008 */
    }
    if (jumped) { has_start = has_end = 0; }                      /* the start dict KEEPS its 'end_jobs' key */
    else if (has_start) {
      /* The start event dict is consumed now.  If an earlier tie left an 'end_jobs' key on it and a jump
       * was served in between, those jobs complete HERE, whatever their state is by now (quirk Q25;
       * reference :720-727 looks only at the key).  A fresh tie has just overwritten the key above. */
      if (stale_n > 0) { elist = stale_jobs; ecount = stale_n; has_end = 1; }
      stale_n = 0;
    }
    /* completions */
    if (has_end) {
      for (int64_t i = 0; i < ecount; ++i) {
        int32_t j = elist[i];
        jb[j].status = ST_END;
        jobs_out[j].start = jb[j].start; jobs_out[j].end = event_time;
        jobs_out[j].jct = D[j]; jobs_out[j].preempt = jb[j].resume;
        finish_order_out[nfin++] = j; ++events;
        list_remove(runnable, &rn, j);
        list_remove(queue[jb[j].q_id], &qn[jb[j].q_id], j);
      }
    }
    /* arrivals */
    if (has_start) {
      while (next_row < n && arrive[next_row] == event_time) {
        int32_t j = (int32_t)next_row++;
        jb[j].status = ST_PENDING; jb[j].last_check = event_time; jb[j].q_id = 0; jb[j].remaining = D[j];
        runnable[rn++] = j; queue[0][qn[0]++] = j; ++events;
      }
    }
    /* counters */
    for (int64_t i = 0; i < rn; ++i) {
      pjob_t *r = &jb[runnable[i]]; int32_t j = runnable[i];
      int dt = event_time - r->last_check;
      r->last_check = event_time;
      if (r->status == ST_RUNNING) {
        r->total_exec += dt; r->exec += dt; r->remaining = D[j] - r->total_exec;
        if (policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU) {
          double j_gt = gputime ? (double)r->exec * gpus[j] : (double)r->exec;
          if (r->q_id < nq - 1 && j_gt >= pol->queue_limit[r->q_id]) {        /* demotion :752-759 */
            list_remove(queue[r->q_id], &qn[r->q_id], j);
            r->q_id += 1; queue[r->q_id][qn[r->q_id]++] = j;
          }
        } else if (policy == GS_SCHED_GITTINS) {
          r->rank = gittins_index(pol, (double)r->exec * gpus[j]);            /* :1040-1053 */
        }
      } else {
        r->pending += dt;
        if (r->exec > 0) r->last_pending += dt;
        if (policy == GS_SCHED_GITTINS) r->rank = gittins_index(pol, (double)r->exec);   /* :1071-1073 */
      }
    }
    /* order, empty the cluster, greedy re-admission */
    int64_t nrun = 0, npre = 0, busy = 0; int64_t mem_busy = 0;
    int32_t *run_jobs = tmp;                                     /* tmp[0..nrun) run list, grows up */
    int32_t *pre_jobs = tmp + (n - 1 >= 0 ? n - 1 : 0);          /* preempt list grows down */
    if (policy == GS_SCHED_SJF) {
      stable_sort_int(runnable, rn, gpus);                       /* :237 */
      for (int nd = 0; nd < cl.M; ++nd) { cl.idle[nd] = cl.G; cl.kfree[nd] = cl.K; }
      for (int64_t i = 0; i < rn; ++i) {
        int32_t j = runnable[i];
        if (yarn_place(&cl, gpus[j], gpc[j], mem_bytes[j] < fit_limit)) {
          if (jb[j].start < 0) jb[j].start = event_time;
          if (jb[j].status == ST_PENDING) run_jobs[nrun++] = j;
          busy += gpus[j]; mem_busy += (int64_t)gpus[j] * (mem_bytes[j] < cap_bytes ? mem_bytes[j] : cap_bytes);
        } else if (jb[j].status == ST_RUNNING) { pre_jobs[-(npre++)] = j; }
      }
    } else {
      if (policy == GS_SCHED_GITTINS) { stable_sort_rank(runnable, rn, jb); }   /* :1083 */
      int free_gpu = total_gpus;
      int nlists = (policy == GS_SCHED_GITTINS) ? 1 : nq;
      for (int q = 0; q < nlists; ++q) {
        int32_t *lst = (policy == GS_SCHED_GITTINS) ? runnable : queue[q];
        int64_t ln = (policy == GS_SCHED_GITTINS) ? rn : qn[q];
        for (int64_t i = 0; i < ln; ++i) {
          int32_t j = lst[i];
          if (free_gpu >= gpus[j]) {
            if (jb[j].status == ST_PENDING) run_jobs[nrun++] = j;
            free_gpu -= gpus[j];
            busy += gpus[j]; mem_busy += (int64_t)gpus[j] * (mem_bytes[j] < cap_bytes ? mem_bytes[j] : cap_bytes);
          } else if (jb[j].status == ST_RUNNING) { pre_jobs[-(npre++)] = j; }
        }
      }
    }
    for (int64_t i = 0; i < npre; ++i) { pjob_t *r = &jb[pre_jobs[-i]]; r->status = ST_PENDING; r->preempt += 1; ++events; }
    for (int64_t i = 0; i < nrun; ++i) {
      pjob_t *r = &jb[run_jobs[i]]; r->status = ST_RUNNING; r->resume += 1; ++events;
      if (r->start < 0) r->start = event_time;
    }
    if (policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU) {  /* pending jobs go behind, :838-848 */
      for (int q = 0; q < nq; ++q) {
        int64_t w = 0, pn = 0;
        for (int64_t i = 0; i < qn[q]; ++i) { int32_t j = queue[q][i]; if (jb[j].status == ST_PENDING) tmp[pn++] = j; else queue[q][w++] = j; }
        for (int64_t i = 0; i < pn; ++i) queue[q][w++] = tmp[i];
      }
    }
    /* next completion: the earliest end among RUNNING jobs, ties in runnable order */
    end_time = T_INF; en = 0;
    for (int64_t i = 0; i < rn; ++i) {
      int32_t j = runnable[i];
      if (jb[j].status != ST_RUNNING) continue;
      int e = event_time + (D[j] - jb[j].total_exec);
      if (e < end_time) { end_time = e; en = 0; end_jobs[en++] = j; }
      else if (e == end_time) end_jobs[en++] = j;
    }
    /* next queue jump (dlas) :925-935 */
    if (policy == GS_SCHED_DLAS || policy == GS_SCHED_DLAS_GPU) {
      next_job_jump = T_INF;
      for (int64_t i = 0; i < rn; ++i) {
        int32_t j = runnable[i]; pjob_t *r = &jb[j];
        if (r->status != ST_RUNNING || r->q_id >= nq - 1) continue;
        double lim = pol->queue_limit[r->q_id];
        double jt = gputime ? ceil((lim - (double)r->exec) / (double)gpus[j]) + event_time : lim - (double)r->exec + event_time;
        int jti = jt > 2.0e9 ? T_INF : (int)jt;
        if (jti < next_job_jump) next_job_jump = jti;
      }
    }
    if (policy == GS_SCHED_GITTINS) next_gittins_unit += (double)event_time;      /* :1202, verbatim */
    /* checkpoint row */
    if (ticks >= rows_cap) { rc = GS_ERR_CAPACITY; break; }
    gs_tick_row *row = &rows_out[ticks++];
    memset(row, 0, sizeof(*row));
    row->now = event_time;
    row->busy_gpus = (int32_t)busy; row->idle_gpus = total_gpus - (int32_t)busy;
    row->mem_busy_bytes = mem_busy; row->finished = (int32_t)nfin;
    for (int64_t i = 0; i < rn; ++i) {
      pjob_t *r = &jb[runnable[i]];
      if (r->status == ST_RUNNING) row->running++;
      else { row->queued++; row->pend_sum += r->pending; if (r->pending > row->pend_max) row->pend_max = r->pending; }
    }
    if (policy == GS_SCHED_SJF) {
      for (int nd = 0; nd < cl.M; ++nd) row->busy_nodes += (cl.idle[nd] < cl.G);
      row->idle_nodes = cl.M - row->busy_nodes;
    } else { row->idle_nodes = cl.M; }
  }
  *n_finished_out = nfin;
  if (events_out) *events_out = events;
  for (int64_t j = 0; j < n; ++j)
    if (jb[j].status != ST_END && jb[j].start >= 0) { jobs_out[j].start = jb[j].start; jobs_out[j].preempt = jb[j].resume; }
  free(cl.idle); free(cl.kfree); free(jb); free(D); free(runnable); free(end_jobs); free(tmp); free(stale_jobs);
  for (int q = 0; q < GS_MAX_QUEUES; ++q) free(queue[q]);
  return rc < 0 ? rc : ticks;
}
