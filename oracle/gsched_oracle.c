/*
 * gsched_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's LIVE hot path
 * (fifo policy + yarn placement + optional network cost), function by
 * function, each citing the reference file:line it follows (paths relative to
 * /root/reference).  It deliberately keeps the reference's own shape: per-tick
 * re-scans of every node and device for the statistics row, per-tick aging of
 * every queued and running job -- it is the checker, so it favours being
 * obviously equal to the Python over being fast.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * every fixture under tests/golden/, which were produced by running the
 * unmodified reference (tests/golden/make_golden.py).  The network-cost branch
 * (net_cost below) is pinned too: tests/golden/netcost.json holds return values of
 * the reference's own calculate_network_costs, and the netcost / netcost_lat fixtures
 * are reference runs with the three attributes core/jobs/job.py:199-200 and
 * network_service.py:34-36 read (ps_count, model_size, iterations) attached to Job.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.  The product (libgsched.so) never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gsched.h"

typedef struct {
  /* --- Infrastructure / Node / Device (infra/infrastructure.py:45-58, node.py:7-34, device.py:6-14) */
  int M, G;
  int cpu_count, mem_size;
  int64_t cap_bytes;          /* Device.memory in bytes (MiB * 2^20)            */
  int cpu_per_task, mem_per_task;
  int *cpu_used, *mem_used;   /* Node.cpu_used / mem_used                       */
  int *dev_job, *dev_task;    /* Device.running_tasks (non-pack: at most one)   */
  int *node_running;          /* len(Node.running_tasks)                        */
  int *node_placed_tasks;     /* len(Node.placed_tasks)                         */
  int *node_placed_jobs;      /* len(Node.placed_jobs): never cleared on finish */
  /* --- trace */
  int64_t n;
  const int32_t *arrive, *gpus, *gpc;
  const double *duration_in;
  const int64_t *mem_bytes;
  const double *model_mb, *iterations;
  const int32_t *ps_count;
  /* --- Job state (core/jobs/job.py:60-110) */
  double *duration;           /* Job.duration (network cost may grow it)        */
  int32_t *pending;           /* Job.pending_time                               */
  int32_t *processed;         /* max Task.time_processed                        */
  int32_t *start, *end;
  int64_t *task_off;          /* Job.tasks_running_on as a flat array            */
  int32_t *task_node;
  uint64_t *task_mask;        /* devices each task landed on (for the RNG column replay) */
  /* --- JobQueueManager.queues[0] (head == top of this stack) and JobsManager dicts */
  int32_t *queue; int64_t qlen;
  int32_t *running; int64_t rlen;
  int32_t *finished; int64_t flen;
  int64_t next_row;           /* JobTraceReader: first row with generated == 0  */
  int64_t evals;
} sim_t;

/* ---------------------------------------------------------------- node.py */

/* Node.get_free_devices(pack=False)                        infra/node.py:99-107 */
static int node_free_devices(const sim_t *s, int nd) {
  int cnt = 0;
  for (int d = 0; d < s->G; ++d) cnt += (s->dev_job[nd * s->G + d] < 0);
  return cnt;
}

/* Node.is_free                                             infra/node.py:59-60 */
static int node_is_free(const sim_t *s, int nd) {
  return (s->cpu_count - s->cpu_used[nd]) > 0 || (s->mem_size - s->mem_used[nd]) > 0;
}

/* Node.can_fit_num_task(tasks) with len(tasks) == remaining     node.py:109-127 */
static int node_can_fit_num_task(const sim_t *s, int nd, int gpc, int remaining) {
  int g = node_free_devices(s, nd) / gpc - remaining;
  int c = (s->cpu_count - s->cpu_used[nd]) / s->cpu_per_task - remaining;
  int m = (s->mem_size - s->mem_used[nd]) / s->mem_per_task - remaining;
  int ng = g >= 0 ? remaining : remaining + g;
  int nc = c >= 0 ? remaining : remaining + c;
  int nm = m >= 0 ? remaining : remaining + m;
  int r = nc < nm ? nc : nm;
  return r < ng ? r : ng;
}

/* Node.can_fit(task, pack=False)                               node.py:146-171 */
static int node_can_fit(const sim_t *s, int nd, int gpc) {
  int cpu_off = (s->cpu_count - s->cpu_used[nd]) - s->cpu_per_task;
  int mem_off = (s->mem_size - s->mem_used[nd]) - s->mem_per_task;
  if (cpu_off < 0 || mem_off < 0) return 0;
  return node_free_devices(s, nd) - gpc >= 0;
}

/* Device.can_fit on an EMPTY device: cap - (0 + task_mem) > 500 MiB
 * (device.py:67-77).  All quantities are multiples of 2^-20 MiB, so the float
 * comparison of the reference is exact in integer bytes.                      */
static int device_can_fit(const sim_t *s, int64_t task_bytes) {
  return s->cap_bytes - task_bytes > (int64_t)500 * (1 << 20);
}

/* Node.try_reserve_and_placed_task(task)                       node.py:200-221
 * NOTE the reference charges cpu/mem BEFORE walking the devices and never
 * refunds them when the devices refuse the task (quirk Q21): kept.            */
static int node_try_reserve_and_placed_task(sim_t *s, int nd, int job, int task) {
  int gpc = s->gpc[job];
  if (!node_can_fit(s, nd, gpc)) return 0;
  s->cpu_used[nd] += s->cpu_per_task;
  s->mem_used[nd] += s->mem_per_task;
  int should = gpc;
  for (int d = 0; d < s->G && should > 0; ++d) {
    int *own = &s->dev_job[nd * s->G + d];
    /* Device.add_task: can_fit first, then refuse a non-empty device (device.py:19-24) */
    if (!device_can_fit(s, s->mem_bytes[job])) continue;
    if (*own >= 0) continue;
    *own = job;
    s->dev_task[nd * s->G + d] = task;
    if (s->task_mask) s->task_mask[s->task_off[job] + task] |= (uint64_t)1 << d;
    --should;
  }
  if (should == 0) s->node_placed_tasks[nd] += 1;
  return should == 0;
}

/* Node.release_allocated_resources(task)                        node.py:71-91 */
static void node_release(sim_t *s, int nd, int job, int task) {
  s->cpu_used[nd] -= s->cpu_per_task;
  s->mem_used[nd] -= s->mem_per_task;
  for (int d = 0; d < s->G; ++d) {
    int i = nd * s->G + d;
    if (s->dev_job[i] == job && s->dev_task[i] == task) { s->dev_job[i] = -1; s->dev_task[i] = -1; }
  }
}
/* rollback only: forget which devices the task had */
static void task_mask_clear(sim_t *s, int job, int task) {
  if (s->task_mask) s->task_mask[s->task_off[job] + task] = 0;
}

/* ---------------------------------------------------------- algorithm.py */

static int job_tasks(const sim_t *s, int j) { return s->gpus[j] / s->gpc[j]; }   /* job.py:96-98 */

/* try_single_node_alloc_ms                      core/scheduling/algorithm.py:396-417
 * with Node.try_alloc_job(job, True)                       infra/node.py:245-275 */
static int try_single_node_alloc_ms(sim_t *s, int j) {
  int tasks = job_tasks(s, j), gpc = s->gpc[j];
  for (int nd = 0; nd < s->M; ++nd) {
    s->evals++;
    if (!node_is_free(s, nd)) continue;                      /* get_free_nodes */
    if (!(node_free_devices(s, nd) >= s->gpus[j] &&
          (s->cpu_count - s->cpu_used[nd]) >= s->cpu_per_task * tasks &&
          (s->mem_size - s->mem_used[nd]) >= s->mem_per_task * tasks)) continue;
    /* try_alloc_job */
    if (node_can_fit_num_task(s, nd, gpc, tasks) < tasks) continue;
    int placed = 0;
    for (int t = 0; t < tasks; ++t) {
      if (node_try_reserve_and_placed_task(s, nd, j, t)) {
        s->task_node[s->task_off[j] + t] = nd;
        ++placed;
      }
    }
    if (placed == 0) continue;       /* devices refused every task: resources leaked, next node */
    s->node_placed_jobs[nd] += 1;    /* try_reserve_and_placed_job      node.py:223-243 */
    return 1;
  }
  return 0;
}

/* try_cross_node_alloc_ms                       core/scheduling/algorithm.py:301-393 */
static int try_cross_node_alloc_ms(sim_t *s, int j) {
  int tasks = job_tasks(s, j), gpc = s->gpc[j];
  int assigned = 0, n_nodes = 0;
  int least = (s->gpus[j] + s->G - 1) / s->G;                 /* :310 */
  int *touched = (int *)malloc(sizeof(int) * (size_t)s->M);
  for (int nd = 0; nd < s->M; ++nd) {
    s->evals++;
    if (!node_is_free(s, nd)) continue;                       /* :326 */
    if (assigned == tasks) break;                             /* :328 */
    int can = node_can_fit_num_task(s, nd, gpc, tasks - assigned);   /* :331 */
    if (can == 0) continue;
    int worker_count = 0, check_next = 0;
    for (int t = assigned; t < tasks; ++t) {                  /* :338-361, unassigned tasks in id order */
      if (!(worker_count <= can)) continue;                   /* the `<=` over-try, quirk Q7 */
      ++worker_count;
      if (!node_try_reserve_and_placed_task(s, nd, j, t)) {
        --worker_count;
        check_next = 1;
        break;
      }
      s->task_node[s->task_off[j] + t] = nd;
    }
    if (worker_count > 0) {                                   /* :364-368 */
      assigned += worker_count;
      s->node_placed_jobs[nd] += 1;
      touched[n_nodes++] = nd;
    }
    if (check_next) continue;
    if (n_nodes >= least && assigned == tasks) break;         /* :373 */
  }
  int ok = (assigned == tasks && n_nodes >= least);
  if (!ok) {                                                  /* rollback :378-387 */
    for (int k = 0; k < n_nodes; ++k) {
      int nd = touched[k];
      s->node_placed_jobs[nd] -= 1;
      for (int t = 0; t < tasks; ++t) {
        if (t < assigned && s->task_node[s->task_off[j] + t] == nd) {
          s->node_placed_tasks[nd] -= 1;
          node_release(s, nd, j, t);
          task_mask_clear(s, j, t);
        }
      }
    }
    for (int t = 0; t < tasks; ++t) s->task_node[s->task_off[j] + t] = -1;
  }
  free(touched);
  return ok;
}

/* ms_yarn_placement                              core/scheduling/algorithm.py:28-32 */
static int ms_yarn_placement(sim_t *s, int j) {
  return s->gpus[j] > s->G ? try_cross_node_alloc_ms(s, j) : try_single_node_alloc_ms(s, j);
}

/* calculate_network_costs                   core/network/network_service.py:3-39
 * The live Job creates only 'worker*' tasks (job.py:100), so ps_nodes is empty
 * and the symmetric difference is the set of distinct worker nodes.
 * Pinned (see header).  Exact operation order of :34-37 is kept.               */
static double net_cost(const sim_t *s, const gs_cluster *c, int j) {
  if (!s->ps_count || !(s->ps_count[j] > 1)) return 0.0;      /* is_distributed, job.py:199-200 */
  int tasks = job_tasks(s, j), cross = 0;
  for (int t = 0; t < tasks; ++t) {
    int nd = s->task_node[s->task_off[j] + t], seen = 0;
    for (int u = 0; u < t; ++u) seen |= (s->task_node[s->task_off[j] + u] == nd);
    cross += !seen;
  }
  if (cross == 0) return 0.0;
  volatile double model_per_sec = s->model_mb[j] / c->bandwidth;
  volatile double nodes_induced = (double)cross * c->internode_latency;
  volatile double round_trip = s->iterations[j] * 2.0;
  volatile double sum = model_per_sec + nodes_induced;
  return sum * round_trip;
}

/* ------------------------------------------------------- jobs_manager.py */

/* gen_jobs + insert for fifo: the new batch lands at indices 0..k-1 of the
 * list, i.e. AHEAD of everything queued (jobs_manager.py:52-55,134-135,228-241;
 * job_queue_manager.py:154).  `queue` is kept as a stack whose top is the list
 * head, so the batch is pushed last-to-first.                                  */
static int gen_jobs(sim_t *s, int delta) {
  int64_t a = s->next_row, b = a;
  while (b < s->n && s->arrive[b] <= delta) ++b;              /* normalized_time <= delta, job_generator.py:203 */
  for (int64_t k = b - 1; k >= a; --k) s->queue[s->qlen++] = (int32_t)k;
  s->next_row = b;
  return (int)(b - a);
}

/* JobsManager.step                                        jobs_manager.py:143-148 */
static void step(sim_t *s) {
  for (int64_t q = 0; q < s->qlen; ++q) s->pending[s->queue[q]] += 1;      /* add_pending_time :65-70 */
  for (int64_t r = 0; r < s->rlen; ++r) s->processed[s->running[r]] += 1;  /* Job.step job.py:183-188 */
}

/* prepare_finish_tasks + release_finished_jobs   jobs_manager.py:243-250, schedule.py:141-162 */
static void release_finished_jobs(sim_t *s, int now) {
  int64_t w = 0;
  for (int64_t r = 0; r < s->rlen; ++r) {
    int j = s->running[r];
    /* time_processed() < get_duration(): get_duration = max(task.duration, job.duration), job.py:206-210 */
    double dur = s->duration[j] > s->duration_in[j] ? s->duration[j] : s->duration_in[j];
    if ((double)s->processed[j] < dur) { s->running[w++] = j; continue; }
    int tasks = job_tasks(s, j);
    for (int t = 0; t < tasks; ++t) {
      int nd = s->task_node[s->task_off[j] + t];
      s->node_running[nd] -= 1;
      node_release(s, nd, j, t);
    }
    s->end[j] = now;                                          /* Job.try_finished job.py:143-152 */
    s->finished[s->flen++] = j;
  }
  s->rlen = w;
}

/* Scheduler._construct_info + pending_time_infos   schedule.py:95-133, jobs_manager.py:72-87 */
static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}
static void construct_info(const sim_t *s, int now, gs_tick_row *row, int32_t *scratch) {
  memset(row, 0, sizeof(*row));
  row->now = now;
  for (int nd = 0; nd < s->M; ++nd) {
    int idle = (s->node_running[nd] + s->node_placed_tasks[nd] + s->node_placed_jobs[nd]) == 0; /* node.py:93-97 */
    row->idle_nodes += idle;
    row->busy_nodes += !idle;
    for (int d = 0; d < s->G; ++d) {
      int j = s->dev_job[nd * s->G + d];
      if (j < 0) { row->idle_gpus++; continue; }
      row->busy_gpus++;
      int64_t mb = s->mem_bytes[j];                           /* Device.get_current_memory device.py:56-62 */
      row->mem_busy_bytes += mb < s->cap_bytes ? mb : s->cap_bytes;
    }
  }
  for (int64_t q = 0; q < s->qlen; ++q) {
    int32_t p = s->pending[s->queue[q]];
    scratch[q] = p;
    row->pend_sum += p;
    if (p > row->pend_max) row->pend_max = p;
  }
  if (s->qlen > 0) {                                          /* np.median: mean of the two middle values */
    qsort(scratch, (size_t)s->qlen, sizeof(int32_t), cmp_i32);
    row->pend_med_lo = scratch[(s->qlen - 1) / 2];
    row->pend_med_hi = scratch[s->qlen / 2];
  }
  row->running = (int32_t)s->rlen;
  row->queued = (int32_t)s->qlen;
  row->finished = (int32_t)s->flen;
}

/* ------------------------------------------------------------ schedule.py */

/* Scheduler.start with _schedule and schedule_fifo inlined.
 * core/scheduling/schedule.py:40-60,178-215; core/scheduling/algorithm.py:189-202.
 * Returns the number of ticks simulated, or a negative gs_status.              */
int64_t oracle_run_fifo(const gs_cluster *c, int64_t n, const int32_t *arrive,
                        const int32_t *gpus, const int32_t *gpc, const double *duration,
                        const int64_t *mem_bytes, const double *model_mb,
                        const double *iterations, const int32_t *ps_count,
                        gs_tick_row *rows_out, int64_t rows_cap, gs_job_rec *jobs_out,
                        int32_t *finish_order_out, int64_t *n_finished_out,
                        const int64_t *task_off, int32_t *task_node_out,
                        uint64_t *task_mask_out, int64_t *events_out, int64_t *evals_out) {
  sim_t s;
  memset(&s, 0, sizeof(s));
  s.M = c->num_switch * c->num_node_p_switch;
  s.G = c->num_gpu_p_node;
  if (s.M <= 0 || s.G <= 0 || s.G > GS_MAX_GPUS_PER_NODE) return GS_ERR_ARG;
  s.cpu_count = c->num_cpu_p_node;
  s.mem_size = c->mem_p_node;
  s.cap_bytes = (int64_t)c->gpu_mem_cap_mib << 20;
  s.cpu_per_task = c->cpu_per_task;
  s.mem_per_task = c->mem_per_task;
  s.n = n; s.arrive = arrive; s.gpus = gpus; s.gpc = gpc; s.duration_in = duration;
  s.mem_bytes = mem_bytes; s.model_mb = model_mb; s.iterations = iterations; s.ps_count = ps_count;
  size_t MG = (size_t)s.M * (size_t)s.G, N = (size_t)(n > 0 ? n : 1);
  s.cpu_used = (int *)calloc((size_t)s.M, sizeof(int));
  s.mem_used = (int *)calloc((size_t)s.M, sizeof(int));
  s.node_running = (int *)calloc((size_t)s.M, sizeof(int));
  s.node_placed_tasks = (int *)calloc((size_t)s.M, sizeof(int));
  s.node_placed_jobs = (int *)calloc((size_t)s.M, sizeof(int));
  s.dev_job = (int *)malloc(MG * sizeof(int));
  s.dev_task = (int *)malloc(MG * sizeof(int));
  for (size_t i = 0; i < MG; ++i) { s.dev_job[i] = -1; s.dev_task[i] = -1; }
  s.duration = (double *)malloc(N * sizeof(double));
  s.pending = (int32_t *)calloc(N, sizeof(int32_t));
  s.processed = (int32_t *)calloc(N, sizeof(int32_t));
  s.start = (int32_t *)malloc(N * sizeof(int32_t));
  s.end = (int32_t *)malloc(N * sizeof(int32_t));
  s.queue = (int32_t *)malloc(N * sizeof(int32_t));
  s.running = (int32_t *)malloc(N * sizeof(int32_t));
  s.finished = finish_order_out ? finish_order_out : (int32_t *)malloc(N * sizeof(int32_t));
  int32_t *scratch = (int32_t *)malloc(N * sizeof(int32_t));
  s.task_off = (int64_t *)task_off;
  s.task_node = task_node_out;
  s.task_mask = task_mask_out;
  if (task_mask_out) memset(task_mask_out, 0, sizeof(uint64_t) * (size_t)task_off[n]);
  for (int64_t j = 0; j < n; ++j) { s.duration[j] = duration[j]; s.start[j] = -1; s.end[j] = -1; }
  for (int64_t t = 0; t < task_off[n]; ++t) task_node_out[t] = -1;

  int64_t events = 0, ticks = 0;
  int delta = 0;
  int64_t remaining = n;                         /* remaining_jobs(delta)      schedule.py:181 */
  int64_t running_jobs = 0;
  int64_t rc = 0;
  while (remaining + running_jobs > 0) {         /* :185 -- the queue is NOT counted (quirk Q4) */
    events += gen_jobs(&s, delta);               /* :187 */
    if (s.qlen > 0) {                            /* :188-190 -> _schedule */
      int free_nodes = 0;
      for (int nd = 0; nd < s.M; ++nd) free_nodes += node_is_free(&s, nd);
      if (free_nodes >= 1) {                     /* :41 */
        int j = s.queue[s.qlen - 1];             /* get_next_job: queue head  jobs_manager.py:32-37 */
        if (ms_yarn_placement(&s, j)) {          /* schedule_fifo             algorithm.py:197 */
          s.qlen -= 1;                           /* jobs_manager.pop          algorithm.py:199 */
          if (c->enable_network_costs)           /* schedule.py:49-52 */
            s.duration[j] += net_cost(&s, c, j);
          int tasks = job_tasks(&s, j);          /* add_to_running -> start_job -> execute_job */
          for (int t = 0; t < tasks; ++t) {
            int nd = s.task_node[s.task_off[j] + t];
            s.node_placed_tasks[nd] -= 1;        /* placed_tasks.pop -> running_tasks  node.py:191-195 */
            s.node_running[nd] += 1;
          }
          s.start[j] = delta;                    /* Job.try_execute           job.py:160-175 */
          s.running[s.rlen++] = j;
          events += 1;
        }
      } else {
        s.evals += s.M;   /* metric only: the engine still scores all M nodes (none can fit) */
      }
    }
    remaining = n - s.next_row;                  /* :191 */
    delta += 1;                                  /* :193 */
    step(&s);                                    /* :194 */
    int64_t f0 = s.flen;
    release_finished_jobs(&s, delta);            /* :195 */
    events += s.flen - f0;
    running_jobs = s.rlen;                       /* :196 */
    if (ticks >= rows_cap) { rc = GS_ERR_CAPACITY; break; }
    construct_info(&s, delta, &rows_out[ticks], scratch);   /* :204-205 */
    ++ticks;
  }
  for (int64_t j = 0; j < n; ++j) {
    jobs_out[j].start = s.start[j];
    jobs_out[j].end = s.end[j];
    jobs_out[j].jct = s.end[j] >= 0 ? s.processed[j] : 0;    /* time_processed()  log_manager.py:151 */
    jobs_out[j].preempt = s.start[j] >= 0 ? 1 : 0;           /* migration_count   job.py:171 */
    jobs_out[j].duration = s.duration[j];
  }
  *n_finished_out = s.flen;
  if (events_out) *events_out = events;
  if (evals_out) *evals_out = s.evals;
  free(s.cpu_used); free(s.mem_used); free(s.node_running); free(s.node_placed_tasks);
  free(s.node_placed_jobs); free(s.dev_job); free(s.dev_task); free(s.duration);
  free(s.pending); free(s.processed); free(s.start); free(s.end); free(s.queue);
  free(s.running); free(scratch);
  if (!finish_order_out) free(s.finished);
  return rc < 0 ? rc : ticks;
}

/* Stateless single-job placement against a caller-supplied cluster state: the
 * unit-level checker for gs_place_batch.  Follows ms_yarn_placement exactly
 * (on a private copy of the state) and reports the per-task nodes.            */
int oracle_place_one(const gs_cluster *c, const gs_node *nodes, int32_t m,
                     const gs_jobreq *job, int32_t *first_node, int32_t *nodes_used,
                     int32_t *task_node /* gpus/gpc entries */) {
  sim_t s;
  memset(&s, 0, sizeof(s));
  s.M = m; s.G = c->num_gpu_p_node;
  s.cpu_count = c->num_cpu_p_node; s.mem_size = c->mem_p_node;
  s.cap_bytes = (int64_t)c->gpu_mem_cap_mib << 20;
  s.cpu_per_task = c->cpu_per_task; s.mem_per_task = c->mem_per_task;
  size_t MG = (size_t)s.M * (size_t)s.G;
  s.cpu_used = (int *)malloc(sizeof(int) * (size_t)m);
  s.mem_used = (int *)malloc(sizeof(int) * (size_t)m);
  s.node_placed_tasks = (int *)calloc((size_t)m, sizeof(int));
  s.node_placed_jobs = (int *)calloc((size_t)m, sizeof(int));
  s.dev_job = (int *)malloc(MG * sizeof(int));
  s.dev_task = (int *)malloc(MG * sizeof(int));
  for (int nd = 0; nd < m; ++nd) {
    s.cpu_used[nd] = nodes[nd].cpu_used;
    s.mem_used[nd] = nodes[nd].mem_used;
    for (int d = 0; d < s.G; ++d) {
      int busy = (int)((nodes[nd].busy_mask >> d) & 1u);
      s.dev_job[nd * s.G + d] = busy ? 1 : -1;     /* job 1 = "someone else" */
      s.dev_task[nd * s.G + d] = busy ? 0 : -1;
    }
  }
  int32_t gpus = job->gpus, gpc = job->gpu_per_task;
  int64_t mem = job->mem_bytes, off[2] = {0, gpus / gpc};
  s.n = 1; s.gpus = &gpus; s.gpc = &gpc; s.mem_bytes = &mem;
  s.task_off = off; s.task_node = task_node;
  for (int t = 0; t < gpus / gpc; ++t) task_node[t] = -1;
  int ok = ms_yarn_placement(&s, 0);
  *first_node = -1; *nodes_used = 0;
  if (ok) {
    *first_node = task_node[0];
    for (int nd = 0; nd < m; ++nd) *nodes_used += (s.node_placed_jobs[nd] > 0);
  }
  free(s.cpu_used); free(s.mem_used); free(s.node_placed_tasks); free(s.node_placed_jobs);
  free(s.dev_job); free(s.dev_task);
  return ok;
}

/* Public wrapper of net_cost for the unit checker of gs_net_cost: general form
 * with PS marks (network_service.py:16-24): |ps_nodes symmetric-difference wk_nodes|. */
double oracle_net_cost(const gs_cluster *c, int32_t n_tasks, const int32_t *task_node,
                       const uint8_t *is_ps, int32_t ps_count, double model_mb, double iterations) {
  if (!(ps_count > 1)) return 0.0;
  int cross = 0;
  for (int t = 0; t < n_tasks; ++t) {
    int nd = task_node[t], first = 1, in_ps = 0, in_wk = 0;
    for (int u = 0; u < t; ++u) first &= (task_node[u] != nd);
    if (!first) continue;
    for (int u = 0; u < n_tasks; ++u)
      if (task_node[u] == nd) { if (is_ps && is_ps[u]) in_ps = 1; else in_wk = 1; }
    cross += (in_ps != in_wk);
  }
  if (cross == 0) return 0.0;
  volatile double model_per_sec = model_mb / c->bandwidth;
  volatile double nodes_induced = (double)cross * c->internode_latency;
  volatile double round_trip = iterations * 2.0;
  volatile double sum = model_per_sec + nodes_induced;
  return sum * round_trip;
}
