// gs_common.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// Device-visible records (trace row, per-job state, per-replica descriptor) and lane-mask helpers.
#pragma once

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
  int *stalej;                // end list a start event inherited from a tie it lost to a jump (quirk Q25)
  const double *git_data, *git_index;                      // device copies of the gittins tables
  double queue_limit[GS_MAX_QUEUES];
  double gittins_delta, next_gittins_unit;
  int num_queue, git_n, rn, en, end_time, next_job_jump, stale_n, qn[GS_MAX_QUEUES];
  // ---- loop state (persisted)
  int delta, p, top, running, finished, ever, busy_gpus, done, status, need_init;
  long long mem_busy, sum_arr, span_used, events, evals, started, ticks, row_first;
};

#define EVER_BIT 0x80000000u

