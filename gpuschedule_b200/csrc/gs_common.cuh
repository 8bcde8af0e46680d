// gs_common.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// Device-visible records (trace row, per-job state, per-replica descriptor) and lane-mask helpers.
#pragma once

#ifndef GS_TICK_MINBLOCKS
#define GS_TICK_MINBLOCKS 28     // 72 registers: no spills (64 spill), 28 warps per SM; measured best (profiles/r02_kernel_versions.md)
#endif
#ifndef GS_POLICY_MINBLOCKS
#define GS_POLICY_MINBLOCKS 20   // event-driven policy kernels (one warp per replica): resident warps per SM the registers are cut for (96 / 94 registers);
                                 // measured 16 / 20 / 24: sjf 8.9 / 8.8 / 9.5e8, dlas-gpu 1.23 / 1.32 / 1.31e9, gittins 1.31 / 1.39 / 1.40e9 events/s at 148 x that many replicas
#endif
#define FULL 0xffffffffu

// ------------------------------------------------------------------ device state

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
  gs_job_rec *rec;            // event-driven policies: full 24-byte record per job
  int *jstart;                // fifo: start tick per job (-1 = never started)
  double *dur2;               // fifo with network costs: job.duration after the cost was added
  struct JobState2 *jst2;     // fifo: release record of a running job
  int *stack, *fin, *wheel_head;
  long long *wheel_mem;       // per finish-tick bucket: memory share of the jobs ending there
  void *spans;                // start order; gs_cspan (8 B) when G <= 32, else gs_span with bit 31 of ntasks = first span of a job
  gs_tick_row *rows;          // event-driven policies: one row per event
  gs_evrow *evrows;           // fifo: one record per tick on which a counter changed
  gs_qrow *qrows;             // fifo: queue statistics beside the records taken with a non-empty queue
  gs_nodeev *nodeev;          // fifo: (tick, nodes that ever hosted a job) whenever that count grows / at the start of a window
  unsigned long long *nbusy;  // persisted node table (between launches)
  int *nk;                    // bit31 = node ever hosted a placement (node.py:93-97, never cleared)
  long long span_cap, rows_cap, qrows_cap;
  // ---- event-driven policies (sjf / dlas / dlas-gpu / gittins): scratch + parameters
  struct PJob *pj;            // per-job dynamic state
  int *runnable, *queues, *endj, *tmpl, *cidle, *ckfree;   // queues: num_queue lists of n entries
  int *stalej;                // end list a start event inherited from a tie it lost to a jump (quirk Q25)
  const double *git_data, *git_index;                      // device copies of the gittins tables
  const double *git_direct;                                // index value for every integer attained service below git_direct_n (or null)
  long long git_direct_n;
  double queue_limit[GS_MAX_QUEUES];
  double gittins_delta, next_gittins_unit;
  int num_queue, git_n, rn, en, end_time, next_job_jump, stale_n, qn[GS_MAX_QUEUES];
  // ---- sharded single simulation (gs_comm_init): the gittins rank evaluation of every event is split over the GPUs
  // of one box; each rank stores the ranks it computed straight into every peer's receive buffer over NVLink
  // (peer stores) and publishes an event counter; nothing else crosses the link
  int comm_rank, comm_n;                     // comm_n <= 1: not sharded
  int comm_min_runnable, comm_pad;           // events with at most this many runnable jobs are evaluated locally, without an exchange
  long long comm_cap;                        // rank values per receive buffer (>= n)
  double *comm_rk_in;                        // local receive buffers: 2 x comm_cap doubles, selected by event parity
  unsigned long long *comm_flags;            // local flags[comm_n]: the event counter last published by each rank
  double *comm_peer_rk[GS_MAX_RANKS];        // every rank's receive buffers (this rank's own at [comm_rank])
  unsigned long long *comm_peer_flags[GS_MAX_RANKS];
  unsigned long long comm_epoch;             // events exchanged so far (continues across launches)
  long long comm_wait_cycles;                // SM cycles between publishing and seeing every peer's counter, summed
  // ---- loop state (persisted)
  int delta, p, top, running, finished, ever, busy_gpus, done, status, need_init;
  int blocked;                // fifo: the queue head did not fit and nothing has changed since
  int nev, nq;                // fifo: records / queue records written by the last launch
  int nne;                    // fifo: node events written by the last launch
  long long mem_busy, sum_arr, span_used, events, evals, started, ticks, row_first;
};

#define EVER_BIT 0x80000000u

