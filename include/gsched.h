/*
 * gsched.h -- C ABI of libgsched.so, the sm_100a discrete-event engine that
 * replaces the per-tick hot path of matthewygf/GPUSchedule.
 *
 * The reference has no FFI: its "operator API" is the Python loop object and
 * three dict registries.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference root):
 *
 *   gs_run            Scheduler.start()            core/scheduling/schedule.py:178-215
 *                     (gen_jobs jobs_manager.py:228-241, _schedule schedule.py:40-60,
 *                      schedule_fifo algorithm.py:189-202, ms_yarn_placement :28-32,
 *                      step jobs_manager.py:143-148, release_finished_jobs
 *                      schedule.py:141-162, _construct_info schedule.py:95-133)
 *   gs_place_batch    placement_algorithms['yarn'](infrastructure, job, scheme)
 *                     core/scheduling/algorithm.py:28-32,301-393,396-417
 *   gs_net_cost       calculate_network_costs(infrastructure, job)
 *                     core/network/network_service.py:3-39
 *   gs_load_trace     JobsManager.gen_jobs' per-row Job(...) construction
 *                     core/jobs/jobs_manager.py:233-239 (rows arrive pre-sorted by
 *                     the host ingest, job_generator.py:181-193)
 *   gs_config_sim     Infrastructure(FLAGS)        infra/infrastructure.py:26-58
 *
 * Conventions: every function returns 0 on success and a negative gs_status
 * on failure (message via gs_last_error); nothing throws or exits across the
 * boundary.  The caller owns every host buffer and passes sizes explicitly;
 * the library owns all device memory and its CUDA stream.  One handle is
 * bound to one CUDA device and must be driven by one host thread at a time;
 * handles are independent.  There is NO CPU fallback: without a usable CUDA
 * device gs_create fails with GS_ERR_CUDA.
 */
#ifndef GSCHED_H_
#define GSCHED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 4
#define GS_MAX_QUEUES 8
#define GS_MAX_GPUS_PER_NODE 64
#define GS_MAX_RANKS 8          /* GPUs of one box that may share one simulation (gs_comm_init) */

typedef enum {
  GS_OK = 0,
  GS_ERR_ARG = -1,      /* bad argument / unsupported configuration          */
  GS_ERR_CUDA = -2,     /* CUDA runtime or driver error                      */
  GS_ERR_STATE = -3,    /* call order (e.g. gs_run before gs_load_trace)     */
  GS_ERR_CAPACITY = -4, /* an output buffer is too small                     */
  GS_ERR_COMM = -5      /* sharded multi-GPU path: a peer did not answer in time / IPC error */
} gs_status;

/* scheduling policies (run_sim.py:37-49 names) and placement schemes (:25-36) */
enum { GS_SCHED_FIFO = 0, GS_SCHED_SJF = 1, GS_SCHED_DLAS = 2,
       GS_SCHED_DLAS_GPU = 3, GS_SCHED_GITTINS = 4 };
enum { GS_SCHEME_YARN = 0, GS_SCHEME_COUNT = 1 };

/* Cluster description == the flags Infrastructure reads (infrastructure.py:26-43). */
typedef struct gs_cluster {
  int32_t num_switch;
  int32_t num_node_p_switch;
  int32_t num_gpu_p_node;        /* G <= GS_MAX_GPUS_PER_NODE                  */
  int32_t num_cpu_p_node;
  int32_t mem_p_node;
  int32_t gpu_mem_cap_mib;       /* gpu_memory_capacity * 1024 (infrastructure.py:36) */
  int32_t enable_network_costs;  /* schedule.py:49                             */
  int32_t cpu_per_task;          /* 12 in the reference (job.py:105)           */
  int32_t mem_per_task;          /* 60 in the reference (job.py:106)           */
  int32_t reserved0;
  double bandwidth;              /* MB/s   (run_sim.py:59)                     */
  double internode_latency;      /* s      (run_sim.py:65)                     */
} gs_cluster;

typedef struct gs_policy {
  int32_t schedule;              /* GS_SCHED_*                                 */
  int32_t scheme;                /* GS_SCHEME_*                                */
  int32_t num_queue;             /* dlas MLFQ depth (<= GS_MAX_QUEUES)          */
  int32_t gittins_n;             /* entries in the gittins tables              */
  double queue_limit[GS_MAX_QUEUES]; /* dlas thresholds (README.md:57-62)      */
  double gittins_delta;          /* service quantum for the index (3250)       */
  const double *gittins_data;    /* host ptr, sorted sample + sentinel         */
  const double *gittins_index;   /* host ptr, index per sample + 0.0           */
} gs_policy;

/* One row of integer aggregates per simulated tick: everything LogInfo
 * (log_manager.py:5-30) carries except the RNG column, as integers so that the
 * host can apply the reference's own float expressions.  64 bytes.            */
typedef struct gs_tick_row {
  int32_t now;            /* 'delta' column (already incremented, schedule.py:193) */
  int32_t idle_nodes;     /* nodes that never hosted a placement (node.py:93-97)  */
  int32_t busy_nodes;
  int32_t busy_gpus;
  int32_t idle_gpus;
  int32_t running;
  int32_t queued;
  int32_t finished;
  int64_t mem_busy_bytes; /* sum over busy devices of min(cap, task memory_max)   */
  int64_t pend_sum;       /* sum of pending ticks over the queue                  */
  int32_t pend_max;       /* 0 when the queue is empty                            */
  int32_t pend_med_lo;    /* the two middle pending values (equal when odd)       */
  int32_t pend_med_hi;
  int32_t reserved;
} gs_tick_row;

/* Compact form of the same information, as the fifo engine writes it.  On a tick where nothing arrives,
 * starts or finishes no LogInfo counter changes except `delta` and the pending times, which move linearly
 * with the tick, so the engine writes one gs_evrow per tick on which a counter DID change (and for the first
 * tick of every gs_run window), beside it one gs_qrow while the queue is non-empty, and one gs_nodeev whenever
 * the number of nodes that ever hosted a job grows (and for the first record of a window); the row of any tick
 * v in [now_k, now_k+1) follows from record k:
 *   delta = v, pending sum = queued*v - arrive_sum, max/median pending = v - oldest / middle arrivals
 * (jobs_manager.py:72-87), busy nodes = the last gs_nodeev at or before v, queue statistics = the gs_qrow with
 * the same `now` (both streams are ordered by `now`).  gs_fetch_rows does this expansion on the device;
 * gs_fetch_compact / gs_fetch_results hand out the records themselves (about a quarter of the bytes of the rows
 * on the BASELINE trace).  24 + 24 + 8 bytes.                                                              */
typedef struct gs_evrow {
  int32_t now;              /* 'delta' of the first row this record describes                          */
  int32_t queued;
  int32_t finished;
  uint16_t busy_gpus;       /* (the engine requires M*G <= 65535)                                      */
  uint16_t running;
  int64_t mem_busy_bytes;
} gs_evrow;

typedef struct gs_qrow {
  int32_t now;              /* the gs_evrow this record belongs to                                     */
  int32_t oldest_arrive;    /* arrival tick of the job that has waited longest                         */
  int32_t med_lo_arrive;    /* arrival ticks of the two middle jobs of the queue (equal when odd)      */
  int32_t med_hi_arrive;
  int64_t arrive_sum;       /* sum of the arrival ticks of the queued jobs                             */
} gs_qrow;

typedef struct gs_nodeev {
  int32_t now;              /* from this row on ...                                                    */
  int32_t busy_nodes;       /* ... this many nodes have hosted a placement (node.py:93-97, never decreases) */
} gs_nodeev;

/* Compact per-job result of the fifo engine: the start tick (-1: the job never started).  fifo never preempts,
 * so the rest of job.csv follows from the trace: run length = max(1, ceil(job.duration)) ticks (quirk Q11; with
 * network costs job.duration is the value gs_fetch_compact returns in duration_out), end = start + run length,
 * jct = run length, preempt = 1 (quirk Q12).  4 bytes per job.                                          */
typedef int32_t gs_job_start;

/* Compact (job, node) record of the fifo engine for clusters with at most 32 GPUs per node (every BASELINE
 * cluster); wider nodes keep the 16-byte gs_span with GS_SPAN_FIRST in ntasks.  8 bytes.                  */
typedef struct gs_cspan {
  uint32_t where;           /* node (bits 0-19) | (ntasks - 1) << 20 | first record of a job << 31           */
  uint32_t devmask;         /* devices of that node held by the job                                          */
} gs_cspan;
#define GS_CSPAN_NODE(w) ((w) & 0xfffffu)
#define GS_CSPAN_NTASKS(w) ((((w) >> 20) & 0x3fu) + 1u)
#define GS_CSPAN_FIRST(w) ((w) >> 31)

/* What the last gs_run window of one replica holds (sizes for gs_fetch_compact).                        */
typedef struct gs_window_info {
  int64_t row_first;        /* tick index of the window's first row                                    */
  int64_t ticks;            /* rows produced so far (the window is [row_first, ticks))                 */
  int64_t ev_rows, q_rows;  /* records of the window                                                   */
  int64_t node_events;      /* gs_nodeev records of the window (at least 1 once a tick has run)         */
  int64_t spans_used;       /* (job, node) records so far, start order                                 */
  int64_t admitted;         /* trace rows consumed so far: gs_job_start is defined for jobs below this  */
  int64_t finished;
  int64_t n;
} gs_window_info;

/* Per-job result, 24 bytes: what LogManager.jcts prints (log_manager.py:143-153). */
typedef struct gs_job_rec {
  int32_t start;          /* start_time; -1 if the job never started             */
  int32_t end;            /* end_time;   -1 if it never finished                 */
  int32_t jct;            /* time_processed() at completion                      */
  int32_t preempt;        /* migration_count (1 for a never-preempted job)       */
  double duration;        /* job.duration after network cost (job.py:196-197)    */
} gs_job_rec;

/* Where a job's tasks ran: one record per (job, node).  16 bytes.               */
#define GS_SPAN_FIRST 0x80000000u  /* gs_fetch_compact only: set in ntasks on the first record of every job */
typedef struct gs_span {
  int32_t node;           /* 0-based node index (reference node_id = node + 1)   */
  int32_t ntasks;
  uint64_t devmask;       /* devices of that node held by the job                */
} gs_span;

typedef struct gs_run_stats {
  int64_t ticks;          /* rows produced so far                                */
  int64_t events;         /* arrivals + starts + completions (+preempt/resume)   */
  int64_t finished;       /* jobs completed                                      */
  int64_t started;
  int64_t placement_evals;/* (job,node) candidate evaluations                    */
  int32_t done;           /* 1 when the loop's exit condition was reached        */
  int32_t status;         /* 0 or a gs_status raised inside the kernel           */
  double kernel_ms;       /* CUDA-event time of the engine kernel(s)             */
  double h2d_ms, d2h_ms;  /* CUDA-event time of the copies of the last calls     */
} gs_run_stats;

/* cluster state record for the stateless placement entry point. 16 bytes.       */
typedef struct gs_node {
  uint64_t busy_mask;     /* bit d set <=> device d has a task                   */
  int32_t cpu_used;
  int32_t mem_used;
} gs_node;

typedef struct gs_jobreq {
  int32_t gpus;           /* Job.gpus                                            */
  int32_t gpu_per_task;   /* Job.gpu_per_worker                                  */
  int64_t mem_bytes;      /* memory_max                                          */
} gs_jobreq;

/* One trace row as the device stores it (32 bytes, one DRAM sector).               */
typedef struct gs_jobin {
  int32_t arrive_tick;    /* first tick with normalized_time <= tick             */
  int32_t gpus;
  int32_t gpu_per_task;
  int32_t ps_count;       /* 0 when the trace has no network columns             */
  int64_t mem_bytes;
  double duration;        /* minutes * scale_factor                              */
} gs_jobin;

typedef struct gs_engine *gs_handle;

int gs_abi_version(void);
const char *gs_build_tag(void);          /* "cuda:sm_100a": the package refuses any other build of these entry points */
const char *gs_last_error(gs_handle h);  /* h may be NULL: last creation error    */

/* A handle simulates `nsims` independent replicas on CUDA device `device`.      */
int gs_create(int device, int nsims, gs_handle *out);
void gs_destroy(gs_handle h);

int gs_config_sim(gs_handle h, int sim, const gs_cluster *cluster, const gs_policy *policy);

/* Trace of one replica, rows in admission order (arrive_tick non-decreasing).
 * model_mb / iterations / ps_count may be NULL (no network term).               */
int gs_load_trace(gs_handle h, int sim, int64_t n,
                  const int32_t *arrive_tick, const int32_t *gpus,
                  const int32_t *gpu_per_task, const double *duration,
                  const int64_t *mem_bytes, const double *model_mb,
                  const double *iterations, const int32_t *ps_count);

/* The same trace already packed as gs_jobin records (no column gather on the host). */
int gs_load_trace_packed(gs_handle h, int sim, int64_t n, const gs_jobin *jobs,
                         const double *model_mb, const double *iterations);

/* Advance every replica by at most max_ticks ticks (<=0: until done).  Record
 * storage on the device is sized by rows_cap per replica at the first call (fifo: that many gs_evrow
 * and gs_qrow records; event-driven policies: that many rows); a launch also ends when it is full,
 * so any value >= 1 is safe -- drain with the fetch calls and call again.                            */
int gs_run(gs_handle h, int64_t max_ticks, int64_t rows_cap);
/* Separate capacity for the gs_qrow stream of replicas prepared afterwards (0 = same as rows_cap).   */
int gs_set_queue_rows_cap(gs_handle h, int64_t qrows_cap);

int gs_stats(gs_handle h, int sim, gs_run_stats *out);

/* Restart every replica at tick 0 on the traces already resident in device
 * memory (sweeps, benchmarking); also clears the accumulated timers.           */
int gs_reset(gs_handle h);

/* Kernel mapping of the event-driven policies: 0 = warp-cooperative kernels (default), 2 = one thread
 * per replica (first version, kept as a cross-check).  The fifo engine has one mapping (a warp per replica). */
int gs_set_engine(gs_handle h, int mode);

/* Span-pool sizing for traces loaded afterwards: 0 (default) = worst case, never overflows;
 * x > 0 = min(worst case, x * n + 4096) records per replica (overflow -> GS_ERR_CAPACITY). */
int gs_set_span_budget(gs_handle h, double spans_per_job);

/* Number of CUDA kernels this handle has launched so far.                       */
int64_t gs_launch_count(gs_handle h);

/* Page-locked host memory for DMA-speed gs_load_trace / gs_fetch_* transfers.  */
int gs_host_alloc(size_t bytes, void **out);
int gs_host_free(void *p);

/* Copy results of one replica to caller-owned host buffers (any may be NULL).  */
int gs_fetch_rows(gs_handle h, int sim, int64_t first, int64_t count, gs_tick_row *rows_out);
int gs_fetch_jobs(gs_handle h, int sim, gs_job_rec *jobs_out /* n */,
                  int32_t *finish_order_out /* n, first `finished` valid */);
int gs_fetch_spans(gs_handle h, int sim, int64_t *span_off_out /* n+1 */,
                   gs_span *spans_out, int64_t spans_cap, int64_t *spans_used);

/* All results of one replica in one call (any output may be NULL); rows [first, first+count)
 * must lie in the last gs_run window.                                           */
int gs_fetch_all(gs_handle h, int sim, int64_t first, int64_t count, gs_tick_row *rows_out,
                 gs_job_rec *jobs_out, int32_t *finish_order_out, int64_t *span_off_out,
                 gs_span *spans_out, int64_t spans_cap, int64_t *spans_used);

/* ---- compact, asynchronous result path (fifo engine) ------------------------------------------------
 * gs_window_info: sizes of what the last gs_run left.  gs_fetch_compact enqueues the copies of every
 * non-NULL output on the handle's stream and returns; gs_sync waits for them.  Buffers from
 * gs_host_alloc make the copies true DMA.  Sizes: ev_out ev_rows, q_out q_rows, nodeev_out node_events, jobs_out n,
 * duration_out n (only with network costs, else left untouched), finish_order_out finished,
 * spans_out spans_used records of gs_result_layout's span_bytes each (8: gs_cspan, 16: gs_span; start order,
 * the first-of-job flag marks job boundaries; jobs in start order = jobs_out sorted by start, start ticks are
 * unique -- one start per tick, schedule.py:188-190).                                                   */
int gs_window(gs_handle h, int sim, gs_window_info *out);
int gs_fetch_compact(gs_handle h, int sim, gs_evrow *ev_out, gs_qrow *q_out, gs_nodeev *nodeev_out, gs_job_start *jobs_out,
                     double *duration_out, int32_t *finish_order_out, void *spans_out /* gs_cspan[] or gs_span[], see gs_result_layout */);
int gs_sync(gs_handle h);
/* The same, for many replicas with ONE copy each way (what bench.py's end-to-end path uses).  The replicas of a
 * handle keep their results in blocks of one layout, side by side on the device:
 *   gs_load_traces_packed  all traces of the handle from one host block, trace i at jobs + i * pitch_bytes
 *                          (n_each[i] records): one strided upload
 *   gs_result_layout       byte offsets of the arrays inside a replica's result block, their capacities, and
 *                          block_bytes, the length of the block
 *   gs_fetch_results       the blocks of replicas [first, first + count) into out + i * out_pitch: one strided
 *                          copy, asynchronous (gs_sync waits); gs_window tells how many entries of each array are valid */
typedef struct gs_result_layout_t {
  int64_t block_bytes;
  int64_t off_ev, off_q, off_nodeev, off_jobs, off_duration /* -1 without network costs */, off_finish_order, off_spans;
  int64_t cap_ev, cap_q, cap_nodeev, cap_spans, n;
  int64_t span_bytes;       /* 8: gs_cspan records (at most 32 GPUs per node), 16: gs_span records               */
} gs_result_layout_t;
int gs_load_traces_packed(gs_handle h, const gs_jobin *jobs, size_t pitch_bytes, const int64_t *n_each);
int gs_result_layout(gs_handle h, int sim, gs_result_layout_t *out);
int gs_fetch_results(gs_handle h, int first, int count, void *out, size_t out_pitch);
/* 1: gs_load_trace_packed from a page-locked buffer enqueues the upload without staging and returns
 * before it completes -- the caller keeps the buffer unchanged until the next gs_run / gs_sync on
 * this handle returns.  0 (default): every call copies and completes before returning.              */
int gs_set_async(gs_handle h, int on);

/* ---- one simulation on several GPUs of one box (BASELINE config C4) ----------------------
 * The north-star sketches "the trace sharded across the 8 GPUs with one allreduce per tick".  What is worth
 * sharding in this path is the per-event work that is O(runnable jobs): for the gittins policy that is the index
 * evaluation (run_sim.py:1040-1078 -> get_gittins_index :949-954, a table search per runnable job per event).
 * Every rank holds the (small) cluster / queue state and the trace; rank r evaluates the chunks c of the
 * runnable list with c mod nranks == r and STORES the results straight into every peer's receive buffer over
 * NVLink (peer memory opened with CUDA IPC), then publishes an event counter in the peers' flag words and waits
 * for theirs -- one exchange per event, inside the persistent kernel, no host round trip and no NCCL call on the
 * data path (torch.distributed only carries the 64-byte IPC handles at start-up).  Order, admission and the
 * statistics are then computed redundantly and identically on every rank: results are bit-identical to one GPU.
 *
 *   gs_comm_prepare  allocates this handle's exchange buffer (for traces of up to max_jobs jobs) and returns its IPC handle
 *   gs_comm_init     receives the handles of ALL ranks (own at [rank]), opens the peers', arms the sharded mode for
 *                    gittins replicas prepared afterwards; the handle must hold exactly one replica
 *   gs_comm_stats    exchanges done and their mean cost in microseconds (publish -> all peers seen, incl. skew)
 * A peer that does not answer within ~5 s ends the run with GS_ERR_COMM on every waiting rank.            */
typedef struct gs_comm_handle { unsigned char bytes[64]; } gs_comm_handle;
int gs_comm_prepare(gs_handle h, int64_t max_jobs, gs_comm_handle *out);
int gs_comm_init(gs_handle h, int rank, int nranks, const gs_comm_handle *all_ranks);
int gs_comm_stats(gs_handle h, int64_t *exchanges, double *mean_us);
/* Events with at most k runnable jobs are evaluated by every rank itself -- the replicated state guarantees identical
 * results without sending anything -- and only longer lists are split and exchanged.  k = 0: always exchange.
 * DEFAULT: no list is long enough (k = INT32_MAX).  An exchange costs 2-4 us, and since the index of a job is one table
 * load (the direct table of gs_config_sim) no list we measured -- up to ~7000 runnable jobs -- is evaluated slower by
 * one GPU than an exchange takes (DESIGN section 7 has the numbers); the call is how a caller opts in. */
int gs_comm_set_min_runnable(gs_handle h, int k);

/* Stateless candidate scoring: evaluate b jobs against ONE cluster state.
 * first_node[i] = node of a single-node first fit, or the first node of a
 * cross-node fill, or -1 if the job cannot be placed; task_node (optional)
 * receives the node of each task at task_off[i] .. task_off[i+1].               */
int gs_place_batch(gs_handle h, const gs_cluster *cluster, const gs_node *nodes, int32_t m,
                   const gs_jobreq *jobs, int64_t b, int32_t *first_node,
                   int32_t *nodes_used, const int64_t *task_off, int32_t *task_node,
                   double *kernel_ms);

/* ---- legacy switch-local yarn placement with parameter-server traffic accounting --------
 * _Cluster.ms_yarn_placement (infra/cluster.py:888-898) over _Switch.try_cross_node_alloc / try_single_node_alloc
 * (infra/switch.py:38-167,190-206): all gpus of a job come from ONE switch; a job wider than a node takes
 * floor(g/G) completely idle nodes plus one node for the remainder, is charged 6 cpus per gpu and
 * (ps_mem + g*p_w_mem + worker_mem) memory per gpu, and every node's network load grows by
 *   round(model*k, 1), then per PS shard on the node:  += ps*(g-k);  -= ps*k;  round(., 1)     (switch.py:98-108)
 * (Python's round, reproduced exactly).  `ncl` independent clusters per call; the jobs of a cluster are placed in
 * order and change its node table (in/out).  spans: caller-sized, job j writes at span_off (floor(g/G)+1 slots);
 * network is NaN on the single-node path (the reference records none there).                              */
typedef struct gs_switch_cluster {
  int32_t num_switch, num_node_p_switch, num_gpu_p_node, reserved;
  int64_t node_off;        /* first record of this cluster in `nodes` (num_switch * num_node_p_switch records)   */
  int64_t job_off, job_cnt;
} gs_switch_cluster;
typedef struct gs_switch_node { int32_t free_gpus, free_cpus; double free_mem; double net_in; } gs_switch_node;
typedef struct gs_switch_job {
  int32_t num_gpu, n_ps;   /* job['num_gpu'], len(job['ps_network'])                                             */
  int64_t ps_off;          /* first entry of the job's ps_network in `ps_network`                                */
  int64_t span_off;        /* where the job's per-node records go in `spans`                                     */
  double model_size;       /* job['model']['total_size']                                                         */
} gs_switch_job;
typedef struct gs_switch_ans { int32_t n_nodes; int32_t sw; } gs_switch_ans;   /* n_nodes == 0: not placed      */
typedef struct gs_switch_span { int32_t node, num_gpu, num_cpu, reserved; double mem, network; } gs_switch_span;
int gs_switch_yarn(gs_handle h, int32_t ncl, const gs_switch_cluster *clusters, gs_switch_node *nodes, int64_t n_nodes,
                   const gs_switch_job *jobs, int64_t n_jobs, const double *ps_network, int64_t n_ps,
                   double worker_mem, double ps_mem, double p_w_mem,
                   gs_switch_ans *ans, gs_switch_span *spans, int64_t n_spans);

/* ---- host side of the log writer (no device work): the sampled avg_gpu_utilization column of cluster.csv.
 * The reference draws one np.random.normal per busy device and tick, nodes in id order, devices 0..G-1
 * (infra/device.py:48-54, core/scheduling/schedule.py:103-120), from numpy's sequential global stream; the caller draws the
 * standard-normal values from numpy, these entry points consume them in that order.  A holding = one (job, device)
 * pair counted on rows first..last; holdings are passed sorted by `first`, key = node * G + device.                */
typedef struct gs_logcol_s *gs_logcol;
gs_logcol gs_logcol_open(int64_t n_rows, int32_t width, int64_t n_hold, const int64_t *first, const int64_t *last,
                         const int32_t *key, const int32_t *job);
void gs_logcol_close(gs_logcol c);
int gs_logcol_counts(gs_logcol c, int64_t *counts);                  /* values consumed by each row                  */
int gs_logcol_rows(gs_logcol c, int64_t r_end, const double *loc, const double *scale, const double *z, int64_t n_z,
                   double *acc, int32_t *unclipped);                 /* rows [current, r_end): sums in draw order     */

/* PS<->worker transfer time for a batch of placed jobs.  task_node holds the
 * node of every task (segments given by task_off), is_ps marks PS tasks.        */
int gs_net_cost(gs_handle h, const gs_cluster *cluster, int64_t b,
                const int64_t *task_off, const int32_t *task_node, const uint8_t *is_ps,
                const int32_t *ps_count, const double *model_mb, const double *iterations,
                double *extra_out);

#ifdef __cplusplus
}
#endif
#endif /* GSCHED_H_ */
