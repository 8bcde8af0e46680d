// gs_aux.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// Replica reset (event-driven policies), stateless placement scoring, network-cost kernels.
#pragma once

// One launch (re)initialises every event-driven replica flagged need_init: job records (never-started
// jobs report start=end=-1, jct=preempt=0 and their input duration) and the per-job policy state.
// (The fifo engine resets its own small tables at the start of gs_tick2_kernel.)
__global__ void gs_init_kernel(SimDev *sims, int nsims) {
  const int sim = blockIdx.y;
  if (sim >= nsims) return;
  const SimDev &S = sims[sim];
  if (!S.need_init || S.policy == GS_SCHED_FIFO) return;
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < S.n; i += stride) {
    gs_job_rec r; r.start = -1; r.end = -1; r.jct = 0; r.preempt = 0; r.duration = S.jobs[i].dur;
    S.rec[i] = r;
    PJob z; memset(&z, 0, sizeof(z)); z.start = -1; S.pj[i] = z;
  }
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

