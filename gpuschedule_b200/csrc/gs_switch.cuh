// gs_switch.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// Legacy switch-local yarn placement with parameter-server traffic accounting (second half):
//   _Cluster.ms_yarn_placement   infra/cluster.py:888-898    switches in order, the first that takes the whole job
//   _Switch.ms_yarn_alloc_res    infra/switch.py:190-206     num_gpu > gpus per node -> cross-node, else single node
//   _Switch.try_cross_node_alloc infra/switch.py:38-139      floor(g/G) completely idle nodes + one node for the rest;
//                                                            per node: cpus 6/gpu, memory (ps_mem + g*p_w_mem + worker_mem)/gpu,
//                                                            network = round(model*k, 1) then per PS shard
//                                                            += ps*(g-k); -= ps*k; round(., 1)          (:98-108,122-133)
//   _Switch.try_single_node_alloc infra/switch.py:142-167    first node with the gpus, 2 or 6 cpus per gpu, worker_mem
// One warp per cluster; jobs are placed in order (each placement changes the node table); lanes stripe over the nodes
// of a switch: the k first qualifying nodes are a ballot prefix, every chosen node's traffic chain runs in its own lane.
#pragma once

// Python's round(x, 1): nearest multiple of 0.1 to the exact binary value, ties to even, then the nearest double
// (CPython float.__round__).  x*10 = p + e exactly (e from one fma), so the half-way comparison is exact.
__device__ __forceinline__ double py_round1(double x) {
  const double ax = fabs(x);
  if (!(ax < 1.0e12)) return x;
  const double p = __dmul_rn(ax, 10.0), e = __fma_rn(ax, 10.0, -p);
  double q = floor(p);
  const double d = __dsub_rn(__dsub_rn(p, q), 0.5);
  if (d > 0.0 || (d == 0.0 && (e > 0.0 || (e == 0.0 && fmod(q, 2.0) != 0.0)))) q += 1.0;
  return copysign(__ddiv_rn(q, 10.0), x);
}

__device__ __forceinline__ double ps_traffic(double model_size, int k, int need_gpu, const double *ps, int n_ps, int idx0) {
  double traffic = py_round1(__dmul_rn(model_size, (double)k));
  for (int i = 0; i < k; ++i) {
    const double v = (idx0 + i) < n_ps ? ps[idx0 + i] : 0.0;
    traffic = __dadd_rn(traffic, __dmul_rn(v, (double)(need_gpu - k)));
    traffic = __dsub_rn(traffic, __dmul_rn(v, (double)k));
    traffic = py_round1(traffic);
  }
  return traffic;
}

__global__ void __launch_bounds__(32) gs_switch_yarn_kernel(int ncl, const gs_switch_cluster *__restrict__ cls, gs_switch_node *nodes,
                                                           const gs_switch_job *__restrict__ jobs, const double *__restrict__ ps_all,
                                                           double worker_mem, double ps_mem_c, double p_w_mem,
                                                           gs_switch_ans *ans, gs_switch_span *spans) {
  const int c = blockIdx.x, lane = threadIdx.x;
  if (c >= ncl) return;
  const gs_switch_cluster cl = cls[c];
  const int S = cl.num_switch, P = cl.num_node_p_switch, G = cl.num_gpu_p_node;
  gs_switch_node *tab = nodes + cl.node_off;
  const unsigned lt = (1u << lane) - 1u;
  for (long long jj = 0; jj < cl.job_cnt; ++jj) {
    const gs_switch_job jb = jobs[cl.job_off + jj];
    const int g = jb.num_gpu, n_ps = jb.n_ps;
    const double *ps = ps_all + jb.ps_off;
    gs_switch_span *out = spans + jb.span_off;
    int placed = 0, sw = -1;
    for (int s = 0; s < S && placed == 0; ++s) {
      gs_switch_node *nd = tab + (long long)s * P;
      if (g > G) {
        const int num_full = g / G, last_gpu = g % G, last_cpu = last_gpu * 6, idle_cpu = G * 6;
        const double ps_w_mem = __dadd_rn(__dadd_rn(ps_mem_c, __dmul_rn((double)g, p_w_mem)), worker_mem);
        const double full_mem = __dmul_rn(ps_w_mem, (double)G), last_mem = __dmul_rn(ps_w_mem, (double)last_gpu);
        // pass 1: are there num_full completely idle nodes, and a node for the rest that is not one of them?
        int nfull = 0, last = -1;
        for (int base = 0; base < P; base += 32) {
          const int i = base + lane;
          bool full = false, rest = false;
          if (i < P) {
            const gs_switch_node v = nd[i];
            full = v.free_gpus == G && v.free_cpus >= idle_cpu && v.free_mem >= full_mem;
            rest = last_gpu != 0 && v.free_gpus >= last_gpu && v.free_cpus >= last_cpu && v.free_mem >= last_mem;
          }
          const unsigned fb = __ballot_sync(FULL, full);
          const int take = min(__popc(fb), num_full - nfull);                 // the first `take` full nodes of this chunk are used
          const bool chosen = full && __popc(fb & lt) < take;
          nfull += take;
          const unsigned rb = __ballot_sync(FULL, rest && !chosen);
          if (last < 0 && rb) last = base + __ffs(rb) - 1;
        }
        if (nfull < num_full || (last_gpu != 0 && last < 0)) continue;          // this switch cannot take it (:66-67,76-77)
        // pass 2: commit, same walk
        int k = 0;
        for (int base = 0; base < P && k < num_full; base += 32) {
          const int i = base + lane;
          bool full = false;
          gs_switch_node v; v.free_gpus = 0; v.free_cpus = 0; v.free_mem = 0.0; v.net_in = 0.0;
          if (i < P) { v = nd[i]; full = v.free_gpus == G && v.free_cpus >= idle_cpu && v.free_mem >= full_mem; }
          const unsigned fb = __ballot_sync(FULL, full);
          const int take = min(__popc(fb), num_full - k);
          const int rank = __popc(fb & lt);
          if (full && rank < take) {
            const int slot = k + rank;
            const double traffic = ps_traffic(jb.model_size, G, g, ps, n_ps, slot * G);
            v.free_gpus -= G; v.free_cpus -= idle_cpu; v.free_mem = __dsub_rn(v.free_mem, full_mem); v.net_in = __dadd_rn(v.net_in, traffic);
            nd[i] = v;
            gs_switch_span sp; sp.node = i; sp.num_gpu = G; sp.num_cpu = idle_cpu; sp.reserved = 0; sp.mem = full_mem; sp.network = traffic;
            out[slot] = sp;
          }
          k += take;
        }
        __syncwarp();
        if (last_gpu != 0 && lane == 0) {
          gs_switch_node v = nd[last];
          const double traffic = ps_traffic(jb.model_size, last_gpu, g, ps, n_ps, num_full * G);
          v.free_gpus -= last_gpu; v.free_cpus -= last_cpu; v.free_mem = __dsub_rn(v.free_mem, last_mem); v.net_in = __dadd_rn(v.net_in, traffic);
          nd[last] = v;
          gs_switch_span sp; sp.node = last; sp.num_gpu = last_gpu; sp.num_cpu = last_cpu; sp.reserved = 0; sp.mem = last_mem; sp.network = traffic;
          out[num_full] = sp;
        }
        placed = num_full + (last_gpu != 0 ? 1 : 0); sw = s;
      } else {
        const int need_cpu = (n_ps == 0 && g == 1) ? g * 2 : g * 6;
        int found = -1;
        for (int base = 0; base < P && found < 0; base += 32) {
          const int i = base + lane;
          bool fit = false;
          if (i < P) { const gs_switch_node v = nd[i]; fit = v.free_gpus >= g && v.free_cpus >= need_cpu && v.free_mem >= worker_mem; }
          const unsigned b = __ballot_sync(FULL, fit);
          if (b) found = base + __ffs(b) - 1;
        }
        if (found < 0) continue;
        if (lane == 0) {
          gs_switch_node v = nd[found];
          v.free_gpus -= g; v.free_cpus -= need_cpu; v.free_mem = __dsub_rn(v.free_mem, worker_mem);
          nd[found] = v;
          gs_switch_span sp; sp.node = found; sp.num_gpu = g; sp.num_cpu = need_cpu; sp.reserved = 0; sp.mem = worker_mem;
          sp.network = __longlong_as_double(0x7ff8000000000000LL);        // the single-node path records no traffic (switch.py:161-162)
          out[0] = sp;
        }
        placed = 1; sw = s;
      }
      __syncwarp();
    }
    if (lane == 0) { gs_switch_ans a; a.n_nodes = placed; a.sw = sw; ans[cl.job_off + jj] = a; }
    __syncwarp();
  }
}
