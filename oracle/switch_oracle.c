/*
 * switch_oracle.c -- TEST INFRASTRUCTURE (CPU checker), NOT PRODUCT CODE.
 *
 * Restatement of the reference's legacy switch-local yarn placement with parameter-server traffic
 * accounting (second half; all paths relative to the reference root):
 *   _Cluster.ms_yarn_placement        infra/cluster.py:888-898   switches in order, first that takes the job
 *   _Switch.ms_yarn_alloc_res         infra/switch.py:190-206    num_gpu > gpus/node -> cross, else single
 *   _Switch.try_cross_node_alloc      infra/switch.py:38-139
 *   _Switch.try_single_node_alloc     infra/switch.py:142-167
 * PINNED: tests/golden/switch_yarn.json holds the answers of those four methods executed verbatim (taken out of the
 * reference with `ast`, tests/golden/make_switch_golden.py) under the documented stubs for what the repository
 * never defines (_Node, the job_queue constants 5 / 8 / 0.2 of core/models.py:24-26, the placement recorders);
 * tests/test_switch_oracle.py compares this file with them value for value (doubles bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/gsched.h"

/* Python's round(x, 1) (float.__round__ -> correctly rounded decimal, ties to even, Objects/floatobject.c double_round):
 * nearest multiple of 0.1 to the EXACT binary value, then the nearest double.  x*10 is split into the rounded product p
 * and its exact error e (fma), so the comparison with the half-way point is exact. */
double switch_round1(double x) {
  if (!(fabs(x) < 1.0e12)) return x;      /* traffic figures are MB; beyond 2^52/10 the half-way test below would not be exact */
  const double ax = fabs(x);
  const double p = ax * 10.0, e = fma(ax, 10.0, -p);
  double q = floor(p);
  const double d = (p - q) - 0.5;        /* exact */
  /* |d| is 0 or at least one ulp(p) > |e|, so the sign of d decides unless d == 0; then e does; an exact tie goes to even */
  if (d > 0.0 || (d == 0.0 && (e > 0.0 || (e == 0.0 && fmod(q, 2.0) != 0.0)))) q += 1.0;
  return copysign(q / 10.0, x);
}

/* One job against the cluster (S switches x P nodes, G gpus per node).  Node tables are updated on success.
 * Outputs (capacity P + 1 entries): node index within the switch, gpus, cpus, memory, network (NaN: the single-node path
 * records no traffic, switch.py:161-162).  Returns the number of nodes used (0: not placed), *switch_out the switch. */
int switch_yarn_place(int S, int P, int G, int32_t *free_gpus, int32_t *free_cpus, double *free_mem, double *net_in,
                      int need_gpu, double model_size, const double *ps_network, int n_ps,
                      double worker_mem, double ps_mem_c, double p_w_mem,
                      int32_t *switch_out, int32_t *node_out, int32_t *gpu_out, int32_t *cpu_out, double *mem_out, double *net_out) {
  for (int s = 0; s < S; ++s) {                                            /* cluster.py:892-897 */
    int32_t *fg = free_gpus + (size_t)s * P, *fc = free_cpus + (size_t)s * P;
    double *fm = free_mem + (size_t)s * P, *ni = net_in + (size_t)s * P;
    if (need_gpu > G) {                                                    /* switch.py:201-202 */
      const int num_full = need_gpu / G, last_gpu = need_gpu % G;          /* :47-48 */
      const int last_cpu = last_gpu * 6, idle_cpu = G * 6;                 /* :49,51 */
      const double ps_mem = ps_mem_c + need_gpu * p_w_mem;                 /* :55 */
      const double ps_w_mem = ps_mem + worker_mem;                         /* :56 */
      int nfull = 0, last = -1;
      for (int nd = 0; nd < P && nfull < num_full; ++nd)                   /* :59-65 */
        if (fg[nd] == G && fc[nd] >= idle_cpu && fm[nd] >= ps_w_mem * G) node_out[nfull++] = nd;
      if (nfull < num_full) continue;                                      /* :66-67 -> next switch */
      if (last_gpu != 0) {                                                 /* :69-77 */
        for (int nd = 0; nd < P; ++nd) {
          int in_full = 0;
          for (int k = 0; k < nfull; ++k) in_full |= node_out[k] == nd;
          if (in_full) continue;
          if (fg[nd] >= last_gpu && fc[nd] >= last_cpu && fm[nd] >= ps_w_mem * last_gpu) { last = nd; break; }
        }
        if (last < 0) continue;
      }
      int idx = 0, used = 0;
      for (int k = 0; k < nfull; ++k) {                                    /* :83-113 */
        const int nd = node_out[k];
        fg[nd] -= G; fc[nd] -= idle_cpu;                                   /* alloc_job_res */
        fm[nd] -= ps_w_mem * G;
        double traffic = switch_round1(model_size * G);                    /* :99 */
        for (int i = 0; i < G; ++i) {                                      /* :101-107 */
          const double ps = idx < n_ps ? ps_network[idx] : 0.0;
          traffic += ps * (need_gpu - G);
          traffic -= ps * G;
          traffic = switch_round1(traffic);
          ++idx;
        }
        ni[nd] += traffic;
        gpu_out[used] = G; cpu_out[used] = idle_cpu; mem_out[used] = ps_w_mem * G; net_out[used] = traffic;
        ++used;
      }
      if (last_gpu != 0) {                                                 /* :115-138 */
        fg[last] -= last_gpu; fc[last] -= last_cpu;
        fm[last] -= ps_w_mem * last_gpu;
        double traffic = switch_round1(model_size * last_gpu);
        for (int i = 0; i < last_gpu; ++i) {
          const double ps = idx < n_ps ? ps_network[idx] : 0.0;
          traffic += ps * (need_gpu - last_gpu);
          traffic -= ps * last_gpu;
          traffic = switch_round1(traffic);
          ++idx;
        }
        ni[last] += traffic;
        node_out[used] = last; gpu_out[used] = last_gpu; cpu_out[used] = last_cpu; mem_out[used] = ps_w_mem * last_gpu; net_out[used] = traffic;
        ++used;
      }
      *switch_out = s;
      return used;
    }
    /* try_single_node_alloc, switch.py:142-167 */
    const int need_cpu = (n_ps == 0 && need_gpu == 1) ? need_gpu * 2 : need_gpu * 6;   /* :149-152 */
    for (int nd = 0; nd < P; ++nd) {
      if (fg[nd] >= need_gpu && fc[nd] >= need_cpu && fm[nd] >= worker_mem) {          /* :155 */
        fg[nd] -= need_gpu; fc[nd] -= need_cpu;
        fm[nd] = fm[nd] - worker_mem;                                                  /* :159 */
        node_out[0] = nd; gpu_out[0] = need_gpu; cpu_out[0] = need_cpu; mem_out[0] = worker_mem; net_out[0] = NAN;
        *switch_out = s;
        return 1;
      }
    }
  }
  return 0;
}
