/*
 * tight_cpu.c -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A tight single-thread CPU implementation of the SAME algorithm the CUDA tick kernel runs
 * (O(1) incremental counters, LIFO queue, timing wheel keyed by finish tick, (idle, free slot)
 * node table, no per-tick re-scan) -- i.e. what a careful C programmer would write instead of
 * transliterating the reference's Python.  It exists so that the GPU numbers can be read next to
 * a strong CPU competitor and not only next to the literal restatement (oracle/gsched_oracle.c,
 * which keeps the reference's per-tick re-scans on purpose).  SURVEY 6 / 8(d) ask for exactly
 * this comparison.  It is validated bit-for-bit against the pinned oracle in
 * tests/test_tight_cpu.py and is reported by bench.py as `cpu_tight`.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gsched.h"

typedef struct { int32_t next, node0, ntasks0, span_cnt; uint64_t mask0; int64_t span_first; } tstate_t;

static inline uint64_t take_lowest(uint64_t idle, int cnt) {
  uint64_t m = idle;
  for (int i = 0; i < cnt; ++i) m &= m - 1;
  return idle ^ m;
}

int64_t tight_run_fifo(const gs_cluster *c, int64_t n, const int32_t *arrive, const int32_t *gpus,
                       const int32_t *gpc, const double *duration, const int64_t *mem_bytes,
                       gs_tick_row *rows_out, int64_t rows_cap, gs_job_rec *jobs_out,
                       int32_t *finish_order_out, int64_t *n_finished_out, gs_span *spans_out,
                       int64_t spans_cap, int64_t *span_first_out, int32_t *span_cnt_out, int64_t *events_out) {
  const int M = c->num_switch * c->num_node_p_switch, G = c->num_gpu_p_node;
  const int kc = c->num_cpu_p_node / c->cpu_per_task, km = c->mem_p_node / c->mem_per_task;
  const int K = kc < km ? kc : km;
  const int64_t cap_bytes = (int64_t)c->gpu_mem_cap_mib << 20, fit_limit = cap_bytes - ((int64_t)500 << 20);
  const uint64_t gmask = G >= 64 ? ~0ull : ((1ull << G) - 1ull);
  uint64_t *busy = (uint64_t *)calloc((size_t)M, sizeof(uint64_t));
  int32_t *idle = (int32_t *)malloc(sizeof(int32_t) * (size_t)M), *kfree = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
  uint8_t *everf = (uint8_t *)calloc((size_t)M, 1);
  for (int i = 0; i < M; ++i) { idle[i] = G; kfree[i] = K; }
  size_t N = (size_t)(n > 0 ? n : 1);
  double maxd = 1.0;
  for (int64_t j = 0; j < n; ++j) if (duration[j] > maxd) maxd = duration[j];
  int W = 64; while (W < (int)maxd + 3) W <<= 1;
  const int wmask = W - 1;
  int32_t *wh = (int32_t *)malloc(sizeof(int32_t) * (size_t)W), *wt = (int32_t *)malloc(sizeof(int32_t) * (size_t)W);
  for (int i = 0; i < W; ++i) { wh[i] = -1; wt[i] = -1; }
  tstate_t *st = (tstate_t *)malloc(sizeof(tstate_t) * N);
  int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * N);
  for (int64_t j = 0; j < n; ++j) {
    jobs_out[j].start = -1; jobs_out[j].end = -1; jobs_out[j].jct = 0; jobs_out[j].preempt = 0; jobs_out[j].duration = duration[j];
    if (span_cnt_out) { span_cnt_out[j] = 0; span_first_out[j] = 0; }
  }
  int64_t p = 0, top = 0, running = 0, finished = 0, span_used = 0, ticks = 0, started = 0;
  int64_t mem_busy = 0, sum_arr = 0;
  int ever = 0, busy_gpus = 0, delta = 0, lo = 0;
  int64_t rc = 0;
  while ((n - p) + running > 0 || ticks == 0) {
    if (n == 0) break;
    /* A: arrivals */
    if (p < n && arrive[p] <= delta) {
      int64_t b = p;
      while (b < n && arrive[b] <= delta) ++b;
      for (int64_t i = b - 1; i >= p; --i) stack[top++] = (int32_t)i;
      sum_arr += (b - p) * (int64_t)delta;
      p = b;
    }
    /* B: one attempt on the head */
    if (top > 0) {
      const int32_t j = stack[top - 1];
      const int hg = gpus[j], hc = gpc[j], tasks = hg / hc;
      const int placeable = mem_bytes[j] < fit_limit;
      int ok = 0, nspans = 0, first_node = -1;
      uint64_t mask0 = 0;
      const int64_t sf = span_used;
      if (hg <= G) {
        for (int nd = lo; nd < M; ++nd) {
          if (idle[nd] >= hg && kfree[nd] >= tasks) {
            if (!placeable) { kfree[nd] -= tasks; continue; }
            uint64_t tk = take_lowest(~busy[nd] & gmask, hg);
            busy[nd] |= tk; idle[nd] -= hg; kfree[nd] -= tasks;
            if (!everf[nd]) { everf[nd] = 1; ++ever; }
            if (span_used + 1 > spans_cap) { rc = GS_ERR_CAPACITY; goto out; }
            spans_out[sf].node = nd; spans_out[sf].ntasks = tasks; spans_out[sf].devmask = tk;
            ok = 1; nspans = 1; first_node = nd; mask0 = tk;
            break;
          }
        }
      } else {
        int cum = 0, last = -1;
        for (int nd = lo; nd < M; ++nd) {
          int cp = idle[nd] / hc; if (kfree[nd] < cp) cp = kfree[nd];
          if (cp <= 0) continue;
          if (!placeable) { kfree[nd] -= 1; continue; }
          cum += cp;
          if (cum >= tasks) { last = nd; break; }
        }
        if (last >= 0) {
          int rem = tasks;
          for (int nd = lo; nd <= last; ++nd) {
            int cp = idle[nd] / hc; if (kfree[nd] < cp) cp = kfree[nd];
            if (cp <= 0) continue;
            int take = cp < rem ? cp : rem;
            uint64_t tk = take_lowest(~busy[nd] & gmask, take * hc);
            busy[nd] |= tk; idle[nd] -= take * hc; kfree[nd] -= take;
            if (!everf[nd]) { everf[nd] = 1; ++ever; }
            if (span_used + nspans + 1 > spans_cap) { rc = GS_ERR_CAPACITY; goto out; }
            spans_out[sf + nspans].node = nd; spans_out[sf + nspans].ntasks = take; spans_out[sf + nspans].devmask = tk;
            if (nspans == 0) { first_node = nd; mask0 = tk; }
            ++nspans; rem -= take;
          }
          ok = 1;
        }
      }
      if (ok) {
        while (lo < M && idle[lo] == 0) ++lo;
        double cl = ceil(duration[j]);
        int need = cl < 1.0 ? 1 : (int)cl;
        int endt = delta + need;
        span_used += nspans;
        jobs_out[j].start = delta; jobs_out[j].end = endt; jobs_out[j].jct = need; jobs_out[j].preempt = 1;
        if (span_cnt_out) { span_first_out[j] = sf; span_cnt_out[j] = nspans; }
        st[j].next = -1; st[j].node0 = first_node; st[j].mask0 = mask0; st[j].ntasks0 = tasks; st[j].span_cnt = nspans; st[j].span_first = sf;
        int sl = endt & wmask, tl = wt[sl];
        if (tl < 0) wh[sl] = j; else st[tl].next = j;
        wt[sl] = j;
        --top; sum_arr -= arrive[j]; ++running; ++started;
        busy_gpus += hg;
        mem_busy += (int64_t)hg * (mem_bytes[j] < cap_bytes ? mem_bytes[j] : cap_bytes);
      }
    }
    /* D/E: completions */
    const int now = delta + 1;
    {
      int sl = now & wmask, h = wh[sl];
      if (h >= 0) {
        wh[sl] = -1; wt[sl] = -1;
        while (h >= 0) {
          const tstate_t *s = &st[h];
          if (s->span_cnt == 1) {
            int nd = s->node0;
            busy[nd] &= ~s->mask0; idle[nd] += gpus[h]; kfree[nd] += s->ntasks0;
            if (nd < lo) lo = nd;
          } else {
            for (int i = 0; i < s->span_cnt; ++i) {
              const gs_span *sp = &spans_out[s->span_first + i];
              busy[sp->node] &= ~sp->devmask; idle[sp->node] += sp->ntasks * gpc[h]; kfree[sp->node] += sp->ntasks;
              if (sp->node < lo) lo = sp->node;
            }
          }
          finish_order_out[finished++] = h; --running;
          busy_gpus -= gpus[h];
          mem_busy -= (int64_t)gpus[h] * (mem_bytes[h] < cap_bytes ? mem_bytes[h] : cap_bytes);
          h = s->next;
        }
      }
    }
    /* H: row */
    if (ticks >= rows_cap) { rc = GS_ERR_CAPACITY; goto out; }
    {
      gs_tick_row *r = &rows_out[ticks];
      r->now = now; r->idle_nodes = M - ever; r->busy_nodes = ever; r->busy_gpus = busy_gpus;
      r->idle_gpus = M * G - busy_gpus; r->running = (int32_t)running; r->queued = (int32_t)top; r->finished = (int32_t)finished;
      r->mem_busy_bytes = mem_busy; r->pend_sum = top * (int64_t)now - sum_arr;
      r->pend_max = 0; r->pend_med_lo = 0; r->pend_med_hi = 0; r->reserved = 0;
      if (top > 0) {
        r->pend_max = now - arrive[stack[0]];
        r->pend_med_lo = now - arrive[stack[top - 1 - (top - 1) / 2]];
        r->pend_med_hi = now - arrive[stack[top - 1 - top / 2]];
      }
    }
    ++ticks; delta = now;
    if ((n - p) + running == 0) break;
  }
out:
  *n_finished_out = finished;
  if (events_out) *events_out = p + started + finished;
  free(busy); free(idle); free(kfree); free(everf); free(wh); free(wt); free(st); free(stack);
  return rc < 0 ? rc : ticks;
}
