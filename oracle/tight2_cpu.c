/*
 * tight2_cpu.c -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The event-stepped algorithm of the CUDA fifo engine (gpuschedule_b200/csrc/gs_tick2.cuh) as tight
 * single-thread C, statement for statement where a warp is not needed: ticks on which nothing arrives,
 * starts or finishes are jumped over, the queue head waits in registers, the timing wheel is pushed at the
 * front and released newest-first (finish order repaired afterwards), a head that did not fit is not tried
 * again until something could change the answer, and the output is the compact record stream of
 * include/gsched.h (gs_evrow / gs_qrow / start ticks / start-ordered spans) in resumable windows.
 *
 * Two uses: (1) tests/test_tight2_cpu.py checks it -- through the package's own record decoders --
 * against the pinned oracle (oracle/gsched_oracle.c) on the reference fixtures and on random cases, which
 * validates the algorithm and the decoders on a box without a GPU; (2) bench.py times it as `cpu_tight`,
 * the strongest CPU competitor of the GPU engine we could write (it is NOT the reference's algorithm).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gsched.h"

#define T2_INF 0x7fffffff
typedef struct { int32_t start, run_ticks; } t2_job_run;   /* internal; the engine's compact result is the start tick alone */

typedef struct { int32_t next, where; uint64_t mask0; } t2_state;

typedef struct tight2 {
  int M, G, K, W, wmask;
  int64_t cap_bytes, fit_limit, n;
  uint64_t gmask;
  const int32_t *arrive, *gpus, *gpc;
  const double *duration;
  const int64_t *mem_bytes;
  uint64_t *busy; int32_t *idle, *kfree; uint8_t *everf;
  int32_t *whead; int64_t *wmem;
  t2_state *st;
  int32_t *stack_job, *stack_arr;
  t2_job_run *rec2;
  int32_t *fin;
  gs_span *spans; int64_t span_cap;
  /* loop state */
  int delta, top, scount, running, finished, ever, busy_gpus, blocked, done, status;
  int64_t p, span_used, mem_busy, sum_arr, evals;
  int bottom_arr;
  int hvalid, hjob, harr;
  int next_fin, nf_head;
  int64_t row_first;
} tight2;

static inline uint64_t take_lowest(uint64_t idle, int cnt) {
  uint64_t m = idle;
  for (int i = 0; i < cnt; ++i) m &= m - 1;
  return idle ^ m;
}

static inline int need_of(double dur) {
  const double cl = ceil(dur);
  return cl < 1.0 ? 1 : (cl > 1.0e9 ? T2_INF : (int)cl);
}

void tight2_free(tight2 *t) {
  if (!t) return;
  free(t->busy); free(t->idle); free(t->kfree); free(t->everf); free(t->whead); free(t->wmem); free(t->st);
  free(t->stack_job); free(t->stack_arr); free(t->rec2); free(t->fin); free(t->spans);
  free(t);
}

static void tight2_reset(tight2 *t) {
  for (int i = 0; i < t->M; ++i) { t->busy[i] = 0; t->idle[i] = t->G; t->kfree[i] = t->K; t->everf[i] = 0; }
  for (int i = 0; i < t->W; ++i) { t->whead[i] = -1; t->wmem[i] = 0; }
  t->delta = t->top = t->scount = t->running = t->finished = t->ever = t->busy_gpus = t->blocked = t->status = 0;
  t->done = t->n == 0;
  t->p = t->span_used = t->mem_busy = t->sum_arr = t->evals = 0;
  t->bottom_arr = 0; t->hvalid = 0; t->hjob = -1; t->harr = 0;
  t->next_fin = T2_INF; t->nf_head = -1; t->row_first = 0;
}

tight2 *tight2_create(const gs_cluster *c, int64_t n, const int32_t *arrive, const int32_t *gpus, const int32_t *gpc,
                      const double *duration, const int64_t *mem_bytes, int64_t span_cap) {
  tight2 *t = (tight2 *)calloc(1, sizeof(tight2));
  t->M = c->num_switch * c->num_node_p_switch; t->G = c->num_gpu_p_node;
  const int kc = c->num_cpu_p_node / c->cpu_per_task, km = c->mem_p_node / c->mem_per_task;
  t->K = kc < km ? kc : km;
  t->cap_bytes = (int64_t)c->gpu_mem_cap_mib << 20; t->fit_limit = t->cap_bytes - ((int64_t)500 << 20);
  t->gmask = t->G >= 64 ? ~0ull : ((1ull << t->G) - 1ull);
  t->n = n; t->arrive = arrive; t->gpus = gpus; t->gpc = gpc; t->duration = duration; t->mem_bytes = mem_bytes;
  double maxd = 1.0;
  int64_t worst = 0;
  for (int64_t j = 0; j < n; ++j) {
    if (duration[j] > maxd) maxd = duration[j];
    const int64_t tasks = gpus[j] / gpc[j];
    worst += tasks < t->M ? tasks : t->M;
  }
  int W = 256; while (W < (int)maxd + 3) W <<= 1;
  t->W = W; t->wmask = W - 1;
  const size_t N = (size_t)(n > 0 ? n : 1), M = (size_t)t->M;
  t->busy = (uint64_t *)calloc(M, 8); t->idle = (int32_t *)malloc(4 * M); t->kfree = (int32_t *)malloc(4 * M);
  t->everf = (uint8_t *)calloc(M, 1);
  t->whead = (int32_t *)malloc(4 * (size_t)W); t->wmem = (int64_t *)malloc(8 * (size_t)W);
  t->st = (t2_state *)malloc(sizeof(t2_state) * N);
  t->stack_job = (int32_t *)malloc(4 * (N + 1)); t->stack_arr = (int32_t *)malloc(4 * (N + 1));
  t->rec2 = (t2_job_run *)malloc(sizeof(t2_job_run) * N);
  t->fin = (int32_t *)malloc(4 * N);
  t->span_cap = span_cap > 0 ? span_cap : (worst > 0 ? worst : 1);
  t->spans = (gs_span *)malloc(sizeof(gs_span) * (size_t)t->span_cap);
  tight2_reset(t);
  return t;
}

void tight2_restart(tight2 *t) { tight2_reset(t); }

/* arrival tick of queue entry i (0 = bottom); the head may live in registers */
static inline int q_arrive(const tight2 *t, int i) {
  if (t->hvalid && i == t->top - 1) return t->harr;
  return t->stack_arr[i];
}

static void probe_next_fin(tight2 *t, int from) {
  t->next_fin = T2_INF; t->nf_head = -1;
  if (t->running <= 0) return;
  for (int v = from;; ++v) {
    const int h = t->whead[v & t->wmask];
    if (h >= 0) { t->next_fin = v; t->nf_head = h; return; }
  }
}

/* One window: at most max_ticks ticks (<= 0: no limit), at most capA / capB records.  Returns 0 or a gs_status. */
int tight2_run(tight2 *t, int64_t max_ticks, int64_t capA, int64_t capB, gs_evrow *ev, gs_qrow *qr, gs_nodeev *ne,
               int64_t *nev_out, int64_t *nq_out, int64_t *nne_out) {
  const int M = t->M, G = t->G;
  const int64_t n = t->n;
  int64_t na = 0, nb = 0, nne = 0;
  int ever_rec = -1;                 /* busy-node count of the last node event of this window (-1: none yet) */
  int64_t budget = max_ticks > 0 ? max_ticks : 0x7fffffffLL;
  if (budget > 0x7fffffffLL - t->delta - 2) budget = 0x7fffffffLL - t->delta - 2;
  const int t_end = t->delta + (int)budget;
  t->row_first = t->delta;
  int force = 1;
  /* a resumed window starts with the head on the stack */
  while (!t->done && t->status == 0) {
    if (na >= capA || nb >= capB) break;
    if (!force && (t->top == 0 || t->blocked)) {
      const int next_arr = t->p < n ? t->arrive[t->p] : T2_INF;
      const int nf1 = t->next_fin == T2_INF ? T2_INF : t->next_fin - 1;
      const int t_next = next_arr < nf1 ? next_arr : nf1;
      if (t_next > t->delta) {
        const int t_to = t_next < t_end ? t_next : t_end;
        if (t->blocked) t->evals += (int64_t)(t_to - t->delta) * M;
        t->delta = t_to;
      }
    }
    if (t->delta >= t_end) break;
    int changed = force;
    force = 0;
    const int delta = t->delta;
    /* A */
    if (t->p < n && t->arrive[t->p] <= delta) {
      int64_t b = t->p;
      while (b < n && t->arrive[b] <= delta) ++b;
      const int cnt = (int)(b - t->p);
      if (t->hvalid) { t->stack_job[t->scount] = t->hjob; t->stack_arr[t->scount] = t->harr; t->scount += 1; }
      for (int i = 0; i < cnt - 1; ++i) { t->stack_job[t->scount + i] = (int32_t)(t->p + cnt - 1 - i); t->stack_arr[t->scount + i] = delta; }
      t->scount += cnt - 1;
      if (t->top == 0) t->bottom_arr = delta;
      t->hjob = (int)t->p; t->harr = delta; t->hvalid = 1;
      t->top += cnt; t->p += cnt;
      t->sum_arr += (int64_t)cnt * delta;
      t->blocked = 0; changed = 1;
    }
    /* B */
    if (t->top > 0 && t->blocked) t->evals += M;     /* the reference tries (and fails) on this tick too */
    if (t->top > 0 && !t->blocked) {
      if (!t->hvalid) { t->scount -= 1; t->hjob = t->stack_job[t->scount]; t->harr = t->stack_arr[t->scount]; t->hvalid = 1; }
      const int j = t->hjob;
      const int hg = t->gpus[j], hc = t->gpc[j], tasks = hg / hc;
      const int placeable = t->mem_bytes[j] < t->fit_limit;
      int ok = 0, nspans = 0, where = 0;
      uint64_t mask0 = 0;
      const int64_t sf = t->span_used;
      if (hg <= G) {
        int found = -1;
        for (int nd = 0; nd < M; ++nd) {
          if (t->idle[nd] >= hg && t->kfree[nd] >= tasks) {
            if (!placeable) { t->kfree[nd] -= tasks; continue; }
            found = nd; break;
          }
        }
        if (found >= 0 && t->span_used + 1 > t->span_cap) { t->status = GS_ERR_CAPACITY; found = -1; }
        if (found >= 0) {
          const uint64_t tk = take_lowest(~t->busy[found] & t->gmask, hg);
          t->busy[found] |= tk; t->idle[found] -= hg; t->kfree[found] -= tasks;
          if (!t->everf[found]) { t->everf[found] = 1; t->ever += 1; }
          t->spans[sf].node = found; t->spans[sf].ntasks = (int32_t)((uint32_t)tasks | GS_SPAN_FIRST); t->spans[sf].devmask = tk;
          ok = 1; nspans = 1; mask0 = tk; where = found | ((tasks - 1) << 20);
          t->evals += found + 1;
        } else t->evals += M;
      } else {
        int cum = 0, last = -1;
        if (placeable) {
          for (int nd = 0; nd < M; ++nd) {
            int cp = t->idle[nd] / hc; if (t->kfree[nd] < cp) cp = t->kfree[nd];
            if (cp <= 0) continue;
            cum += cp;
            if (cum >= tasks) { last = nd; break; }
          }
        } else {
          for (int nd = 0; nd < M; ++nd) {
            int cp = t->idle[nd] / hc; if (t->kfree[nd] < cp) cp = t->kfree[nd];
            if (cp > 0) t->kfree[nd] -= 1;
          }
        }
        if (last >= 0 && t->span_used + (tasks < M ? tasks : M) > t->span_cap) { t->status = GS_ERR_CAPACITY; last = -1; }
        if (last >= 0) {
          int rem = tasks, last_node = 0;
          for (int nd = 0; nd <= last; ++nd) {
            int cp = t->idle[nd] / hc; if (t->kfree[nd] < cp) cp = t->kfree[nd];
            if (cp <= 0) continue;
            const int take = cp < rem ? cp : rem;
            if (take <= 0) continue;
            const uint64_t tk = take_lowest(~t->busy[nd] & t->gmask, take * hc);
            t->busy[nd] |= tk; t->idle[nd] -= take * hc; t->kfree[nd] -= take;
            if (!t->everf[nd]) { t->everf[nd] = 1; t->ever += 1; }
            gs_span *sp = &t->spans[sf + nspans];
            sp->node = nd; sp->ntasks = (int32_t)((uint32_t)take | (nspans == 0 ? GS_SPAN_FIRST : 0u)); sp->devmask = tk;
            ++nspans; rem -= take; last_node = nd;
          }
          ok = 1;
          where = (int)(0x80000000u | (uint32_t)sf);
          mask0 = (uint64_t)(uint32_t)nspans | ((uint64_t)(uint32_t)hg << 32);
          t->evals += last_node + 1;
        } else t->evals += M;
      }
      if (ok) {
        int need = need_of(t->duration[j]);
        if (need > t->wmask) { t->status = GS_ERR_ARG; need = t->wmask; }
        const int endt = delta + need, bk = endt & t->wmask;
        t->span_used += nspans;
        t2_state js; js.where = where; js.mask0 = mask0; js.next = -1;
        if (endt < t->next_fin) { t->next_fin = endt; t->nf_head = j; }
        else if (endt == t->next_fin) { js.next = t->nf_head; t->nf_head = j; }
        else js.next = t->whead[bk];
        t->whead[bk] = j;
        t->wmem[bk] += (int64_t)hg * (t->mem_bytes[j] < t->cap_bytes ? t->mem_bytes[j] : t->cap_bytes);
        t->st[j] = js;
        t->rec2[j].start = delta; t->rec2[j].run_ticks = need;
        t->top -= 1; t->sum_arr -= t->harr; t->running += 1;
        t->busy_gpus += hg;
        t->mem_busy += (int64_t)hg * (t->mem_bytes[j] < t->cap_bytes ? t->mem_bytes[j] : t->cap_bytes);
        t->hvalid = 0; changed = 1;
      } else if (placeable) t->blocked = 1;
    }
    /* D/E */
    const int now = delta + 1;
    if (t->next_fin == now) {
      const int sl = now & t->wmask;
      int h = t->nf_head;
      const int f0 = t->finished;
      int c = 0;
      while (h >= 0) {
        const t2_state js = t->st[h];
        if (js.where >= 0) {
          const int nd = js.where & 0xfffff, nt = ((js.where >> 20) & 63) + 1;
          t->busy[nd] &= ~js.mask0; t->idle[nd] += __builtin_popcountll(js.mask0); t->kfree[nd] += nt;
          t->busy_gpus -= __builtin_popcountll(js.mask0);
        } else {
          const int first = js.where & 0x7fffffff, scnt = (int)(uint32_t)(js.mask0 & 0xffffffffull);
          for (int i = 0; i < scnt; ++i) {
            const gs_span *sp = &t->spans[first + i];
            t->busy[sp->node] &= ~sp->devmask; t->idle[sp->node] += __builtin_popcountll(sp->devmask);
            t->kfree[sp->node] += (int)((uint32_t)sp->ntasks & ~GS_SPAN_FIRST);
          }
          t->busy_gpus -= (int)(uint32_t)(js.mask0 >> 32);
        }
        t->fin[f0 + c] = h;
        c += 1;
        h = js.next;
      }
      t->finished += c; t->running -= c;
      t->mem_busy -= t->wmem[sl];
      t->whead[sl] = -1; t->wmem[sl] = 0;
      for (int i = 0; i < c / 2; ++i) { const int32_t a = t->fin[f0 + i]; t->fin[f0 + i] = t->fin[f0 + c - 1 - i]; t->fin[f0 + c - 1 - i] = a; }
      probe_next_fin(t, now + 1);
      t->blocked = 0; changed = 1;
    }
    /* H */
    if (changed) {
      if (t->top > 0) {
        const int top = t->top;
        const int ilo = top - 1 - ((top - 1) >> 1), ihi = top - 1 - (top >> 1);
        gs_qrow *q = &qr[nb];
        q->now = now; q->arrive_sum = t->sum_arr; q->oldest_arrive = t->bottom_arr;
        q->med_lo_arrive = q_arrive(t, ilo); q->med_hi_arrive = q_arrive(t, ihi);
        nb += 1;
      }
      if (t->ever != ever_rec) { ne[nne].now = now; ne[nne].busy_nodes = t->ever; nne += 1; ever_rec = t->ever; }
      gs_evrow *e = &ev[na];
      e->now = now; e->queued = t->top; e->finished = t->finished;
      e->busy_gpus = (uint16_t)t->busy_gpus; e->running = (uint16_t)t->running;
      e->mem_busy_bytes = t->mem_busy;
      na += 1;
    }
    t->delta = now;
    t->done = (n - t->p) + t->running == 0;
  }
  if (t->hvalid) { t->stack_job[t->scount] = t->hjob; t->stack_arr[t->scount] = t->harr; t->scount += 1; t->hvalid = 0; }
  for (int i = 0; i < t->scount; ++i) { t->rec2[t->stack_job[i]].start = -1; t->rec2[t->stack_job[i]].run_ticks = 0; }
  *nev_out = na; *nq_out = nb; *nne_out = nne;
  return t->status;
}

void tight2_info(const tight2 *t, gs_window_info *w, int64_t nev, int64_t nq, int64_t nne, int64_t *events, int64_t *evals, int32_t *done) {
  w->row_first = t->row_first; w->ticks = t->delta; w->ev_rows = nev; w->q_rows = nq; w->node_events = nne; w->spans_used = t->span_used;
  w->admitted = t->p; w->finished = t->finished; w->n = t->n;
  *events = t->p + 2 * (int64_t)t->finished + t->running; *evals = t->evals; *done = t->done;
}

const t2_job_run *tight2_jobs(const tight2 *t) { return t->rec2; }
const int32_t *tight2_finish_order(const tight2 *t) { return t->fin; }
const gs_span *tight2_spans(const tight2 *t) { return t->spans; }
