// gs_tick_warp.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// fifo + yarn tick engine, one warp (or half warp) per replica.
#pragma once

__device__ __forceinline__ int meta_cap(unsigned mt, int gpc) {
  // tasks a node can still take: min(idle devices / gpus per task, free task slots)
  const int idle = (int)(mt & 0xffu), kfree = (int)(mt >> 16);
  return min(gpc == 1 ? idle : idle / gpc, kfree);
}

// Shared-memory geometry shared by both tick kernels
#define LW 128          // finish-tick buckets kept in shared memory (ticks, power of two)
#define SCACHE 4        // cached top-of-queue entries (power of two)
#define WARP_EXTRA_BYTES (LW * 8 + SCACHE * 8)

__device__ __forceinline__ unsigned long long take_lowest(unsigned long long idle, int cnt, int G) {
  // the `cnt` lowest set bits of `idle` (devices are claimed in index order, node.py:208-216)
  if (cnt == 1) return idle & (~idle + 1ull);
  if (G <= 32) {
    unsigned m = (unsigned)idle;
    if (cnt >= __popc(m)) return idle;
    for (int i = 0; i < cnt; ++i) m &= m - 1u;
    return (unsigned long long)((unsigned)idle ^ m);
  }
  unsigned long long m = idle;
  if (cnt >= __popcll(m)) return idle;
  for (int i = 0; i < cnt; ++i) m &= m - 1ull;
  return idle ^ m;
}

// One warp advances one replica.  Lanes stripe over the node table for placement; all other
// state is warp-uniform.  Per warp in shared memory: node table (busy mask u64 + slot counter
// i32 per node), a LW-tick window of the finish wheel (head, tail) and the top SCACHE queue
// entries.  Every global load on the per-tick path is issued one tick before its value is
// needed (pending wheel bucket, release record of the next tick's first finisher) or comes
// from a register-resident 32-record window of the trace, so the loop body has no dependent
// DRAM/L2 round trip in the common case.
template <int SUB>
__global__ void __launch_bounds__(32, GS_TICK_MINBLOCKS) gs_tick_kernel(SimDev *sims, int nsims, long long max_ticks, int smem_stride) {
  // SUB lanes own one replica: 32 = a whole warp, 16 = two replicas per warp sharing the
  // (mostly warp-uniform) instruction stream.  Every collective uses the group's own lane
  // mask, so the groups of a warp may diverge freely.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wl = threadIdx.x & 31;              // lane within the warp
  const int lane = wl & (SUB - 1);              // lane within the replica's group
  const int gbase = wl & ~(SUB - 1);            // first warp lane of the group
  const unsigned GM = (SUB == 32) ? 0xffffffffu : (((1u << SUB) - 1u) << gbase);
  const int grp = wl / SUB;
  const int sim = blockIdx.x * (32 / SUB) + grp;
  if (sim >= nsims) return;
  SimDev &S = sims[sim];
  if (S.done || S.status != 0 || S.policy != GS_SCHED_FIFO) return;

  const int M = S.M, G = S.G, K = S.K, n = S.n;
  unsigned long long *busy = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)grp * smem_stride);
  int *kk = reinterpret_cast<int *>(busy + M);
  int2 *sstk = reinterpret_cast<int2 *>(kk + M + (M & 1));   // 8-byte aligned

  const JobIn *__restrict__ jobs = S.jobs;
  gs_job_rec *rec = S.rec;
  JobState *jst = S.jst;
  int2 *stack = reinterpret_cast<int2 *>(S.stack);
  int *fin = S.fin, *gwh = S.wheel_head, *gwt = S.wheel_tail;
  gs_span *spans = S.spans;
  const int wmask = S.wheel_mask;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int netcost = S.netcost;
  const long long span_cap = S.span_cap;      // SimDev lives in global memory: keep loop-invariant fields in registers
  int2 *sref = S.sref;
  const unsigned long long gmask = (G >= 64) ? ~0ull : ((1ull << G) - 1ull);

  int delta = S.delta, p = S.p, top = S.top, running = S.running, finished = S.finished;
  int ever = S.ever, busy_gpus = S.busy_gpus, status = 0;
  long long mem_busy = S.mem_busy, sum_arr = S.sum_arr, span_used = S.span_used;
  long long evals = S.evals, started = S.started, ticks = S.ticks;
  const long long row_first = ticks;
  gs_tick_row *rows = S.rows;
  const long long rows_cap = S.rows_cap;
  long long budget_ll = max_ticks > 0 ? max_ticks : 0x7fffffffLL;
  if (budget_ll > rows_cap) budget_ll = rows_cap;     // the row window bounds a launch anyway
  int budget = (int)(budget_ll > 0x7fffffffLL ? 0x7fffffffLL : budget_ll);
  int4 *rowp = reinterpret_cast<int4 *>(rows);
  int tick_i = 0;                                     // ticks done in this launch

  // ---- stage persistent state: node table, wheel window, queue top
  for (int i = lane; i < M; i += SUB) {
    const unsigned long long bz = S.nbusy[i];
    const unsigned kv = (unsigned)S.nk[i];
    busy[i] = bz;
    kk[i] = (int)((unsigned)(G - __popcll(bz)) | ((kv & EVER_BIT) ? 0x100u : 0u) | ((unsigned)(K - (int)(kv & ~EVER_BIT)) << 16));
  }
  for (int i = max(top - SCACHE, 0) + lane; i < top; i += SUB) sstk[i & (SCACHE - 1)] = stack[i];
  int cache_lo = max(top - SCACHE, 0);           // queue entries [cache_lo, top) are cached
  __syncwarp(GM);

  // trace window: lane l holds record wbase + l
  int wbase = p & ~(SUB - 1);
  JobIn wj;
  wj.arrive = 0x7fffffff; wj.gpus = 1; wj.gpc = 1; wj.ps = 0; wj.memb = 0; wj.dur = 0.0;
  if (wbase + lane < n) wj = jobs[wbase + lane];
  // arrival tick of the next trace record: most ticks admit nothing and only compare this scalar
  int next_arr = 0x7fffffff;
  if (p < n) next_arr = __shfl_sync(GM, wj.arrive, gbase + (p - wbase));

  // queue head (cached while it stays the head)
  int head = -1, hg = 1, hgpc = 1, htasks = 1, hps = 0, harr = 0;
  long long hmemb = 0;
  double hdur = 0.0;
  bool head_valid = false;
  int bottom_arr = (top > 0) ? stack[0].y : 0;


  bool done = (n == 0);

  while (!done && tick_i < budget && status == 0) {
    // ---------------- A. admit arrivals (gen_jobs + head insert)
    if (next_arr <= delta) {
      int cnt = 0, q = p;
      while (q < n) {
        const int idx = wbase + lane;
        const unsigned b = GBALLOT(idx >= q && idx < n && wj.arrive <= delta);
        const int c = __popc(b);
        if (c > 0 && cnt == 0) {
          // the batch's first job becomes the queue head: take its record out of the window now
          const int src = p - wbase;
          hg = __shfl_sync(GM, wj.gpus, gbase + src); hgpc = __shfl_sync(GM, wj.gpc, gbase + src);
          harr = delta;
          if (netcost) hps = __shfl_sync(GM, wj.ps, gbase + src);
          hmemb = __shfl_sync(GM, wj.memb, gbase + src);
          hdur = __longlong_as_double(__shfl_sync(GM, __double_as_longlong(wj.dur), gbase + src));
        }
        cnt += c; q += c;
        if (q < wbase + SUB || q >= n) break;
        wbase += SUB;
        wj.arrive = 0x7fffffff;
        if (wbase + lane < n) wj = jobs[wbase + lane];
      }
      if (cnt > 0) {
        // batch [p, p+cnt) lands AHEAD of the queue, first of the batch on top (quirk Q2)
        if (cnt == 1) {
          if (lane == 0) { const int2 e = make_int2(p, delta); stack[top] = e; sstk[top & (SCACHE - 1)] = e; }
        } else {
          for (int i = lane; i < cnt; i += SUB) {
            const int2 e = make_int2(p + cnt - 1 - i, delta);
            stack[top + i] = e;
            if (i >= cnt - SCACHE) sstk[(top + i) & (SCACHE - 1)] = e;
          }
        }
        if (top == 0) bottom_arr = delta;
        head = p; head_valid = true; htasks = hgpc == 1 ? hg : hg / hgpc;
        top += cnt; p += cnt;
        if (top - cache_lo > SCACHE) cache_lo = top - SCACHE;
        sum_arr += (long long)cnt * delta;
        __syncwarp(GM);
      }
      next_arr = 0x7fffffff;
      if (p < n) next_arr = __shfl_sync(GM, wj.arrive, gbase + (p - wbase));
    }
    // ---------------- B. one scheduling attempt on the queue head (quirks Q1, Q3)
    if (top > 0) {
      if (!head_valid) {
        if (top - 1 >= cache_lo) head = sstk[(top - 1) & (SCACHE - 1)].x;
        else { head = stack[top - 1].x; cache_lo = top; }
        if (head >= wbase && head < wbase + SUB) {
          const int src = head - wbase;
          hg = __shfl_sync(GM, wj.gpus, gbase + src); hgpc = __shfl_sync(GM, wj.gpc, gbase + src);
          hps = __shfl_sync(GM, wj.ps, gbase + src); harr = __shfl_sync(GM, wj.arrive, gbase + src);
          hmemb = __shfl_sync(GM, wj.memb, gbase + src);
          hdur = __longlong_as_double(__shfl_sync(GM, __double_as_longlong(wj.dur), gbase + src));
        } else {
          const JobIn jr = jobs[head];
          hg = jr.gpus; hgpc = jr.gpc; hmemb = jr.memb; hdur = jr.dur; hps = jr.ps; harr = jr.arrive;
        }
        htasks = hgpc == 1 ? hg : hg / hgpc;
        head_valid = true;
      }
      const bool placeable = hmemb < fit_limit;   // Device.can_fit on an empty device
      bool ok = false;
      int first_node = -1, nspans = 0;
      const int span_first = (int)span_used;
      unsigned long long mask0 = 0;
      if (hg <= G) {
        // try_single_node_alloc_ms: first node (id order) that fits the whole job
        int found = -1;
        for (int base = 0; base < M; base += SUB) {
          const int nd = base + lane;
          bool fit = false;
          if (nd < M) {
            const unsigned mt = (unsigned)kk[nd];
            fit = (int)(mt & 0xffu) >= hg && (int)(mt >> 16) >= htasks;
          }
          const unsigned b = GBALLOT(fit);
          if (!placeable) {          // quirk Q21: cpu/mem charged for every task, never refunded
            if (fit) kk[nd] -= htasks << 16;
            continue;
          }
          if (b) { found = base + __ffs(b) - 1; break; }
        }
        if (found >= 0 && span_used + 1 > span_cap) { status = GS_ERR_CAPACITY; found = -1; }
        if (found >= 0) {
          ok = true; first_node = found; nspans = 1;
          bool fresh = false;
          if (lane == (found & (SUB - 1))) {
            const unsigned long long take = take_lowest(~busy[found] & gmask, hg, G);
            busy[found] |= take;
            const unsigned kv = (unsigned)kk[found];
            kk[found] = (int)((kv - (unsigned)hg - ((unsigned)htasks << 16)) | 0x100u);
            mask0 = take;
            fresh = !(kv & 0x100u);
            gs_span sp; sp.node = found; sp.ntasks = htasks; sp.devmask = take;
            spans[span_first] = sp;
          }
          mask0 = __shfl_sync(GM, mask0, gbase + (found & (SUB - 1)));
          ever += __popc(GBALLOT(fresh));
          evals += found + 1;
        } else {
          evals += M;
        }
      } else {
        // try_cross_node_alloc_ms: walk nodes in id order, each takes what it can hold
        int cum = 0, last_base = -1;
        if (placeable) {
          for (int base = 0; base < M; base += SUB) {
            const int nd = base + lane;
            const int c = (nd < M) ? meta_cap((unsigned)kk[nd], hgpc) : 0;
            cum += __reduce_add_sync(GM, c);
            if (cum >= htasks) { last_base = base; break; }
          }
        } else {
          for (int base = 0; base < M; base += SUB) {   // quirk Q21, cross-node flavour: one task charged per node
            const int nd = base + lane;
            if (nd < M && meta_cap((unsigned)kk[nd], hgpc) > 0) kk[nd] -= 1 << 16;
          }
        }
        if (last_base >= 0 && span_used + min(htasks, M) > span_cap) { status = GS_ERR_CAPACITY; last_base = -1; }
        if (last_base >= 0) {
          // pass 1 proved the job fits: commit (a failed walk is rolled back exactly by the
          // reference, algorithm.py:378-387, so no state changes in that case)
          ok = true;
          int rem = htasks, last_node = 0;
          for (int base = 0; base <= last_base; base += SUB) {
            const int nd = base + lane;
            const int c = (nd < M) ? meta_cap((unsigned)kk[nd], hgpc) : 0;
            int incl = c;
            #pragma unroll
            for (int o = 1; o < SUB; o <<= 1) { int v = __shfl_up_sync(GM, incl, o, SUB); if (lane >= o) incl += v; }
            const int take = min(c, max(rem - (incl - c), 0));
            const unsigned tb = GBALLOT(take > 0);
            bool fresh = false;
            if (take > 0) {
              const unsigned long long tk = take_lowest(~busy[nd] & gmask, take * hgpc, G);
              busy[nd] |= tk;
              const unsigned kv = (unsigned)kk[nd];
              kk[nd] = (int)((kv - (unsigned)(take * hgpc) - ((unsigned)take << 16)) | 0x100u);
              fresh = !(kv & 0x100u);
              const int slot = nspans + __popc(tb & ((1u << lane) - 1u));
              gs_span sp; sp.node = nd; sp.ntasks = take; sp.devmask = tk;
              spans[span_first + slot] = sp;
              if (slot == 0) { mask0 = tk; first_node = nd; }
            }
            ever += __popc(GBALLOT(fresh));
            if (tb) last_node = base + 31 - __clz(tb);
            nspans += __popc(tb);
            const int tot = __shfl_sync(GM, incl, gbase + SUB - 1);
            rem -= min(rem, tot);
          }
          {  // first-span fields live in whichever lane owned slot 0
            const int src = __ffs(GBALLOT(first_node >= 0)) - 1;
            mask0 = __shfl_sync(GM, mask0, gbase + src);
            first_node = __shfl_sync(GM, first_node, gbase + src);
          }
          evals += last_node + 1;
        } else {
          evals += M;
        }
      }
      __syncwarp(GM);
      if (ok) {
        // ---- commit: pop, network cost, start (algorithm.py:198-200, schedule.py:49-54,164-167)
        const int j = head;
        double dur2 = hdur;
        if (netcost && hps > 1) {
          // (model_size/bandwidth + cross*latency) * (iterations*2.0), network_service.py:34-37
          const double mps = __ddiv_rn(S.model_mb[j], S.bandwidth);
          const double nis = __dmul_rn((double)nspans, S.latency);
          const double rt = __dmul_rn(S.iters[j], 2.0);
          dur2 = __dadd_rn(hdur, __dmul_rn(__dadd_rn(mps, nis), rt));
        }
        const double eff = dur2 > hdur ? dur2 : hdur;               // Job.get_duration (job.py:206-210)
        const double cl = ceil(eff);
        int need = cl < 1.0 ? 1 : (cl > 1.0e9 ? 0x7fffffff : (int)cl);   // quirk Q11
        if (need > wmask) { status = GS_ERR_ARG; need = wmask; }
        const int endt = delta + need;
        span_used += nspans;
        const long long memc = (long long)hg * (hmemb < cap_bytes ? hmemb : cap_bytes);
        JobState js; js.next = -1; js.node0 = nspans == 1 ? first_node : span_first; js.mask0 = mask0;
        js.memc = memc; js.gpus = hg | ((nspans == 1 ? htasks : 0) << 24); js.cnt_gpc = nspans | (hgpc << 24);
        // append to the finish-tick bucket of the timing wheel (start order)
        const int gs_ = endt & wmask;
        const int tl = gwt[gs_];
        if (lane == 0) {
          if (tl < 0) gwh[gs_] = j;
          gwt[gs_] = j;
          gs_job_rec r; r.start = delta; r.end = endt; r.jct = need; r.preempt = 1; r.duration = dur2;
          rec[j] = r;
          jst[j] = js;
          sref[j] = make_int2(span_first, nspans);
          if (tl >= 0) jst[tl].next = j;
        }
        top -= 1;
        sum_arr -= harr;
        running += 1; started += 1;
        busy_gpus += hg;
        mem_busy += memc;
        head = -1; head_valid = false;
        __syncwarp(GM);
      }
    }
    // ---------------- D/E. time advances; release jobs whose finish tick is now
    const int now = delta + 1;
    {
      const int sl = now & wmask;
      int h = gwh[sl];
      if (h >= 0) {
        __syncwarp(GM);
        if (lane == 0) { gwh[sl] = -1; gwt[sl] = -1; }
        while (h >= 0) {
          const JobState js = jst[h];
          const int scnt = JS_CNT(js.cnt_gpc), sgpc = JS_GPC(js.cnt_gpc);
          if (scnt == 1) {
            if (lane == 0) { busy[js.node0] &= ~js.mask0; kk[js.node0] += JS_GPUS(js.gpus) + (JS_NT0(js.gpus) << 16); }
          } else {
            for (int i = lane; i < scnt; i += SUB) {
              const gs_span sp = spans[js.node0 + i];
              busy[sp.node] &= ~sp.devmask; kk[sp.node] += sp.ntasks * sgpc + (sp.ntasks << 16);
            }
          }
          if (lane == 0) fin[finished] = h;
          finished += 1; running -= 1;
          busy_gpus -= JS_GPUS(js.gpus);
          mem_busy -= js.memc;
          h = js.next;
        }
        __syncwarp(GM);
      }
    }
    // ---------------- H. statistics row (schedule.py:95-133) from O(1) counters
    {
      int pmax = 0, mlo = 0, mhi = 0;
      if (top > 0) {
        // queue is a stack with non-decreasing arrival ticks bottom->top, so the sorted
        // pending list is the stack read top->bottom: median/max are index look-ups
        const int ilo = top - 1 - ((top - 1) >> 1), ihi = top - 1 - (top >> 1);
        const int a_lo = ilo >= cache_lo ? sstk[ilo & (SCACHE - 1)].y : stack[ilo].y;
        const int a_hi = ihi >= cache_lo ? sstk[ihi & (SCACHE - 1)].y : stack[ihi].y;
        pmax = now - bottom_arr; mlo = now - a_lo; mhi = now - a_hi;
      }
      if (lane == 0) {
        int4 *dst = rowp;
        const int tg = M * G;
        const long long ps = top > 0 ? (long long)top * now - sum_arr : 0;
        dst[0] = make_int4(now, M - ever, ever, busy_gpus);
        dst[1] = make_int4(tg - busy_gpus, running, top, finished);
        dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(ps & 0xffffffffLL), (int)(ps >> 32));
        dst[3] = make_int4(pmax, mlo, mhi, 0);
      }
    }
    tick_i += 1; rowp += 4;
    delta = now;
    done = (n - p) + running == 0;      // schedule.py:185 -- the queue is NOT counted (quirk Q4)
  }

  // ---------------- persist: node table, wheel window and pending bucket go back to global memory
  __syncwarp(GM);
  for (int i = lane; i < M; i += SUB) {
    const unsigned mt = (unsigned)kk[i];
    S.nbusy[i] = busy[i];
    S.nk[i] = (int)((unsigned)(K - (int)(mt >> 16)) | ((mt & 0x100u) ? EVER_BIT : 0u));
  }
  if (lane == 0) {
    S.delta = delta; S.p = p; S.top = top; S.running = running; S.finished = finished;
    S.ever = ever; S.busy_gpus = busy_gpus; S.mem_busy = mem_busy; S.sum_arr = sum_arr;
    S.span_used = span_used; S.events = (long long)p + started + finished; S.evals = evals; S.started = started;
    S.ticks = ticks + tick_i; S.row_first = row_first; S.done = done ? 1 : 0; S.status = status;
  }
}

