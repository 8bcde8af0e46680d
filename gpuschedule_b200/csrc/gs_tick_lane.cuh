// gs_tick_lane.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// fifo + yarn tick engine, one lane per replica (experiment, selectable).
#pragma once

// ------------------------------------------------------------------ lane engine
// One THREAD owns one replica; a warp carries up to 32 unrelated replicas.  This is the
// throughput kernel: a replica's tick is almost all scalar bookkeeping, so a whole warp
// per replica wastes 31/32 of the issue slots, and the number of replicas in flight is
// capped by HBM capacity (~18 MB per 100k-job replica), so per-tick LATENCY decides
// throughput.  Everything the common path touches therefore lives in shared memory or
// registers, and every global load is issued one iteration before its value is needed:
//   * node table        meta word per node: idle devices (0-7) | ever (8) | free slots (16-31),
//                       plus the busy-device bitmap (32 or 64 bit)
//   * wheel window      finish-tick buckets (head, tail) for the next LW ticks; far buckets
//                       stay in the global wheel and are pulled in LW ticks ahead
//   * job ring          the next few 32-byte trace records, refilled one per tick
//   * stack cache       the top 4 queue entries (job, arrival tick)
//   * release record    JobState of the job finishing next tick, prefetched into registers
// Shared memory is laid out [word][lane] so lane l always hits bank l: conflict-free no
// matter which node / slot each lane is looking at.  First fit is a serial scan from `lo`,
// the lowest node with an idle device (first fit packs low ids, so the scan is short).
// Lanes never share data: no warp collectives except the per-tick reconvergence barrier.
#define META_IDLE(m) ((int)((m) & 0xffu))
#define META_EVER 0x100u
#define META_KFREE(m) ((int)((m) >> 16))
#define RING 8          // job-record ring, entries (power of two)
#define LANE_EXTRA_WORDS (2 * LW + RING * 8 + SCACHE * 2)

template <typename MaskT>
__global__ void __launch_bounds__(32) gs_lane_kernel(SimDev *sims, int nsims, long long max_ticks, int Mmax, int L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int sim = blockIdx.x * L + lane;
  const bool in_range = lane < L && sim < nsims;
  SimDev &S = sims[in_range ? sim : 0];
  const bool alive = in_range && !S.done && S.status == 0 && S.policy == GS_SCHED_FIFO;

  const int M = S.M, G = S.G, K = S.K, n = S.n;
  const int MW = (sizeof(MaskT) == 8) ? 3 : 2;
  uint32_t *meta = reinterpret_cast<uint32_t *>(smem_raw) + (lane < L ? lane : 0);   // [nd * L]
  uint32_t *mlo = meta + (size_t)Mmax * L;
  uint32_t *mhi = meta + (size_t)2 * Mmax * L;                 // only when MaskT is 64 bit
  uint32_t *swh = meta + (size_t)MW * Mmax * L;                // [slot * L]  bucket head
  uint32_t *swt = swh + (size_t)LW * L;                        //             bucket tail
  uint32_t *ring = swt + (size_t)LW * L;                       // [(slot * 8 + word) * L]
  uint32_t *sstk = ring + (size_t)RING * 8 * L;                // [(slot * 2 + {job,arrive}) * L]
  const MaskT gmask = (G >= (int)(8 * sizeof(MaskT))) ? (MaskT)~(MaskT)0 : (MaskT)(((MaskT)1 << G) - 1);

  const JobIn *__restrict__ jobs = S.jobs;
  gs_job_rec *rec = S.rec;
  JobState *jst = S.jst;
  int2 *sref = S.sref;
  int2 *stack = reinterpret_cast<int2 *>(S.stack);
  int *fin = S.fin, *gwh = S.wheel_head, *gwt = S.wheel_tail;
  gs_span *spans = S.spans;
  const int wmask = S.wheel_mask;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int netcost = S.netcost;

  int delta = S.delta, p = S.p, top = S.top, running = S.running, finished = S.finished;
  int ever = S.ever, busy_gpus = S.busy_gpus, status = 0;
  long long mem_busy = S.mem_busy, sum_arr = S.sum_arr, span_used = S.span_used;
  long long evals = S.evals, started = S.started, ticks = S.ticks;
  const long long row_first = ticks;
  gs_tick_row *rows = S.rows;
  const long long rows_cap = S.rows_cap;
  long long budget = max_ticks > 0 ? max_ticks : 0x7fffffffffffffffLL;

  // ---- stage the persistent state into shared memory
  int lo = M;
  int pf = p;                        // ring holds trace records [max(ring_lo, pf - RING), pf)
  int ring_lo = p;
  int pend_h = -1, pend_t = -1;      // bucket of tick delta + LW, loaded but not yet in the window
  if (alive) {
    for (int nd = 0; nd < M; ++nd) {
      unsigned long long bz = S.nbusy[nd];
      unsigned kv = (unsigned)S.nk[nd];
      int idle = G - __popcll(bz);
      meta[nd * L] = (uint32_t)idle | ((kv & EVER_BIT) ? META_EVER : 0u) | ((uint32_t)(K - (int)(kv & ~EVER_BIT)) << 16);
      mlo[nd * L] = (uint32_t)bz;
      if (sizeof(MaskT) == 8) mhi[nd * L] = (uint32_t)(bz >> 32);
      if (idle > 0 && nd < lo) lo = nd;
    }
    for (int t = delta + 1; t <= delta + LW - 1; ++t) {      // window ticks move from the global wheel
      int gs_ = t & wmask;
      swh[(t & (LW - 1)) * L] = (uint32_t)gwh[gs_]; swt[(t & (LW - 1)) * L] = (uint32_t)gwt[gs_];
      gwh[gs_] = -1; gwt[gs_] = -1;
    }
    { int gs_ = (delta + LW) & wmask; pend_h = gwh[gs_]; pend_t = gwt[gs_]; gwh[gs_] = -1; gwt[gs_] = -1; }
    for (; pf < n && pf < p + RING - 2; ++pf) {
      const uint4 *src = reinterpret_cast<const uint4 *>(&jobs[pf]);
      uint4 a0 = src[0], a1 = src[1];
      uint32_t *r = ring + (size_t)((pf & (RING - 1)) * 8) * L;
      r[0] = a0.x; r[L] = a0.y; r[2 * L] = a0.z; r[3 * L] = a0.w;
      r[4 * L] = a1.x; r[5 * L] = a1.y; r[6 * L] = a1.z; r[7 * L] = a1.w;
    }
    for (int i = max(top - SCACHE, 0); i < top; ++i) {
      int2 e = stack[i];
      sstk[((i & (SCACHE - 1)) * 2) * L] = (uint32_t)e.x; sstk[((i & (SCACHE - 1)) * 2 + 1) * L] = (uint32_t)e.y;
    }
  }
  int cache_lo = max(top - SCACHE, 0);     // stack entries [cache_lo, top) are in the cache

  // trace record q -> registers (ring if resident, else global)
  auto load_job = [&](int q) -> JobIn {
    JobIn r;
    if (q < pf && q >= pf - RING && q >= ring_lo) {
      const uint32_t *w = ring + (size_t)((q & (RING - 1)) * 8) * L;
      r.arrive = (int)w[0]; r.gpus = (int)w[L]; r.gpc = (int)w[2 * L]; r.ps = (int)w[3 * L];
      r.memb = (long long)(((unsigned long long)w[5 * L] << 32) | w[4 * L]);
      r.dur = __longlong_as_double((long long)(((unsigned long long)w[7 * L] << 32) | w[6 * L]));
    } else {
      r = jobs[q];
    }
    return r;
  };
  auto arrive_of = [&](int q) -> int {
    if (q >= n) return 0x7fffffff;
    if (q < pf && q >= pf - RING && q >= ring_lo) return (int)ring[(size_t)((q & (RING - 1)) * 8) * L];
    return jobs[q].arrive;
  };

  int next_arrive = alive ? arrive_of(p) : 0x7fffffff;
  int head = -1, htasks = 1;
  JobIn hj;
  hj.arrive = 0; hj.gpus = 1; hj.gpc = 1; hj.ps = 0; hj.memb = 0; hj.dur = 0.0;
  int bottom_arr = (alive && top > 0) ? stack[0].y : 0;
  // pipelined loads: issued at the end of iteration d, consumed in iteration d + 1
  bool rp_valid = false; uint4 rp0 = make_uint4(0, 0, 0, 0), rp1 = make_uint4(0, 0, 0, 0);   // trace record pf
  int pre_h = -1; JobState pre_js;                                                          // release record
  pre_js.next = -1; pre_js.node0 = 0; pre_js.mask0 = 0; pre_js.memc = 0; pre_js.gpus = 0; pre_js.cnt_gpc = 1;
  int com_j = -1; JobState com_js = pre_js;                                                 // last commit
  bool done = !alive || (n == 0);

  // Every tick starts with the whole warp reconverged (no break/return inside the body,
  // explicit barrier): otherwise independent thread scheduling lets the replicas drift
  // apart and run one at a time.
  while (true) {
    const bool go = !done && status == 0 && budget > 0 && (ticks - row_first) < rows_cap;
    if (!__any_sync(FULL, go)) break;
    if (go) {
      // ---------------- A. admit arrivals: the batch lands ahead of the queue, first job on top (Q2)
      if (next_arrive <= delta) {
        const int a = p;
        int b = p, na;
        do { ++b; na = arrive_of(b); } while (na <= delta);
        if (top == 0) bottom_arr = delta;
        for (int i = b - 1; i >= a; --i) {
          stack[top] = make_int2(i, delta);
          sstk[((top & (SCACHE - 1)) * 2) * L] = (uint32_t)i; sstk[((top & (SCACHE - 1)) * 2 + 1) * L] = (uint32_t)delta;
          ++top;
        }
        if (top - cache_lo > SCACHE) cache_lo = top - SCACHE;
        sum_arr += (long long)(b - a) * delta;
        head = a; hj = load_job(a); htasks = hj.gpc == 1 ? hj.gpus : hj.gpus / hj.gpc;
        p = b; next_arrive = na;
      }
      // ---------------- B. one attempt on the queue head (Q1, Q3)
      com_j = -1;
      if (top > 0) {
        if (head < 0) {
          if (top - 1 >= cache_lo) head = (int)sstk[(((top - 1) & (SCACHE - 1)) * 2) * L];
          else { head = stack[top - 1].x; cache_lo = top; }      // cache exhausted: deeper entries are global only
          hj = load_job(head); htasks = hj.gpc == 1 ? hj.gpus : hj.gpus / hj.gpc;
        }
        const int hg = hj.gpus, hgpc = hj.gpc;
        const bool placeable = hj.memb < fit_limit;
        bool ok = false;
        int first_node = -1, nspans = 0;
        const int span_first = (int)span_used;
        MaskT mask0 = 0;
        if (hg <= G) {
          int found = -1;
          for (int nd = lo; nd < M; ++nd) {
            uint32_t mt = meta[nd * L];
            if (META_IDLE(mt) >= hg && META_KFREE(mt) >= htasks) {
              if (!placeable) { meta[nd * L] = mt - ((uint32_t)htasks << 16); continue; }   // Q21 leak
              found = nd; break;
            }
          }
          if (found >= 0 && span_used + 1 > S.span_cap) { status = GS_ERR_CAPACITY; found = -1; }
          if (found >= 0) {
            uint32_t mt = meta[found * L];
            MaskT bz = (MaskT)mlo[found * L];
            if (sizeof(MaskT) == 8) bz |= (MaskT)((unsigned long long)mhi[found * L] << 32);
            MaskT m = (MaskT)(~bz & gmask), take = 0;
            for (int i = 0; i < hg; ++i) { MaskT bit = (MaskT)(m & (MaskT)(~m + 1)); take |= bit; m ^= bit; }
            bz |= take;
            mlo[found * L] = (uint32_t)bz;
            if (sizeof(MaskT) == 8) mhi[found * L] = (uint32_t)((unsigned long long)bz >> 32);
            if (!(mt & META_EVER)) ever += 1;
            meta[found * L] = (mt - (uint32_t)hg - ((uint32_t)htasks << 16)) | META_EVER;
            gs_span sp; sp.node = found; sp.ntasks = htasks; sp.devmask = (unsigned long long)take;
            spans[span_first] = sp;
            ok = true; first_node = found; nspans = 1; mask0 = take;
            evals += found + 1;
          } else {
            evals += M;
          }
        } else {
          int cum = 0, last = -1;
          for (int nd = lo; nd < M; ++nd) {
            uint32_t mt = meta[nd * L];
            int idle = META_IDLE(mt);
            int c = min(hgpc == 1 ? idle : idle / hgpc, META_KFREE(mt));
            if (c <= 0) continue;
            if (!placeable) { meta[nd * L] = mt - (1u << 16); continue; }     // Q21 leak, one task per node
            cum += c;
            if (cum >= htasks) { last = nd; break; }
          }
          if (last >= 0 && span_used + min(htasks, M) > S.span_cap) { status = GS_ERR_CAPACITY; last = -1; }
          if (last >= 0) {
            int rem = htasks;
            for (int nd = lo; nd <= last; ++nd) {
              uint32_t mt = meta[nd * L];
              int idle = META_IDLE(mt);
              int c = min(hgpc == 1 ? idle : idle / hgpc, META_KFREE(mt));
              if (c <= 0) continue;
              int take_n = min(c, rem);
              MaskT bz = (MaskT)mlo[nd * L];
              if (sizeof(MaskT) == 8) bz |= (MaskT)((unsigned long long)mhi[nd * L] << 32);
              MaskT m = (MaskT)(~bz & gmask), take = 0;
              for (int i = 0; i < take_n * hgpc; ++i) { MaskT bit = (MaskT)(m & (MaskT)(~m + 1)); take |= bit; m ^= bit; }
              bz |= take;
              mlo[nd * L] = (uint32_t)bz;
              if (sizeof(MaskT) == 8) mhi[nd * L] = (uint32_t)((unsigned long long)bz >> 32);
              if (!(mt & META_EVER)) ever += 1;
              meta[nd * L] = (mt - (uint32_t)(take_n * hgpc) - ((uint32_t)take_n << 16)) | META_EVER;
              gs_span sp; sp.node = nd; sp.ntasks = take_n; sp.devmask = (unsigned long long)take;
              spans[span_first + nspans] = sp;
              if (nspans == 0) { first_node = nd; mask0 = take; }
              ++nspans;
              rem -= take_n;
            }
            ok = true;
            evals += last + 1;
          } else {
            evals += M;
          }
        }
        if (ok) {
          // ---- commit: pop, network cost, start (algorithm.py:198-200, schedule.py:49-54,164-167)
          while (lo < M && META_IDLE(meta[lo * L]) == 0) ++lo;
          const int j = head;
          double dur2 = hj.dur;
          if (netcost && hj.ps > 1) {
            double mps = __ddiv_rn(S.model_mb[j], S.bandwidth);
            double nis = __dmul_rn((double)nspans, S.latency);
            double rt = __dmul_rn(S.iters[j], 2.0);
            dur2 = __dadd_rn(hj.dur, __dmul_rn(__dadd_rn(mps, nis), rt));
          }
          double eff = dur2 > hj.dur ? dur2 : hj.dur;
          double cl = ceil(eff);
          int need = cl < 1.0 ? 1 : (cl > 1.0e9 ? 0x7fffffff : (int)cl);
          if (need > wmask) { status = GS_ERR_ARG; need = wmask; }
          const int endt = delta + need;
          span_used += nspans;
          gs_job_rec r; r.start = delta; r.end = endt; r.jct = need; r.preempt = 1; r.duration = dur2;
          rec[j] = r;
          sref[j] = make_int2(span_first, nspans);
          const long long memc = (long long)hg * (hj.memb < cap_bytes ? hj.memb : cap_bytes);
          JobState js; js.next = -1; js.node0 = nspans == 1 ? first_node : span_first;
          js.mask0 = (unsigned long long)mask0; js.memc = memc; js.gpus = hg | ((nspans == 1 ? htasks : 0) << 24); js.cnt_gpc = nspans | (hgpc << 24);
          jst[j] = js;
          com_j = j; com_js = js;
          // append to the finish-tick bucket (start order): window / pending register / global wheel
          int tl;
          if (need <= LW - 1) {
            const int sl = (endt & (LW - 1)) * L;
            tl = (int)swt[sl];
            if (tl < 0) swh[sl] = (uint32_t)j;
            swt[sl] = (uint32_t)j;
          } else if (need == LW) {
            tl = pend_t;
            if (tl < 0) pend_h = j;
            pend_t = j;
          } else {
            const int gs_ = endt & wmask;
            tl = gwt[gs_];
            if (tl < 0) gwh[gs_] = j;
            gwt[gs_] = j;
          }
          if (tl >= 0) {
            jst[tl].next = j;
            if (tl == pre_h) pre_js.next = j;
          }
          top -= 1;
          sum_arr -= hj.arrive;
          running += 1; started += 1;
          busy_gpus += hg;
          mem_busy += memc;
          head = -1;
        }
      }
      // ---------------- D/E. release jobs whose finish tick is now
      const int now = delta + 1;
      {
        const int sl = (now & (LW - 1)) * L;
        int h = (int)swh[sl];
        if (h >= 0) {
          swh[sl] = 0xffffffffu; swt[sl] = 0xffffffffu;
          do {
            JobState js;
            if (h == pre_h) js = pre_js;
            else if (h == com_j) js = com_js;
            else js = jst[h];
            const int scnt = JS_CNT(js.cnt_gpc), sgpc = JS_GPC(js.cnt_gpc);
            if (scnt == 1) {
              const int nd = js.node0;
              mlo[nd * L] &= ~(uint32_t)js.mask0;
              if (sizeof(MaskT) == 8) mhi[nd * L] &= ~(uint32_t)(js.mask0 >> 32);
              meta[nd * L] += (uint32_t)JS_GPUS(js.gpus) + ((uint32_t)JS_NT0(js.gpus) << 16);
              if (nd < lo) lo = nd;
            } else {
              for (int i = 0; i < scnt; ++i) {
                gs_span sp = spans[js.node0 + i];
                mlo[sp.node * L] &= ~(uint32_t)sp.devmask;
                if (sizeof(MaskT) == 8) mhi[sp.node * L] &= ~(uint32_t)(sp.devmask >> 32);
                meta[sp.node * L] += (uint32_t)(sp.ntasks * sgpc) + ((uint32_t)sp.ntasks << 16);
                if (sp.node < lo) lo = sp.node;
              }
            }
            fin[finished] = h;
            finished += 1; running -= 1;
            busy_gpus -= JS_GPUS(js.gpus);
            mem_busy -= js.memc;
            h = js.next;
          } while (h >= 0);
        }
      }
      // ---------------- H. statistics row (schedule.py:95-133) from O(1) counters
      {
        int pmax = 0, mlo_p = 0, mhi_p = 0;
        if (top > 0) {
          const int ilo = top - 1 - (top - 1) / 2, ihi = top - 1 - top / 2;
          const int alo = ilo >= cache_lo ? (int)sstk[((ilo & (SCACHE - 1)) * 2 + 1) * L] : stack[ilo].y;
          const int ahi = ihi >= cache_lo ? (int)sstk[((ihi & (SCACHE - 1)) * 2 + 1) * L] : stack[ihi].y;
          pmax = now - bottom_arr; mlo_p = now - alo; mhi_p = now - ahi;
        }
        int4 *dst = reinterpret_cast<int4 *>(&rows[ticks - row_first]);
        const int tg = M * G;
        const long long ps = top > 0 ? (long long)top * now - sum_arr : 0;
        dst[0] = make_int4(now, M - ever, ever, busy_gpus);
        dst[1] = make_int4(tg - busy_gpus, running, top, finished);
        dst[2] = make_int4((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32), (int)(ps & 0xffffffffLL), (int)(ps >> 32));
        dst[3] = make_int4(pmax, mlo_p, mhi_p, 0);
      }
      // ---------------- pipeline stage: retire last iteration's loads, issue the next ones
      {
        // bucket of tick delta + LW enters the window (its slot held tick delta, consumed last iteration)
        const int sl = (delta & (LW - 1)) * L;
        swh[sl] = (uint32_t)pend_h; swt[sl] = (uint32_t)pend_t;
        const int gs_ = (delta + 1 + LW) & wmask;
        pend_h = gwh[gs_]; pend_t = gwt[gs_];
        gwh[gs_] = -1; gwt[gs_] = -1;
        // trace record ring: one record per tick
        if (rp_valid) {
          uint32_t *r = ring + (size_t)((pf & (RING - 1)) * 8) * L;
          r[0] = rp0.x; r[L] = rp0.y; r[2 * L] = rp0.z; r[3 * L] = rp0.w;
          r[4 * L] = rp1.x; r[5 * L] = rp1.y; r[6 * L] = rp1.z; r[7 * L] = rp1.w;
          ++pf;
        }
        // never run further ahead than RING - 2 past p (the 2 slots behind p keep the records of the
        // jobs admitted last, which are the ones popped next), never fall behind p
        rp_valid = false;
        if (pf < p) { pf = p; ring_lo = p; }         // a burst outran the ring: restart it at p
        if (pf < n && pf < p + RING - 2) {
          const uint4 *src = reinterpret_cast<const uint4 *>(&jobs[pf]);
          rp0 = src[0]; rp1 = src[1]; rp_valid = true;
        }
        // release record of the job that heads the bucket of tick now + 1
        pre_h = (int)swh[((now + 1) & (LW - 1)) * L];
        if (pre_h >= 0) pre_js = jst[pre_h];
      }
      ticks += 1; budget -= 1;
      delta = now;
      done = (n - p) + running == 0;      // schedule.py:185 -- the queue is NOT counted (Q4)
    }   // go
    __syncwarp();
  }

  if (!alive) return;
  // ---- persist: node table, wheel window and pending bucket go back to global memory
  for (int nd = 0; nd < M; ++nd) {
    uint32_t mt = meta[nd * L];
    unsigned long long bz = mlo[nd * L];
    if (sizeof(MaskT) == 8) bz |= (unsigned long long)mhi[nd * L] << 32;
    S.nbusy[nd] = bz;
    S.nk[nd] = (int)((uint32_t)(K - META_KFREE(mt)) | ((mt & META_EVER) ? EVER_BIT : 0u));
  }
  for (int t = delta + 1; t <= delta + LW - 1; ++t) {
    gwh[t & wmask] = (int)swh[(t & (LW - 1)) * L]; gwt[t & wmask] = (int)swt[(t & (LW - 1)) * L];
  }
  gwh[(delta + LW) & wmask] = pend_h; gwt[(delta + LW) & wmask] = pend_t;
  S.delta = delta; S.p = p; S.top = top; S.running = running; S.finished = finished;
  S.ever = ever; S.busy_gpus = busy_gpus; S.mem_busy = mem_busy; S.sum_arr = sum_arr;
  S.span_used = span_used; S.events = (long long)p + started + finished; S.evals = evals; S.started = started;
  S.ticks = ticks; S.row_first = row_first; S.done = done ? 1 : 0; S.status = status;
}

