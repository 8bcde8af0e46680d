// gs_tick2.cuh -- part of libgsched.so (single translation unit, included from gsched.cu).
// fifo + yarn tick engine, one warp per replica, event stepped.
//
// What the reference does every tick (Scheduler.start, core/scheduling/schedule.py:185-209) is
// restated here per EVENT tick: a tick on which nothing arrives, nothing can start and nothing
// finishes changes no counter of the statistics row (schedule.py:95-133) except the ones that are
// linear in the tick number, so such ticks are jumped over in one step and leave no record.  The
// device therefore emits
//   * one 24-byte gs_evrow per tick on which a counter changed (plus the first tick of a launch),
//   * one 24-byte gs_qrow beside it while the queue is non-empty (arrival-tick sum, oldest arrival,
//     the two middle arrivals -- the pending statistics of jobs_manager.py:72-87 are `now - arrival`), and
//   * one 8-byte gs_nodeev whenever the count of nodes that ever hosted a job grows (node.py:93-97),
// from which gs_expand_rows_kernel (or the host) rebuilds every gs_tick_row bit for bit.  Per job it writes the start
// tick (4 bytes; the rest of job.csv follows from the trace) and per (job, node) an 8-byte gs_cspan.
//
// Per-tick path, in the order of the reference loop:
//   A  admit arrivals  (jobs_manager.py:228-241; head insert, quirk Q2): a 32-record register window of
//      the trace; the batch's first job stays in registers as the queue head, the rest go to the stack
//   B  one attempt on the head  (schedule.py:40-60, algorithm.py:189-202,301-417): ballot first fit /
//      prefix-sum cross-node fill over the shared-memory node table; a head that did not fit is not
//      tried again until a completion or a new head could change the answer (the reference re-tries
//      every tick with the same outcome; the evaluation counter is advanced in closed form)
//   E  release the jobs whose finish tick is now  (schedule.py:141-162): timing wheel keyed by finish
//      tick; the next non-empty bucket, its first job and that job's release record are kept in
//      registers, so the common release has no dependent global load
//   H  the record(s)
#pragma once

#define SCACHE 4        // cached top-of-stack entries (power of two)

struct __align__(16) JobState2 {   // 16 B, written at start, read once at completion
  int next;                  // next job of the same finish-tick bucket (reverse start order)
  int where;                 // one span:  node (bits 0-19) | (tasks - 1) << 20;   several: bit 31 | first index in the span pool
  unsigned long long mask0;  // one span:  devices held;                            several: span count | gpus << 32
};

// Shared memory through 32-bit shared-window addresses held in registers: generic pointers make the compiler rebuild the
// window base (S2UR CgaCtaId / ULEA ...) at every access once registers are tight (ncu, profiles/r02_tick2_v1_ncu_full.txt)
__device__ __forceinline__ unsigned lds32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v)); }
__device__ __forceinline__ unsigned long long lds64(unsigned a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v)); }
__device__ __forceinline__ int2 ldsv2(unsigned a) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void stsv2(unsigned a, int2 v) { asm volatile("st.shared.v2.s32 [%0], {%1, %2};" :: "r"(a), "r"(v.x), "r"(v.y)); }

// Device masks: 32-bit words when a node has at most 32 GPUs (every BASELINE cluster), 64-bit otherwise
template <bool G64> struct MaskOps;
template <> struct MaskOps<true> {
  typedef unsigned long long T;
  typedef gs_span Span;                                     // 16-byte span records
  static __device__ __forceinline__ Span make_span(int node, int ntasks, bool first, T mask) {
    Span sp; sp.node = node; sp.ntasks = ntasks | (first ? (int)0x80000000 : 0); sp.devmask = mask; return sp;
  }
  static __device__ __forceinline__ void st_span(Span *p, const Span &v) { __stcs(reinterpret_cast<int4 *>(p), *reinterpret_cast<const int4 *>(&v)); }
  static __device__ __forceinline__ int span_node(const Span &v) { return v.node; }
  static __device__ __forceinline__ int span_ntasks(const Span &v) { return v.ntasks & 0x7fffffff; }
  static __device__ __forceinline__ T span_mask(const Span &v) { return v.devmask; }
  static __device__ __forceinline__ T ld(unsigned a) { return lds64(a); }
  static __device__ __forceinline__ void st(unsigned a, T v) { sts64(a, v); }
  static __device__ __forceinline__ int popc(T v) { return __popcll(v); }
};
template <> struct MaskOps<false> {
  typedef unsigned T;
  typedef gs_cspan Span;                                    // 8-byte span records
  static __device__ __forceinline__ Span make_span(int node, int ntasks, bool first, T mask) {
    Span sp; sp.where = (unsigned)node | ((unsigned)(ntasks - 1) << 20) | (first ? 0x80000000u : 0u); sp.devmask = mask; return sp;
  }
  static __device__ __forceinline__ void st_span(Span *p, const Span &v) { __stcs(reinterpret_cast<int2 *>(p), *reinterpret_cast<const int2 *>(&v)); }
  static __device__ __forceinline__ int span_node(const Span &v) { return (int)(v.where & 0xfffffu); }
  static __device__ __forceinline__ int span_ntasks(const Span &v) { return (int)((v.where >> 20) & 0x3fu) + 1; }
  static __device__ __forceinline__ T span_mask(const Span &v) { return v.devmask; }
  static __device__ __forceinline__ T ld(unsigned a) { return lds32(a); }
  static __device__ __forceinline__ void st(unsigned a, T v) { sts32(a, v); }
  static __device__ __forceinline__ int popc(T v) { return __popc(v); }
};

template <typename T>
__device__ __forceinline__ T take_lowest(T idle, int cnt) {
  // the `cnt` lowest set bits of `idle` (devices are claimed in index order, node.py:208-216)
  const T low = idle & (~idle + (T)1);              // lowest set bit
  if (cnt == 1) return low;
  // idle devices are mostly a contiguous run (claimed lowest first): `cnt` consecutive set bits starting at the lowest one
  const T run = (cnt >= (int)(8 * sizeof(T))) ? ~(T)0 : (low << cnt) - low;
  if ((idle & run) == run) return run;
  T m = idle;
  for (int i = 0; i < cnt && m; ++i) m &= m - (T)1;
  return idle ^ m;
}

__device__ __forceinline__ int meta_cap(unsigned mt, int gpc) {
  // tasks a node can still take: min(idle devices / gpus per task, free task slots)
  const int idle = (int)(mt & 0xffu), kfree = (int)(mt >> 16);
  return min(gpc == 1 ? idle : idle / gpc, kfree);
}

__device__ __forceinline__ int need_of(double dur) {     // quirk Q11: run length = max(1, ceil(duration)) ticks
  const double cl = ceil(dur);
  return cl < 1.0 ? 1 : (cl > 1.0e9 ? 0x7fffffff : (int)cl);
}

#define GS_INF 0x7fffffff

template <bool NET, bool G64>
__global__ void __launch_bounds__(32, GS_TICK_MINBLOCKS) gs_tick2_kernel(SimDev *sims, int nsims, long long max_ticks, int smem_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef MaskOps<G64> MO;
  typedef typename MO::T MaskT;
  unsigned lane_u;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_u));      // read once; the compiler would otherwise re-read the special register
  const int lane = (int)lane_u;
  const int sim = blockIdx.x;
  if (sim >= nsims) return;
  SimDev &S = sims[sim];
  if (S.done || S.status != 0 || S.policy != GS_SCHED_FIFO || (S.netcost != 0) != NET || (S.G > 32) != G64) return;

  const int M = S.M, G = S.G, n = S.n;
  // shared memory, per warp: busy[M] device masks (8 bytes each) | kk[M] idle devices (bits 0-7), ever placed (bit 8),
  // free task slots << 16 | sstk[SCACHE] top of the stack
  unsigned sb_busy = (unsigned)__cvta_generic_to_shared(smem_raw);
  asm volatile("" : "+r"(sb_busy));
  unsigned sb_kk = sb_busy + 8u * (unsigned)M;
  asm volatile("" : "+r"(sb_kk));
  const unsigned sb_stk = sb_kk + 4u * (unsigned)(M + (M & 1));   // 8-byte aligned
#define BUSY_A(nd_) (sb_busy + 8u * (unsigned)(nd_))
#define KK_A(nd_) (sb_kk + 4u * (unsigned)(nd_))
#define STK_A(i_) (sb_stk + 8u * (unsigned)((i_) & (SCACHE - 1)))

  const JobIn *__restrict__ jobs = S.jobs;
  int *jstart = S.jstart;
  JobState2 *jst = S.jst2;
  int2 *stack = reinterpret_cast<int2 *>(S.stack);
  int *fin = S.fin, *whead = S.wheel_head;
  long long *wmem = S.wheel_mem;
  typedef typename MO::Span SpanT;
  SpanT *spans = reinterpret_cast<SpanT *>(S.spans);
  int2 *rowA = reinterpret_cast<int2 *>(S.evrows), *rowB = reinterpret_cast<int2 *>(S.qrows);   // 24-byte records = 3 x 8 bytes
  int2 *nodeev = reinterpret_cast<int2 *>(S.nodeev);
  // everything a replica owns lives in global memory: let the compiler emit ld/st.global instead of generic accesses
  __builtin_assume(__isGlobal(jobs)); __builtin_assume(__isGlobal(jstart)); __builtin_assume(__isGlobal(jst));
  __builtin_assume(__isGlobal(stack)); __builtin_assume(__isGlobal(fin)); __builtin_assume(__isGlobal(whead));
  __builtin_assume(__isGlobal(wmem)); __builtin_assume(__isGlobal(spans)); __builtin_assume(__isGlobal(rowA));
  __builtin_assume(__isGlobal(rowB)); __builtin_assume(__isGlobal(nodeev));
  const int wmask = S.wheel_mask;
  const long long cap_bytes = S.cap_bytes, fit_limit = S.fit_limit;
  const int span_cap = (int)(S.span_cap > 0x7fffffffLL ? 0x7fffffffLL : S.span_cap);
  MaskT gmask = (G >= (int)(8 * sizeof(MaskT))) ? ~(MaskT)0 : (((MaskT)1 << G) - (MaskT)1);
  if constexpr (G64) asm volatile("" : "+l"(gmask));                // keep it in a register instead of rebuilding it at every use
  else asm volatile("" : "+r"(gmask));

  int delta = S.delta, p = S.p, top = S.top, running = S.running, finished = S.finished;
  int ever = S.ever, busy_gpus = S.busy_gpus, status = 0, blocked = S.blocked;
  int span_used = (int)S.span_used;
  long long mem_busy = S.mem_busy, sum_arr = S.sum_arr, evals = S.evals;
  const int delta0 = delta;
  const int capA = (int)(S.rows_cap > 0x7fffffffLL ? 0x7fffffffLL : S.rows_cap);
  const int capB = (int)(S.qrows_cap > 0x7fffffffLL ? 0x7fffffffLL : S.qrows_cap);
  int na = 0, nb = 0, nne = 0;
  int ever_rec = -1;                              // busy-node count of this launch's last node event (-1: none yet)
  long long budget_ll = max_ticks > 0 ? max_ticks : 0x7fffffffLL;
  if (budget_ll > 0x7fffffffLL - delta - 2) budget_ll = 0x7fffffffLL - delta - 2;
  const int t_end = delta + (int)(budget_ll > 0 ? budget_ll : 0);      // first tick this launch does NOT process

  // ---- stage the node table (a fresh replica starts idle), the wheel and the top of the stack
  if (S.need_init) {
    const int K = S.K;
    for (int i = lane; i < M; i += 32) { sts64(BUSY_A(i), 0ull); sts32(KK_A(i), (unsigned)G | ((unsigned)K << 16)); }
    for (int i = lane; i <= wmask; i += 32) { whead[i] = -1; wmem[i] = 0; }
  } else {
    const int K = S.K;
    for (int i = lane; i < M; i += 32) {
      const unsigned long long bz = S.nbusy[i];
      const unsigned kv = (unsigned)S.nk[i];
      sts64(BUSY_A(i), bz);
      sts32(KK_A(i), (unsigned)(G - __popcll(bz)) | ((kv & EVER_BIT) ? 0x100u : 0u) | ((unsigned)(K - (int)(kv & ~EVER_BIT)) << 16));
    }
  }
  int scount = top;                              // entries in the stack array; the head may live in registers instead
  for (int i = max(scount - SCACHE, 0) + lane; i < scount; i += 32) stsv2(STK_A(i), stack[i]);
  int cache_lo = max(scount - SCACHE, 0);        // stack entries [cache_lo, scount) are cached in shared memory
  int bottom_arr = (top > 0) ? stack[0].y : 0;
  __syncwarp();

  // ---- trace window: lane l holds record wbase + l, reduced to what the loop needs
  int wbase = p & ~31;
  int warr = GS_INF, wpk = 1 | (1 << 24), wneed = 1, wps = 0;
  long long wmemc = 0;
  double wdur = 0.0;
#define LOAD_WINDOW()                                                                         \
  do {                                                                                        \
    warr = GS_INF;                                                                            \
    if (wbase + lane < n) {                                                                   \
      JobIn r_;                                                                               \
      {                                                                                       \
        const int4 *src_ = reinterpret_cast<const int4 *>(&jobs[wbase + lane]);               \
        const int4 v0_ = __ldcs(src_), v1_ = __ldcs(src_ + 1);      /* streamed once */       \
        r_.arrive = v0_.x; r_.gpus = v0_.y; r_.gpc = v0_.z; r_.ps = v0_.w;                     \
        r_.memb = (long long)(((unsigned long long)(unsigned)v1_.y << 32) | (unsigned)v1_.x); \
        r_.dur = __hiloint2double(v1_.w, v1_.z);                                              \
      }                                                                                       \
      warr = r_.arrive; wpk = r_.gpus | (r_.gpc << 24);                                       \
      wneed = need_of(r_.dur); if (!(r_.memb < fit_limit)) wneed = -wneed;                    \
      wmemc = (long long)r_.gpus * (r_.memb < cap_bytes ? r_.memb : cap_bytes);               \
      if (NET) { wdur = r_.dur; wps = r_.ps; }                                                \
    }                                                                                         \
  } while (0)
  LOAD_WINDOW();
  int next_arr = GS_INF;
  if (p < n) next_arr = __shfl_sync(FULL, warr, p - wbase);

  // ---- queue head, kept in registers while it waits
  bool hvalid = false;
  int hjob = -1, harr = 0, hpk = 1 | (1 << 24), hneed = 1, hps = 0;
  long long hmemc = 0;
  double hdur = 0.0;
#define HEAD_FROM_WINDOW(src_)                                                                \
  do {                                                                                        \
    hpk = __shfl_sync(FULL, wpk, (src_)); hneed = __shfl_sync(FULL, wneed, (src_));           \
    hmemc = __shfl_sync(FULL, wmemc, (src_));                                                 \
    if (NET) { hdur = __longlong_as_double(__shfl_sync(FULL, __double_as_longlong(wdur), (src_))); hps = __shfl_sync(FULL, wps, (src_)); } \
  } while (0)

  // ---- next completion: tick, first job of that bucket and its release record stay in registers
  int next_fin = GS_INF, nf_head = -1;
  JobState2 nf_js; nf_js.next = -1; nf_js.where = 0; nf_js.mask0 = 0ull;
  if (running > 0) {
    int base = delta + 1;
    while (true) {
      const int ph = whead[(base + lane) & wmask];
      const unsigned b = __ballot_sync(FULL, ph >= 0);
      if (b) { const int pos = __ffs(b) - 1; next_fin = base + pos; nf_head = __shfl_sync(FULL, ph, pos); break; }
      base += 32;
    }
    nf_js = jst[nf_head];
  }

  bool done = (n == 0);
  bool force = true;                              // the first tick of a launch always leaves a record

  while (!done && status == 0) {
    if (na >= capA || nb >= capB) break;          // no room for the records of another event tick
    // ---------------- ticks on which nothing happens: jump
    if (!force && (top == 0 || blocked)) {
      const int t_next = min(next_arr, next_fin - 1);
      if (t_next > delta) {
        const int t_to = min(t_next, t_end);
        if (blocked) evals += (long long)(t_to - delta) * M;       // the reference re-tries the head every tick
        delta = t_to;
      }
    }
    if (delta >= t_end) break;
    bool changed = force;
    force = false;
    // ---------------- A. admit arrivals (gen_jobs + head insert)
    if (next_arr <= delta) {
      int cnt = 0, q = p;
      while (true) {
        const int idx = wbase + lane;
        const unsigned b = __ballot_sync(FULL, idx >= q && warr <= delta);
        const int c = __popc(b);
        if (c > 0 && cnt == 0) {
          if (hvalid) {      // the waiting head goes back under the new batch
            const int2 e = make_int2(hjob, harr);
            stack[scount] = e; stsv2(STK_A(scount), e);
            scount += 1;
          }
          HEAD_FROM_WINDOW(p - wbase);           // the batch's first job becomes the head (quirk Q2)
        }
        cnt += c; q += c;
        if (q < wbase + 32 || q >= n) break;
        wbase += 32;
        LOAD_WINDOW();
      }
      if (cnt > 0) {
        // jobs p+1 .. p+cnt-1 land under the new head, p+1 on top of them
        const int m = cnt - 1;
        for (int i = lane; i < m; i += 32) {
          const int2 e = make_int2(p + cnt - 1 - i, delta);
          stack[scount + i] = e;
          if (i >= m - SCACHE) stsv2(STK_A(scount + i), e);
        }
        scount += m;
        if (scount - cache_lo > SCACHE) cache_lo = scount - SCACHE;
        if (top == 0) bottom_arr = delta;
        hjob = p; harr = delta; hvalid = true;
        top += cnt; p += cnt;
        sum_arr += (long long)cnt * delta;
        blocked = 0; changed = true;
        __syncwarp();
      }
      next_arr = GS_INF;
      if (p < n) next_arr = __shfl_sync(FULL, warr, p - wbase);
    }
    // ---------------- B. one scheduling attempt on the queue head (quirks Q1, Q3)
    if (top > 0 && blocked) evals += M;           // the reference tries (and fails) on this tick too
    if (top > 0 && !blocked) {
      if (!hvalid) {         // the head was started or the launch just resumed: pop the stack
        int2 e;
        if (scount - 1 >= cache_lo) e = ldsv2(STK_A(scount - 1));
        else e = stack[scount - 1];
        scount -= 1;
        if (cache_lo > scount) cache_lo = scount;
        hjob = e.x; harr = e.y; hvalid = true;
        if (hjob >= wbase && hjob < wbase + 32) {
          HEAD_FROM_WINDOW(hjob - wbase);
        } else {
          const JobIn jr = jobs[hjob];
          hpk = jr.gpus | (jr.gpc << 24);
          hneed = need_of(jr.dur); if (!(jr.memb < fit_limit)) hneed = -hneed;
          hmemc = (long long)jr.gpus * (jr.memb < cap_bytes ? jr.memb : cap_bytes);
          if (NET) { hdur = jr.dur; hps = jr.ps; }
        }
      }
      const int hg = hpk & 0xffffff, hgpc = (int)((unsigned)hpk >> 24);
      const int htasks = hgpc == 1 ? hg : hg / hgpc;
      const bool placeable = hneed > 0;            // Device.can_fit on an empty device
      // the wheel bucket this job would finish in is known before the placement runs (without network costs): issue the
      // two loads of its current head / memory sum now, so that their L2 latency passes during the node scan
      int pre_bk = 0, pre_head = -1;
      long long pre_wm = 0;
      if (!NET && placeable) {
        pre_bk = (delta + min(hneed, wmask)) & wmask;
        pre_head = whead[pre_bk];
        pre_wm = wmem[pre_bk];
      }
      bool ok = false;
      int nspans = 0, where = 0;
      const int span_first = span_used;
      unsigned long long mask0 = 0;
      if (hg <= G) {
        // try_single_node_alloc_ms: first node (id order) that fits the whole job
        int found = -1;
        for (int base = 0; base < M; base += 32) {
          const int nd = base + lane;
          bool fit = false;
          if (nd < M) {
            const unsigned mt = lds32(KK_A(nd));
            fit = (int)(mt & 0xffu) >= hg && (int)(mt >> 16) >= htasks;
          }
          if (!placeable) {          // quirk Q21: cpu/mem charged for every task, never refunded
            if (fit) sts32(KK_A(nd), lds32(KK_A(nd)) - ((unsigned)htasks << 16));
            continue;
          }
          const unsigned b = __ballot_sync(FULL, fit);
          if (b) { found = base + __ffs(b) - 1; break; }
        }
        if (found >= 0 && span_used + 1 > span_cap) { status = GS_ERR_CAPACITY; found = -1; }
        if (found >= 0) {
          ok = true; nspans = 1;
          // warp-uniform update: every lane reads the same words and writes the same values (no lane-0 branch, no sync)
          const MaskT bz = MO::ld(BUSY_A(found));
          const MaskT take = take_lowest<MaskT>(~bz & gmask, hg);
          const unsigned kv = lds32(KK_A(found));
          MO::st(BUSY_A(found), bz | take);
          sts32(KK_A(found), (kv - (unsigned)hg - ((unsigned)htasks << 16)) | 0x100u);
          MO::st_span(&spans[span_first], MO::make_span(found, htasks, true, take));
          mask0 = take;
          where = found | ((htasks - 1) << 20);
          ever += (kv & 0x100u) ? 0 : 1;
          evals += found + 1;
        } else {
          evals += M;
        }
      } else {
        // try_cross_node_alloc_ms: walk nodes in id order, each takes what it can hold
        int cum = 0, last_base = -1;
        if (placeable) {
          for (int base = 0; base < M; base += 32) {
            const int nd = base + lane;
            const int c = (nd < M) ? meta_cap(lds32(KK_A(nd)), hgpc) : 0;
            cum += __reduce_add_sync(FULL, c);
            if (cum >= htasks) { last_base = base; break; }
          }
        } else {
          for (int base = 0; base < M; base += 32) {   // quirk Q21, cross-node flavour: one task charged per node
            const int nd = base + lane;
            if (nd < M) { const unsigned mt = lds32(KK_A(nd)); if (meta_cap(mt, hgpc) > 0) sts32(KK_A(nd), mt - (1u << 16)); }
          }
        }
        if (last_base >= 0 && (long long)span_used + min(htasks, M) > (long long)span_cap) { status = GS_ERR_CAPACITY; last_base = -1; }
        if (last_base >= 0) {
          // pass 1 proved the job fits: commit (a failed walk is rolled back exactly by the
          // reference, algorithm.py:378-387, so no state changes in that case)
          ok = true;
          int rem = htasks, last_node = 0;
          for (int base = 0; base <= last_base; base += 32) {
            const int nd = base + lane;
            const unsigned kv = (nd < M) ? lds32(KK_A(nd)) : 0u;
            const int c = meta_cap(kv, hgpc);
            int incl = c;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
            const int take = min(c, max(rem - (incl - c), 0));
            const unsigned tb = __ballot_sync(FULL, take > 0);
            bool fresh = false;
            if (take > 0) {
              const MaskT bz = MO::ld(BUSY_A(nd));
              const MaskT tk = take_lowest<MaskT>(~bz & gmask, take * hgpc);
              MO::st(BUSY_A(nd), bz | tk);
              sts32(KK_A(nd), (kv - (unsigned)(take * hgpc) - ((unsigned)take << 16)) | 0x100u);
              fresh = !(kv & 0x100u);
              const int slot = nspans + __popc(tb & ((1u << lane) - 1u));
              MO::st_span(&spans[span_first + slot], MO::make_span(nd, take, slot == 0, tk));
            }
            ever += __popc(__ballot_sync(FULL, fresh));
            if (tb) last_node = base + 31 - __clz(tb);
            nspans += __popc(tb);
            const int tot = __shfl_sync(FULL, incl, 31);
            rem -= min(rem, tot);
          }
          where = (int)0x80000000 | span_first;
          mask0 = (unsigned long long)(unsigned)nspans | ((unsigned long long)(unsigned)hg << 32);
          evals += last_node + 1;
          __syncwarp();        // lanes updated different nodes
        } else {
          evals += M;
          __syncwarp();        // (the unplaceable walk charges nodes lane by lane)
        }
      }
      if (ok) {
        // ---- commit: pop, network cost, start (algorithm.py:198-200, schedule.py:49-54,164-167)
        const int j = hjob;
        int need = hneed;
        if (NET) {
          double dur2 = hdur;
          if (hps > 1) {
            // (model_size/bandwidth + cross*latency) * (iterations*2.0), network_service.py:34-37
            const double mps = __ddiv_rn(S.model_mb[j], S.bandwidth);
            const double nis = __dmul_rn((double)nspans, S.latency);
            const double rt = __dmul_rn(S.iters[j], 2.0);
            dur2 = __dadd_rn(hdur, __dmul_rn(__dadd_rn(mps, nis), rt));
          }
          const double eff = dur2 > hdur ? dur2 : hdur;             // Job.get_duration (job.py:206-210)
          need = need_of(eff);
          S.dur2[j] = dur2;
        }
        if (need > wmask) { status = GS_ERR_ARG; need = wmask; }
        const int endt = delta + need;
        const int bk = endt & wmask;
        span_used += nspans;
        JobState2 js; js.where = where; js.mask0 = mask0; js.next = -1;
        // push on the finish-tick bucket of the timing wheel (released in start order, see E)
        if (endt < next_fin) { next_fin = endt; nf_head = j; nf_js = js; }
        else if (endt == next_fin) { js.next = nf_head; nf_head = j; nf_js = js; }
        else js.next = NET ? whead[bk] : pre_head;
        const long long wm = NET ? wmem[bk] : pre_wm;
        whead[bk] = j;                                   // warp-uniform stores (same address, same value from every lane)
        wmem[bk] = wm + hmemc;
        *reinterpret_cast<int4 *>(&jst[j]) = *reinterpret_cast<const int4 *>(&js);
        __stcs(&jstart[j], delta);
        top -= 1;
        sum_arr -= harr;
        running += 1;
        busy_gpus += hg;
        mem_busy += hmemc;
        hvalid = false; changed = true;
      } else if (placeable) {
        blocked = 1;           // nothing can change the outcome before a completion or a new head
      }
    }
    // ---------------- D/E. time advances; release jobs whose finish tick is now
    const int now = delta + 1;
    if (next_fin == now) {
      const int sl = now & wmask;
      int ph = whead[(now + 1 + lane) & wmask];      // start looking for the following bucket right away
      const long long wm = wmem[sl];
      int h = nf_head;
      JobState2 js = nf_js;
      const int f0 = finished;
      int c = 0;
      while (true) {
        if (js.where >= 0) {
          const int nd = js.where & 0xfffff, nt = ((js.where >> 20) & 63) + 1;
          const int gp = __popcll(js.mask0);
          const MaskT bz = MO::ld(BUSY_A(nd));           // warp-uniform read-modify-write
          const unsigned kv = lds32(KK_A(nd));
          MO::st(BUSY_A(nd), bz & ~(MaskT)js.mask0);
          sts32(KK_A(nd), kv + (unsigned)gp + ((unsigned)nt << 16));
          busy_gpus -= gp;
        } else {
          const int first = js.where & 0x7fffffff, scnt = (int)(unsigned)(js.mask0 & 0xffffffffull);
          for (int i = lane; i < scnt; i += 32) {
            const SpanT sp = spans[first + i];
            const int snd = MO::span_node(sp);
            const MaskT smk = MO::span_mask(sp);
            MO::st(BUSY_A(snd), MO::ld(BUSY_A(snd)) & ~smk);
            sts32(KK_A(snd), lds32(KK_A(snd)) + (unsigned)MO::popc(smk) + ((unsigned)MO::span_ntasks(sp) << 16));
          }
          busy_gpus -= (int)(unsigned)(js.mask0 >> 32);
          __syncwarp();          // lanes updated different nodes
        }
        fin[f0 + c] = h;
        c += 1;
        h = js.next;
        if (h < 0) break;
        js = jst[h];
      }
      finished += c; running -= c;
      mem_busy -= wm;
      whead[sl] = -1; wmem[sl] = 0;
      if (c >= 2) {          // the bucket was walked newest first; job.csv lists equal finish ticks in start order
        for (int i = lane; i < (c >> 1); i += 32) { const int a = fin[f0 + i], b2 = fin[f0 + c - 1 - i]; fin[f0 + i] = b2; fin[f0 + c - 1 - i] = a; }
        __syncwarp();
      }
      next_fin = GS_INF; nf_head = -1;
      if (running > 0) {
        int base = now + 1;
        while (true) {
          const unsigned b = __ballot_sync(FULL, ph >= 0);
          if (b) { const int pos = __ffs(b) - 1; next_fin = base + pos; nf_head = __shfl_sync(FULL, ph, pos); break; }
          base += 32;
          ph = whead[(base + lane) & wmask];
        }
        nf_js = jst[nf_head];
      }
      blocked = 0; changed = true;
    }
    // ---------------- H. statistics record (schedule.py:95-133) from O(1) counters
    if (changed) {
      if (top > 0) {
        // the queue is a stack with non-decreasing arrival ticks bottom->top, so the sorted pending
        // list is the stack read top->bottom: median/max are index look-ups
        const int ilo = top - 1 - ((top - 1) >> 1), ihi = top - 1 - (top >> 1);
        int a_lo, a_hi;
        if (hvalid && ilo == top - 1) a_lo = harr;
        else a_lo = ilo >= cache_lo ? ldsv2(STK_A(ilo)).y : stack[ilo].y;
        if (hvalid && ihi == top - 1) a_hi = harr;
        else a_hi = ihi >= cache_lo ? ldsv2(STK_A(ihi)).y : stack[ihi].y;
        {
          int2 *qb = rowB + 3 * nb;
          __stcs(qb, make_int2(now, bottom_arr));
          __stcs(qb + 1, make_int2(a_lo, a_hi));
          __stcs(qb + 2, make_int2((int)(sum_arr & 0xffffffffLL), (int)(sum_arr >> 32)));
        }
        nb += 1;
      }
      if (ever != ever_rec) { __stcs(&nodeev[nne], make_int2(now, ever)); nne += 1; ever_rec = ever; }
      {
        int2 *ab = rowA + 3 * na;
        __stcs(ab, make_int2(now, top));
        __stcs(ab + 1, make_int2(finished, busy_gpus | (running << 16)));
        __stcs(ab + 2, make_int2((int)(mem_busy & 0xffffffffLL), (int)(mem_busy >> 32)));
      }
      na += 1;
    }
    delta = now;
    done = (n - p) + running == 0;      // schedule.py:185 -- the queue is NOT counted (quirk Q4)
  }
#undef LOAD_WINDOW
#undef HEAD_FROM_WINDOW

  // ---------------- persist: the head goes back on the stack, node table back to global memory
  __syncwarp();
  if (hvalid) { if (lane == 0) stack[scount] = make_int2(hjob, harr); scount += 1; }
  __syncwarp();
  // jobs still queued have not started: their result record says so (start = -1), whatever ran before
  for (int i = lane; i < scount; i += 32) jstart[stack[i].x] = -1;
  {
    const int K = S.K;
    for (int i = lane; i < M; i += 32) {
      const unsigned mt = lds32(KK_A(i));
      S.nbusy[i] = (unsigned long long)MO::ld(BUSY_A(i));
      S.nk[i] = (int)((unsigned)(K - (int)(mt >> 16)) | ((mt & 0x100u) ? EVER_BIT : 0u));
    }
  }
  if (lane == 0) {
    S.delta = delta; S.p = p; S.top = top; S.running = running; S.finished = finished;
    S.ever = ever; S.busy_gpus = busy_gpus; S.mem_busy = mem_busy; S.sum_arr = sum_arr;
    S.span_used = span_used; S.started = (long long)finished + running;
    S.events = (long long)p + finished + running + finished; S.evals = evals;
    S.ticks = delta; S.row_first = delta0; S.nev = na; S.nq = nb; S.nne = nne; S.blocked = blocked;
    S.done = done ? 1 : 0; S.status = status; S.need_init = 0;
  }
#undef BUSY_A
#undef KK_A
#undef STK_A
}

// ------------------------------------------------------------------ record -> row expansion
// One thread per event record: it writes the gs_tick_row of its own tick and of every jumped tick up to
// the next record (on those only `now` and the pending statistics move, linearly in the tick number).
__global__ void gs_expand_rows_kernel(const SimDev *sims, int sim, int M, int G, gs_tick_row *__restrict__ out) {
  const SimDev &S = sims[sim];
  const int na = S.nev;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= na) return;
  const gs_evrow a = S.evrows[k];
  const int t_first = a.now;                                           // `now` of this record
  const int t_last = (k + 1 < na) ? S.evrows[k + 1].now - 1 : (int)S.ticks;   // last `now` this record covers
  const int queued = a.queued, busy_gpus = a.busy_gpus, running = a.running;
  int nodes = 0;
  {                       // last node event at or before this record (the first record of a window always has one)
    int lo = 0, hi = S.nne - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S.nodeev[mid].now <= t_first) lo = mid; else hi = mid - 1; }
    if (S.nne > 0) nodes = S.nodeev[lo].busy_nodes;
  }
  long long sum_arr = 0; int bottom = 0, a_lo = 0, a_hi = 0;
  if (queued > 0) {       // the queue record taken on the same tick
    int lo = 0, hi = S.nq - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (S.qrows[mid].now < t_first) lo = mid + 1; else hi = mid; }
    const gs_qrow b = S.qrows[lo];
    sum_arr = b.arrive_sum; bottom = b.oldest_arrive; a_lo = b.med_lo_arrive; a_hi = b.med_hi_arrive;
  }
  const long long row0 = S.row_first;                                  // tick index of the window's first row
  for (int v = t_first; v <= t_last; ++v) {
    int4 *dst = reinterpret_cast<int4 *>(&out[(long long)v - 1 - row0]);
    const long long ps = queued > 0 ? (long long)queued * v - sum_arr : 0;
    dst[0] = make_int4(v, M - nodes, nodes, busy_gpus);
    dst[1] = make_int4(M * G - busy_gpus, running, queued, a.finished);
    dst[2] = make_int4((int)(a.mem_busy_bytes & 0xffffffffLL), (int)(a.mem_busy_bytes >> 32), (int)(ps & 0xffffffffLL), (int)(ps >> 32));
    dst[3] = queued > 0 ? make_int4(v - bottom, v - a_lo, v - a_hi, 0) : make_int4(0, 0, 0, 0);
  }
}

// Legacy gs_job_rec view of the compact per-job result (the start tick): fifo never preempts, so the run length is
// max(1, ceil(job.duration)) (quirk Q11), end = start + run length, jct = run length, preempt (migration_count) = 1 (Q12).
__global__ void gs_expand_jobs_kernel(const SimDev *sims, int sim, gs_job_rec *__restrict__ out) {
  const SimDev &S = sims[sim];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S.n) return;
  gs_job_rec r; r.start = -1; r.end = -1; r.jct = 0; r.preempt = 0;
  r.duration = S.jobs[j].dur;
  if (j < S.p) {
    const int st = S.jstart[j];
    if (st >= 0) {
      const double dur_in = r.duration;
      if (S.netcost) r.duration = S.dur2[j];
      const int need = need_of(r.duration > dur_in ? r.duration : dur_in);      // Job.get_duration (job.py:206-210)
      r.start = st; r.end = st + need; r.jct = need; r.preempt = 1;
    }
  }
  out[j] = r;
}
