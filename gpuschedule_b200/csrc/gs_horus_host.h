// Host-side helper of the utilisation-aware engine: numpy's legacy normal sampler over a raw MT19937 word stream.
//
// numpy.random.normal(loc, scale) = loc + scale * legacy_gauss(); legacy_gauss is the polar (Marsaglia) method:
// two 53-bit uniforms from four 32-bit words (a >> 5, b >> 6), rejected unless 0 < r2 < 1, then
// f = sqrt(-2 log(r2) / r2); it returns f * x2 and keeps f * x1 for the next call.  Under horus+ the k-means
// integer draws (one word each) interleave with the samples, so the word a sample starts at is only known
// while the simulation runs.  The device has no libm-identical log(), so the host tabulates, for EVERY start
// position p, what a call starting there returns (ret), keeps (keep) and where the stream continues (next);
// the kernel then only looks values up.  Built with the host's libm -- the library numpy itself calls.
#pragma once
#include <math.h>
#include <stdint.h>

static inline void gs_horus_build_gauss_tables(const uint32_t *w, long long n, double *ret, double *keep, int *next) {
  for (long long p = n - 1; p >= 0; --p) {
    ret[p] = 0.0; keep[p] = 0.0; next[p] = -1;                       // not enough words left: the run reports GS_ERR_CAPACITY
    if (p + 4 > n) continue;
    const int32_t a1 = (int32_t)(w[p] >> 5), b1 = (int32_t)(w[p + 1] >> 6), a2 = (int32_t)(w[p + 2] >> 5), b2 = (int32_t)(w[p + 3] >> 6);
    const double d1 = (a1 * 67108864.0 + b1) / 9007199254740992.0, d2 = (a2 * 67108864.0 + b2) / 9007199254740992.0;
    const double x1 = 2.0 * d1 - 1.0, x2 = 2.0 * d2 - 1.0;
    const double r2 = x1 * x1 + x2 * x2;
    if (r2 >= 1.0 || r2 == 0.0) {                                    // rejected: the loop tries again four words on
      if (p + 4 < n) { ret[p] = ret[p + 4]; keep[p] = keep[p + 4]; next[p] = next[p + 4]; }
      continue;
    }
    const double f = sqrt(-2.0 * log(r2) / r2);
    keep[p] = f * x1; ret[p] = f * x2; next[p] = (int)(p + 4);
  }
}

// Index over the same tables for the warp-cooperative scoring.  An attempt advances four words, so from any start
// position the polar method only ever visits positions of one residue class mod 4.  acc lists, class by class
// (cls_off[c] .. cls_off[c + 1]), the accepted start positions in increasing order; rank[p] = how many accepted
// positions of p's class lie before p.  "The m-th pair drawn from start position p" is then acc[cls_off[p & 3] +
// rank[p] + m]: one indexed load instead of a walk along next[].
static inline void gs_horus_build_gauss_index(const int *next, long long n, int *acc, int *rank, int cls_off[5]) {
  long long w = 0;
  for (int c = 0; c < 4; ++c) {
    cls_off[c] = (int)w;
    int seen = 0;
    for (long long p = c; p < n; p += 4) {
      rank[p] = seen;
      if (next[p] == (int)(p + 4)) { acc[w++] = (int)p; ++seen; }     // accepted exactly here (not inherited from p + 4)
    }
  }
  cls_off[4] = (int)w;
}

// Task records of a freshly loaded trace (one per (job, worker)); shared by the library and by the host build of the
// device functions in tests/emu, so that both start from the same bytes.
template <class Job, class Task>
static inline void gs_horus_init_tasks(const Job *jobs, long long n, long long cap_b, Task *tasks) {
  for (long long j = 0; j < n; ++j)
    for (int k = 0; k < jobs[j].ntasks; ++k) {
      Task &t = tasks[jobs[j].first_task + k];
      t.duration = t.original = jobs[j].duration;
      t.util_avg = jobs[j].util_avg;
      t.half_spread = (jobs[j].util_max - jobs[j].util_avg) / 2;
      t.quarter_spread = (jobs[j].util_max - jobs[j].util_avg) / 4;
      t.mem_clamped = jobs[j].mem_b < cap_b ? jobs[j].mem_b : cap_b;
      t.job = (int)j; t.time_processed = 0; t.placed_node = -1; t.run_node = -1;
      t.interfered = t.running = t.finished = t.pad = 0; t.pad2 = 0;
    }
}
