// gs_logcol.cpp -- host side of the log writer: the one stochastic cluster.csv column (avg_gpu_utilization).
//
// The reference draws, on every tick, one np.random.normal(loc, scale, size=1) per busy device, walking the nodes in id
// order and the devices 0..G-1 (infra/device.py:48-54, core/scheduling/schedule.py:103-120), clips each at 100 and adds
// them up left to right.  The draws come from numpy's global legacy stream, which is sequential by construction, so the
// column is rebuilt on the host AFTER the run from where every job ran (the engine's span records) and when: the caller
// (gpuschedule_b200/rngcol.py) draws the standard-normal values from numpy and this file walks the ticks, keeps the set
// of busy devices as a bitmap + owner table, and consumes the values in the reference's order.  No device work here.
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "gsched.h"

struct gs_logcol_s {
  int64_t n_rows = 0, n_hold = 0;
  int32_t width = 0;
  std::vector<int64_t> first, last;
  std::vector<int32_t> key, job;
  std::vector<int64_t> end_head, end_next;     // holdings whose last counted row is r - 1, chained per row r
  std::vector<int32_t> owner;                  // job on each device (valid where the bitmap is set)
  std::vector<uint64_t> busy;
  int64_t next_start = 0, row = 0;
};

extern "C" gs_logcol gs_logcol_open(int64_t n_rows, int32_t width, int64_t n_hold, const int64_t *first, const int64_t *last,
                                    const int32_t *key, const int32_t *job) {
  if (n_rows < 0 || width <= 0 || n_hold < 0 || (n_hold > 0 && (!first || !last || !key || !job))) return nullptr;
  gs_logcol_s *c = new (std::nothrow) gs_logcol_s;
  if (!c) return nullptr;
  try {
    c->n_rows = n_rows; c->width = width;
    c->end_head.assign((size_t)n_rows + 1, -1);
    c->owner.assign((size_t)width, -1);
    c->busy.assign(((size_t)width + 63) / 64, 0);
    int64_t prev = -1;
    for (int64_t i = 0; i < n_hold; ++i) {
      if (first[i] < prev || key[i] < 0 || key[i] >= width) { delete c; return nullptr; }     // sorted by first row, devices in range
      prev = first[i];
      if (first[i] >= n_rows || last[i] < first[i] || first[i] < 0) continue;                  // never counted
      const int64_t lz = last[i] < n_rows ? last[i] : n_rows - 1;
      c->first.push_back(first[i]); c->last.push_back(lz); c->key.push_back(key[i]); c->job.push_back(job[i]);
    }
    c->n_hold = (int64_t)c->first.size();
    c->end_next.assign((size_t)c->n_hold, -1);
    for (int64_t i = 0; i < c->n_hold; ++i) { const int64_t r = c->last[i] + 1; c->end_next[i] = c->end_head[r]; c->end_head[r] = i; }
  } catch (...) { delete c; return nullptr; }
  return c;
}

extern "C" void gs_logcol_close(gs_logcol c) { delete c; }

// busy devices per row (= values the row consumes), rows 0 .. n_rows-1
extern "C" int gs_logcol_counts(gs_logcol c, int64_t *counts) {
  if (!c || (!counts && c->n_rows > 0)) return GS_ERR_ARG;
  try {
    std::vector<int64_t> diff((size_t)c->n_rows + 1, 0);
    for (int64_t i = 0; i < c->n_hold; ++i) { diff[c->first[i]] += 1; diff[c->last[i] + 1] -= 1; }
    int64_t run = 0;
    for (int64_t r = 0; r < c->n_rows; ++r) { run += diff[r]; counts[r] = run; }
  } catch (...) { return GS_ERR_STATE; }      // host allocation failed: nothing may be thrown across the C ABI
  return GS_OK;
}

// Rows [the current row, r_end): value = loc[job] + scale[job] * z (two roundings, like numpy's legacy normal), clipped at
// 100, summed left to right in device order.  acc / unclipped are indexed from the current row.  z must hold exactly
// the values those rows consume (gs_logcol_counts).
extern "C" int gs_logcol_rows(gs_logcol c, int64_t r_end, const double *loc, const double *scale, const double *z, int64_t n_z,
                              double *acc, int32_t *unclipped) {
  if (!c || r_end < c->row || r_end > c->n_rows || !loc || !scale || (n_z > 0 && !z)) return GS_ERR_ARG;
  const int words = (int)c->busy.size();
  int64_t p = 0;
  for (int64_t r = c->row; r < r_end; ++r) {
    for (int64_t i = c->end_head[r]; i >= 0; i = c->end_next[i]) {
      const int32_t k = c->key[i];
      if (c->owner[k] == c->job[i]) { c->busy[k >> 6] &= ~(1ull << (k & 63)); c->owner[k] = -1; }
    }
    while (c->next_start < c->n_hold && c->first[c->next_start] == r) {
      const int64_t i = c->next_start++;
      const int32_t k = c->key[i];
      c->owner[k] = c->job[i]; c->busy[k >> 6] |= 1ull << (k & 63);
    }
    double a = 0.0; int32_t nu = 0;
    for (int w = 0; w < words; ++w) {
      uint64_t m = c->busy[w];
      while (m) {
        const int k = (w << 6) + __builtin_ctzll(m);
        m &= m - 1;
        if (p >= n_z) return GS_ERR_CAPACITY;
        const int32_t j = c->owner[k];
        const double x = loc[j] + scale[j] * z[p++];
        if (x >= 100.0) a += 100.0; else { a += x; nu += 1; }
      }
    }
    acc[r - c->row] = a; unclipped[r - c->row] = nu;
  }
  c->row = r_end;
  return p == n_z ? GS_OK : GS_ERR_ARG;
}
