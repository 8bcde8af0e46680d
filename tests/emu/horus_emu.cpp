// TEST INFRASTRUCTURE -- host build of the device functions in gpuschedule_b200/csrc/gs_horus_core.cuh.
//
// The horus / gandiva engine is written as plain scalar __host__ __device__ functions; this shim compiles
// the SAME header with g++ (GS_HD = inline) so that the logic the CUDA kernel runs can be checked against
// the pinned oracle and the reference fixtures on a box without a GPU (tests/test_horus_emu.py).  It is
// built into tests/emu/_build/ by the test itself, is never loaded by the package, and is not a CPU
// fallback: gs_horus_create still fails without a CUDA device.
#include <cstdlib>
#include <cstring>
#include <vector>

#define GS_HD static inline
#include "gs_horus_core.cuh"
#include "gs_horus_host.h"

extern "C" long long emu_run_horus(const gs_cluster *c, const gs_horus_params *par, long long n, const int *arrive,
                                   const int *gpus, const int *gpc, const double *duration, const long long *mem_bytes,
                                   const double *util_avg, const double *util_max, const double *mem_avg_mib,
                                   const double *gauss, long long gauss_n, const unsigned int *words, long long words_n,
                                   gs_tick_row *rows, double *util, unsigned char *util_arr, long long rows_cap,
                                   gs_horus_job_rec *recs, int *fin, long long *nfin, long long *events, long long *draws,
                                   long long max_ticks_per_call, int cooperative) {
  const int M = c->num_switch * c->num_node_p_switch, G = c->num_gpu_p_node;
  const size_t N = n ? (size_t)n : 1;
  std::vector<HJob> jobs(N);
  long long first = 0; int maxg = 1;
  for (long long j = 0; j < n; ++j) {
    HJob &o = jobs[(size_t)j];
    o.arrive = arrive[j]; o.gpus = gpus[j]; o.gpc = gpc[j]; o.ntasks = gpus[j] / gpc[j]; o.first_task = (int)first; o.pad = 0;
    o.mem_b = mem_bytes[j]; o.util_avg = util_avg[j]; o.util_max = util_max[j]; o.duration = duration[j];
    o.mem_avg_mib = mem_avg_mib ? mem_avg_mib[j] : 0.0;
    first += o.ntasks; if (gpus[j] > maxg) maxg = gpus[j];
  }
  const size_t NT = first ? (size_t)first : 1;
  std::vector<HJobState> js(N); std::vector<HTask> tasks(NT); std::vector<int> tron(NT, -1), troo(NT, 0);
  std::vector<HNode> nodes((size_t)M); std::vector<HDev> devs((size_t)M * G);
  const int pjw = (M + 63) / 64;
  std::vector<unsigned long long> pj(N * (size_t)pjw, 0);
  const int nq = par->schedule == GS_HSCHED_HORUS_PLUS ? (par->num_queue > 0 ? par->num_queue : 1) : 1;
  const size_t nb = (size_t)(par->num_buffer > 0 ? par->num_buffer : 1);
  std::vector<int> queue((N + 1) * (size_t)nq), running(N), finv(N), look(nb), lookq(nb), work(N), res((size_t)M);
  std::vector<int> kall(N), kas(N), kold(N); std::vector<double> ksc(N);
  std::vector<int> sccnt((size_t)M * G), scoff((size_t)M * G); std::vector<double> sccost((size_t)M * G);
  std::vector<double> gret, gkeep; std::vector<int> gnext, gacc, grank; int cls_off[5] = {0, 0, 0, 0, 0};
  if (words) {                        // what gs_horus_load_words does on the host side of the library
    const size_t W = words_n ? (size_t)words_n : 1;
    gret.resize(W); gkeep.resize(W); gnext.resize(W);
    gs_horus_build_gauss_tables(words, words_n, gret.data(), gkeep.data(), gnext.data());
    gacc.resize(W); grank.resize(W);
    gs_horus_build_gauss_index(gnext.data(), words_n, gacc.data(), grank.data(), cls_off);
  }
  std::vector<int> mn((size_t)maxg * maxg), mo((size_t)maxg * maxg), mc((size_t)maxg), ok((size_t)maxg), di((size_t)maxg);
  std::vector<HCand> heap((size_t)maxg + 2); std::vector<long long> mskip((size_t)maxg);
  memset(js.data(), 0, sizeof(HJobState) * N); memset(nodes.data(), 0, sizeof(HNode) * (size_t)M); memset(devs.data(), 0, sizeof(HDev) * (size_t)M * G);
  memset(tasks.data(), 0, sizeof(HTask) * NT);
  gs_horus_init_tasks(jobs.data(), n, (long long)c->gpu_mem_cap_mib << 20, tasks.data());
  HSim s; memset(&s, 0, sizeof(s));
  s.M = M; s.G = G; s.S = c->num_switch; s.P = c->num_node_p_switch; s.cpu_cap = c->num_cpu_p_node; s.mem_cap = c->mem_p_node;
  s.placement = par->placement;
  s.scheme = par->score; s.schedule = par->schedule; s.num_buffer = par->num_buffer; s.n = (int)n; s.maxg = maxg; s.pjw = pjw;
  s.cap_b = (long long)c->gpu_mem_cap_mib << 20;
  s.jobs = jobs.data(); s.js = js.data(); s.tasks = tasks.data(); s.tro_node = tron.data(); s.tro_order = troo.data();
  s.nodes = nodes.data(); s.devs = devs.data(); s.pj_bits = pj.data(); s.queue = queue.data(); s.running = running.data(); s.fin = finv.data();
  s.nq = nq; s.look_q = lookq.data(); s.km_all = kall.data(); s.km_assign = kas.data(); s.km_old = kold.data(); s.km_score = ksc.data();
  if (words) {
    s.words = words; s.words_n = words_n; s.gv_ret = gret.data(); s.gv_keep = gkeep.data(); s.gv_next = gnext.data();
    s.gv_acc = gacc.data(); s.gv_rank = grank.data(); for (int c = 0; c < 5; ++c) s.gv_cls_off[c] = cls_off[c];
  }
  s.sc_cnt = sccnt.data(); s.sc_off = scoff.data(); s.sc_cost = sccost.data();
  s.look = look.data(); s.work = work.data(); s.res_nodes = res.data(); s.map_node = mn.data(); s.map_order = mo.data(); s.map_n = mc.data();
  s.ok = ok.data(); s.distinct = di.data(); s.heap = heap.data(); s.map_skip = mskip.data();
  s.gauss = gauss; s.gauss_n = gauss_n; s.gauss_pos = 0;
  s.rows = rows; s.util = util; s.util_arr = util_arr; s.recs = recs; s.rows_cap = rows_cap;
  s.current_remaining = n; s.running_jobs = 0;
  while (!s.done && s.status == 0) {                                    // > 0: exercises the resume-between-launches path
    if (cooperative) h_run_coop(s, max_ticks_per_call); else h_run(s, max_ticks_per_call);
  }
  for (int i = 0; i < s.nfin; ++i) fin[i] = finv[(size_t)i];
  *nfin = s.nfin; *events = s.events; *draws = s.draws;
  return s.status < 0 ? s.status : s.ticks;
}
