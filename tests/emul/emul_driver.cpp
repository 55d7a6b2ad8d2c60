// emul_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the real kernel source (repeatafterme_b200/csrc/extend_kernel.cuh) and the real host
// flatten logic (csrc/flatten.h) with g++ against cuda_emul.h and exposes one flat C entry point,
// so that the CPU test-suite can diff the kernel's logic against the oracle without a GPU.
#include <pthread.h>
#include <stdio.h>
#include <vector>
#include <functional>

#include "cuda_emul.h"
#include "../../repeatafterme_b200/csrc/extend_kernel.cuh"
#include "../../repeatafterme_b200/csrc/aux_kernels.cuh"
#include "../../repeatafterme_b200/csrc/flatten.h"

thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;
thread_local EmulCta *emul_cta;
thread_local EmulWarp *emul_warp;

using namespace ramcu;

typedef void (*extend_fn)(const FamDev *, const int *, int *);
typedef void (*pack_fn)(const uint8_t *, unsigned long long, uint32_t *, uint32_t *, unsigned long long);

struct LaunchArgs
{
  int tid, bid, block, grid;
  EmulCta *cta;
  extend_fn fn;
  const FamDev *fams;
  const int *cta_fam;
  int *ticket;
  // pack
  pack_fn pfn;
  const uint8_t *bytes;
  unsigned long long len, n_chunks;
  uint32_t *lib2, *libn;
  // any other kernel, bound with its arguments
  const std::function<void()> *gen;
};

static void *thread_main(void *p)
{
  LaunchArgs *a = (LaunchArgs *) p;
  threadIdx = {(unsigned) a->tid, 0, 0};
  blockIdx = {(unsigned) a->bid, 0, 0};
  blockDim = {(unsigned) a->block, 1, 1};
  gridDim = {(unsigned) a->grid, 1, 1};
  emul_cta = a->cta;
  emul_warp = &a->cta->warps[a->tid >> 5];
  if (a->gen)
    (*a->gen)();
  else if (a->fn)
    a->fn(a->fams, a->cta_fam, a->ticket);
  else
    a->pfn(a->bytes, a->len, a->lib2, a->libn, a->n_chunks);
  return nullptr;
}

static int launch(LaunchArgs proto, int grid, int block)
{
  std::vector<EmulCta> ctas(grid);
  std::vector<std::vector<EmulWarp>> warps(grid);
  const int nw = (block + 31) / 32;
  for (int b = 0; b < grid; ++b)
  {
    warps[b].resize(nw);
    for (int w = 0; w < nw; ++w)
      pthread_barrier_init(&warps[b][w].bar, nullptr, std::min(32, block - 32 * w));
    pthread_barrier_init(&ctas[b].bar, nullptr, block);
    ctas[b].warps = warps[b].data();
  }
  std::vector<LaunchArgs> args((size_t) grid * block, proto);
  std::vector<pthread_t> th((size_t) grid * block);
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, 512 * 1024);
  for (int b = 0; b < grid; ++b)
    for (int t = 0; t < block; ++t)
    {
      LaunchArgs &a = args[(size_t) b * block + t];
      a.tid = t;
      a.bid = b;
      a.block = block;
      a.grid = grid;
      a.cta = &ctas[b];
      if (pthread_create(&th[(size_t) b * block + t], &at, thread_main, &a) != 0)
        return -1;
    }
  for (auto &t : th)
    pthread_join(t, nullptr);
  return 0;
}

struct Cfg { int C, G, WS; extend_fn fn, fn2, fnx; };   // lean single pass, two passes, single pass with extras
#define ECFG(C_, G_, WS_) \
  {C_, G_, WS_, extend_kernel<C_, G_, WS_, 1024, 1, false>, extend_kernel<C_, G_, WS_, 1024, 2, false>, \
   extend_kernel<C_, G_, WS_, 1024, 1, true>}
static const Cfg kCfgs[] = {
  ECFG(11, 1, 11), ECFG(6, 2, 11), ECFG(3, 4, 11), ECFG(4, 3, 11),      // B = 5, static band
  ECFG(29, 1, 29), ECFG(15, 2, 29), ECFG(8, 4, 29), ECFG(4, 8, 29),     // B = 14
  ECFG(21, 4, 81), ECFG(11, 8, 81), ECFG(41, 2, 81), ECFG(27, 3, 81), ECFG(17, 5, 81), ECFG(81, 1, 81),   // B = 40
  ECFG(10, 3, 29), ECFG(5, 6, 29),
  ECFG(4, 4, 0), ECFG(8, 4, 0), ECFG(16, 4, 0), ECFG(32, 4, 0), ECFG(16, 16, 0), ECFG(16, 32, 0), ECFG(8, 5, 0), ECFG(16, 7, 0),
};

extern "C" int emul_num_cfgs(void) { return (int) (sizeof(kCfgs) / sizeof(kCfgs[0])); }
extern "C" void emul_cfg(int i, int *C, int *G, int *WS) { *C = kCfgs[i].C; *G = kCfgs[i].G; *WS = kCfgs[i].WS; }

// Same contract as ramcu_extend() / ramcu_debug_rows() (include/ramcu.h), plus an explicit band
// tiling (C, G, WS) and CTA size. dbg_* and forced may be NULL.
static int emul_extend_impl(int C, int G, int WS, int threads, const uint8_t *seq, uint64_t seq_len,
                            int direction, int n_cores, const ramcu_core_t *cores,
                            const ramcu_scoring_t *sc, const ramcu_params_t *pr,
                            uint8_t *master_out, int32_t *ext_len_out, int32_t *score_add_out,
                            ramcu_result_t *res, int dbg_nrows, const uint8_t *forced,
                            int32_t *rows_out, int32_t *best_out, int32_t *sums_out,
                            uint8_t *path_out, int32_t *pbest_out, int32_t *pidx_out,
                            int verbose_rows = 0, int32_t *cand_out = nullptr, int32_t *vsums_out = nullptr,
                            int32_t *cells_out = nullptr)
{
  const Cfg *cfg = nullptr;
  for (const Cfg &c : kCfgs)
    if (c.C == C && c.G == G && c.WS == WS)
      cfg = &c;
  const int W = 2 * pr->bandwidth + 1;
  if (!cfg || (WS && WS != W) || C * G < W || threads % 32 || threads > 1024)
    return -2;

  // library packing through the real pack kernel
  const unsigned long long n_chunks = (seq_len + 31) / 32;
  std::vector<uint32_t> a2(n_chunks * 2 + 8, 0), an(n_chunks + 8, 0xffffffffu);
  {
    LaunchArgs p = {};
    p.pfn = pack_library_kernel;
    p.bytes = seq;
    p.len = seq_len;
    p.n_chunks = n_chunks;
    p.lib2 = a2.data() + 4;
    p.libn = an.data() + 4;
    if (launch(p, 2, 32))
      return -1;
  }

  std::vector<InstDev> inst;
  flatten_cores(direction ? 1 : 0, n_cores, cores, seq_len, inst);
  const int n_ext = (int) inst.size();
  memset(res, 0, sizeof(*res));
  res->n_extendable = n_ext;
  if (n_ext == 0 || pr->L <= 0)
    return -3;
  ramcu_params_t p2 = *pr;
  if (dbg_nrows > 0)
  {
    p2.L = dbg_nrows;
    p2.when_to_stop = 1 << 30;
  }
  // `threads` is the largest CTA the caller wants; CTAs get equal shares of the instances
  const int per_cta_max = std::max(1, (threads / 32) * inst_per_warp(G));
  const int n_ctas = (n_ext + per_cta_max - 1) / per_cta_max;
  threads = cta_threads(n_ext, n_ctas, G);
  std::vector<int> orig(n_ext), thigh(n_ext), tpos(n_ext), result(8, 0), cta_fam(n_ctas, 0);
  for (int i = 0; i < n_ext; ++i)
    orig[i] = inst[i].orig;
  std::vector<unsigned char> master(p2.L + 1);
  std::vector<unsigned long long> slots((size_t) (p2.L + 1) * 4 * kRedStride, 0);
  std::vector<int> drows, dbest, dsums;
  FamDev f;
  fill_family(f, sc, &p2, n_ext, n_ctas);
  f.lib2 = a2.data() + 4;
  f.libn = an.data() + 4;
  f.lib_len = (long long) seq_len;
  f.inst = inst.data();
  f.master = master.data();
  f.thigh = thigh.data();
  f.tpos = tpos.data();
  f.result = result.data();
  f.slots = slots.data();
  if (dbg_nrows > 0)
  {
    drows.assign((size_t) dbg_nrows * n_ext * W * 2, 0);
    dbest.assign((size_t) dbg_nrows * n_ext * 4, 0);
    dsums.assign((size_t) dbg_nrows * 4, 0);
    f.dbg_rows = drows.data();
    f.dbg_best = dbest.data();
    f.dbg_sums = dsums.data();
    f.forced = forced;
    f.dbg_nrows = dbg_nrows;
  }
  std::vector<unsigned char> dpath;
  std::vector<int> didx;
  if (path_out)   // -outmat mode (ramcu_extend_paths)
  {
    dpath.assign((size_t) p2.L * n_ext * W, 0);
    didx.assign((size_t) p2.L * n_ext, 0);
    dbest.assign((size_t) p2.L * n_ext * 4, 0);
    dsums.assign((size_t) p2.L * 4, 0);
    f.dbg_path = dpath.data();
    f.dbg_idx = didx.data();
    f.dbg_best = dbest.data();
    f.dbg_sums = dsums.data();
    f.dbg_nrows = p2.L;
  }
  std::vector<int> dcand, dcells;
  if (verbose_rows > 0)   // verbose dumps (ramcu_extend_verbose)
  {
    dcand.assign((size_t) verbose_rows * n_ext * 12, 0);
    dbest.assign((size_t) verbose_rows * n_ext * 4, 0);
    dsums.assign((size_t) verbose_rows * 4, 0);
    f.dbg_cand = dcand.data();
    f.dbg_best = dbest.data();
    f.dbg_sums = dsums.data();
    f.dbg_nrows = verbose_rows;
    if (cells_out)
    {
      dcells.assign((size_t) verbose_rows * n_ext * 4 * W * 2, 0);
      f.dbg_cells = dcells.data();
    }
  }
  int ticket = 0;
  LaunchArgs a = {};
  // as in ramcu.cu: the lean instantiation unless a diagnostic / per-row output is asked for
  a.fn = (f.forced || f.dbg_rows || f.dbg_best || f.dbg_path || f.seed_rows || f.dbg_cand) ? cfg->fnx : cfg->fn;
  a.fams = &f;
  a.cta_fam = cta_fam.data();
  a.ticket = &ticket;
  if (launch(a, n_ctas, threads))
    return -1;

  res->extended_bp = result[0] + 1;
  res->rows_processed = result[1];
  res->stop_row = result[2];
  res->max_score = result[3];
  res->hit_limit = (result[2] >= 0 && result[2] == p2.L - 1) ? 1 : 0;
  res->lanes_per_instance = G;
  if (master_out)
    memcpy(master_out, master.data(), result[1]);
  map_results(n_cores, orig, thigh.data(), tpos.data(), ext_len_out, score_add_out);
  if (path_out)
    for (int r = 0; r < result[1]; ++r)
      for (int e = 0; e < n_ext; ++e)
      {
        memcpy(path_out + ((size_t) r * n_cores + orig[e]) * W, dpath.data() + ((size_t) r * n_ext + e) * W, W);
        pbest_out[(size_t) r * n_cores + orig[e]] = dbest[((size_t) r * n_ext + e) * 4 + (master[r] & 3)];
        pidx_out[(size_t) r * n_cores + orig[e]] = didx[(size_t) r * n_ext + e];
      }
  for (int r = 0; r < std::min(verbose_rows, result[1]); ++r)
  {
    for (int e = 0; e < n_ext; ++e)
    {
      for (int a4 = 0; a4 < 4; ++a4)
      {
        int32_t *o = cand_out + (((size_t) r * n_cores + orig[e]) * 4 + a4) * 4;
        o[0] = dbest[((size_t) r * n_ext + e) * 4 + a4];
        for (int q = 0; q < 3; ++q)
          o[1 + q] = dcand[(((size_t) r * n_ext + e) * 4 + a4) * 3 + q];
      }
      if (cells_out)
        memcpy(cells_out + ((size_t) r * n_cores + orig[e]) * 4 * W * 2, dcells.data() + ((size_t) r * n_ext + e) * 4 * W * 2,
               (size_t) 4 * W * 2 * sizeof(int));
    }
    memcpy(vsums_out + (size_t) r * 4, dsums.data() + (size_t) r * 4, 4 * sizeof(int));
  }
  for (int r = 0; r < dbg_nrows; ++r)
  {
    for (int e = 0; e < n_ext; ++e)
    {
      if (rows_out)
        memcpy(rows_out + ((size_t) r * n_cores + orig[e]) * W * 2,
               drows.data() + ((size_t) r * n_ext + e) * W * 2, (size_t) W * 2 * sizeof(int));
      if (best_out)
        memcpy(best_out + ((size_t) r * n_cores + orig[e]) * 4,
               dbest.data() + ((size_t) r * n_ext + e) * 4, 4 * sizeof(int));
    }
    if (sums_out)
      memcpy(sums_out + (size_t) r * 4, dsums.data() + (size_t) r * 4, 4 * sizeof(int));
  }
  return 0;
}

extern "C" int emul_extend(int C, int G, int WS, int threads, const uint8_t *seq, uint64_t seq_len,
                           int direction, int n_cores, const ramcu_core_t *cores,
                           const ramcu_scoring_t *sc, const ramcu_params_t *pr,
                           uint8_t *master_out, int32_t *ext_len_out, int32_t *score_add_out,
                           ramcu_result_t *res, int dbg_nrows, const uint8_t *forced,
                           int32_t *rows_out, int32_t *best_out, int32_t *sums_out)
{
  return emul_extend_impl(C, G, WS, threads, seq, seq_len, direction, n_cores, cores, sc, pr, master_out,
                          ext_len_out, score_add_out, res, dbg_nrows, forced, rows_out, best_out, sums_out,
                          nullptr, nullptr, nullptr);
}

// Same contract as ramcu_extend_paths() (include/ramcu.h)
extern "C" int emul_extend_paths(int C, int G, int WS, int threads, const uint8_t *seq, uint64_t seq_len,
                                 int direction, int n_cores, const ramcu_core_t *cores,
                                 const ramcu_scoring_t *sc, const ramcu_params_t *pr,
                                 uint8_t *master_out, int32_t *ext_len_out, int32_t *score_add_out,
                                 ramcu_result_t *res, uint8_t *path_out, int32_t *pbest_out, int32_t *pidx_out)
{
  return emul_extend_impl(C, G, WS, threads, seq, seq_len, direction, n_cores, cores, sc, pr, master_out,
                          ext_len_out, score_add_out, res, 0, nullptr, nullptr, nullptr, nullptr, path_out,
                          pbest_out, pidx_out);
}

// Same contract as ramcu_extend_verbose() (include/ramcu.h)
extern "C" int emul_extend_verbose(int C, int G, int WS, int threads, const uint8_t *seq, uint64_t seq_len,
                                   int direction, int n_cores, const ramcu_core_t *cores,
                                   const ramcu_scoring_t *sc, const ramcu_params_t *pr,
                                   uint8_t *master_out, int32_t *ext_len_out, int32_t *score_add_out,
                                   ramcu_result_t *res, int n_rows, int32_t *cand_out, int32_t *sums_out,
                                   int32_t *cells_out)
{
  return emul_extend_impl(C, G, WS, threads, seq, seq_len, direction, n_cores, cores, sc, pr, master_out,
                          ext_len_out, score_add_out, res, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, n_rows, cand_out, sums_out, cells_out);
}

// Same contract as the device part of ramcu_library_from_twobit() (include/ramcu.h): 2bit packed DNA
// + window list -> lib2 words (twobit_windows_kernel + twobit_boundary_kernel). lib2_out holds
// 2 * ceil(total / 32) words, zeroed by the caller.
extern "C" int emul_twobit(const uint8_t *blob, unsigned long long blob_bytes, const ramcu_window_t *windows,
                           int n_windows, unsigned long long total, uint32_t *lib2_out)
{
  static_assert(sizeof(ramcu_window_t) == sizeof(WindowDev), "window layout");
  const WindowDev *win = reinterpret_cast<const WindowDev *>(windows);
  const unsigned long long n_chunks = (total + 31) / 32;
  // the fast path needs a 16-byte aligned blob with whole 16-byte loads readable
  std::vector<uint8_t> store(blob_bytes + 64, 0);
  uint8_t *ab = store.data() + ((16 - (reinterpret_cast<uintptr_t>(store.data()) & 15)) & 15);
  memcpy(ab, blob, blob_bytes);
  std::vector<unsigned long long> slow((size_t) n_windows + 4, 0);
  unsigned int nslow = 0;
  {
    const std::function<void()> k = [&]() {
      twobit_windows_kernel(ab, blob_bytes, win, n_windows, total, lib2_out, n_chunks, slow.data(), &nslow, (unsigned) slow.size());
    };
    LaunchArgs a = {};
    a.gen = &k;
    if (launch(a, 3, 64))
      return -1;
  }
  {
    const std::function<void()> k = [&]() {
      twobit_boundary_kernel(ab, blob_bytes, win, n_windows, total, lib2_out, n_chunks, slow.data(), &nslow, (unsigned) slow.size());
    };
    LaunchArgs a = {};
    a.gen = &k;
    if (launch(a, 2, 32))
      return -1;
  }
  return (int) nslow;
}

// Same contract as ramcu_extend_pair() (include/ramcu.h): both passes of a family in the same threads
// (extend_kernel<..., NP = 2>), the right pass as launch unit, the left pass as its mate.
extern "C" int emul_extend_pair(int C, int G, int WS, int threads, const uint8_t *seq, uint64_t seq_len,
                                int n_cores, const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                const ramcu_params_t *pr, uint8_t *master_r, int32_t *ext_len_r,
                                int32_t *score_add_r, ramcu_result_t *res_r, uint8_t *master_l,
                                int32_t *ext_len_l, int32_t *score_add_l, ramcu_result_t *res_l)
{
  const Cfg *cfg = nullptr;
  for (const Cfg &c : kCfgs)
    if (c.C == C && c.G == G && c.WS == WS)
      cfg = &c;
  const int W = 2 * pr->bandwidth + 1;
  if (!cfg || (WS && WS != W) || C * G < W || threads % 32 || threads > 1024)
    return -2;
  const unsigned long long n_chunks = (seq_len + 31) / 32;
  std::vector<uint32_t> a2(n_chunks * 2 + 8, 0), an(n_chunks + 8, 0xffffffffu);
  {
    LaunchArgs p = {};
    p.pfn = pack_library_kernel;
    p.bytes = seq;
    p.len = seq_len;
    p.n_chunks = n_chunks;
    p.lib2 = a2.data() + 4;
    p.libn = an.data() + 4;
    if (launch(p, 2, 32))
      return -1;
  }
  std::vector<InstDev> inst[2];
  std::vector<int> orig[2], thigh[2], tpos[2], result[2];
  std::vector<unsigned char> master[2];
  std::vector<unsigned long long> slots[2];
  FamDev f[2];
  int n_ext_max = 0;
  for (int p = 0; p < 2; ++p)
  {
    flatten_cores(p == 0 ? 1 : 0, n_cores, cores, seq_len, inst[p]);
    n_ext_max = std::max(n_ext_max, (int) inst[p].size());
    if (inst[p].empty() || pr->L <= 0)
      return -3;
  }
  const int per_cta_max = std::max(1, (threads / 32) * inst_per_warp(G));
  const int n_ctas = (n_ext_max + per_cta_max - 1) / per_cta_max;
  threads = cta_threads(n_ext_max, n_ctas, G);
  for (int p = 0; p < 2; ++p)
  {
    const int n_ext = (int) inst[p].size();
    orig[p].resize(n_ext);
    for (int i = 0; i < n_ext; ++i)
      orig[p][i] = inst[p][i].orig;
    thigh[p].assign(n_ext, 0);
    tpos[p].assign(n_ext, 0);
    result[p].assign(8, 0);
    master[p].assign(pr->L + 1, 0);
    slots[p].assign((size_t) (pr->L + 1) * 4 * kRedStride, 0);
    fill_family(f[p], sc, pr, n_ext, n_ctas);
    f[p].lib2 = a2.data() + 4;
    f[p].lib_len = (long long) seq_len;
    f[p].libn = an.data() + 4;
    f[p].inst = inst[p].data();
    f[p].master = master[p].data();
    f[p].thigh = thigh[p].data();
    f[p].tpos = tpos[p].data();
    f[p].result = result[p].data();
    f[p].slots = slots[p].data();
  }
  f[0].mate = 1;
  std::vector<int> cta_fam(n_ctas, 0);
  int ticket = 0;
  LaunchArgs a = {};
  a.fn = cfg->fn2;
  a.fams = f;
  a.cta_fam = cta_fam.data();
  a.ticket = &ticket;
  if (launch(a, n_ctas, threads))
    return -1;
  uint8_t *mo[2] = {master_r, master_l};
  int32_t *eo[2] = {ext_len_r, ext_len_l}, *so[2] = {score_add_r, score_add_l};
  ramcu_result_t *ro[2] = {res_r, res_l};
  for (int p = 0; p < 2; ++p)
  {
    memset(ro[p], 0, sizeof(ramcu_result_t));
    ro[p]->n_extendable = (int) inst[p].size();
    ro[p]->extended_bp = result[p][0] + 1;
    ro[p]->rows_processed = result[p][1];
    ro[p]->stop_row = result[p][2];
    ro[p]->max_score = result[p][3];
    ro[p]->hit_limit = (result[p][2] >= 0 && result[p][2] == pr->L - 1) ? 1 : 0;
    ro[p]->lanes_per_instance = G;
    memcpy(mo[p], master[p].data(), result[p][1]);
    map_results(n_cores, orig[p], thigh[p].data(), tpos[p].data(), eo[p], so[p]);
  }
  return 0;
}

// One seed-extension pass (ramcu_seed_extend_batch's unit of work, ramcu.cu): a one-instance family with a
// given consensus (values 0..3, 4 = N) of `rows` rows; returns compute_nw_row()'s value and sequence index
// per row (FamDev::seed_rows).
extern "C" int emul_seed_rows(int C, int G, int WS, const uint8_t *seq, uint64_t seq_len, int direction,
                              const ramcu_core_t *core, const ramcu_scoring_t *sc, const int32_t *n_row, int bandwidth,
                              int rows, const uint8_t *forced, int32_t *rows_out)
{
  const Cfg *cfg = nullptr;
  for (const Cfg &c : kCfgs)
    if (c.C == C && c.G == G && c.WS == WS)
      cfg = &c;
  const int W = 2 * bandwidth + 1;
  if (!cfg || (WS && WS != W) || C * G < W)
    return -2;
  const unsigned long long n_chunks = (seq_len + 31) / 32;
  std::vector<uint32_t> a2(n_chunks * 2 + 8, 0), an(n_chunks + 8, 0xffffffffu);
  {
    LaunchArgs p = {};
    p.pfn = pack_library_kernel;
    p.bytes = seq;
    p.len = seq_len;
    p.n_chunks = n_chunks;
    p.lib2 = a2.data() + 4;
    p.libn = an.data() + 4;
    if (launch(p, 2, 32))
      return -1;
  }
  std::vector<InstDev> inst;
  flatten_cores(direction ? 1 : 0, 1, core, seq_len, inst);
  if (inst.size() != 1 || rows <= 0)
    return -3;
  ramcu_params_t pr;
  memset(&pr, 0, sizeof(pr));
  pr.bandwidth = bandwidth;
  pr.L = rows;
  pr.when_to_stop = 1 << 30;
  std::vector<int> thigh(1), tpos(1), result(8, 0), cta_fam(1, 0), srows((size_t) 2 * rows, 0);
  std::vector<unsigned char> master(rows + 1);
  std::vector<unsigned long long> slots((size_t) (rows + 1) * 4 * kRedStride, 0);
  FamDev f;
  fill_family(f, sc, &pr, 1, 1);
  for (int code = 0; code < RAMCU_NCODES; ++code)
    f.mN[code] = n_row[code];
  f.lib2 = a2.data() + 4;
  f.libn = an.data() + 4;
  f.lib_len = (long long) seq_len;
  f.inst = inst.data();
  f.master = master.data();
  f.thigh = thigh.data();
  f.tpos = tpos.data();
  f.result = result.data();
  f.slots = slots.data();
  f.forced = forced;
  f.seed_rows = srows.data();
  int ticket = 0;
  LaunchArgs a = {};
  a.fn = cfg->fnx;
  a.fams = &f;
  a.cta_fam = cta_fam.data();
  a.ticket = &ticket;
  if (launch(a, 1, 32))
    return -1;
  memcpy(rows_out, srows.data(), srows.size() * sizeof(int));
  return 0;
}
