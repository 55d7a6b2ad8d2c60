// ramcu.cu -- extern "C" ABI (include/ramcu.h) over the sm_100a extension kernels.
//
// Host-side responsibilities (all cheap, O(N)): flatten the caller's cores into per-direction
// instance descriptors (geometry of bnw_extend.c:778-788,824-837), pick a band tiling
// (G lanes x C cells), size the grid so that every CTA of a family can be co-resident, launch,
// and map results back (ram_extend.c:1234-1247). No DP arithmetic happens on the host and there
// is no CPU fallback: without a CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ramcu.h"
#include "extend_kernel.cuh"
#include "aux_kernels.cuh"
#include "flatten.h"
#include "kernels.h"

using namespace ramcu;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                         \
  do                                                                                     \
  {                                                                                      \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(RAMCU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

// Runs its function when the scope is left unless it was dismissed: an early `return fail(...)` of CK() then
// releases what the entry point had allocated so far.
template <class F>
struct Unwind
{
  F f;
  bool armed = true;
  ~Unwind()
  {
    if (armed)
      f();
  }
};
template <class F>
static Unwind<F> unwind(F f)
{
  return Unwind<F>{f};
}

// ------------------------------------------------------------------------------------------
// kernel registry: band tilings
// ------------------------------------------------------------------------------------------
static int batch_threads_init()
{
  const char *e = getenv("RAMCU_BATCH_THREADS");   // tuning aid
  const int v = e ? atoi(e) : 0;
  return (v >= 32 && v <= 1024 && v % 32 == 0) ? v : 128;
}
static const int kBatchThreads = batch_threads_init();   // CTA size used when many families share a launch

// how a family is laid out on the grid
struct Plan
{
  int n_ctas;
  int threads;  // CTA size for a launch of this family alone
};

static std::vector<KernelCfg *> g_cfgs;
static void collect_cfgs()
{
  if (!g_cfgs.empty())
    return;
  const CfgList lists[] = {cfgs_b14(), cfgs_b20(), cfgs_b40(), cfgs_b60(), cfgs_b80(), cfgs_b100(), cfgs_generic()};
  for (const CfgList &l : lists)
    for (int i = 0; i < l.n; ++i)
      g_cfgs.push_back(&l.cfgs[i]);
}

// ------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------
constexpr int kTickets = 256;   // ring of launch tickets: a launch keeps its own word while it runs

struct ramcu_ctx
{
  int device;
  int sm_count, cc_major, cc_minor;
  cudaStream_t own_stream, stream;
  cudaEvent_t ev0, ev1;
  int *d_ticket;            // kTickets words
  unsigned ticket_next;
  unsigned long long launch_gen;   // ev0/ev1 belong to the jobs timed in this generation
  bool coop_ok;             // cudaDevAttrCooperativeLaunch
  std::map<size_t, std::vector<void *>> pinned_free;   // pinned result blocks, by size class
};

struct ramcu_library
{
  uint64_t length;
  uint32_t *d_lib2;  // padded by 2 words at both ends; pointer is to base 0
  uint32_t *d_libn;
  void *alloc2, *allocn;
};

struct ramcu_job
{
  ramcu_library *lib;
  int direction, n_cores, n_ext;
  ramcu_scoring_t sc;
  ramcu_params_t pr;
  const KernelCfg *cfg = nullptr;
  int n_ctas = 0;             // > 0 when there is anything to launch (== solo.n_ctas)
  Plan solo = {0, 0};         // layout for ramcu_job_run
  int batch_ctas = 0;         // layout for ramcu_batch_run (batch_threads() per CTA)
  int max_ctas = 0;           // size of d_cta_fam / d_dbg_wait
  // the family's other pass (ramcu_job_create_pair) and the layout of both passes in the same threads
  ramcu_job *mate = nullptr;
  const KernelCfg *pair_cfg = nullptr;
  Plan pair_solo = {0, 0};
  int pair_batch_ctas = 0;
  std::vector<int> orig;      // ext index -> core index
  FamDev fam;                 // host copy (device pointers inside)
  bool fam_dirty = false;     // d_fam no longer holds the solo layout
  // device
  InstDev *d_inst = nullptr;
  unsigned char *d_out = nullptr;   // [8 ints result][thigh n_ext][tpos n_ext][master L bytes]
  unsigned char *h_out = nullptr;   // pinned host copy, filled in stream order behind every launch
  size_t h_out_cap = 0;
  unsigned long long *d_slots = nullptr;
  FamDev *d_fam = nullptr;    // single-family launch
  int *d_cta_fam = nullptr;   // zeros
  unsigned char *d_forced = nullptr;
  int *d_dbg_rows = nullptr, *d_dbg_best = nullptr, *d_dbg_sums = nullptr;
  unsigned char *d_dbg_path = nullptr;   // -outmat mode (negative dbg_rows in job_create_impl)
  int *d_dbg_idx = nullptr;
  int *d_seed_rows = nullptr;   // seed extension: [L][n_ext][2]
  int *d_dbg_cand = nullptr, *d_dbg_cells = nullptr;   // verbose dumps (ramcu_extend_verbose)
  long long *d_dbg_wait = nullptr;   // RAMCU_STATS=1 diagnostics
  const KernelCfg *ran_cfg = nullptr;
  int ran_passes = 1;         // passes per thread of the launch that ran the job
  int stats_ctas = 0;
  bool ran = false, in_batch = false;
  // the job asks for a diagnostic / per-row output: it runs the instantiation that carries that code (cfg->fnx)
  bool extras = false;
  double plan_cost = 0;       // the cost model's figure for cfg (a column; of both passes when planned as a duo)
  unsigned long long timed_gen = 0;
};

// ------------------------------------------------------------------------------------------
// basic entry points
// ------------------------------------------------------------------------------------------
extern "C" void ramcu_destroy(ramcu_ctx *c);
extern "C" const char *ramcu_last_error(void) { return g_err; }
extern "C" int ramcu_abi_version(void) { return RAMCU_ABI_VERSION; }

extern "C" int ramcu_init(int device, ramcu_ctx **out)
{
  if (!out)
    return fail(RAMCU_ERR_ARG, "ramcu_init: out == NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(RAMCU_ERR_CUDA, "ramcu_init: no CUDA device (%s); there is no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n)
    return fail(RAMCU_ERR_ARG, "ramcu_init: device %d out of range (0..%d)", device, n - 1);
  CK(cudaSetDevice(device));
  ramcu_ctx *c = new ramcu_ctx();
  auto undo = unwind([&]() { ramcu_destroy(c); });
  c->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(cudaMalloc(&c->d_ticket, kTickets * sizeof(int)));
  c->ticket_next = 0;
  c->launch_gen = 0;
  {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    c->coop_ok = coop != 0;
  }
  {
    // per-call buffers come from the stream-ordered pool; keep freed memory cached in the pool
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  collect_cfgs();
  {
    // Is the image (built for sm_100a only) loadable here? One small kernel answers that; the band
    // tilings are loaded when a family first needs them (loading all ~100 of them costs 1.7 s, more than
    // a whole single-family run).
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, (const void *) int32_peak_kernel);
    if (e != cudaSuccess)
      return fail(RAMCU_ERR_CUDA,
                  "ramcu_init: kernel image not loadable on this device (built for sm_100a): %s",
                  cudaGetErrorString(e));
  }
  undo.armed = false;
  *out = c;
  return RAMCU_OK;
}

extern "C" void ramcu_destroy(ramcu_ctx *c)
{
  if (!c)
    return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->d_ticket);
  for (auto &kv : c->pinned_free)
    for (void *p : kv.second)
      cudaFreeHost(p);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->own_stream);
  delete c;
}

extern "C" int ramcu_set_stream(ramcu_ctx *c, void *s)
{
  if (!c)
    return fail(RAMCU_ERR_ARG, "ramcu_set_stream: ctx == NULL");
  c->stream = s ? (cudaStream_t) s : c->own_stream;
  return RAMCU_OK;
}

extern "C" int ramcu_device_info(ramcu_ctx *c, int32_t *sm, int32_t *maj, int32_t *min)
{
  if (!c)
    return fail(RAMCU_ERR_ARG, "ramcu_device_info: ctx == NULL");
  if (sm) *sm = c->sm_count;
  if (maj) *maj = c->cc_major;
  if (min) *min = c->cc_minor;
  return RAMCU_OK;
}

// Built-in tables: facts of score_system.c:207-376 (matrix[consensus][sequence]); N scores -1
// (score_system.c:384-395); "repeatscout" (score_system.c:91-172) has gapopen 0, gapextn = gap
// and every non-identical pair (N included) = mismatch.
extern "C" int ramcu_scoring_by_name(const char *name, int rs_match, int rs_mismatch, int rs_gap,
                                     ramcu_scoring_t *out)
{
  static const struct
  {
    const char *name;
    int go, ge;
    int m[4][4];
  } T[] = {
    {"14p43g", -33, -7, {{9, -18, -10, -21}, {-18, 11, -18, -7}, {-7, -18, 11, -18}, {-21, -10, -18, 9}}},
    {"18p43g", -30, -6, {{9, -15, -8, -18}, {-16, 10, -16, -5}, {-5, -16, 10, -16}, {-18, -8, -15, 9}}},
    {"20p43g", -28, -5, {{9, -15, -8, -17}, {-15, 10, -15, -4}, {-4, -15, 10, -15}, {-17, -8, -15, 9}}},
    {"25p43g", -25, -5, {{8, -13, -6, -15}, {-13, 9, -13, -2}, {-2, -13, 9, -13}, {-15, -6, -13, 8}}},
  };
  if (!name || !out)
    return fail(RAMCU_ERR_ARG, "ramcu_scoring_by_name: NULL argument");
  memset(out, 0, sizeof(*out));
  if (strcmp(name, "repeatscout") == 0)
  {
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < RAMCU_NCODES; ++b)
        out->matrix[a][b] = (a == b) ? rs_match : rs_mismatch;
    out->gapopen = 0;
    out->gapextn = rs_gap;
    return RAMCU_OK;
  }
  for (size_t i = 0; i < sizeof(T) / sizeof(T[0]); ++i)
    if (strcmp(name, T[i].name) == 0)
    {
      for (int a = 0; a < 4; ++a)
      {
        for (int b = 0; b < 4; ++b)
          out->matrix[a][b] = T[i].m[a][b];
        out->matrix[a][4] = -1;
      }
      out->gapopen = T[i].go;
      out->gapextn = T[i].ge;
      return RAMCU_OK;
    }
  return fail(RAMCU_ERR_ARG, "Matrix name not found: %s", name);
}

// Measured INT32 ALU peak (instruction-lanes per second) of this device: the roofline denominator
// of the extension kernel. Runs ~iters*16 fused add+max ops per thread on a grid
// that fills every SM.
extern "C" int ramcu_int32_peak(ramcu_ctx *c, double *gops, double *ms_out)
{
  if (!c || !gops)
    return fail(RAMCU_ERR_ARG, "ramcu_int32_peak: NULL argument");
  CK(cudaSetDevice(c->device));
  int *d = nullptr;
  CK(cudaMalloc(&d, sizeof(int)));
  const int iters = 20000, grid = c->sm_count * 8, block = 256;
  int32_peak_kernel<<<grid, block, 0, c->stream>>>(d, 100, 3);   // warm-up
  CK(cudaEventRecord(c->ev0, c->stream));
  int32_peak_kernel<<<grid, block, 0, c->stream>>>(d, iters, 3);
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  cudaFree(d);
  *gops = (double) grid * block * (double) iters * kPeakChains / (ms * 1e-3) / 1e9;
  if (ms_out)
    *ms_out = ms;
  return RAMCU_OK;
}

// ------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------
static int library_pack(ramcu_ctx *c, const uint8_t *d_bytes, uint64_t length, ramcu_library **out)
{
  ramcu_library *lib = new ramcu_library();
  auto undo = unwind([&]() { ramcu_library_free(c, lib); });
  lib->length = length;
  const uint64_t n_chunks = (length + 31) / 32;  // 32 bases per chunk
  const size_t pad = 4;                           // words of slack on both sides
  CK(cudaMallocAsync(&lib->alloc2, (n_chunks * 2 + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMallocAsync(&lib->allocn, (n_chunks + 2 * pad) * sizeof(uint32_t), c->stream));
  lib->d_lib2 = (uint32_t *) lib->alloc2 + pad;
  lib->d_libn = (uint32_t *) lib->allocn + pad;
  // the kernels write every word of [0, n_chunks); only the slack on both sides is initialised here
  // (bases: A, mask: N)
  CK(cudaMemsetAsync(lib->alloc2, 0, pad * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->d_lib2 + n_chunks * 2, 0, pad * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->allocn, 0xff, pad * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->d_libn + n_chunks, 0xff, pad * sizeof(uint32_t), c->stream));
  if (n_chunks > 0)
  {
    const unsigned long long want = (n_chunks + 255) / 256;
    const int grid = (int) std::min<unsigned long long>(want, (unsigned long long) c->sm_count * 16);
    pack_library_kernel<<<grid, 256, 0, c->stream>>>(d_bytes, length, lib->d_lib2, lib->d_libn,
                                                     n_chunks);
    CK(cudaGetLastError());
  }
  undo.armed = false;
  *out = lib;
  return RAMCU_OK;
}

extern "C" int ramcu_library_upload(ramcu_ctx *c, const uint8_t *seq, uint64_t length,
                                    ramcu_library **out)
{
  if (!c || !out || (!seq && length))
    return fail(RAMCU_ERR_ARG, "ramcu_library_upload: NULL argument");
  CK(cudaSetDevice(c->device));
  uint8_t *d_bytes = nullptr;
  *out = nullptr;
  auto undo = unwind([&]() {
    if (d_bytes)
      cudaFreeAsync(d_bytes, c->stream);
    if (*out)
      ramcu_library_free(c, *out);
    *out = nullptr;
  });
  CK(cudaMallocAsync(&d_bytes, length ? length : 1, c->stream));
  CK(cudaMemcpyAsync(d_bytes, seq, length, cudaMemcpyHostToDevice, c->stream));
  int rc = library_pack(c, d_bytes, length, out);
  if (rc)
    return rc;
  CK(cudaStreamSynchronize(c->stream));   // the caller may reuse / free its host buffer
  cudaFreeAsync(d_bytes, c->stream);
  undo.armed = false;
  return RAMCU_OK;
}

extern "C" int ramcu_library_from_device(ramcu_ctx *c, const uint8_t *d_seq, uint64_t length,
                                         ramcu_library **out)
{
  if (!c || !out || (!d_seq && length))
    return fail(RAMCU_ERR_ARG, "ramcu_library_from_device: NULL argument");
  CK(cudaSetDevice(c->device));
  return library_pack(c, d_seq, length, out);
}

static_assert(sizeof(ramcu_window_t) == sizeof(WindowDev), "window layout");

extern "C" int ramcu_library_from_twobit(ramcu_ctx *c, const uint8_t *blob, uint64_t blob_bytes,
                                         int n_windows, const ramcu_window_t *windows,
                                         int n_nranges, const uint64_t *n_start,
                                         const uint64_t *n_len, uint64_t total, ramcu_library **out)
{
  if (!c || !out || (!blob && blob_bytes) || (n_windows > 0 && !windows) || n_windows < 0 ||
      n_nranges < 0 || (n_nranges > 0 && (!n_start || !n_len)))
    return fail(RAMCU_ERR_ARG, "ramcu_library_from_twobit: invalid argument");
  uint64_t expect = 0;
  for (int i = 0; i < n_windows; ++i)
  {
    if (windows[i].dst != expect)
      return fail(RAMCU_ERR_ARG, "ramcu_library_from_twobit: windows must tile the library in order "
                  "(window %d starts at %llu, expected %llu)", i, (unsigned long long) windows[i].dst,
                  (unsigned long long) expect);
    if (windows[i].src_byte + (windows[i].src_base + windows[i].len + 3) / 4 > blob_bytes)
      return fail(RAMCU_ERR_ARG, "ramcu_library_from_twobit: window %d reads past the blob", i);
    expect += windows[i].len;
  }
  if (expect != total)
    return fail(RAMCU_ERR_ARG, "ramcu_library_from_twobit: windows cover %llu of %llu bases",
                (unsigned long long) expect, (unsigned long long) total);
  for (int i = 0; i < n_nranges; ++i)
    if (n_start[i] + n_len[i] > total)
      return fail(RAMCU_ERR_ARG, "ramcu_library_from_twobit: N range %d outside the library", i);
  CK(cudaSetDevice(c->device));
  ramcu_library *lib = new ramcu_library();
  lib->length = total;
  const uint64_t n_chunks = (total + 31) / 32;
  const size_t pad = 4;
  uint8_t *d_blob = nullptr;
  WindowDev *d_win = nullptr;
  unsigned long long *d_ns = nullptr, *d_nl = nullptr, *d_slow = nullptr;
  unsigned int *d_nslow = nullptr;
  auto undo = unwind([&]() {
    void *tmp[] = {d_blob, d_win, d_ns, d_nl, d_slow, d_nslow};
    for (void *t : tmp)
      if (t)
        cudaFreeAsync(t, c->stream);
    ramcu_library_free(c, lib);
  });
  CK(cudaMallocAsync(&lib->alloc2, (n_chunks * 2 + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMallocAsync(&lib->allocn, (n_chunks + 2 * pad) * sizeof(uint32_t), c->stream));
  lib->d_lib2 = (uint32_t *) lib->alloc2 + pad;
  lib->d_libn = (uint32_t *) lib->allocn + pad;
  CK(cudaMemsetAsync(lib->alloc2, 0, pad * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->d_lib2 + n_chunks * 2, 0, pad * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->allocn, 0xff, pad * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->d_libn + n_chunks, 0xff, pad * sizeof(uint32_t), c->stream));
  if (n_chunks > 0)
  {
    // every word of lib2 is written by the decode kernels; the mask starts empty and the N ranges
    // (and the positions past the library inside the last word) are OR-ed in
    CK(cudaMemsetAsync(lib->d_libn, 0, n_chunks * sizeof(uint32_t), c->stream));
    CK(cudaMallocAsync(&d_blob, blob_bytes ? blob_bytes : 1, c->stream));
    CK(cudaMemcpyAsync(d_blob, blob, blob_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMallocAsync(&d_win, (size_t) std::max(n_windows, 1) * sizeof(WindowDev), c->stream));
    CK(cudaMemcpyAsync(d_win, windows, (size_t) n_windows * sizeof(WindowDev), cudaMemcpyHostToDevice, c->stream));
    const unsigned long long want = ((n_chunks + 3) / 4 + 255) / 256;   // one thread per 128 bases
    const int grid = (int) std::min<unsigned long long>(want, (unsigned long long) c->sm_count * 16);
    // groups of 128 bases that are not inside one window: at most one per window boundary + the tail
    const unsigned int slow_cap = 2u * (unsigned int) n_windows + 8u;
    CK(cudaMallocAsync(&d_slow, (size_t) slow_cap * sizeof(unsigned long long), c->stream));
    CK(cudaMallocAsync(&d_nslow, sizeof(unsigned int), c->stream));
    CK(cudaMemsetAsync(d_nslow, 0, sizeof(unsigned int), c->stream));
    twobit_windows_kernel<<<grid, 256, 0, c->stream>>>(d_blob, blob_bytes, d_win, n_windows, total,
                                                       lib->d_lib2, n_chunks, d_slow, d_nslow, slow_cap);
    CK(cudaGetLastError());
    {
      const unsigned long long wantb = (4ull * slow_cap + 127) / 128;
      const int gridb = (int) std::min<unsigned long long>(wantb, (unsigned long long) c->sm_count * 16);
      twobit_boundary_kernel<<<gridb, 128, 0, c->stream>>>(d_blob, blob_bytes, d_win, n_windows, total,
                                                           lib->d_lib2, n_chunks, d_slow, d_nslow, slow_cap);
      CK(cudaGetLastError());
    }
    CK(cudaFreeAsync(d_slow, c->stream));
    d_slow = nullptr;
    CK(cudaFreeAsync(d_nslow, c->stream));
    d_nslow = nullptr;
    // positions past `total` inside the last chunk are N
    if (total & 31)
    {
      nmask_one_range_kernel<<<1, 32, 0, c->stream>>>(total, 32 - (total & 31), lib->d_libn);
      CK(cudaGetLastError());
    }
    if (n_nranges > 0)
    {
      CK(cudaMallocAsync(&d_ns, (size_t) n_nranges * sizeof(unsigned long long), c->stream));
      CK(cudaMallocAsync(&d_nl, (size_t) n_nranges * sizeof(unsigned long long), c->stream));
      CK(cudaMemcpyAsync(d_ns, n_start, (size_t) n_nranges * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(d_nl, n_len, (size_t) n_nranges * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
      nmask_ranges_kernel<<<n_nranges, 128, 0, c->stream>>>(d_ns, d_nl, n_nranges, lib->d_libn);
      CK(cudaGetLastError());
      CK(cudaFreeAsync(d_ns, c->stream));
      d_ns = nullptr;
      CK(cudaFreeAsync(d_nl, c->stream));
      d_nl = nullptr;
    }
    CK(cudaStreamSynchronize(c->stream));   // the caller may reuse its host buffers
    CK(cudaFreeAsync(d_blob, c->stream));
    d_blob = nullptr;
    CK(cudaFreeAsync(d_win, c->stream));
    d_win = nullptr;
  }
  undo.armed = false;
  *out = lib;
  return RAMCU_OK;
}

extern "C" int ramcu_library_download(ramcu_ctx *c, ramcu_library *lib, uint32_t *lib2_out, uint32_t *libn_out)
{
  if (!c || !lib)
    return fail(RAMCU_ERR_ARG, "ramcu_library_download: NULL argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const uint64_t n_chunks = (lib->length + 31) / 32;
  if (lib2_out)
    CK(cudaMemcpy(lib2_out, lib->d_lib2, n_chunks * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (libn_out)
    CK(cudaMemcpy(libn_out, lib->d_libn, n_chunks * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return RAMCU_OK;
}

extern "C" int ramcu_library_free(ramcu_ctx *c, ramcu_library *lib)
{
  if (!lib)
    return RAMCU_OK;
  if (c)
  {
    cudaSetDevice(c->device);
    cudaFreeAsync(lib->alloc2, c->stream);
    cudaFreeAsync(lib->allocn, c->stream);
  }
  else
  {
    cudaFree(lib->alloc2);
    cudaFree(lib->allocn);
  }
  delete lib;
  return RAMCU_OK;
}

// ------------------------------------------------------------------------------------------
// jobs
// ------------------------------------------------------------------------------------------
// np (kernel variant): 0 = one pass per CTA, 1 = both passes of a family in the same threads, 2 = one pass per
// CTA with the diagnostic / per-row outputs compiled in (jobs with `extras`)
static extend_fn fn_of(const KernelCfg *k, int np) { return np == 2 ? k->fnx : k->fn[np]; }
static int maxt_of(const KernelCfg *k, int np) { return k->maxt[np == 2 ? 0 : np]; }
// Occupancy and stack size of a kernel do not change: asked once per (kernel, CTA size), then remembered (a
// 2 000-family round plans 4 000 jobs over a dozen candidate tilings each; every GPU of a box is the same model).
static std::mutex g_plan_mutex;
static int ctas_per_sm(const KernelCfg *k, int np, int threads)
{
  const extend_fn fn = fn_of(k, np);
  if (!fn)
    return 0;
  static std::map<std::pair<const void *, int>, int> cache;
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  const auto key = std::make_pair((const void *) fn, threads);
  const auto it = cache.find(key);
  if (it != cache.end())
    return it->second;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *) fn, threads, 0) != cudaSuccess)
  {
    (void) cudaGetLastError();
    return 0;
  }
  cache[key] = nb;
  return nb;
}
// stack bytes (ptxas spills) and registers of a kernel; false if it cannot be loaded
static bool kernel_stack_and_regs(extend_fn fn, int *stack_bytes, int *regs)
{
  static std::map<const void *, std::pair<int, int>> cache;
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  const auto it = cache.find((const void *) fn);
  if (it == cache.end())
  {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void *) fn) != cudaSuccess)
    {
      (void) cudaGetLastError();
      return false;
    }
    cache[(const void *) fn] = std::make_pair((int) fa.localSizeBytes, fa.numRegs);
    *stack_bytes = (int) fa.localSizeBytes;
    *regs = fa.numRegs;
    return true;
  }
  *stack_bytes = it->second.first;
  *regs = it->second.second;
  return true;
}

static int batch_threads(const KernelCfg *k, int np) { return std::min(kBatchThreads, maxt_of(k, np)); }

// Layout of one launch unit (a pass, or the two passes of a family) launched on its own: a single
// CTA when it fits (no grid exchange at all), else 4-warp CTAs on distinct SMs, else the same number
// of equally loaded CTAs on every SM (the column time is set by the most loaded scheduler, and
// every CTA of the family waits for it).
// duo: the unit is one of two single passes (the right and the left pass of a family) that will share the
// device, each with this layout (ramcu_batch_run): twice the CTAs must be resident.
static bool plan_solo(const ramcu_ctx *c, const KernelCfg *k, int np, int n_ext, Plan *p, bool duo = false)
{
  const int ipw = inst_per_warp(k->G);
  const long long warps = ((long long) n_ext + ipw - 1) / ipw;
  const int sm = c->sm_count;
  const int units = duo ? 2 : 1;
  if (!fn_of(k, np))
    return false;
  if (warps * 32 <= maxt_of(k, np))
  {
    p->n_ctas = 1;
    p->threads = (int) warps * 32;
    return true;
  }
  // RAMCU_DUO_SPREAD (tuning aid): from how many warps per pass, in per cent of the SM count (0 = never)
  static const int spread_from = getenv("RAMCU_DUO_SPREAD") ? atoi(getenv("RAMCU_DUO_SPREAD")) : 100;
  // RAMCU_DUO_CTA_WARPS (tuning aid): CTA size, in warps, of the two single-pass layouts below (0 = the default, 2)
  static const int duo_cta_warps = getenv("RAMCU_DUO_CTA_WARPS") ? atoi(getenv("RAMCU_DUO_CTA_WARPS")) : 0;
  if (duo && spread_from > 0 && warps >= (long long) spread_from * sm / 100 && warps <= (long long) sm * 4)
  {
    // both passes on every SM: small CTAs of FULL warps, so that the block scheduler spreads both passes over
    // all SMs and each SM has the other pass's rows to compute while one pass's column sums travel. Config 4
    // (313 warps per pass): 157 CTAs of 2 warps per pass 16.8 ms; one 3-warp CTA of each pass per SM (the third
    // warp holds 4 of the CTA's 68 instances but costs a whole warp's instructions) 17.3 ms; 4-warp CTAs on
    // distinct SMs, where most SMs hold one pass only, 18.5 ms.
    const int w = duo_cta_warps > 0 ? duo_cta_warps : 2;
    p->n_ctas = (int) ((warps + w - 1) / w);
    p->threads = cta_threads(n_ext, p->n_ctas, k->G);
    if (p->threads <= maxt_of(k, np) && (long long) units * p->n_ctas <= (long long) ctas_per_sm(k, np, p->threads) * sm)
      return true;
    p->n_ctas = sm;
    p->threads = cta_threads(n_ext, p->n_ctas, k->G);
    if (p->threads <= maxt_of(k, np) && ctas_per_sm(k, np, p->threads) >= 2)
      return true;
  }
  if (warps <= (long long) sm * 4)
  {
    p->n_ctas = (int) std::min<long long>((warps + 3) / 4, n_ext);
    p->threads = cta_threads(n_ext, p->n_ctas, k->G);
    if (p->threads > maxt_of(k, np))
      return false;
    return !duo || (long long) units * p->n_ctas <= (long long) ctas_per_sm(k, np, p->threads) * sm;
  }
  for (int per_sm = 1; per_sm <= 16; ++per_sm)
  {
    const long long n = (long long) sm * per_sm;
    if (n > n_ext)
      break;
    const int t = cta_threads(n_ext, (int) n, k->G);
    if (t > maxt_of(k, np))
      continue;
    if (ctas_per_sm(k, np, t) < units * per_sm)
      continue;
    p->n_ctas = (int) n;
    p->threads = t;
    return true;
  }
  return false;
}

// Layout when the unit shares a launch with others: fixed CTA size.
static bool plan_batch(const ramcu_ctx *c, const KernelCfg *k, int np, int n_ext, int *n_ctas)
{
  if (!fn_of(k, np))
    return false;
  const int threads = batch_threads(k, np);
  const int per_cta = (threads / 32) * inst_per_warp(k->G);
  const long long n = ((long long) n_ext + per_cta - 1) / per_cta;
  if (n > (long long) ctas_per_sm(k, np, threads) * c->sm_count)
    return false;   // its CTAs could not all be resident: the column exchange would deadlock
  *n_ctas = (int) n;
  return true;
}

// Rough cycles per column of a unit running alone (tuned on B200): a warp executes
// about C*18 + 320 instructions per row and pass (C*15 + 250 when one lane holds the whole band);
// one warp alone issues one every ~3.4 cycles (dependent max-plus chains), a loaded scheduler one
// every ~2.2 cycles; the busiest of the four schedulers of the busiest SM sets the pace. (Fitted to
// tools/duo_sweep.py on B200: bandwidth 20-100, 50-10 000 instances; the choice is within a few per cent
// of the best forced tiling except where neither two single-pass layouts nor a two-pass tiling fit.)
// duo: cost of a column of BOTH passes when two single-pass units with this layout share the device.
static double solo_cost(const ramcu_ctx *c, const KernelCfg *k, int np, int n_ext, const Plan &p, bool duo = false)
{
  const double passes = np == 1 ? 2.0 : 1.0;
  // a tiling squeezed below ~2.3 registers per cell and pass keeps part of its rows in local memory; so does one
  // whose stack frame (ptxas spills) is large: the 41-cell lanes at 168 registers
  const double need = passes * (2.28 * k->C + 30.0) + 40.0;
  double squeeze = (65536.0 / maxt_of(k, np)) < need ? 1.5 : 1.0;
  if (squeeze == 1.0)
  {
    int stack = 0, regs = 1;
    if (kernel_stack_and_regs(fn_of(k, np), &stack, &regs))
      squeeze = stack > 2 * regs ? 1.3 : stack > regs ? 1.1 : 1.0;   // stack bytes per register
  }
  const double instr = passes * ((k->G == 1 ? k->C * 15.0 + 250.0 : k->C * 18.0 + 320.0)) * squeeze;
  const int warps_per_cta = p.threads / 32;
  const int ctas_on_sm = ((duo ? 2 : 1) * p.n_ctas + c->sm_count - 1) / c->sm_count;
  const double per_sched = std::ceil(warps_per_cta * ctas_on_sm / 4.0);
  // one pass per CTA waits for every column's exchange; two passes hide each other's
  const double sync = p.n_ctas > 1 ? ((np == 1 || duo) ? 600.0 : 2400.0) : 0.0;
  (void) n_ext;
  return std::max(instr * 3.4, per_sched * instr * 2.2) + sync;
}

static const KernelCfg *pick_cfg(const ramcu_ctx *c, int np, int B, int n_ext, int force_G,
                                 int concurrent_instances, int force_maxt, Plan *solo, int *batch_ctas,
                                 bool duo = false, double *cost_out = nullptr)
{
  const int W = 2 * B + 1;
  std::vector<const KernelCfg *> cand;
  for (const KernelCfg *k : g_cfgs)
    if (k->B == B && fn_of(k, np))
      cand.push_back(k);
  if (cand.empty())
    for (const KernelCfg *k : g_cfgs)
      if (k->B < 0 && fn_of(k, np) && k->C * k->G >= W && (k->G - 1) * k->C < W)
        cand.push_back(k);
  if (cand.empty())
    for (const KernelCfg *k : g_cfgs)  // narrow bands: trailing lanes may be entirely padding
      if (k->B < 0 && fn_of(k, np) && k->C * k->G >= W)
      {
        cand.push_back(k);
        break;
      }
  std::sort(cand.begin(), cand.end(),
            [](const KernelCfg *a, const KernelCfg *b) { return a->G < b->G; });
  const KernelCfg *best = nullptr;
  double best_cost = 0;
  const bool crowded = concurrent_instances >= c->sm_count * 128;
  bool have_maxt = false;   // cta_thread_limit is a preference: ignored when no tiling of this band has it
  for (const KernelCfg *k : cand)
    have_maxt = have_maxt || (maxt_of(k, np) == force_maxt && (force_G <= 0 || k->G == force_G));
  for (const KernelCfg *k : cand)
  {
    if (force_G > 0 && k->G != force_G)
      continue;
    if (force_maxt > 0 && maxt_of(k, np) != force_maxt && have_maxt)
      continue;
    Plan p;
    int nb = 0;
    // (a tiling that fits alone but not twice still serves: the passes then run one after the other)
    const bool duo_fits = duo && plan_solo(c, k, np, n_ext, &p, true);
    if (!duo_fits && !plan_solo(c, k, np, n_ext, &p))
      continue;
    if (!plan_batch(c, k, np, n_ext, &nb))
      nb = 0;   // too large to share a launch: ramcu_batch_run launches it on its own
    // a crowded device is throughput bound: fewest lanes per instance = fewest instructions
    const double cost = crowded ? (double) k->G + (maxt_of(k, np) > (np == 1 ? maxt2_for(k->C) : maxt_for(k->C)) ? 0.5 : 0.0)
                                : solo_cost(c, k, np, n_ext, p, duo_fits) * ((duo && !duo_fits) ? 2.0 : 1.0);
    if (!best || cost < best_cost)
    {
      best = k;
      best_cost = cost;
      *solo = p;
      *batch_ctas = nb;
    }
  }
  if (cost_out)
    *cost_out = best_cost;
  return best;
}

static size_t slots_bytes(int L) { return ((size_t) L + 1) * 4 * kRedStride * sizeof(unsigned long long); }

// results block of a job, on the device and (pinned) on the host: [8 ints][thigh n_ext][tpos n_ext][master L bytes]
static size_t out_bytes(int n_ext, int L) { return (8 + 2 * (size_t) n_ext) * sizeof(int) + (size_t) std::max(L, 1); }

static void *pinned_get(ramcu_ctx *c, size_t bytes, size_t *cap)
{
  size_t cls = 4096;
  while (cls < bytes)
    cls <<= 1;
  *cap = cls;
  auto &fl = c->pinned_free[cls];
  if (!fl.empty())
  {
    void *p = fl.back();
    fl.pop_back();
    return p;
  }
  void *p = nullptr;
  if (cudaHostAlloc(&p, cls, cudaHostAllocDefault) != cudaSuccess)
    return nullptr;
  return p;
}
static void pinned_put(ramcu_ctx *c, void *p, size_t cap)
{
  if (!p)
    return;
  if (c)
    c->pinned_free[cap].push_back(p);
  else
    cudaFreeHost(p);
}

static void free_job_device(ramcu_ctx *c, ramcu_job *j)
{
  void *ptrs[] = {j->d_inst, j->d_out, j->d_slots, j->d_fam, j->d_cta_fam, j->d_forced, j->d_dbg_rows,
                  j->d_dbg_best, j->d_dbg_sums, j->d_dbg_wait, j->d_dbg_path, j->d_dbg_idx, j->d_seed_rows,
                  j->d_dbg_cand, j->d_dbg_cells};
  for (void *p : ptrs)
    if (p)
    {
      if (c)
        cudaFreeAsync(p, c->stream);
      else
        cudaFree(p);
    }
  pinned_put(c, j->h_out, j->h_out_cap);
  j->h_out = nullptr;
}

static void destroy_job(ramcu_ctx *c, ramcu_job *j)
{
  if (!j)
    return;
  if (j->mate)
  {
    j->mate->mate = nullptr;
    j->mate->pair_cfg = nullptr;
  }
  free_job_device(c, j);
  delete j;
}

// CUDA call inside job creation: on failure the half-built job and its buffers are released
#define CKJ(call)                                                                        \
  do                                                                                     \
  {                                                                                      \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
    {                                                                                    \
      destroy_job(c, j);                                                                 \
      return fail(RAMCU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                   \
    }                                                                                    \
  } while (0)

static int job_create_impl(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                           const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                           const ramcu_params_t *pr, int dbg_rows, const uint8_t *forced,
                           ramcu_job **out, const int32_t *n_row = nullptr, bool duo = false,
                           int verbose_rows = 0, bool verbose_cells = false)
{
  if (!c || !lib || !sc || !pr || !out || (n_cores > 0 && !cores))
    return fail(RAMCU_ERR_ARG, "ramcu_job_create: NULL argument");
  if (n_cores < 0 || pr->L < 0)
    return fail(RAMCU_ERR_ARG, "ramcu_job_create: negative size");
  if (pr->bandwidth < 0 || pr->bandwidth > RAMCU_MAX_BANDWIDTH)
    return fail(RAMCU_ERR_UNSUPPORTED, "bandwidth %d outside 0..%d", pr->bandwidth,
                RAMCU_MAX_BANDWIDTH);
  if (pr->L > (1 << 24))
    return fail(RAMCU_ERR_UNSUPPORTED, "L %d larger than %d", pr->L, 1 << 24);
  {
    // Every value the reference computes stays above its own -1000000000 floor (bnw_extend.c:802-804) when
    // the out-of-bounds sentinel -987654321 cannot drift that far: a path adds at most `worst` per step and
    // has at most L + 2B + 2 steps. The kernel relies on that (it never applies the floor); the reference's
    // matrices and any sensible -gapopen/-gapext/-match/-mismatch are three orders of magnitude inside it.
    long long worst = (long long) llabs((long long) sc->gapopen) + llabs((long long) sc->gapextn);
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < RAMCU_NCODES; ++b)
        worst = std::max(worst, (long long) llabs((long long) sc->matrix[a][b]));
    if (worst * ((long long) pr->L + 2LL * pr->bandwidth + 3) > 12000000LL)
      return fail(RAMCU_ERR_UNSUPPORTED,
                  "scores of magnitude %lld over L + band = %lld steps can cross the reference's -1000000000 "
                  "floor (limit: magnitude x steps <= 12000000)", worst, (long long) pr->L + 2LL * pr->bandwidth + 3);
  }
  CK(cudaSetDevice(c->device));

  ramcu_job *j = new ramcu_job();
  j->lib = lib;
  j->direction = direction ? 1 : 0;
  j->n_cores = n_cores;
  j->sc = *sc;
  j->pr = *pr;

  const int B = pr->bandwidth;
  std::vector<InstDev> inst;
  flatten_cores(j->direction, n_cores, cores, lib->length, inst);
  j->n_ext = (int) inst.size();
  j->orig.resize(inst.size());
  for (size_t i = 0; i < inst.size(); ++i)
    j->orig[i] = inst[i].orig;

  if (j->n_ext > 0 && pr->L > 0)
  {
    // which outputs are asked for is known from the arguments: lay the job out for the kernel variant it will run
    j->extras = dbg_rows != 0 || forced != nullptr || verbose_rows > 0 || getenv("RAMCU_STATS") != nullptr;
    j->cfg = pick_cfg(c, j->extras ? 2 : 0, B, j->n_ext, pr->lanes_per_instance, pr->concurrent_instances, pr->cta_thread_limit,
                      &j->solo, &j->batch_ctas, duo, &j->plan_cost);
    if (!j->cfg)
    {
      destroy_job(c, j);
      return fail(RAMCU_ERR_CAPACITY,
                  "no band tiling for bandwidth %d with %d instances fits this device "
                  "(lanes_per_instance=%d)", B, (int) inst.size(), pr->lanes_per_instance);
    }
    j->n_ctas = j->solo.n_ctas;
    // a pair layout may use up to as many CTAs as the larger of the two passes needs on its own
    const int max_ctas = std::max(std::max(j->solo.n_ctas, j->batch_ctas), 1) * 2 + 8;

    const int W = 2 * B + 1;
    const size_t ob = out_bytes(j->n_ext, pr->L);
    CKJ(cudaMallocAsync(&j->d_inst, inst.size() * sizeof(InstDev), c->stream));
    CKJ(cudaMemcpyAsync(j->d_inst, inst.data(), inst.size() * sizeof(InstDev), cudaMemcpyHostToDevice, c->stream));
    CKJ(cudaMallocAsync(&j->d_out, ob, c->stream));
    j->h_out = (unsigned char *) pinned_get(c, ob, &j->h_out_cap);
    if (!j->h_out)
    {
      destroy_job(c, j);
      return fail(RAMCU_ERR_CUDA, "ramcu_job_create: cudaHostAlloc of %zu bytes failed", ob);
    }
    CKJ(cudaMallocAsync(&j->d_slots, slots_bytes(pr->L), c->stream));
    CKJ(cudaMallocAsync(&j->d_fam, 2 * sizeof(FamDev), c->stream));
    CKJ(cudaMallocAsync(&j->d_cta_fam, max_ctas * sizeof(int), c->stream));
    CKJ(cudaMemsetAsync(j->d_cta_fam, 0, max_ctas * sizeof(int), c->stream));
    j->max_ctas = max_ctas;
    if (dbg_rows < 0)   // -outmat mode: path states of -dbg_rows rows, no forced consensus
    {
      const size_t rows = (size_t) -dbg_rows;
      CKJ(cudaMallocAsync(&j->d_dbg_path, rows * j->n_ext * W, c->stream));
      CKJ(cudaMemsetAsync(j->d_dbg_path, 0, rows * j->n_ext * W, c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_idx, rows * j->n_ext * sizeof(int), c->stream));
      CKJ(cudaMemsetAsync(j->d_dbg_idx, 0, rows * j->n_ext * sizeof(int), c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_best, rows * j->n_ext * 4 * sizeof(int), c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_sums, rows * 4 * sizeof(int), c->stream));
    }
    if (verbose_rows > 0)   // verbose dumps: per-candidate row values of the first verbose_rows rows
    {
      const size_t rows = (size_t) verbose_rows;
      CKJ(cudaMallocAsync(&j->d_dbg_cand, rows * j->n_ext * 12 * sizeof(int), c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_best, rows * j->n_ext * 4 * sizeof(int), c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_sums, rows * 4 * sizeof(int), c->stream));
      if (verbose_cells)
        CKJ(cudaMallocAsync(&j->d_dbg_cells, rows * j->n_ext * 4 * W * 2 * sizeof(int), c->stream));
    }
    if (dbg_rows == 0 && forced)   // seed extension: the query is the consensus, row values come back
    {
      CKJ(cudaMallocAsync(&j->d_forced, std::max(pr->L, 1), c->stream));
      CKJ(cudaMemcpyAsync(j->d_forced, forced, pr->L, cudaMemcpyHostToDevice, c->stream));
      CKJ(cudaMallocAsync(&j->d_seed_rows, (size_t) std::max(pr->L, 1) * j->n_ext * 2 * sizeof(int), c->stream));
    }
    if (dbg_rows > 0)
    {
      CKJ(cudaMallocAsync(&j->d_dbg_rows, (size_t) dbg_rows * j->n_ext * W * 2 * sizeof(int), c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_best, (size_t) dbg_rows * j->n_ext * 4 * sizeof(int), c->stream));
      CKJ(cudaMallocAsync(&j->d_dbg_sums, (size_t) dbg_rows * 4 * sizeof(int), c->stream));
      CKJ(cudaMemsetAsync(j->d_dbg_rows, 0, (size_t) dbg_rows * j->n_ext * W * 2 * sizeof(int), c->stream));
      if (forced)
      {
        CKJ(cudaMallocAsync(&j->d_forced, dbg_rows, c->stream));
        CKJ(cudaMemcpyAsync(j->d_forced, forced, dbg_rows, cudaMemcpyHostToDevice, c->stream));
      }
    }
    if (getenv("RAMCU_STATS"))
    {
      const size_t nw = (size_t) (2 + 3 * kStatCols) * max_ctas;
      CKJ(cudaMallocAsync(&j->d_dbg_wait, nw * sizeof(long long), c->stream));
      CKJ(cudaMemsetAsync(j->d_dbg_wait, 0, nw * sizeof(long long), c->stream));
    }
    FamDev &f = j->fam;
    fill_family(f, sc, pr, j->n_ext, j->n_ctas);
    int *o = reinterpret_cast<int *>(j->d_out);
    f.dbg_wait = j->d_dbg_wait;
    f.lib2 = lib->d_lib2;
    f.libn = lib->d_libn;
    f.lib_len = (long long) lib->length;
    f.inst = j->d_inst;
    f.result = o;
    f.thigh = o + 8;
    f.tpos = o + 8 + j->n_ext;
    f.master = reinterpret_cast<unsigned char *>(o + 8 + 2 * (size_t) j->n_ext);
    f.slots = j->d_slots;
    f.forced = j->d_forced;
    f.dbg_rows = j->d_dbg_rows;
    f.dbg_path = j->d_dbg_path;
    f.dbg_idx = j->d_dbg_idx;
    f.dbg_best = j->d_dbg_best;
    f.dbg_sums = j->d_dbg_sums;
    f.seed_rows = j->d_seed_rows;
    f.dbg_cand = j->d_dbg_cand;
    f.dbg_cells = j->d_dbg_cells;
    f.dbg_nrows = verbose_rows > 0 ? verbose_rows : std::abs(dbg_rows);
    if (n_row)
      for (int code = 0; code < RAMCU_NCODES; ++code)
        f.mN[code] = n_row[code];
    // pageable source: the copy is staged before the call returns, `f` and `inst` may go away
    CKJ(cudaMemcpyAsync(j->d_fam, &f, sizeof(f), cudaMemcpyHostToDevice, c->stream));
  }
  *out = j;
  return RAMCU_OK;
}

extern "C" int ramcu_job_create(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                                const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                const ramcu_params_t *pr, ramcu_job **out)
{
  return job_create_impl(c, lib, direction, n_cores, cores, sc, pr, 0, nullptr, out);
}

static int *next_ticket(ramcu_ctx *c)
{
  int *t = c->d_ticket + (c->ticket_next++ % kTickets);
  return t;
}

// device results -> the job's pinned host block, in stream order behind the launch
static int enqueue_fetch(ramcu_ctx *c, ramcu_job *j)
{
  CK(cudaMemcpyAsync(j->h_out, j->d_out, out_bytes(j->n_ext, j->pr.L), cudaMemcpyDeviceToHost, c->stream));
  return RAMCU_OK;
}

// A launch whose grid fits the device at once is launched cooperatively: the runtime then guarantees
// that every CTA is resident together (or fails the launch), which is what the column exchange of a
// multi-CTA family needs; it does not use grid.sync().
static int launch_unit(ramcu_ctx *c, extend_fn fn, int grid, int threads, const FamDev *d_fams, const int *d_map,
                       bool needs_residency, int ctas_per_sm_limit)
{
  int *ticket = next_ticket(c);
  CK(cudaMemsetAsync(ticket, 0, sizeof(int), c->stream));
  if (needs_residency && grid <= ctas_per_sm_limit * c->sm_count && c->coop_ok)
  {
    void *args[] = {(void *) &d_fams, (void *) &d_map, (void *) &ticket};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *) fn, dim3(grid), dim3(threads), args, 0, c->stream);
    if (e == cudaSuccess)
      return RAMCU_OK;
    (void) cudaGetLastError();
    if (e != cudaErrorCooperativeLaunchTooLarge && e != cudaErrorNotSupported)
      return fail(RAMCU_ERR_CUDA, "cooperative launch failed: %s", cudaGetErrorString(e));
  }
  fn<<<grid, threads, 0, c->stream>>>(d_fams, d_map, ticket);
  CK(cudaGetLastError());
  return RAMCU_OK;
}

static int launch_single(ramcu_ctx *c, ramcu_job *j)
{
  if (j->n_ctas > 1)
    CK(cudaMemsetAsync(j->d_slots, 0, slots_bytes(j->pr.L), c->stream));
  if (j->fam_dirty)
  {
    j->fam.n_ctas = j->solo.n_ctas;
    j->fam.cta0 = 0;
    j->fam.mate = -1;
    CK(cudaMemcpyAsync(j->d_fam, &j->fam, sizeof(FamDev), cudaMemcpyHostToDevice, c->stream));
    j->fam_dirty = false;
  }
  j->ran_cfg = j->cfg;
  j->ran_passes = 1;
  j->stats_ctas = j->solo.n_ctas;
  const int v = j->extras ? 2 : 0;
  int rc = launch_unit(c, fn_of(j->cfg, v), j->solo.n_ctas, j->solo.threads, j->d_fam, j->d_cta_fam, j->solo.n_ctas > 1,
                       ctas_per_sm(j->cfg, v, j->solo.threads));
  if (rc)
    return rc;
  return enqueue_fetch(c, j);
}

extern "C" int ramcu_job_run(ramcu_ctx *c, ramcu_job *j)
{
  if (!c || !j)
    return fail(RAMCU_ERR_ARG, "ramcu_job_run: NULL argument");
  CK(cudaSetDevice(c->device));
  j->ran = true;
  j->timed_gen = 0;
  if (j->n_ctas == 0)
    return RAMCU_OK;
  CK(cudaEventRecord(c->ev0, c->stream));
  int rc = launch_single(c, j);
  if (rc)
    return rc;
  CK(cudaEventRecord(c->ev1, c->stream));
  j->timed_gen = ++c->launch_gen;
  return RAMCU_OK;
}

extern "C" int ramcu_job_fetch(ramcu_ctx *c, ramcu_job *j, uint8_t *master_out, int32_t *ext_len_out,
                               int32_t *score_add_out, ramcu_result_t *res)
{
  if (!c || !j || !res)
    return fail(RAMCU_ERR_ARG, "ramcu_job_fetch: NULL argument");
  if (!j->ran)
    return fail(RAMCU_ERR_ARG, "ramcu_job_fetch: job was never run");
  CK(cudaSetDevice(c->device));
  memset(res, 0, sizeof(*res));
  res->n_extendable = j->n_ext;
  map_results(j->n_cores, std::vector<int>(), nullptr, nullptr, ext_len_out, score_add_out);
  if (j->n_ctas == 0)
  {
    // No instance to extend: every column sums to 0, so 'A' is chosen and the loop stops after
    // WHEN_TO_STOP rows without ever setting a maximum (ram_extend.c:973-974,1194,1216).
    const int L = j->pr.L, stop = j->pr.when_to_stop;
    int rows = 0, stop_row = -1, max_row = -1;
    for (int r = 0; r < L; ++r)
    {
      rows = r + 1;
      if (0 >= 0 + abs(max_row - r) * j->pr.minimprovement)
        max_row = r;
      if (r - max_row >= stop)
      {
        stop_row = r;
        break;
      }
    }
    if (master_out)
      memset(master_out, 0, rows);
    res->extended_bp = max_row + 1;
    res->rows_processed = rows;
    res->stop_row = stop_row;
    res->hit_limit = (stop_row >= 0 && stop_row == L - 1) ? 1 : 0;
    return RAMCU_OK;
  }
  {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
      return fail(RAMCU_ERR_CUDA, "extension kernel failed: %s%s", cudaGetErrorString(e),
                  e == cudaErrorLaunchFailure ? " (a column exchange that waited too long traps: is another "
                                                "process or kernel holding SMs of this device?)" : "");
  }
  const int *r4 = reinterpret_cast<const int *>(j->h_out);
  const int *th = r4 + 8, *tp = r4 + 8 + j->n_ext;
  const unsigned char *hm = reinterpret_cast<const unsigned char *>(r4 + 8 + 2 * (size_t) j->n_ext);
  res->extended_bp = r4[0] + 1;
  res->rows_processed = r4[1];
  res->stop_row = r4[2];
  res->max_score = r4[3];
  res->hit_limit = (r4[2] >= 0 && r4[2] == j->pr.L - 1) ? 1 : 0;
  res->lanes_per_instance = j->ran_cfg ? j->ran_cfg->G : j->cfg->G;
  res->passes_per_thread = j->ran_passes;
  if (j->timed_gen != 0 && j->timed_gen == c->launch_gen)
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess)
      res->kernel_ms = ms;
  }
  if (j->d_dbg_wait && j->stats_ctas > 0)
  {
    const int n_stat = j->stats_ctas;
    std::vector<long long> w((size_t) (2 + 3 * kStatCols) * n_stat);
    CK(cudaMemcpy(w.data(), j->d_dbg_wait, w.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long wmin = 1LL << 62, wmax = 0, wsum = 0, tmax = 0;
    int n_seen = 0;
    for (int i = 0; i < n_stat; ++i)
    {
      if (w[2 * i + 1] == 0)
        continue;
      ++n_seen;
      wmin = std::min(wmin, w[2 * i]);
      wmax = std::max(wmax, w[2 * i]);
      wsum += w[2 * i];
      tmax = std::max(tmax, w[2 * i + 1]);
    }
    fprintf(stderr, "[ramcu stats] ctas=%d rows=%d cycles/row: total %.0f, warp 0 waiting for the column sums: min %.0f avg %.0f max %.0f\n",
            n_seen, r4[1], (double) tmax / std::max(r4[1], 1), (double) wmin / std::max(r4[1], 1),
            (double) wsum / std::max(n_seen, 1) / std::max(r4[1], 1), (double) wmax / std::max(r4[1], 1));
    if (const char *dump = getenv("RAMCU_STATS_DUMP"))
    {
      FILE *fh = fopen(dump, "ab");
      if (fh)
      {
        const long long hdr[2] = {n_stat, kStatCols};
        fwrite(hdr, sizeof(long long), 2, fh);
        fwrite(w.data(), sizeof(long long), w.size(), fh);
        fclose(fh);
      }
    }
    if (n_stat > 1 && r4[1] > kStatCol0 + kStatCols)
    {
      // exchange of the sampled columns: when the CTAs sent (spread = skew between CTAs), when the
      // pollers arrived, when the sums were there
      double skew = 0, lat_first = 0, lat_mean = 0, poll_lead = 0;
      for (int cidx = 0; cidx < kStatCols; ++cidx)
      {
        const long long *t = w.data() + 2 * n_stat + (size_t) cidx * n_stat * 3;
        long long smin = 1LL << 62, smax = 0, rmin = 1LL << 62;
        for (int i = 0; i < n_stat; ++i)
        {
          smin = std::min(smin, t[3 * i]);
          smax = std::max(smax, t[3 * i]);
          rmin = std::min(rmin, t[3 * i + 2]);
        }
        double rsum = 0, psum = 0;
        for (int i = 0; i < n_stat; ++i)
        {
          rsum += (double) (t[3 * i + 2] - smax);
          psum += (double) (t[3 * i + 1] - smax);
        }
        skew += (double) (smax - smin);
        lat_first += (double) (rmin - smax);
        lat_mean += rsum / n_stat;
        poll_lead += psum / n_stat;
      }
      fprintf(stderr, "[ramcu stats] exchange of %d columns (ns): send skew %.0f, last send -> first ready %.0f, -> mean ready %.0f, "
              "mean poll start - last send %.0f\n", kStatCols, skew / kStatCols, lat_first / kStatCols, lat_mean / kStatCols,
              poll_lead / kStatCols);
    }
  }
  if (master_out && r4[1] > 0)
    memcpy(master_out, hm, r4[1]);
  if (ext_len_out || score_add_out)
    map_results(j->n_cores, j->orig, th, tp, ext_len_out, score_add_out);
  return RAMCU_OK;
}

extern "C" int ramcu_job_free(ramcu_ctx *c, ramcu_job *j)
{
  if (!j)
    return RAMCU_OK;
  if (c)
    cudaSetDevice(c->device);
  destroy_job(c, j);
  return RAMCU_OK;
}

extern "C" int ramcu_extend(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                            const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                            const ramcu_params_t *pr, uint8_t *master_out, int32_t *ext_len_out,
                            int32_t *score_add_out, ramcu_result_t *res)
{
  ramcu_job *j = nullptr;
  int rc = ramcu_job_create(c, lib, direction, n_cores, cores, sc, pr, &j);
  if (rc)
    return rc;
  rc = ramcu_job_run(c, j);
  if (!rc)
    rc = ramcu_job_fetch(c, j, master_out, ext_len_out, score_add_out, res);
  ramcu_job_free(c, j);
  return rc;
}

// ------------------------------------------------------------------------------------------
// batches: many independent families sharing the device
// ------------------------------------------------------------------------------------------
// A launch unit is a job on its own (one pass per CTA) or the two linked jobs of
// ramcu_job_create_pair when both are in the list (both passes in the same threads). Units are
// grouped by kernel instantiation, one launch per group, longest first: CTAs take logical ids from
// a ticket in start order, so the CTAs of a multi-CTA unit are resident together or ahead of every
// CTA that could wait for them, and the SMs freed by finished units are back-filled by the next
// ones of the group. Groups run one after the other on the context's stream: two launches that each
// hold a partly resident multi-CTA family could wait for each other forever.
struct Unit
{
  ramcu_job *a, *b;   // b != NULL: two passes per CTA
  const KernelCfg *cfg;
  int np, n_ctas;
  long long work;
  int threads = 0;    // > 0: CTA size of this unit's launch (its own layout) instead of batch_threads()
};

extern "C" int ramcu_batch_run(ramcu_ctx *c, int n_jobs, ramcu_job **jobs)
{
  if (!c || (n_jobs > 0 && !jobs))
    return fail(RAMCU_ERR_ARG, "ramcu_batch_run: NULL argument");
  CK(cudaSetDevice(c->device));
  std::vector<Unit> units;
  std::vector<ramcu_job *> alone;   // too large to share a launch
  for (int i = 0; i < n_jobs; ++i)
  {
    if (!jobs[i])
      return fail(RAMCU_ERR_ARG, "ramcu_batch_run: jobs[%d] == NULL", i);
    jobs[i]->ran = true;
    jobs[i]->timed_gen = 0;
    jobs[i]->in_batch = true;
  }
  // A crowded device hides every family's column exchange behind other families' rows: there the
  // single-pass tilings with the fewest lanes per instance win (measured on the config-5 mix: 1485 against
  // 1297 GCUPS with two passes per thread). Two passes per thread are for a device that is not full.
  long long batch_instances = 0;
  for (int i = 0; i < n_jobs; ++i)
    batch_instances += jobs[i]->n_ext;
  const bool crowded = batch_instances >= (long long) c->sm_count * 256 && !getenv("RAMCU_PAIR_ALWAYS");
  for (int i = 0; i < n_jobs; ++i)
  {
    ramcu_job *j = jobs[i];
    if (j->n_ctas == 0 || !j->in_batch)
      continue;
    ramcu_job *m = crowded ? nullptr : j->mate;
    if (m && m->in_batch && m->n_ctas > 0 && j->pair_cfg && j->pair_batch_ctas > 0)
    {
      ramcu_job *a = j->direction ? j : m, *b = j->direction ? m : j;
      units.push_back(Unit{a, b, j->pair_cfg, 1, j->pair_batch_ctas,
                           (long long) (a->n_ext + b->n_ext) * j->pr.L});
      j->in_batch = m->in_batch = false;
      continue;
    }
    if (m && m->in_batch && m->n_ctas > 0 && j->pair_cfg && j->pair_batch_ctas == 0 && j->pair_solo.n_ctas > 0)
    {
      // the pair does not fit next to others: launched on its own below, still two passes per CTA
      ramcu_job *a = j->direction ? j : m, *b = j->direction ? m : j;
      units.push_back(Unit{a, b, j->pair_cfg, 1, -j->pair_solo.n_ctas, (long long) (a->n_ext + b->n_ext) * j->pr.L});
      j->in_batch = m->in_batch = false;
      continue;
    }
    if (j->batch_ctas > 0)
      units.push_back(Unit{j, nullptr, j->cfg, j->extras ? 2 : 0, j->batch_ctas, (long long) j->n_ext * j->pr.L});
    else
      alone.push_back(j);
    j->in_batch = false;
  }
  if (units.size() == 1 && alone.empty())
  {
    // nothing shares the device with this unit: lay it out as a launch of its own
    if (units[0].np != 1)
    {
      alone.push_back(units[0].a);
      units.clear();
    }
    else
      units[0].n_ctas = -units[0].a->pair_solo.n_ctas;
  }
  if (units.size() + alone.size() == 2 && (units.empty() || units[0].np != 1) && (units.size() < 2 || units[1].np != 1))
  {
    // two single passes with the device to themselves (the right and the left pass of a family whose tiling
    // holds one row state per thread): each keeps the layout it would have alone -- equally loaded CTAs on every
    // SM -- and the two share one launch, so that one pass computes while the other's column sums travel
    ramcu_job *d[2];
    int n = 0;
    for (const Unit &u : units)
      d[n++] = u.a;
    for (ramcu_job *j : alone)
      d[n++] = j;
    if (d[0]->cfg == d[1]->cfg && d[0]->extras == d[1]->extras && d[0]->solo.threads == d[1]->solo.threads &&
        d[0]->solo.n_ctas + d[1]->solo.n_ctas <=
            ctas_per_sm(d[0]->cfg, d[0]->extras ? 2 : 0, d[0]->solo.threads) * c->sm_count)
    {
      units.clear();
      alone.clear();
      for (ramcu_job *j : d)
      {
        Unit u{j, nullptr, j->cfg, j->extras ? 2 : 0, j->solo.n_ctas, (long long) j->n_ext * j->pr.L};
        u.threads = j->solo.threads;
        units.push_back(u);
      }
    }
  }
  std::stable_sort(units.begin(), units.end(), [](const Unit &x, const Unit &y) {
    if ((x.n_ctas < 0) != (y.n_ctas < 0))
      return x.n_ctas < 0;
    if (x.cfg != y.cfg)
      return x.cfg < y.cfg;
    if (x.np != y.np)
      return x.np > y.np;
    return x.work > y.work;
  });
  CK(cudaEventRecord(c->ev0, c->stream));
  for (ramcu_job *j : alone)
  {
    int rc = launch_single(c, j);
    if (rc)
      return rc;
  }
  size_t p = 0;
  while (p < units.size())
  {
    size_t q = p;
    const KernelCfg *cfg = units[p].cfg;
    const int np = units[p].np;
    const bool solo = units[p].n_ctas < 0;
    std::vector<FamDev> fams;
    std::vector<int> cta_fam;
    bool multi = false;
    int threads = solo ? units[p].a->pair_solo.threads : units[p].threads > 0 ? units[p].threads : batch_threads(cfg, np);
    while (q < units.size() && units[q].cfg == cfg && units[q].np == np && (units[q].n_ctas < 0) == solo)
    {
      const Unit &u = units[q];
      const int n_ctas = std::abs(u.n_ctas);
      FamDev f = u.a->fam;
      f.cta0 = (int) cta_fam.size();
      f.n_ctas = n_ctas;
      f.mate = u.b ? (int) fams.size() + 1 : -1;
      for (int k = 0; k < n_ctas; ++k)
        cta_fam.push_back((int) fams.size());
      fams.push_back(f);
      u.a->ran_cfg = cfg;
      u.a->ran_passes = u.b ? 2 : 1;
      u.a->stats_ctas = std::min(n_ctas, u.a->max_ctas);
      if (u.b)
      {
        FamDev g = u.b->fam;
        g.cta0 = f.cta0;
        g.n_ctas = n_ctas;
        g.mate = -1;
        fams.push_back(g);
        u.b->ran_cfg = cfg;
        u.b->ran_passes = 2;
        u.b->stats_ctas = 0;
      }
      if (n_ctas > 1)
      {
        multi = true;
        CK(cudaMemsetAsync(u.a->d_slots, 0, slots_bytes(u.a->pr.L), c->stream));
        if (u.b)
          CK(cudaMemsetAsync(u.b->d_slots, 0, slots_bytes(u.b->pr.L), c->stream));
      }
      u.a->fam_dirty = true;
      ++q;
      if (solo)
        break;   // one unit per launch
    }
    FamDev *d_fams = nullptr;
    int *d_map = nullptr;
    CK(cudaMallocAsync(&d_fams, fams.size() * sizeof(FamDev), c->stream));
    CK(cudaMallocAsync(&d_map, cta_fam.size() * sizeof(int), c->stream));
    // pageable sources: the copies are staged before the calls return, so the vectors may go away
    CK(cudaMemcpyAsync(d_fams, fams.data(), fams.size() * sizeof(FamDev), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_map, cta_fam.data(), cta_fam.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    int rc = launch_unit(c, fn_of(cfg, np), (int) cta_fam.size(), threads, d_fams, d_map, multi, ctas_per_sm(cfg, np, threads));
    if (rc)
      return rc;
    CK(cudaFreeAsync(d_fams, c->stream));
    CK(cudaFreeAsync(d_map, c->stream));
    p = q;
  }
  CK(cudaEventRecord(c->ev1, c->stream));
  ++c->launch_gen;
  for (int i = 0; i < n_jobs; ++i)
  {
    jobs[i]->timed_gen = c->launch_gen;
    if (jobs[i]->n_ctas > 0)
    {
      int rc = enqueue_fetch(c, jobs[i]);
      if (rc)
        return rc;
    }
  }
  return RAMCU_OK;
}

// ------------------------------------------------------------------------------------------
// both passes of one family in one launch
// ------------------------------------------------------------------------------------------
// main() extends right, narrows the bounds of cores that a right extension ran into
// (ram_extend.c:445-499), then extends left with those bounds. When that step leaves every bound
// as it was (the caller checks: ramx_boundary.c compares the bounds it is handed for the left
// pass with the ones used here), the two passes are independent column loops over the same
// cores. Every thread then carries an instance band of BOTH passes and the passes take turns row by
// row: while the column sums of one pass travel between the CTAs, the other pass's row is computed
// (DESIGN.md 3.4).
extern "C" int ramcu_job_create_pair(ramcu_ctx *c, ramcu_library *lib, int n_cores,
                                     const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                     const ramcu_params_t *pr, ramcu_job **right, ramcu_job **left)
{
  if (!c || !pr || !right || !left || (n_cores > 0 && !cores))
    return fail(RAMCU_ERR_ARG, "ramcu_job_create_pair: NULL argument");
  *right = *left = nullptr;
  // each pass is laid out as one of two single-pass units that share the device (ramcu_batch_run) ...
  int rc = job_create_impl(c, lib, 1, n_cores, cores, sc, pr, 0, nullptr, right, nullptr, true);
  if (!rc)
    rc = job_create_impl(c, lib, 0, n_cores, cores, sc, pr, 0, nullptr, left, nullptr, true);
  if (rc)
  {
    ramcu_job_free(c, *right);
    *right = nullptr;
    return rc;
  }
  ramcu_job *a = *right, *b = *left;
  if (a->n_ctas > 0 && b->n_ctas > 0 && !getenv("RAMCU_NO_PAIR"))   // RAMCU_NO_PAIR: tuning aid
  {
    // ... unless both passes in the same threads are expected to be faster: that pays where the column exchange
    // dominates and the family is too large for two single-pass layouts side by side
    Plan solo;
    int nb = 0;
    double pair_cost = 0;
    const KernelCfg *k = pick_cfg(c, 1, pr->bandwidth, std::max(a->n_ext, b->n_ext), pr->lanes_per_instance,
                                  pr->concurrent_instances, pr->cta_thread_limit, &solo, &nb, false, &pair_cost);
    const bool crowded = pr->concurrent_instances >= c->sm_count * 128;
    const bool duo_better = !crowded && !getenv("RAMCU_PAIR_ALWAYS") && a->cfg == b->cfg &&
                            a->solo.threads == b->solo.threads && std::max(a->plan_cost, b->plan_cost) <= pair_cost;
    if (k && !duo_better && solo.n_ctas <= std::min(a->max_ctas, b->max_ctas))
    {
      if (nb > std::min(a->max_ctas, b->max_ctas))
        nb = 0;
      a->mate = b;
      b->mate = a;
      a->pair_cfg = b->pair_cfg = k;
      a->pair_solo = b->pair_solo = solo;
      a->pair_batch_ctas = b->pair_batch_ctas = nb;
    }
  }
  return RAMCU_OK;
}

extern "C" int ramcu_extend_pair(ramcu_ctx *c, ramcu_library *lib, int n_cores, const ramcu_core_t *cores,
                                 const ramcu_scoring_t *sc, const ramcu_params_t *pr,
                                 uint8_t *master_right, int32_t *ext_len_right, int32_t *score_add_right,
                                 ramcu_result_t *res_right, uint8_t *master_left, int32_t *ext_len_left,
                                 int32_t *score_add_left, ramcu_result_t *res_left)
{
  ramcu_job *jobs[2] = {nullptr, nullptr};
  int rc = ramcu_job_create_pair(c, lib, n_cores, cores, sc, pr, &jobs[0], &jobs[1]);
  if (rc)
    return rc;
  rc = ramcu_batch_run(c, 2, jobs);
  if (!rc)
    rc = ramcu_job_fetch(c, jobs[0], master_right, ext_len_right, score_add_right, res_right);
  if (!rc)
    rc = ramcu_job_fetch(c, jobs[1], master_left, ext_len_left, score_add_left, res_left);
  ramcu_job_free(c, jobs[0]);
  ramcu_job_free(c, jobs[1]);
  return rc;
}

// ------------------------------------------------------------------------------------------
// seed extension: the other callers of compute_nw_row() (bnw_extend.c:167-649)
// ------------------------------------------------------------------------------------------
// cons_seed_extend() runs the band row with a GIVEN consensus -- the query -- against one subject
// sequence, left of the seed and right of it, and keeps the best row score and where it was reached.
// Here every extension is a one-instance family with a forced consensus (the column sums play no
// role): 2 x n of them share one ramcu_batch_run, the kernel returns compute_nw_row()'s value and
// sequence index for every row, and the reference's scan over the rows (running maximum, y-drop
// break: bnw_extend.c:338-367, 404-440) is applied to those values here.
static void seed_scan(const int *rows, int n_rows, int y_drop, bool right, int *max_score, int *seq_idx,
                      int *cons_idx, int *rows_used)
{
  int best = 0, sidx = -1, cidx = -1, r = 0;
  for (r = 0; r < n_rows; ++r)
  {
    const int sc = rows[2 * r], pos = rows[2 * r + 1];
    // the reference's right extension compares with the literal 180 whatever y_drop is (bnw_extend.c:429)
    if (y_drop >= 0 && best - sc > (right ? 180 : y_drop))
      break;
    if (sc > best)
    {
      best = sc;
      sidx = pos;
      cidx = r;
    }
  }
  *max_score = best;
  *seq_idx = sidx;
  *cons_idx = cidx;
  *rows_used = r;
}

extern "C" int ramcu_seed_extend_batch(ramcu_ctx *c, ramcu_library *lib, int n, const ramcu_seed_t *seeds,
                                       const ramcu_scoring_t *sc, const int32_t *n_row, int bandwidth, int y_drop,
                                       ramcu_seed_result_t *results)
{
  if (!c || !lib || !sc || !n_row || (n > 0 && (!seeds || !results)) || n < 0)
    return fail(RAMCU_ERR_ARG, "ramcu_seed_extend_batch: invalid argument");
  CK(cudaSetDevice(c->device));
  std::vector<ramcu_job *> jobs((size_t) 2 * n, nullptr);
  auto cleanup = [&]() {
    for (ramcu_job *j : jobs)
      if (j)
        destroy_job(c, j);
  };
  std::vector<uint8_t> forced;
  for (int i = 0; i < n; ++i)
  {
    const ramcu_seed_t &s = seeds[i];
    if (!s.query || s.query_len < 0 || s.query_seed_start < 0 || s.query_seed_end < s.query_seed_start ||
        s.query_seed_end >= s.query_len)
    {
      cleanup();
      return fail(RAMCU_ERR_ARG, "ramcu_seed_extend_batch: seed %d: query seed outside the query", i);
    }
    for (int side = 0; side < 2; ++side)   // 0 = left (direction 0), 1 = right (direction 1)
    {
      const int rows = side ? s.query_len - (s.query_seed_end + 1) : s.query_seed_start;
      if (rows <= 0)
        continue;
      forced.resize(rows);
      for (int r = 0; r < rows; ++r)
      {
        const uint8_t b = side ? s.query[s.query_seed_end + r + 1] : s.query[s.query_seed_start - r - 1];
        forced[r] = b > 3 ? 4 : b;
      }
      ramcu_core_t core;
      memset(&core, 0, sizeof(core));
      core.orient = s.orient ? 1 : 0;
      core.left_extendable = core.right_extendable = 1;
      core.left_seq_pos = s.orient ? s.subj_seed_end : s.subj_seed_start;     // bnw_extend.c:307-313
      core.right_seq_pos = s.orient ? s.subj_seed_start : s.subj_seed_end;
      // the seed itself is out of bounds for both extensions (bnw_extend.c:326-332, 386-391)
      const bool low_side = (side == 0) != (s.orient != 0);
      core.lower_seq_bound = low_side ? s.lower_seq_bound : s.subj_seed_end + 1;
      core.upper_seq_bound = low_side ? s.subj_seed_start - 1 : s.upper_seq_bound;
      ramcu_params_t pr;
      memset(&pr, 0, sizeof(pr));
      pr.bandwidth = bandwidth;
      pr.L = rows;
      pr.when_to_stop = 1 << 30;
      int rc = job_create_impl(c, lib, side, 1, &core, sc, &pr, 0, forced.data(), &jobs[2 * i + side], n_row);
      if (rc)
      {
        cleanup();
        return rc;
      }
    }
  }
  std::vector<ramcu_job *> run;
  for (ramcu_job *j : jobs)
    if (j && j->n_ctas > 0)
      run.push_back(j);
  if (!run.empty())
  {
    int rc = ramcu_batch_run(c, (int) run.size(), run.data());
    if (rc)
    {
      cleanup();
      return rc;
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
    {
      cleanup();
      return fail(RAMCU_ERR_CUDA, "seed extension kernel failed: %s", cudaGetErrorString(e));
    }
  }
  std::vector<int> rows_host;
  for (int i = 0; i < n; ++i)
  {
    const ramcu_seed_t &s = seeds[i];
    int score[2] = {0, 0}, seq_idx[2] = {-1, -1}, cons_idx[2] = {-1, -1}, used[2] = {0, 0};
    for (int side = 0; side < 2; ++side)
    {
      ramcu_job *j = jobs[2 * i + side];
      if (!j || j->n_ctas == 0)
        continue;
      rows_host.resize((size_t) 2 * j->pr.L);
      cudaError_t e = cudaMemcpy(rows_host.data(), j->d_seed_rows, rows_host.size() * sizeof(int), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess)
      {
        cleanup();
        return fail(RAMCU_ERR_CUDA, "ramcu_seed_extend_batch: %s", cudaGetErrorString(e));
      }
      seed_scan(rows_host.data(), j->pr.L, y_drop, side == 1, &score[side], &seq_idx[side], &cons_idx[side], &used[side]);
    }
    ramcu_seed_result_t &o = results[i];
    o.left_score = score[0];
    o.right_score = score[1];
    o.left_rows = used[0];
    o.right_rows = used[1];
    o.query_start = s.query_seed_start - (cons_idx[0] + 1);
    o.query_end = s.query_seed_end + (cons_idx[1] + 1);
    // bnw_extend.c:444-449
    o.subj_start = s.subj_seed_start - ((s.orient ? seq_idx[1] : seq_idx[0]) + 1);
    o.subj_end = s.subj_seed_end + ((s.orient ? seq_idx[0] : seq_idx[1]) + 1);
  }
  cleanup();
  return RAMCU_OK;
}

// global_align() (bnw_extend.c:486-649): one right extension of seq2 with seq1 as the given consensus, from
// the top-left corner of the band to its last row; the global score is the best cell of the last row
// (never below 0), the local score the best row value and where it was reached.
extern "C" int ramcu_global_align(ramcu_ctx *c, ramcu_library *lib, int orient, const uint8_t *seq1, int seq1_len,
                                  int64_t seq2_start, int64_t seq2_end, const ramcu_scoring_t *sc,
                                  const int32_t *n_row, int bandwidth, int32_t *global_score, int32_t *local_score,
                                  int32_t *local_pos)
{
  if (!c || !lib || !sc || !n_row || !seq1 || seq1_len <= 0 || !global_score)
    return fail(RAMCU_ERR_ARG, "ramcu_global_align: invalid argument");
  CK(cudaSetDevice(c->device));
  std::vector<uint8_t> forced(seq1_len);
  for (int r = 0; r < seq1_len; ++r)
    forced[r] = seq1[r] > 3 ? 4 : seq1[r];
  ramcu_core_t core;
  memset(&core, 0, sizeof(core));
  core.orient = orient ? 1 : 0;
  core.left_extendable = 0;
  core.right_extendable = 1;
  core.left_seq_pos = core.right_seq_pos = seq2_start - 1;   // bnw_extend.c:601-607
  core.lower_seq_bound = seq2_start;
  core.upper_seq_bound = seq2_end + 1;
  ramcu_params_t pr;
  memset(&pr, 0, sizeof(pr));
  pr.bandwidth = bandwidth;
  pr.L = seq1_len;
  pr.when_to_stop = 1 << 30;
  ramcu_job *j = nullptr;
  int rc = job_create_impl(c, lib, 1, 1, &core, sc, &pr, 0, forced.data(), &j, n_row);
  if (rc)
    return rc;
  rc = ramcu_job_run(c, j);
  std::vector<int> rows((size_t) 2 * seq1_len);
  if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess)
    rc = fail(RAMCU_ERR_CUDA, "ramcu_global_align: kernel failed");
  if (!rc && cudaMemcpy(rows.data(), j->d_seed_rows, rows.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = fail(RAMCU_ERR_CUDA, "ramcu_global_align: copy failed");
  destroy_job(c, j);
  if (rc)
    return rc;
  int best = 0, pos = -1;
  for (int r = 0; r < seq1_len; ++r)
    if (rows[2 * r] > best)
    {
      best = rows[2 * r];
      pos = rows[2 * r + 1];
    }
  *global_score = std::max(0, rows[2 * (seq1_len - 1)]);
  if (local_score) *local_score = best;
  if (local_pos) *local_pos = pos;
  return RAMCU_OK;
}

// ------------------------------------------------------------------------------------------
// -outmat: one extend_alignment() call that also reports, for every processed row and every
// extendable core, the state of the best path into each band cell, the row's best score and the
// band index of its first best cell (what compute_nw_row() writes into pathString and returns in
// the reference's 5th pass, bnw_extend.c:1020-1044, ram_extend.c:1115-1132)
// ------------------------------------------------------------------------------------------
extern "C" int ramcu_extend_paths(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                                  const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                  const ramcu_params_t *pr, uint8_t *master_out, int32_t *ext_len_out,
                                  int32_t *score_add_out, ramcu_result_t *res, int n_rows, uint8_t *path_state_out,
                                  int32_t *row_best_out, int32_t *row_best_idx_out)
{
  if (!pr || !path_state_out || !row_best_out || !row_best_idx_out || n_rows < 0)
    return fail(RAMCU_ERR_ARG, "ramcu_extend_paths: bad argument");
  if (pr->L <= 0 || n_cores <= 0 || n_rows == 0)
    return ramcu_extend(c, lib, direction, n_cores, cores, sc, pr, master_out, ext_len_out, score_add_out, res);
  n_rows = std::min(n_rows, pr->L);
  const size_t W = 2 * (size_t) pr->bandwidth + 1;
  if ((size_t) n_rows * (size_t) n_cores * W > ((size_t) 16 << 30))
    return fail(RAMCU_ERR_UNSUPPORTED, "ramcu_extend_paths: rows x cores x band = %zu cells of path state (limit 16 Gi)",
                (size_t) n_rows * (size_t) n_cores * W);
  ramcu_job *j = nullptr;
  int rc = job_create_impl(c, lib, direction, n_cores, cores, sc, pr, -n_rows, nullptr, &j);
  if (rc)
    return rc;
  auto undo = unwind([&]() { ramcu_job_free(c, j); });
  if (j->n_ctas > 0)
    rc = ramcu_job_run(c, j);
  ramcu_result_t r0;
  if (!rc)
    rc = ramcu_job_fetch(c, j, master_out, ext_len_out, score_add_out, &r0);
  if (!rc && res)
    *res = r0;
  if (!rc && j->n_ctas > 0 && r0.rows_processed > 0)
  {
    const size_t rows = (size_t) std::min(r0.rows_processed, n_rows), ne = (size_t) j->n_ext;
    std::vector<unsigned char> path(rows * ne * W), cons(rows);
    std::vector<int> best(rows * ne * 4), idx(rows * ne);
    CK(cudaMemcpy(path.data(), j->d_dbg_path, path.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(best.data(), j->d_dbg_best, best.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(idx.data(), j->d_dbg_idx, idx.size() * sizeof(int), cudaMemcpyDeviceToHost));
    memcpy(cons.data(), j->h_out + (8 + 2 * (size_t) j->n_ext) * sizeof(int), rows);
    for (size_t r = 0; r < rows; ++r)
      for (size_t e = 0; e < ne; ++e)
      {
        const size_t i = (size_t) j->orig[e];
        memcpy(path_state_out + (r * (size_t) n_cores + i) * W, path.data() + (r * ne + e) * W, W);
        row_best_out[r * (size_t) n_cores + i] = best[(r * ne + e) * 4 + (cons[r] & 3)];
        row_best_idx_out[r * (size_t) n_cores + i] = idx[r * ne + e];
      }
  }
  return rc;   // the job is released by `undo`
}

// ------------------------------------------------------------------------------------------
// VERBOSE >= 10 dumps: one extend_alignment() call that also reports what compute_nw_row() returned and
// left behind for EVERY candidate base of every processed row (ram_extend.c:992-1090, 1139-1141)
// ------------------------------------------------------------------------------------------
extern "C" int ramcu_extend_verbose(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                                    const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                    const ramcu_params_t *pr, uint8_t *master_out, int32_t *ext_len_out,
                                    int32_t *score_add_out, ramcu_result_t *res, int n_rows, int32_t *cand_out,
                                    int32_t *sums_out, int32_t *cells_out)
{
  if (!pr || !cand_out || !sums_out || n_rows < 0)
    return fail(RAMCU_ERR_ARG, "ramcu_extend_verbose: bad argument");
  if (pr->L <= 0 || n_cores <= 0 || n_rows == 0)
    return ramcu_extend(c, lib, direction, n_cores, cores, sc, pr, master_out, ext_len_out, score_add_out, res);
  const size_t W = 2 * (size_t) pr->bandwidth + 1;
  const size_t per_row = (size_t) n_cores * (cells_out ? 4 * W * 2 + 16 : 16) * sizeof(int);
  if ((size_t) n_rows * per_row > ((size_t) 16 << 30))
    return fail(RAMCU_ERR_UNSUPPORTED, "ramcu_extend_verbose: %d rows x %d cores = %zu bytes of dump state (limit 16 GiB)",
                n_rows, n_cores, (size_t) n_rows * per_row);
  ramcu_job *j = nullptr;
  int rc = job_create_impl(c, lib, direction, n_cores, cores, sc, pr, 0, nullptr, &j, nullptr, false, n_rows,
                           cells_out != nullptr);
  if (rc)
    return rc;
  auto undo = unwind([&]() { ramcu_job_free(c, j); });
  if (j->n_ctas > 0)
    rc = ramcu_job_run(c, j);
  ramcu_result_t r0;
  if (!rc)
    rc = ramcu_job_fetch(c, j, master_out, ext_len_out, score_add_out, &r0);
  if (!rc && res)
    *res = r0;
  if (!rc && j->n_ctas > 0 && r0.rows_processed > 0)
  {
    const size_t rows = (size_t) std::min(r0.rows_processed, n_rows), ne = (size_t) j->n_ext;
    std::vector<int> cand(rows * ne * 12), best(rows * ne * 4);
    CK(cudaMemcpy(cand.data(), j->d_dbg_cand, cand.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(best.data(), j->d_dbg_best, best.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sums_out, j->d_dbg_sums, rows * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> cells;
    if (cells_out)
    {
      cells.resize(rows * ne * 4 * W * 2);
      CK(cudaMemcpy(cells.data(), j->d_dbg_cells, cells.size() * sizeof(int), cudaMemcpyDeviceToHost));
    }
    for (size_t r = 0; r < rows; ++r)
      for (size_t e = 0; e < ne; ++e)
      {
        const size_t i = (size_t) j->orig[e];
        for (int a = 0; a < 4; ++a)
        {
          int32_t *o = cand_out + ((r * (size_t) n_cores + i) * 4 + a) * 4;
          o[0] = best[(r * ne + e) * 4 + a];
          o[1] = cand[((r * ne + e) * 4 + a) * 3];
          o[2] = cand[((r * ne + e) * 4 + a) * 3 + 1];
          o[3] = cand[((r * ne + e) * 4 + a) * 3 + 2];
        }
        if (cells_out)
          memcpy(cells_out + (r * (size_t) n_cores + i) * 4 * W * 2, cells.data() + (r * ne + e) * 4 * W * 2,
                 4 * W * 2 * sizeof(int));
      }
  }
  return rc;   // the job is released by `undo`
}

// ------------------------------------------------------------------------------------------
// test-only: forced-consensus rows (compute_nw_row boundary)
// ------------------------------------------------------------------------------------------
extern "C" int ramcu_debug_rows(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                                const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                const ramcu_params_t *pr, int n_rows, const uint8_t *forced,
                                int32_t *rows_out, int32_t *best_out, int32_t *sums_out)
{
  if (n_rows <= 0)
    return fail(RAMCU_ERR_ARG, "ramcu_debug_rows: n_rows <= 0");
  ramcu_params_t p = *pr;
  p.L = n_rows;
  p.when_to_stop = 1 << 30;
  ramcu_job *j = nullptr;
  int rc = job_create_impl(c, lib, direction, n_cores, cores, sc, &p, n_rows, forced, &j);
  if (rc)
    return rc;
  const int W = 2 * p.bandwidth + 1;
  if (j->n_ctas > 0)
  {
    rc = ramcu_job_run(c, j);
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess)
      rc = fail(RAMCU_ERR_CUDA, "ramcu_debug_rows: kernel failed: %s",
                cudaGetErrorString(cudaGetLastError()));
    if (!rc)
    {
      std::vector<int> rows((size_t) n_rows * j->n_ext * W * 2), best((size_t) n_rows * j->n_ext * 4);
      cudaMemcpy(rows.data(), j->d_dbg_rows, rows.size() * sizeof(int), cudaMemcpyDeviceToHost);
      cudaMemcpy(best.data(), j->d_dbg_best, best.size() * sizeof(int), cudaMemcpyDeviceToHost);
      if (sums_out)
        cudaMemcpy(sums_out, j->d_dbg_sums, (size_t) n_rows * 4 * sizeof(int), cudaMemcpyDeviceToHost);
      for (int r = 0; r < n_rows; ++r)
        for (int e = 0; e < j->n_ext; ++e)
        {
          const int i = j->orig[e];
          if (rows_out)
            memcpy(rows_out + ((size_t) r * n_cores + i) * W * 2,
                   rows.data() + ((size_t) r * j->n_ext + e) * W * 2, (size_t) W * 2 * sizeof(int));
          if (best_out)
            memcpy(best_out + ((size_t) r * n_cores + i) * 4,
                   best.data() + ((size_t) r * j->n_ext + e) * 4, 4 * sizeof(int));
        }
    }
  }
  ramcu_job_free(c, j);
  return rc;
}
