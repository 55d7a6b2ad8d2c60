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
struct ramcu_ctx
{
  int device;
  int sm_count, cc_major, cc_minor;
  cudaStream_t own_stream, stream;
  cudaEvent_t ev0, ev1;
  int *d_ticket;
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
  const KernelCfg *cfg;
  int n_ctas;                 // > 0 when there is anything to launch (== solo.n_ctas)
  Plan solo;                  // layout for ramcu_job_run
  int batch_ctas;             // layout for ramcu_batch_run (kBatchThreads per CTA)
  std::vector<int> orig;      // ext index -> core index
  FamDev fam;                 // host copy (device pointers inside)
  // device
  InstDev *d_inst;
  unsigned char *d_master;
  int *d_thigh, *d_tpos, *d_result;
  unsigned long long *d_slots;
  FamDev *d_fam;      // single-family launch
  int *d_cta_fam;     // zeros, n_ctas
  unsigned char *d_forced;
  int *d_dbg_rows, *d_dbg_best, *d_dbg_sums;
  unsigned char *d_dbg_path;   // -outmat mode (negative dbg_rows in job_create_impl)
  int *d_dbg_del, *d_dbg_idx;
  long long *d_dbg_wait;   // RAMCU_STATS=1 diagnostics
  bool ran, timed;
};

// ------------------------------------------------------------------------------------------
// basic entry points
// ------------------------------------------------------------------------------------------
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
  CK(cudaMalloc(&c->d_ticket, sizeof(int)));
  {
    // per-call buffers come from the stream-ordered pool; keep freed memory cached in the pool
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  collect_cfgs();
  for (KernelCfg *k : g_cfgs)
  {
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, (const void *) k->fn);
    if (e != cudaSuccess)
    {
      delete c;
      return fail(RAMCU_ERR_CUDA,
                  "ramcu_init: kernel image not loadable on this device (built for sm_100a): %s",
                  cudaGetErrorString(e));
    }
    k->regs = fa.numRegs;
  }
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
  lib->length = length;
  const uint64_t n_chunks = (length + 31) / 32;  // 32 bases per chunk
  const size_t pad = 4;                           // words of slack on both sides
  CK(cudaMallocAsync(&lib->alloc2, (n_chunks * 2 + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMallocAsync(&lib->allocn, (n_chunks + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->alloc2, 0, (n_chunks * 2 + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->allocn, 0xff, (n_chunks + 2 * pad) * sizeof(uint32_t), c->stream));
  lib->d_lib2 = (uint32_t *) lib->alloc2 + pad;
  lib->d_libn = (uint32_t *) lib->allocn + pad;
  if (n_chunks > 0)
  {
    const unsigned long long want = (n_chunks + 255) / 256;
    const int grid = (int) std::min<unsigned long long>(want, (unsigned long long) c->sm_count * 16);
    pack_library_kernel<<<grid, 256, 0, c->stream>>>(d_bytes, length, lib->d_lib2, lib->d_libn,
                                                     n_chunks);
    CK(cudaGetLastError());
  }
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
  CK(cudaMallocAsync(&d_bytes, length ? length : 1, c->stream));
  CK(cudaMemcpyAsync(d_bytes, seq, length, cudaMemcpyHostToDevice, c->stream));
  int rc = library_pack(c, d_bytes, length, out);
  CK(cudaFreeAsync(d_bytes, c->stream));
  CK(cudaStreamSynchronize(c->stream));   // the caller may reuse / free its host buffer
  return rc;
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
  unsigned long long *d_ns = nullptr, *d_nl = nullptr;
  CK(cudaMallocAsync(&lib->alloc2, (n_chunks * 2 + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMallocAsync(&lib->allocn, (n_chunks + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->alloc2, 0, (n_chunks * 2 + 2 * pad) * sizeof(uint32_t), c->stream));
  CK(cudaMemsetAsync(lib->allocn, 0xff, (n_chunks + 2 * pad) * sizeof(uint32_t), c->stream));
  lib->d_lib2 = (uint32_t *) lib->alloc2 + pad;
  lib->d_libn = (uint32_t *) lib->allocn + pad;
  if (n_chunks > 0)
  {
    // the tail bits of the last mask word stay set (positions past the library read as N)
    CK(cudaMemsetAsync(lib->d_libn, 0, n_chunks * sizeof(uint32_t), c->stream));
    CK(cudaMallocAsync(&d_blob, blob_bytes ? blob_bytes : 1, c->stream));
    CK(cudaMemcpyAsync(d_blob, blob, blob_bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMallocAsync(&d_win, (size_t) std::max(n_windows, 1) * sizeof(WindowDev), c->stream));
    CK(cudaMemcpyAsync(d_win, windows, (size_t) n_windows * sizeof(WindowDev), cudaMemcpyHostToDevice, c->stream));
    const unsigned long long want = ((n_chunks + 3) / 4 + 255) / 256;   // one thread per 128 bases
    const int grid = (int) std::min<unsigned long long>(want, (unsigned long long) c->sm_count * 16);
    // groups of 128 bases that are not inside one window: at most one per window boundary + the tail
    unsigned long long *d_slow = nullptr;
    unsigned int *d_nslow = nullptr;
    CK(cudaMallocAsync(&d_slow, ((size_t) n_windows + 4) * sizeof(unsigned long long), c->stream));
    CK(cudaMallocAsync(&d_nslow, sizeof(unsigned int), c->stream));
    CK(cudaMemsetAsync(d_nslow, 0, sizeof(unsigned int), c->stream));
    twobit_windows_kernel<<<grid, 256, 0, c->stream>>>(d_blob, blob_bytes, d_win, n_windows, total,
                                                       lib->d_lib2, n_chunks, d_slow, d_nslow);
    CK(cudaGetLastError());
    {
      const unsigned long long wantb = (4ull * ((unsigned long long) n_windows + 4) + 127) / 128;
      const int gridb = (int) std::min<unsigned long long>(wantb, (unsigned long long) c->sm_count * 16);
      twobit_boundary_kernel<<<gridb, 128, 0, c->stream>>>(d_blob, blob_bytes, d_win, n_windows, total,
                                                           lib->d_lib2, n_chunks, d_slow, d_nslow);
      CK(cudaGetLastError());
    }
    CK(cudaFreeAsync(d_slow, c->stream));
    CK(cudaFreeAsync(d_nslow, c->stream));
    // positions past `total` inside the last chunk are N
    if (total & 31)
    {
      const unsigned long long s0 = total, l0 = 32 - (total & 31);
      CK(cudaMallocAsync(&d_ns, 2 * sizeof(unsigned long long), c->stream));
      const unsigned long long tail[2] = {s0, l0};
      CK(cudaMemcpyAsync(d_ns, tail, sizeof(tail), cudaMemcpyHostToDevice, c->stream));
      nmask_ranges_kernel<<<1, 128, 0, c->stream>>>(d_ns, d_ns + 1, 1, lib->d_libn);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(c->stream));   // `tail` is on the stack
      CK(cudaFreeAsync(d_ns, c->stream));
      d_ns = nullptr;
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
      CK(cudaFreeAsync(d_nl, c->stream));
    }
    CK(cudaFreeAsync(d_blob, c->stream));
    CK(cudaFreeAsync(d_win, c->stream));
    CK(cudaStreamSynchronize(c->stream));   // the caller may reuse its host buffers
  }
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
static int ctas_per_sm(const KernelCfg *k, int threads)
{
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void *) k->fn, threads, 0) != cudaSuccess)
    return 0;
  return nb;
}

// Layout of one family launched on its own: a single CTA when it fits (no grid exchange at all),
// else 4-warp CTAs on distinct SMs, else the same number of equally loaded CTAs on every SM (the
// column time is set by the most loaded scheduler, and every CTA of the family waits for it).
static bool plan_solo(const ramcu_ctx *c, const KernelCfg *k, int n_ext, Plan *p)
{
  const int ipw = inst_per_warp(k->G);
  const long long warps = ((long long) n_ext + ipw - 1) / ipw;
  const int sm = c->sm_count;
  if (warps * 32 <= k->maxt)
  {
    p->n_ctas = 1;
    p->threads = (int) warps * 32;
    return true;
  }
  if (warps <= (long long) sm * 4)
  {
    p->n_ctas = (int) std::min<long long>((warps + 3) / 4, n_ext);
    p->threads = cta_threads(n_ext, p->n_ctas, k->G);
    return p->threads <= k->maxt;
  }
  for (int per_sm = 1; per_sm <= 16; ++per_sm)
  {
    const long long n = (long long) sm * per_sm;
    if (n > n_ext)
      break;
    const int t = cta_threads(n_ext, (int) n, k->G);
    if (t > k->maxt)
      continue;
    if (ctas_per_sm(k, t) < per_sm)
      continue;
    p->n_ctas = (int) n;
    p->threads = t;
    return true;
  }
  return false;
}

// Layout when the family shares a launch with others: fixed CTA size.
static bool plan_batch(const ramcu_ctx *c, const KernelCfg *k, int n_ext, int *n_ctas)
{
  const int per_cta = (kBatchThreads / 32) * inst_per_warp(k->G);
  const long long n = ((long long) n_ext + per_cta - 1) / per_cta;
  if (n > (long long) ctas_per_sm(k, kBatchThreads) * c->sm_count)
    return false;   // its CTAs could not all be resident: the column exchange would deadlock
  *n_ctas = (int) n;
  return true;
}

// Rough cycles per column of a family running alone (tuned on B200): a warp
// executes about C*35 + 450 instructions per row; one warp alone issues one every ~4.5 cycles
// (dependent max-plus chains), a loaded scheduler one every ~2.2 cycles; the busiest of the
// four schedulers of the busiest SM sets the pace.
static double solo_cost(const ramcu_ctx *c, const KernelCfg *k, int n_ext, const Plan &p)
{
  // a tiling squeezed below ~6 registers per cell keeps part of its rows in local memory
  const double squeeze = (65536.0 / k->maxt) < 6.0 * k->C + 60.0 ? 1.5 : 1.0;
  const double instr = (k->C * 35.0 + 450.0) * squeeze;
  const int warps_per_cta = p.threads / 32;
  const int ctas_on_sm = (p.n_ctas + c->sm_count - 1) / c->sm_count;
  const double per_sched = std::ceil(warps_per_cta * ctas_on_sm / 4.0);
  const double sync = p.n_ctas > 1 ? 2000.0 : 0.0;
  (void) n_ext;
  return std::max(instr * 4.5, per_sched * instr * 2.2) + sync;
}

static const KernelCfg *pick_cfg(const ramcu_ctx *c, int B, int n_ext, int force_G,
                                 int concurrent_instances, int force_maxt, Plan *solo, int *batch_ctas)
{
  const int W = 2 * B + 1;
  std::vector<const KernelCfg *> cand;
  for (const KernelCfg *k : g_cfgs)
    if (k->B == B)
      cand.push_back(k);
  if (cand.empty())
    for (const KernelCfg *k : g_cfgs)
      if (k->B < 0 && k->C * k->G >= W && (k->G - 1) * k->C < W)
        cand.push_back(k);
  if (cand.empty())
    for (const KernelCfg *k : g_cfgs)  // narrow bands: trailing lanes may be entirely padding
      if (k->B < 0 && k->C * k->G >= W)
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
    have_maxt = have_maxt || (k->maxt == force_maxt && (force_G <= 0 || k->G == force_G));
  for (const KernelCfg *k : cand)
  {
    if (force_G > 0 && k->G != force_G)
      continue;
    if (force_maxt > 0 && k->maxt != force_maxt && have_maxt)
      continue;
    Plan p;
    int nb = 0;
    if (!plan_solo(c, k, n_ext, &p))
      continue;
    if (!plan_batch(c, k, n_ext, &nb))
      nb = 0;   // too large to share a launch: ramcu_batch_run launches it on its own
    // a crowded device is throughput bound: fewest lanes per instance = fewest instructions
    const double cost = crowded ? (double) k->G + (k->maxt > maxt_for(k->C) ? 0.5 : 0.0) : solo_cost(c, k, n_ext, p);
    if (!best || cost < best_cost)
    {
      best = k;
      best_cost = cost;
      *solo = p;
      *batch_ctas = nb;
    }
  }
  return best;
}

static size_t slots_bytes(int n_ctas, int L)
{
  const size_t words = n_ctas > kSlotMaxCtas ? ((size_t) L + 1) * 4 * kRedStride : (size_t) 2 * std::max(n_ctas, 1) * 4;
  return words * sizeof(unsigned long long);
}

static void free_job_device(ramcu_ctx *c, ramcu_job *j)
{
  void *ptrs[] = {j->d_inst, j->d_master, j->d_thigh, j->d_tpos, j->d_result, j->d_slots, j->d_fam,
                  j->d_cta_fam, j->d_forced, j->d_dbg_rows, j->d_dbg_best, j->d_dbg_sums, j->d_dbg_wait,
                  j->d_dbg_path, j->d_dbg_del, j->d_dbg_idx};
  for (void *p : ptrs)
    if (p)
    {
      if (c)
        cudaFreeAsync(p, c->stream);
      else
        cudaFree(p);
    }
}

static int job_create_impl(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                           const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                           const ramcu_params_t *pr, int dbg_rows, const uint8_t *forced,
                           ramcu_job **out)
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
  j->ran = false;
  j->timed = false;

  const int B = pr->bandwidth;
  std::vector<InstDev> inst;
  flatten_cores(j->direction, n_cores, cores, lib->length, inst);
  j->n_ext = (int) inst.size();
  j->orig.resize(inst.size());
  for (size_t i = 0; i < inst.size(); ++i)
    j->orig[i] = inst[i].orig;

  j->cfg = nullptr;
  j->n_ctas = 0;
  if (j->n_ext > 0 && pr->L > 0)
  {
    j->cfg = pick_cfg(c, B, j->n_ext, pr->lanes_per_instance, pr->concurrent_instances, pr->cta_thread_limit, &j->solo,
                      &j->batch_ctas);
    if (!j->cfg)
    {
      delete j;
      return fail(RAMCU_ERR_CAPACITY,
                  "no band tiling for bandwidth %d with %d instances fits this device "
                  "(lanes_per_instance=%d)", B, (int) inst.size(), pr->lanes_per_instance);
    }
    j->n_ctas = j->solo.n_ctas;
    const int max_ctas = std::max(j->solo.n_ctas, j->batch_ctas);

    const int W = 2 * B + 1;
    CK(cudaMallocAsync(&j->d_inst, inst.size() * sizeof(InstDev), c->stream));
    CK(cudaMemcpyAsync(j->d_inst, inst.data(), inst.size() * sizeof(InstDev),
                       cudaMemcpyHostToDevice, c->stream));
    CK(cudaMallocAsync(&j->d_master, std::max(pr->L, 1), c->stream));
    CK(cudaMallocAsync(&j->d_thigh, inst.size() * sizeof(int), c->stream));
    CK(cudaMallocAsync(&j->d_tpos, inst.size() * sizeof(int), c->stream));
    CK(cudaMallocAsync(&j->d_result, 8 * sizeof(int), c->stream));
    CK(cudaMallocAsync(&j->d_slots, std::max(slots_bytes(j->solo.n_ctas, pr->L), slots_bytes(j->batch_ctas, pr->L)), c->stream));
    CK(cudaMallocAsync(&j->d_fam, sizeof(FamDev), c->stream));
    CK(cudaMallocAsync(&j->d_cta_fam, max_ctas * sizeof(int), c->stream));
    CK(cudaMemsetAsync(j->d_cta_fam, 0, max_ctas * sizeof(int), c->stream));
    if (dbg_rows < 0)   // -outmat mode: path states of -dbg_rows rows, no forced consensus
    {
      const size_t rows = (size_t) -dbg_rows;
      CK(cudaMallocAsync(&j->d_dbg_path, rows * j->n_ext * W, c->stream));
      CK(cudaMemsetAsync(j->d_dbg_path, 0, rows * j->n_ext * W, c->stream));
      CK(cudaMallocAsync(&j->d_dbg_del, (size_t) j->n_ext * W * sizeof(int), c->stream));
      CK(cudaMallocAsync(&j->d_dbg_idx, rows * j->n_ext * sizeof(int), c->stream));
      CK(cudaMemsetAsync(j->d_dbg_idx, 0, rows * j->n_ext * sizeof(int), c->stream));
      CK(cudaMallocAsync(&j->d_dbg_best, rows * j->n_ext * 4 * sizeof(int), c->stream));
      CK(cudaMallocAsync(&j->d_dbg_sums, rows * 4 * sizeof(int), c->stream));
    }
    if (dbg_rows > 0)
    {
      CK(cudaMallocAsync(&j->d_dbg_rows, (size_t) dbg_rows * j->n_ext * W * 2 * sizeof(int), c->stream));
      CK(cudaMallocAsync(&j->d_dbg_best, (size_t) dbg_rows * j->n_ext * 4 * sizeof(int), c->stream));
      CK(cudaMallocAsync(&j->d_dbg_sums, (size_t) dbg_rows * 4 * sizeof(int), c->stream));
      CK(cudaMemsetAsync(j->d_dbg_rows, 0, (size_t) dbg_rows * j->n_ext * W * 2 * sizeof(int), c->stream));
      if (forced)
      {
        CK(cudaMallocAsync(&j->d_forced, dbg_rows, c->stream));
        CK(cudaMemcpyAsync(j->d_forced, forced, dbg_rows, cudaMemcpyHostToDevice, c->stream));
      }
    }
    if (getenv("RAMCU_STATS"))
    {
      CK(cudaMallocAsync(&j->d_dbg_wait, (size_t) 2 * max_ctas * sizeof(long long), c->stream));
      CK(cudaMemsetAsync(j->d_dbg_wait, 0, (size_t) 2 * max_ctas * sizeof(long long), c->stream));
    }
    FamDev &f = j->fam;
    fill_family(f, sc, pr, j->n_ext, j->n_ctas);
    f.dbg_wait = j->d_dbg_wait;
    f.lib2 = lib->d_lib2;
    f.libn = lib->d_libn;
    f.inst = j->d_inst;
    f.master = j->d_master;
    f.thigh = j->d_thigh;
    f.tpos = j->d_tpos;
    f.result = j->d_result;
    f.slots = j->d_slots;
    f.forced = j->d_forced;
    f.dbg_rows = j->d_dbg_rows;
    f.dbg_path = j->d_dbg_path;
    f.dbg_del = j->d_dbg_del;
    f.dbg_idx = j->d_dbg_idx;
    f.dbg_best = j->d_dbg_best;
    f.dbg_sums = j->d_dbg_sums;
    CK(cudaMemcpyAsync(j->d_fam, &f, sizeof(f), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // `inst` and `f` are stack/heap temporaries
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

static int launch_single(ramcu_ctx *c, ramcu_job *j)
{
  CK(cudaMemsetAsync(c->d_ticket, 0, sizeof(int), c->stream));
  if (j->n_ctas > 1)
    CK(cudaMemsetAsync(j->d_slots, 0, slots_bytes(j->solo.n_ctas, j->pr.L), c->stream));
  j->cfg->fn<<<j->solo.n_ctas, j->solo.threads, 0, c->stream>>>(j->d_fam, j->d_cta_fam, c->d_ticket);
  CK(cudaGetLastError());
  return RAMCU_OK;
}

extern "C" int ramcu_job_run(ramcu_ctx *c, ramcu_job *j)
{
  if (!c || !j)
    return fail(RAMCU_ERR_ARG, "ramcu_job_run: NULL argument");
  CK(cudaSetDevice(c->device));
  j->ran = true;
  j->timed = false;
  if (j->n_ctas == 0)
    return RAMCU_OK;
  CK(cudaEventRecord(c->ev0, c->stream));
  int rc = launch_single(c, j);
  if (rc)
    return rc;
  CK(cudaEventRecord(c->ev1, c->stream));
  j->timed = true;
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
  CK(cudaStreamSynchronize(c->stream));
  int r4[8];
  CK(cudaMemcpy(r4, j->d_result, 4 * sizeof(int), cudaMemcpyDeviceToHost));
  res->extended_bp = r4[0] + 1;
  res->rows_processed = r4[1];
  res->stop_row = r4[2];
  res->max_score = r4[3];
  res->hit_limit = (r4[2] >= 0 && r4[2] == j->pr.L - 1) ? 1 : 0;
  res->lanes_per_instance = j->cfg->G;
  if (j->timed)
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess)
      res->kernel_ms = ms;
  }
  if (j->d_dbg_wait)
  {
    const int n_stat = std::max(j->solo.n_ctas, j->batch_ctas);
    std::vector<long long> w((size_t) 2 * n_stat);
    CK(cudaMemcpy(w.data(), j->d_dbg_wait, w.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemset(j->d_dbg_wait, 0, w.size() * sizeof(long long)));
    long long wmin = 1LL << 62, wmax = 0, wsum = 0, tmax = 0;
    int n_seen = 0;
    for (int i = 0; i < n_stat; ++i)
    {
      if (w[2 * i + 1] == 0)
        continue;   // CTA index not used by the layout that ran
      ++n_seen;
      wmin = std::min(wmin, w[2 * i]);
      wmax = std::max(wmax, w[2 * i]);
      wsum += w[2 * i];
      tmax = std::max(tmax, w[2 * i + 1]);
    }
    fprintf(stderr, "[ramcu stats] ctas=%d rows=%d cycles/row: total %.0f, exchange wait min %.0f avg %.0f max %.0f\n",
            n_seen, r4[1], (double) tmax / std::max(r4[1], 1), (double) wmin / std::max(r4[1], 1),
            (double) wsum / std::max(n_seen, 1) / std::max(r4[1], 1), (double) wmax / std::max(r4[1], 1));
  }
  if (master_out && r4[1] > 0)
    CK(cudaMemcpy(master_out, j->d_master, r4[1], cudaMemcpyDeviceToHost));
  if (ext_len_out || score_add_out)
  {
    std::vector<int> th(j->n_ext), tp(j->n_ext);
    CK(cudaMemcpy(th.data(), j->d_thigh, j->n_ext * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tp.data(), j->d_tpos, j->n_ext * sizeof(int), cudaMemcpyDeviceToHost));
    map_results(j->n_cores, j->orig, th.data(), tp.data(), ext_len_out, score_add_out);
  }
  return RAMCU_OK;
}

extern "C" int ramcu_job_free(ramcu_ctx *c, ramcu_job *j)
{
  if (!j)
    return RAMCU_OK;
  if (c)
    cudaSetDevice(c->device);
  free_job_device(c, j);
  delete j;
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
extern "C" int ramcu_batch_run(ramcu_ctx *c, int n_jobs, ramcu_job **jobs)
{
  if (!c || (n_jobs > 0 && !jobs))
    return fail(RAMCU_ERR_ARG, "ramcu_batch_run: NULL argument");
  CK(cudaSetDevice(c->device));
  // group by kernel instantiation; one launch per group, families ordered by decreasing work
  std::vector<int> order;
  for (int i = 0; i < n_jobs; ++i)
  {
    if (!jobs[i])
      return fail(RAMCU_ERR_ARG, "ramcu_batch_run: jobs[%d] == NULL", i);
    jobs[i]->ran = true;
    jobs[i]->timed = false;
    if (jobs[i]->n_ctas > 0 && jobs[i]->batch_ctas > 0)
      order.push_back(i);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (jobs[a]->cfg != jobs[b]->cfg)
      return jobs[a]->cfg < jobs[b]->cfg;
    const long long wa = (long long) jobs[a]->n_ext * jobs[a]->pr.L;
    const long long wb = (long long) jobs[b]->n_ext * jobs[b]->pr.L;
    return wa > wb;
  });
  CK(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < n_jobs; ++i)
    if (jobs[i]->n_ctas > 0 && jobs[i]->batch_ctas == 0)
    {
      int rc = launch_single(c, jobs[i]);
      if (rc)
        return rc;
    }
  size_t p = 0;
  while (p < order.size())
  {
    size_t q = p;
    const KernelCfg *cfg = jobs[order[p]]->cfg;
    std::vector<FamDev> fams;
    std::vector<int> cta_fam;
    while (q < order.size() && jobs[order[q]]->cfg == cfg)
    {
      ramcu_job *j = jobs[order[q]];
      FamDev f = j->fam;
      f.cta0 = (int) cta_fam.size();
      f.n_ctas = j->batch_ctas;
      for (int k = 0; k < j->batch_ctas; ++k)
        cta_fam.push_back((int) fams.size());
      fams.push_back(f);
      if (j->batch_ctas > 1)
        CK(cudaMemsetAsync(j->d_slots, 0, slots_bytes(j->batch_ctas, j->pr.L), c->stream));
      ++q;
    }
    FamDev *d_fams = nullptr;
    int *d_map = nullptr;
    CK(cudaMallocAsync(&d_fams, fams.size() * sizeof(FamDev), c->stream));
    CK(cudaMallocAsync(&d_map, cta_fam.size() * sizeof(int), c->stream));
    CK(cudaMemcpyAsync(d_fams, fams.data(), fams.size() * sizeof(FamDev), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_map, cta_fam.data(), cta_fam.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->d_ticket, 0, sizeof(int), c->stream));
    // pageable sources: the copies above have completed on return, so the vectors may go away
    cfg->fn<<<(int) cta_fam.size(), kBatchThreads, 0, c->stream>>>(d_fams, d_map, c->d_ticket);
    CK(cudaGetLastError());
    CK(cudaFreeAsync(d_fams, c->stream));
    CK(cudaFreeAsync(d_map, c->stream));
    p = q;
  }
  CK(cudaEventRecord(c->ev1, c->stream));
  for (int i = 0; i < n_jobs; ++i)
    jobs[i]->timed = true;
  return RAMCU_OK;
}

// ------------------------------------------------------------------------------------------
// both passes of one family in one launch
// ------------------------------------------------------------------------------------------
// main() extends right, narrows the bounds of cores that a right extension ran into
// (ram_extend.c:445-499), then extends left with those bounds. When that step leaves every bound
// as it was (the caller checks: ramx_boundary.c compares the bounds it is handed for the left
// pass with the ones used here), the two passes are independent column loops over the same
// cores. Running them as two families of one launch puts CTAs of both on every SM, so one pass
// computes while the other waits for its per-column grid exchange.
extern "C" int ramcu_job_create_pair(ramcu_ctx *c, ramcu_library *lib, int n_cores,
                                     const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                     const ramcu_params_t *pr, ramcu_job **right, ramcu_job **left)
{
  if (!c || !pr || !right || !left || (n_cores > 0 && !cores))
    return fail(RAMCU_ERR_ARG, "ramcu_job_create_pair: NULL argument");
  *right = *left = nullptr;
  int n_r = 0, n_l = 0;
  for (int i = 0; i < n_cores; ++i)
  {
    n_r += cores[i].right_extendable ? 1 : 0;
    n_l += cores[i].left_extendable ? 1 : 0;
  }
  ramcu_params_t p = *pr;
  bool together = false;
  if (n_r > 0 && n_l > 0 && pr->L > 0 && pr->bandwidth >= 0 && pr->bandwidth <= RAMCU_MAX_BANDWIDTH)
  {
    CK(cudaSetDevice(c->device));
    Plan solo;
    int nb = 0;
    const KernelCfg *k0 = pick_cfg(c, pr->bandwidth, std::max(n_r, n_l), pr->lanes_per_instance,
                                   pr->concurrent_instances, pr->cta_thread_limit, &solo, &nb);
    if (k0)
    {
      // same lanes per instance as the family alone, with full registers if both passes fit the
      // device side by side, else the tiling compiled for half the registers; failing that, the
      // cheapest other half-register tiling of this band (by the single-family cost model) that fits
      auto fits = [&](int G, int maxt, double *cost) -> bool {
        Plan sr, sl;
        int br = 0, bl = 0;
        const KernelCfg *kr = pick_cfg(c, pr->bandwidth, n_r, G, pr->concurrent_instances, maxt, &sr, &br);
        const KernelCfg *kl = pick_cfg(c, pr->bandwidth, n_l, G, pr->concurrent_instances, maxt, &sl, &bl);
        if (!kr || kr != kl || kr->maxt != maxt || kr->G != G || br <= 0 || bl <= 0)
          return false;
        if ((long long) br + bl > (long long) ctas_per_sm(kr, kBatchThreads) * c->sm_count)
          return false;
        if (cost)
          *cost = solo_cost(c, kr, std::max(n_r, n_l), n_r >= n_l ? sr : sl);
        return true;
      };
      const int tries[2] = {k0->maxt, 512};
      for (int t = 0; t < 2 && !together; ++t)
      {
        if (pr->cta_thread_limit > 0 && tries[t] != pr->cta_thread_limit)
          continue;
        if (!fits(k0->G, tries[t], nullptr))
          continue;
        p.lanes_per_instance = k0->G;
        p.cta_thread_limit = tries[t];
        together = true;
      }
      if (!together && pr->lanes_per_instance <= 0 && (pr->cta_thread_limit <= 0 || pr->cta_thread_limit == 512))
      {
        double best_cost = 0;
        for (int G = 1; G <= 32; ++G)
        {
          double cost = 0;
          if (G == k0->G || !fits(G, 512, &cost))
            continue;
          if (!together || cost < best_cost)
          {
            best_cost = cost;
            p.lanes_per_instance = G;
            p.cta_thread_limit = 512;
            together = true;
          }
        }
      }
    }
  }
  int rc = job_create_impl(c, lib, 1, n_cores, cores, sc, &p, 0, nullptr, right);
  if (!rc)
    rc = job_create_impl(c, lib, 0, n_cores, cores, sc, &p, 0, nullptr, left);
  if (rc)
  {
    ramcu_job_free(c, *right);
    *right = nullptr;
    return rc;
  }
  if (!together)   // ramcu_batch_run then launches each pass on its own, one after the other
    (*right)->batch_ctas = (*left)->batch_ctas = 0;

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
// -outmat: one extend_alignment() call that also reports, for every processed row and every
// extendable core, the state of the best path into each band cell, the row's best score and the
// band index of its first best cell (what compute_nw_row() writes into pathString and returns in
// the reference's 5th pass, bnw_extend.c:1020-1044, ram_extend.c:1115-1132)
// ------------------------------------------------------------------------------------------
extern "C" int ramcu_extend_paths(ramcu_ctx *c, ramcu_library *lib, int direction, int n_cores,
                                  const ramcu_core_t *cores, const ramcu_scoring_t *sc,
                                  const ramcu_params_t *pr, uint8_t *master_out, int32_t *ext_len_out,
                                  int32_t *score_add_out, ramcu_result_t *res, uint8_t *path_state_out,
                                  int32_t *row_best_out, int32_t *row_best_idx_out)
{
  if (!pr || !path_state_out || !row_best_out || !row_best_idx_out)
    return fail(RAMCU_ERR_ARG, "ramcu_extend_paths: NULL argument");
  if (pr->L <= 0 || n_cores <= 0)
    return ramcu_extend(c, lib, direction, n_cores, cores, sc, pr, master_out, ext_len_out, score_add_out, res);
  const size_t W = 2 * (size_t) pr->bandwidth + 1;
  if ((size_t) pr->L * (size_t) n_cores * W > ((size_t) 4 << 30))
    return fail(RAMCU_ERR_UNSUPPORTED, "ramcu_extend_paths: L x cores x band = %zu cells of path state (limit 4 Gi)",
                (size_t) pr->L * (size_t) n_cores * W);
  ramcu_job *j = nullptr;
  int rc = job_create_impl(c, lib, direction, n_cores, cores, sc, pr, -pr->L, nullptr, &j);
  if (rc)
    return rc;
  if (j->n_ctas > 0)
    rc = ramcu_job_run(c, j);
  ramcu_result_t r0;
  if (!rc)
    rc = ramcu_job_fetch(c, j, master_out, ext_len_out, score_add_out, &r0);
  if (!rc && res)
    *res = r0;
  if (!rc && j->n_ctas > 0 && r0.rows_processed > 0)
  {
    const size_t rows = (size_t) r0.rows_processed, ne = (size_t) j->n_ext;
    std::vector<unsigned char> path(rows * ne * W), cons(rows);
    std::vector<int> best(rows * ne * 4), idx(rows * ne);
    CK(cudaMemcpy(path.data(), j->d_dbg_path, path.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(best.data(), j->d_dbg_best, best.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(idx.data(), j->d_dbg_idx, idx.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cons.data(), j->d_master, rows, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < rows; ++r)
      for (size_t e = 0; e < ne; ++e)
      {
        const size_t i = (size_t) j->orig[e];
        memcpy(path_state_out + (r * (size_t) n_cores + i) * W, path.data() + (r * ne + e) * W, W);
        row_best_out[r * (size_t) n_cores + i] = best[(r * ne + e) * 4 + (cons[r] & 3)];
        row_best_idx_out[r * (size_t) n_cores + i] = idx[r * ne + e];
      }
  }
  ramcu_job_free(c, j);
  return rc;
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
