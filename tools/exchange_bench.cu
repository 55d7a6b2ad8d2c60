// exchange_bench.cu -- development aid: latency of the per-column grid exchange mechanisms in
// isolation (148 CTAs, 4 sums per column, optional fake compute between columns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_bench exchange_bench.cu
//   ./exchange_bench [iters] [delay_cycles] [jitter_cycles]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ u64 ld_relaxed(const u64 *p)
{
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 ld_cg(const u64 *p)
{
  u64 v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_vol(const u64 *p)
{
  u64 v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// MODE 5: flat red.relaxed.gpu (explicit RED) + ld.global.cg polls
// MODE 6: slots, st.global.cg + ld.global.cg (weak, L2)
// MODE 7: slots, one slot word per CTA carries all four sums? no: 4 words, ld.volatile polls
// MODE 0: flat RED (count << 48 | sum) + poll, 4 words in one sector
// MODE 1: two-level RED: groups of GS CTAs, the last arriver of a group forwards the group total
// MODE 2: per-CTA slots (tag << 32 | sum), warp 0 gathers all
// MODE 3: flat RED, 4 words in 4 different 128-byte lines
// MODE 4: two-level, lane-parallel groups: like 1 but level-1 words of different groups 128 B apart
template <int MODE>
__device__ __forceinline__ void exchange(u64 *buf, int r, int n_ctas, int cta, int GS, unsigned (&s)[4], int lane)
{
  if (MODE == 0 || MODE == 3)
  {
    const int stride = (MODE == 3) ? 16 : 1;
    u64 *row = buf + (size_t) r * 4 * stride;
    unsigned lo = 0;
    if (lane < 4)
    {
      const unsigned mine = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      u64 *w = row + lane * stride;
      atomicAdd(w, (1ull << 48) | mine);
      u64 v;
      do { v = ld_relaxed(w); } while ((int) (v >> 48) != n_ctas);
      lo = (unsigned) v;
    }
    for (int a = 0; a < 4; ++a) s[a] = __shfl_sync(kFull, lo, a);
  }
  else if (MODE == 5)
  {
    u64 *row = buf + (size_t) r * 4;
    unsigned lo = 0;
    if (lane < 4)
    {
      const unsigned mine = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      u64 *w = row + lane;
      const u64 val = (1ull << 48) | mine;
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(w), "l"(val) : "memory");
      u64 v;
      do { v = ld_cg(w); } while ((int) (v >> 48) != n_ctas);
      lo = (unsigned) v;
    }
    for (int a = 0; a < 4; ++a) s[a] = __shfl_sync(kFull, lo, a);
  }
  else if (MODE == 6 || MODE == 7)
  {
    const int par = r & 1;
    const unsigned tag = (unsigned) r + 1;
    u64 *base = buf + (size_t) par * n_ctas * 4;
    if (lane < 4)
    {
      const unsigned mine = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      const u64 val = ((u64) tag << 32) | mine;
      if (MODE == 6)
        asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(base + (size_t) cta * 4 + lane), "l"(val) : "memory");
      else
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(base + (size_t) cta * 4 + lane), "l"(val) : "memory");
    }
    unsigned acc[4] = {0, 0, 0, 0};
    for (int b0 = 0; b0 < n_ctas; b0 += 32)
    {
      const int c = b0 + lane;
      if (c < n_ctas)
      {
        const u64 *p = base + (size_t) c * 4;
        u64 w[4];
        bool ok;
        do
        {
          for (int a = 0; a < 4; ++a) w[a] = (MODE == 6) ? ld_cg(p + a) : ld_vol(p + a);
          ok = true;
          for (int a = 0; a < 4; ++a) ok = ok && ((unsigned) (w[a] >> 32) == tag);
        } while (!ok);
        for (int a = 0; a < 4; ++a) acc[a] += (unsigned) w[a];
      }
    }
    for (int a = 0; a < 4; ++a) s[a] = (unsigned) __reduce_add_sync(kFull, (int) acc[a]);
  }
  else if (MODE == 1 || MODE == 4)
  {
    const int ng = (n_ctas + GS - 1) / GS;
    const int g = cta / GS;
    const int gsize = min(GS, n_ctas - g * GS);
    const int l1stride = (MODE == 4) ? 16 : 4;   // u64 words between groups
    // layout per column: [4 top words][pad to 16][ng * l1stride]
    u64 *row = buf + (size_t) r * (16 + 64 * 16);
    unsigned lo = 0;
    if (lane < 4)
    {
      const unsigned mine = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      u64 *w1 = row + 16 + (size_t) g * l1stride + lane;
      const u64 old = atomicAdd(w1, (1ull << 48) | mine);
      if ((int) (old >> 48) == gsize - 1)
      {
        const unsigned tot = (unsigned) old + mine;
        atomicAdd(row + lane, (1ull << 48) | tot);
      }
      u64 v;
      do { v = ld_relaxed(row + lane); } while ((int) (v >> 48) != ng);
      lo = (unsigned) v;
    }
    for (int a = 0; a < 4; ++a) s[a] = __shfl_sync(kFull, lo, a);
  }
  else if (MODE == 2)
  {
    const int par = r & 1;
    const unsigned tag = (unsigned) r + 1;
    u64 *base = buf + (size_t) par * n_ctas * 4;
    if (lane < 4)
    {
      const unsigned mine = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
      st_relaxed(base + (size_t) cta * 4 + lane, ((u64) tag << 32) | mine);
    }
    unsigned acc[4] = {0, 0, 0, 0};
    constexpr int U = 5;
    u64 w[U][4];
    bool done[U];
    for (int u = 0; u < U; ++u) done[u] = (u * 32 + lane) >= n_ctas;
    bool all;
    do
    {
      for (int u = 0; u < U; ++u)
        if (!done[u])
        {
          const u64 *p = base + (size_t) (u * 32 + lane) * 4;
          asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[u][0]), "=l"(w[u][1]) : "l"(p) : "memory");
          asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[u][2]), "=l"(w[u][3]) : "l"(p + 2) : "memory");
        }
      all = true;
      for (int u = 0; u < U; ++u)
        if (!done[u])
        {
          const bool ok = ((unsigned) (w[u][0] >> 32) == tag) && ((unsigned) (w[u][1] >> 32) == tag) &&
                          ((unsigned) (w[u][2] >> 32) == tag) && ((unsigned) (w[u][3] >> 32) == tag);
          if (ok)
          {
            done[u] = true;
            for (int a = 0; a < 4; ++a) acc[a] += (unsigned) w[u][a];
          }
          else
            all = false;
        }
    } while (!all);
    for (int a = 0; a < 4; ++a) s[a] = (unsigned) __reduce_add_sync(kFull, (int) acc[a]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench_kernel(u64 *buf, int iters, int delay, int jitter, int GS,
                                                        long long *out, int *bad)
{
  __shared__ int4 bc[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cta = blockIdx.x, n = gridDim.x;
  long long tw = 0;
  unsigned rng = cta * 2654435761u + 12345u;
  const long long t0 = clock64();
  for (int r = 0; r < iters; ++r)
  {
    rng = rng * 1664525u + 1013904223u;
    const int d = delay + (jitter ? (int) ((rng >> 8) % (unsigned) jitter) : 0);
    const long long ts = clock64();
    while (clock64() - ts < d) { }
    __syncthreads();
    if (warp == 0)
    {
      unsigned s[4] = {(unsigned) (r + cta), (unsigned) (2 * r + cta), 7u, (unsigned) cta};
      const long long te = clock64();
      exchange<MODE>(buf, r, n, cta, GS, s, lane);
      tw += clock64() - te;
      if (lane == 0)
      {
        bc[r & 1] = make_int4((int) s[0], (int) s[1], (int) s[2], (int) s[3]);
        const unsigned tri = (unsigned) (n * (n - 1) / 2);
        if (s[0] != (unsigned) (r * n) + tri || s[1] != (unsigned) (2 * r * n) + tri || s[2] != 7u * n || s[3] != tri)
          atomicAdd(bad, 1);
      }
    }
    __syncthreads();
    volatile int4 *p = bc;
    if (p[r & 1].z == -12345) out[0] = 1;
  }
  if (threadIdx.x == 0)
  {
    out[2 * cta] = clock64() - t0;
    out[2 * cta + 1] = tw;
  }
}

template <int MODE>
static void run(const char *name, int n_ctas, int iters, int delay, int jitter, int GS)
{
  const size_t words = (size_t) iters * (16 + 64 * 16) + 4096;
  u64 *buf;
  long long *out;
  int *bad;
  CK(cudaMalloc(&buf, words * 8));
  CK(cudaMalloc(&out, 2 * n_ctas * sizeof(long long)));
  CK(cudaMalloc(&bad, 4));
  for (int rep = 0; rep < 2; ++rep)
  {
    CK(cudaMemset(buf, 0, words * 8));
    CK(cudaMemset(bad, 0, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench_kernel<MODE><<<n_ctas, 224>>>(buf, iters, delay, jitter, GS, out, bad);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(2 * n_ctas);
    int hb;
    CK(cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    double tot = 0, wait = 0;
    for (int c = 0; c < n_ctas; ++c) { tot += h[2 * c]; wait += h[2 * c + 1]; }
    if (rep == 1)
      printf("%-28s ctas=%d GS=%d delay=%d jitter=%d: %.3f us/col, cycles/col %.0f, exchange %.0f, bad=%d\n", name,
             n_ctas, GS, delay, jitter, 1e3 * ms / iters, tot / n_ctas / iters, wait / n_ctas / iters, hb);
  }
  cudaFree(buf);
  cudaFree(out);
  cudaFree(bad);
}

// MODE 8: thread-block clusters. The CTAs of a cluster push their four sums into the leader's shared memory
// through DSMEM (tagged 64-bit stores, the leader spins on its own shared memory: no cluster barrier), the
// leaders exchange through L2 exactly like MODE 0 (flat RED + poll, n_clusters participants), and each leader
// pushes the column sums back into every member's shared memory (tagged), where the member's warp 0 spins.
__global__ void __launch_bounds__(256, 1) cluster_kernel(u64 *buf, int iters, int delay, int jitter, long long *out, int *bad)
{
  __shared__ u64 inbox[2][16][4];   // leader: [parity][member rank][sum] = tag << 32 | value
  __shared__ u64 result[2][4];      // every CTA: [parity][sum] = tag << 32 | value, written by the leader
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cta = blockIdx.x, n = gridDim.x;
  const int cs = (int) cluster.num_blocks(), rank = (int) cluster.block_rank();
  const int n_clusters = n / cs;
  for (int i = threadIdx.x; i < 2 * 16 * 4; i += blockDim.x)
    (&inbox[0][0][0])[i] = 0;
  for (int i = threadIdx.x; i < 8; i += blockDim.x)
    (&result[0][0])[i] = 0;
  cluster.sync();
  u64 *leader_inbox = cluster.map_shared_rank(&inbox[0][0][0], 0);
  long long tw = 0;
  unsigned rng = cta * 2654435761u + 12345u;
  const long long t0 = clock64();
  for (int r = 0; r < iters; ++r)
  {
    rng = rng * 1664525u + 1013904223u;
    const int d = delay + (jitter ? (int) ((rng >> 8) % (unsigned) jitter) : 0);
    const long long ts = clock64();
    while (clock64() - ts < d) { }
    __syncthreads();
    if (warp == 0)
    {
      unsigned s[4] = {(unsigned) (r + cta), (unsigned) (2 * r + cta), 7u, (unsigned) cta};
      const long long te = clock64();
      const int par = r & 1;
      const u64 tag = (u64) (unsigned) (r + 1) << 32;
      if (lane < 4)
      {
        const unsigned mine = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
        // into the leader's inbox (DSMEM store; the leader's own goes the same way)
        volatile u64 *slot = leader_inbox + ((size_t) par * 16 + rank) * 4 + lane;
        *slot = tag | mine;
      }
      if (rank == 0)
      {
        // gather the cluster's sums from the local inbox: lane = member * 4 + candidate (cs <= 8)
        unsigned part = 0;
        if (lane < cs * 4)
        {
          volatile u64 *slot = &inbox[par][lane >> 2][lane & 3];
          u64 v;
          unsigned spins = 0;
          do { v = *slot; } while ((v >> 32) != (tag >> 32) && ++spins < (1u << 24));
          if (spins >= (1u << 24)) atomicAdd(bad, 1000);
          part = (unsigned) v;
        }
        for (int o = 4; o < 32; o <<= 1)
          part += __shfl_xor_sync(kFull, part, o);   // lanes with equal (lane & 3) hold the cluster sum
        unsigned lo = 0;
        if (lane < 4)
        {
          u64 *w = buf + (size_t) r * 4 + lane;
          atomicAdd(w, (1ull << 48) | part);
          u64 v;
          unsigned spins = 0;
          do { v = ld_relaxed(w); } while ((int) (v >> 48) != n_clusters && ++spins < (1u << 22));
          if (spins >= (1u << 22)) atomicAdd(bad, 1000);
          lo = (unsigned) v;
        }
        // push the column sums to every member (lane = member * 4 + candidate)
        const unsigned val = __shfl_sync(kFull, lo, lane & 3);
        if (lane < cs * 4)
        {
          volatile u64 *dst = cluster.map_shared_rank(&result[0][0], lane >> 2) + par * 4 + (lane & 3);
          *dst = tag | val;
        }
      }
      unsigned got = 0;
      if (lane < 4)
      {
        volatile u64 *slot = &result[par][lane];
        u64 v;
        unsigned spins = 0;
        do { v = *slot; } while ((v >> 32) != (tag >> 32) && ++spins < (1u << 24));
        if (spins >= (1u << 24)) atomicAdd(bad, 1000);
        got = (unsigned) v;
      }
      for (int a = 0; a < 4; ++a) s[a] = __shfl_sync(kFull, got, a);
      tw += clock64() - te;
      if (lane == 0)
      {
        const unsigned tri = (unsigned) (n * (n - 1) / 2);
        if (s[0] != (unsigned) (r * n) + tri || s[1] != (unsigned) (2 * r * n) + tri || s[2] != 7u * n || s[3] != tri)
          atomicAdd(bad, 1);
      }
    }
    __syncthreads();
  }
  cluster.sync();
  if (threadIdx.x == 0)
  {
    out[2 * cta] = clock64() - t0;
    out[2 * cta + 1] = tw;
  }
}

static void run_cluster(int n_ctas, int cs, int iters, int delay, int jitter)
{
  n_ctas -= n_ctas % cs;
  const size_t words = (size_t) iters * 4 + 4096;
  u64 *buf;
  long long *out;
  int *bad;
  CK(cudaMalloc(&buf, words * 8));
  CK(cudaMalloc(&out, 2 * n_ctas * sizeof(long long)));
  CK(cudaMalloc(&bad, 4));
  for (int rep = 0; rep < 2; ++rep)
  {
    CK(cudaMemset(buf, 0, words * 8));
    CK(cudaMemset(bad, 0, 4));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas);
    cfg.blockDim = dim3(224);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, cluster_kernel, &cfg) != cudaSuccess || max_clusters * cs < n_ctas)
    {
      printf("cluster DSMEM: %d CTAs in clusters of %d cannot all be resident (max %d clusters): skipped\n", n_ctas, cs,
             max_clusters);
      cudaGetLastError();
      break;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_kernel, buf, iters, delay, jitter, out, bad);
    if (e != cudaSuccess)
    {
      printf("cluster launch (%d CTAs, clusters of %d) failed: %s\n", n_ctas, cs, cudaGetErrorString(e));
      cudaGetLastError();
      break;
    }
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(2 * n_ctas);
    int hb;
    CK(cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    double tot = 0, wait = 0;
    for (int c = 0; c < n_ctas; ++c) { tot += h[2 * c]; wait += h[2 * c + 1]; }
    if (rep == 1)
      printf("%-28s ctas=%d GS=%d delay=%d jitter=%d: %.3f us/col, cycles/col %.0f, exchange %.0f, bad=%d\n",
             "cluster DSMEM + leaders RED", n_ctas, cs, delay, jitter, 1e3 * ms / iters, tot / n_ctas / iters,
             wait / n_ctas / iters, hb);
  }
  cudaFree(buf);
  cudaFree(out);
  cudaFree(bad);
}

int main(int argc, char **argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  const int delay = argc > 2 ? atoi(argv[2]) : 0;
  const int jitter = argc > 3 ? atoi(argv[3]) : 0;
  const bool quick = argc > 4;
  for (int cs : {2, 4, 8})
  {
    run_cluster(148, cs, iters, delay, jitter);
    run_cluster(296, cs, iters, delay, jitter);
  }
  run<0>("flat RED, one sector", 296, iters, delay, jitter, 0);
  for (int n : {148, 37, 8, 2})
  {
    run<5>("flat red.relaxed + ld.cg", n, iters, delay, jitter, 0);
    run<6>("slots st.cg + ld.cg", n, iters, delay, jitter, 0);
    run<7>("slots volatile", n, iters, delay, jitter, 0);
    if (quick)
    {
      run<0>("flat RED, one sector", n, iters, delay, jitter, 0);
      run<2>("slots, warp gathers", n, iters, delay, jitter, 0);
      continue;
    }
    run<0>("flat RED, one sector", n, iters, delay, jitter, 0);
    run<3>("flat RED, 4 lines", n, iters, delay, jitter, 0);
    for (int gs : {4, 8, 12, 16, 24})
      if (gs < n)
      {
        run<1>("two-level RED", n, iters, delay, jitter, gs);
        run<4>("two-level RED, groups 128B", n, iters, delay, jitter, gs);
      }
    run<2>("slots, warp gathers", n, iters, delay, jitter, 0);
  }
  return 0;
}
