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

int main(int argc, char **argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  const int delay = argc > 2 ? atoi(argv[2]) : 0;
  const int jitter = argc > 3 ? atoi(argv[3]) : 0;
  const bool quick = argc > 4;
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
