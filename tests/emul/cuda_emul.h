// cuda_emul.h -- TEST INFRASTRUCTURE ONLY.
//
// Lock-step CPU emulation of the handful of SIMT primitives that
// repeatafterme_b200/csrc/extend_kernel.cuh uses, so that the *actual kernel source* can be
// compiled with g++ (-DRAMCU_EMUL) and exercised by the CPU test-suite against the oracle before
// it is run on a B200. One OS thread per CUDA thread; warp collectives and __syncthreads are
// pthread barriers. It is never linked into libramcu.so and is not a fallback for anything.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cstdlib>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)

using std::max;
using std::min;

struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
struct dim3e { unsigned x, y, z; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v = {x, y}; return v; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }

struct EmulWarp
{
  pthread_barrier_t bar;
  long long buf[32];
};
struct EmulCta
{
  pthread_barrier_t bar;
  EmulWarp *warps;
  unsigned char shared[4096] __attribute__((aligned(16)));
};

extern thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;
extern thread_local EmulCta *emul_cta;
extern thread_local EmulWarp *emul_warp;

#define RAMCU_SHARED_STRUCT(type, name) \
  static_assert(sizeof(type) <= 4096, "emulated shared memory too small"); \
  type &name = *reinterpret_cast<type *>(emul_cta->shared)

static inline void __syncthreads() { pthread_barrier_wait(&emul_cta->bar); }

template <typename T>
static inline T emul_xchg(T v, int src_lane)
{
  const int lane = threadIdx.x & 31;
  emul_warp->buf[lane] = (long long) v;
  pthread_barrier_wait(&emul_warp->bar);
  const T r = (T) emul_warp->buf[(src_lane >= 0 && src_lane < 32) ? src_lane : lane];
  pthread_barrier_wait(&emul_warp->bar);
  return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_xchg(v, src & 31); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emul_xchg(v, (int) (threadIdx.x & 31) - d); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emul_xchg(v, (int) (threadIdx.x & 31) + d); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_xchg(v, (int) (threadIdx.x & 31) ^ m); }

static inline int __all_sync(unsigned, int pred)
{
  const int lane = threadIdx.x & 31;
  emul_warp->buf[lane] = pred ? 1 : 0;
  pthread_barrier_wait(&emul_warp->bar);
  int all = 1;
  for (int i = 0; i < 32; ++i) all &= (int) emul_warp->buf[i];
  pthread_barrier_wait(&emul_warp->bar);
  return all;
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
  const int lane = threadIdx.x & 31;
  emul_warp->buf[lane] = pred ? 1 : 0;
  pthread_barrier_wait(&emul_warp->bar);
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (unsigned) (emul_warp->buf[i] & 1) << i;
  pthread_barrier_wait(&emul_warp->bar);
  return m;
}
static inline int __reduce_add_sync(unsigned, int v)
{
  const int lane = threadIdx.x & 31;
  emul_warp->buf[lane] = v;
  pthread_barrier_wait(&emul_warp->bar);
  unsigned s = 0;
  for (int i = 0; i < 32; ++i) s += (unsigned) (int) emul_warp->buf[i];
  pthread_barrier_wait(&emul_warp->bar);
  return (int) s;
}
static inline int __reduce_max_sync(unsigned, int v)
{
  const int lane = threadIdx.x & 31;
  emul_warp->buf[lane] = v;
  pthread_barrier_wait(&emul_warp->bar);
  int m = (int) emul_warp->buf[0];
  for (int i = 1; i < 32; ++i) m = ((int) emul_warp->buf[i] > m) ? (int) emul_warp->buf[i] : m;
  pthread_barrier_wait(&emul_warp->bar);
  return m;
}
static inline int __reduce_min_sync(unsigned, int v)
{
  const int lane = threadIdx.x & 31;
  emul_warp->buf[lane] = v;
  pthread_barrier_wait(&emul_warp->bar);
  int m = (int) emul_warp->buf[0];
  for (int i = 1; i < 32; ++i) m = ((int) emul_warp->buf[i] < m) ? (int) emul_warp->buf[i] : m;
  pthread_barrier_wait(&emul_warp->bar);
  return m;
}
static inline unsigned __brev(unsigned x)
{
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(x);
}
// c + a.lo16 * b.byte0 + a.hi16 * b.byte1 (signed)
static inline int __dp2a_lo(int a, int b, int c)
{
  const int lo = (int) (short) (a & 0xffff), hi = (int) (short) ((unsigned) a >> 16);
  const int b0 = (int) (signed char) (b & 0xff), b1 = (int) (signed char) ((b >> 8) & 0xff);
  return (int) ((unsigned) c + (unsigned) (lo * b0) + (unsigned) (hi * b1));
}
static inline void __threadfence_block() { __sync_synchronize(); }
static inline void __syncwarp()
{
  pthread_barrier_wait(&emul_warp->bar);
}
static inline int atomicAdd(int *p, int v) { return __sync_fetch_and_add(p, v); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __sync_fetch_and_add(p, v); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __sync_fetch_and_or(p, v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __sync_fetch_and_add(p, v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
  const unsigned long long v = ((unsigned long long) hi << 32) | lo;
  return (unsigned) (v >> (s & 31));
}
static inline unsigned long long emul_ld_relaxed_u64(const unsigned long long *p)
{
  sched_yield();
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
static inline void emul_st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
  __atomic_store_n(p, v, __ATOMIC_RELAXED);
}
