// aux_kernels.cuh -- the non-template kernels of libramcu.so: library packing (byte codes ->
// 2 bits/base + N bitmask, 128-bit loads) and the INT32 ALU peak microbenchmark.
#pragma once
#include "extend_kernel.cuh"

namespace ramcu {

// Library packing: 1 byte/base (0..3 = ACGT, anything else = N) -> 2 bits/base + N bitmask.
// Each thread converts 32 bases: two 128-bit loads, one 64-bit + one 32-bit store.
__global__ void __launch_bounds__(256)
pack_library_kernel(const uint8_t *__restrict__ bytes, const unsigned long long len,
                    uint32_t *__restrict__ lib2, uint32_t *__restrict__ libn,
                    const unsigned long long n_chunks)
{
  const bool aligned = ((reinterpret_cast<uintptr_t>(bytes) & 15) == 0);
  for (unsigned long long ch = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
       ch < n_chunks; ch += (unsigned long long) gridDim.x * blockDim.x)
  {
    const unsigned long long base = ch * 32;
    uint32_t raw[8];
    if (aligned && base + 32 <= len)
    {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(bytes + base));
      const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(bytes + base + 16));
      raw[0] = v0.x; raw[1] = v0.y; raw[2] = v0.z; raw[3] = v0.w;
      raw[4] = v1.x; raw[5] = v1.y; raw[6] = v1.z; raw[7] = v1.w;
    }
    else
    {
#pragma unroll
      for (int i = 0; i < 8; ++i)
      {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          const unsigned long long p = base + i * 4 + j;
          const uint32_t b = (p < len) ? bytes[p] : 99u;
          w |= b << (8 * j);
        }
        raw[i] = w;
      }
    }
    uint32_t two[2] = {0, 0};
    uint32_t nm = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
#pragma unroll
      for (int j = 0; j < 4; ++j)
      {
        const uint32_t b = (raw[i] >> (8 * j)) & 0xffu;
        const int q = i * 4 + j;
        const uint32_t isn = (b > 3u) ? 1u : 0u;
        nm |= isn << q;
        two[q >> 4] |= (b & 3u) << (2 * (q & 15));
      }
    }
    *reinterpret_cast<uint2 *>(lib2 + ch * 2) = make_uint2(two[0], two[1]);
    libn[ch] = nm;
  }
}


// 2bit-native flank feed: build the packed library straight from the .2bit file's own packed DNA
// (kentsrc/twoBitNew.c:571-582: 4 bases per byte, first base in the top two bits, T0 C1 A2 G3),
// without the 1-byte-per-base expansion of loadSequenceSubsetMinimal (sequence.c:752-822).
// N-blocks are applied by nmask_ranges_kernel afterwards.
struct WindowDev
{
  unsigned long long dst;      // first library position of the window
  unsigned long long len;      // bases
  unsigned long long src_byte; // byte offset of the record's packed DNA inside the blob
  unsigned long long src_base; // index of the window's first base inside the record
};

// one 32-base chunk, base by base: chunks that straddle windows, the library's tail and sources
// too close to the end of the blob for whole 16-byte loads
__device__ __forceinline__ void twobit_chunk_slow(const uint8_t *__restrict__ blob,
                                                  const unsigned long long blob_bytes,
                                                  const WindowDev *__restrict__ win, const int n_win,
                                                  const unsigned long long total,
                                                  uint32_t *__restrict__ lib2, const unsigned long long ch)
{
  {
    const unsigned long long base = ch * 32;
    uint32_t two[2] = {0, 0};
    // window containing `base` (windows tile [0,total) in dst order)
    int lo = 0, hi = n_win - 1;
    while (lo < hi)
    {
      const int mid = (lo + hi + 1) >> 1;
      if (win[mid].dst <= base)
        lo = mid;
      else
        hi = mid - 1;
    }
    int w = lo;
    unsigned long long p = base;
    const unsigned long long end = min(base + 32, total);
    while (p < end)
    {
      while (w + 1 < n_win && win[w + 1].dst <= p)
        ++w;
      const WindowDev W = win[w];
      const unsigned long long wend = min(end, W.dst + W.len);
      if (wend <= p)
      {
        ++p;     // hole between windows (not produced by the loader): leave code 0
        continue;
      }
      // source bases [sb, sb + n) of the record
      const unsigned long long sb = W.src_base + (p - W.dst);
      const int n = (int) (wend - p);
      const unsigned long long b0 = W.src_byte + (sb >> 2);          // first source byte
      const unsigned long long a0 = b0 & ~3ull;                      // aligned down to 4 bytes
      uint32_t raw[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        const unsigned long long a = a0 + 4ull * i;
        raw[i] = (a + 4 <= blob_bytes) ? __ldg(reinterpret_cast<const uint32_t *>(blob + a)) : 0u;
        if (a + 4 > blob_bytes && a < blob_bytes)
        {
          uint32_t v = 0;
          for (unsigned long long q = a; q < blob_bytes; ++q)
            v |= (uint32_t) blob[q] << (8 * (q - a));
          raw[i] = v;
        }
      }
      const int first = (int) ((b0 - a0) * 4 + (sb & 3));            // base index inside raw[]
      for (int i = 0; i < n; ++i)
      {
        const int bi = first + i;
        const uint32_t byte = (raw[bi >> 4] >> (8 * ((bi >> 2) & 3))) & 0xffu;
        const uint32_t t = (byte >> (6 - 2 * (bi & 3))) & 3u;         // T0 C1 A2 G3
        const uint32_t code = (0x87u >> (2 * t)) & 3u;                // -> A0 C1 G2 T3 (3,1,0,2)
        const int q = (int) (p - base) + i;
        two[q >> 4] |= code << (2 * (q & 15));
      }
      p = wend;
    }
    *reinterpret_cast<uint2 *>(lib2 + ch * 2) = make_uint2(two[0], two[1]);
  }
}

// T0 C1 A2 G3 with the first base of a byte in its top bits  ->  A0 C1 G2 T3 with base i of the word
// at bits 2i: reverse the four 2-bit fields of every byte, then recode (hi' = ~(hi ^ lo), lo' = ~hi)
__device__ __forceinline__ uint32_t twobit_word_to_lib2(uint32_t x)
{
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  const uint32_t h = x & 0xaaaaaaaau, l = x & 0x55555555u;
  return (~(h ^ (l << 1)) & 0xaaaaaaaau) | ((~h >> 1) & 0x55555555u);
}

// One thread produces 128 library bases (32 bytes of lib2, two 128-bit stores) from the 33 source
// bytes that cover them, read as three aligned 128-bit loads; consecutive threads read and write
// consecutive addresses. The twelve source words are recoded 16 bases at a time and funnel-shifted
// to the library's alignment. Groups that are not inside one window (at most one per window
// boundary, plus the tail) are only recorded here and decoded base by base by
// twobit_boundary_kernel, so that no warp of this kernel waits for a slow lane.
__global__ void __launch_bounds__(256)
twobit_windows_kernel(const uint8_t *__restrict__ blob, const unsigned long long blob_bytes,
                      const WindowDev *__restrict__ win, const int n_win,
                      const unsigned long long total, uint32_t *__restrict__ lib2,
                      const unsigned long long n_chunks, unsigned long long *__restrict__ slow_list,
                      unsigned int *__restrict__ slow_count, const unsigned int slow_cap)
{
  const unsigned long long n_groups = (n_chunks + 3) / 4;
  for (unsigned long long g = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
       g < n_groups; g += (unsigned long long) gridDim.x * blockDim.x)
  {
    const unsigned long long base = g * 128;
    int lo = 0, hi = n_win - 1;
    while (lo < hi)
    {
      const int mid = (lo + hi + 1) >> 1;
      if (win[mid].dst <= base)
        lo = mid;
      else
        hi = mid - 1;
    }
    bool fast = false;
    unsigned long long a0 = 0;
    int bitoff = 0;
    if (n_win > 0)
    {
      const WindowDev W = win[lo];
      const unsigned long long sb = W.src_base + (base - W.dst);
      const unsigned long long byte0 = W.src_byte + (sb >> 2);
      a0 = byte0 & ~15ull;
      bitoff = (int) (byte0 - a0) * 8 + 2 * (int) (sb & 3);
      fast = (W.dst <= base) && (base + 128 <= W.dst + W.len) && (base + 128 <= total) &&
             (a0 + 48 <= blob_bytes) && ((reinterpret_cast<unsigned long long>(blob) & 15ull) == 0);
    }
    if (fast)
    {
      uint32_t cw[12];
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(blob + a0) + i);
        cw[4 * i + 0] = twobit_word_to_lib2(v.x);
        cw[4 * i + 1] = twobit_word_to_lib2(v.y);
        cw[4 * i + 2] = twobit_word_to_lib2(v.z);
        cw[4 * i + 3] = twobit_word_to_lib2(v.w);
      }
      const int k0 = bitoff >> 5, sh = bitoff & 31;   // k0 in 0..3
      uint32_t t[9];
#pragma unroll
      for (int j = 0; j < 9; ++j)
        t[j] = (k0 == 0) ? cw[j] : (k0 == 1) ? cw[j + 1] : (k0 == 2) ? cw[j + 2] : cw[j + 3];
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = __funnelshift_r(t[j], t[j + 1], sh);
      uint4 *dst = reinterpret_cast<uint4 *>(lib2 + g * 8);
      dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
      dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    else
    {
      // the list holds one group per window boundary plus the tail; callers may also hand in windows
      // whose sources sit in the last bytes of the blob (no whole 16-byte loads): past the list's
      // capacity the group is decoded right here
      const unsigned int idx = atomicAdd(slow_count, 1u);
      if (idx < slow_cap)
        slow_list[idx] = g;
      else
        for (unsigned long long ch = g * 4; ch < g * 4 + 4 && ch < n_chunks; ++ch)
          twobit_chunk_slow(blob, blob_bytes, win, n_win, total, lib2, ch);
    }
  }
}

// the groups twobit_windows_kernel left out: one thread per 32-base chunk
__global__ void __launch_bounds__(128)
twobit_boundary_kernel(const uint8_t *__restrict__ blob, const unsigned long long blob_bytes,
                       const WindowDev *__restrict__ win, const int n_win,
                       const unsigned long long total, uint32_t *__restrict__ lib2,
                       const unsigned long long n_chunks, const unsigned long long *__restrict__ slow_list,
                       const unsigned int *__restrict__ slow_count, const unsigned int slow_cap)
{
  const unsigned long long n = 4ull * min(*slow_count, slow_cap);
  for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long) gridDim.x * blockDim.x)
  {
    const unsigned long long ch = slow_list[i >> 2] * 4 + (i & 3);
    if (ch < n_chunks)
      twobit_chunk_slow(blob, blob_bytes, win, n_win, total, lib2, ch);
  }
}

// set the N bit of every base in [start, start+len) for a list of ranges (library coordinates)
__global__ void __launch_bounds__(128)
nmask_ranges_kernel(const unsigned long long *__restrict__ start,
                    const unsigned long long *__restrict__ len, const int n_ranges,
                    uint32_t *__restrict__ libn)
{
  const int r = blockIdx.x;
  if (r >= n_ranges)
    return;
  const unsigned long long s = start[r], e = s + len[r];
  if (e <= s)
    return;
  const unsigned long long w0 = s >> 5, w1 = (e - 1) >> 5;
  for (unsigned long long w = w0 + threadIdx.x; w <= w1; w += blockDim.x)
  {
    uint32_t m = 0xffffffffu;
    if (w == w0)
      m &= 0xffffffffu << (s & 31);
    if (w == w1)
      m &= 0xffffffffu >> (31 - ((e - 1) & 31));
    atomicOr(libn + w, m);
  }
}

// one range handed over by value (the positions past the library inside its last mask word)
__global__ void nmask_one_range_kernel(const unsigned long long s, const unsigned long long len, uint32_t *__restrict__ libn)
{
  if (threadIdx.x == 0 && blockIdx.x == 0 && len > 0)
  {
    const unsigned long long e = s + len;
    for (unsigned long long w = s >> 5; w <= ((e - 1) >> 5); ++w)
    {
      uint32_t m = 0xffffffffu;
      if (w == (s >> 5))
        m &= 0xffffffffu << (s & 31);
      if (w == ((e - 1) >> 5))
        m &= 0xffffffffu >> (31 - ((e - 1) & 31));
      atomicOr(libn + w, m);
    }
  }
}

// INT32 ALU roofline denominator: kChains independent max-plus chains per thread, one fused
// add+max (VIADDMNMX) per chain per step, on every SM. Counted as ONE op per instruction-lane.
constexpr int kPeakChains = 16;
__global__ void __launch_bounds__(256)
int32_peak_kernel(int *__restrict__ out, const int iters, const int seed)
{
  int x[kPeakChains];
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k)
    x[k] = seed + k * 7 + (int) threadIdx.x;
  const int inc = seed | 1, floor_ = -seed;
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int k = 0; k < kPeakChains; ++k)
      x[k] = max(x[k] + inc, floor_ + k);
  }
  int acc = 0;
#pragma unroll
  for (int k = 0; k < kPeakChains; ++k)
    acc ^= x[k];
  if (acc == 0x7fffffff)
    out[0] = acc;
}

}  // namespace ramcu
