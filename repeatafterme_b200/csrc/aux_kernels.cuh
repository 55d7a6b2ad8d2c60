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
