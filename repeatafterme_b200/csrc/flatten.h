// flatten.h -- host-side flattening of the reference's per-core records into the per-direction
// instance descriptors the kernels read. Plain C++ (no CUDA calls) so that the CPU test-suite
// can exercise exactly this logic (tests/emul).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/ramcu.h"
#include "extend_kernel.cuh"

namespace ramcu {

// struct coreAlignment -> InstDev for `direction` (1 = right, 0 = left); only cores that are
// extendable in that direction are kept (ram_extend.c:984-985). Sorted so that instances that
// run out of sequence at about the same row sit in the same warps.
static inline void flatten_cores(int direction, int n_cores, const ramcu_core_t *cores,
                                 uint64_t lib_len, std::vector<InstDev> &inst)
{
  inst.clear();
  inst.reserve(n_cores);
  for (int i = 0; i < n_cores; ++i)
  {
    const ramcu_core_t &k = cores[i];
    if (!(direction ? k.right_extendable : k.left_extendable))
      continue;
    InstDev d;
    // first base outside the core in the direction of travel (bnw_extend.c:778-788)
    if (direction)
      d.start = k.orient ? k.right_seq_pos - 1 : k.right_seq_pos + 1;
    else
      d.start = k.orient ? k.left_seq_pos + 1 : k.left_seq_pos - 1;
    // sequence index decreases when direction == orient (bnw_extend.c:824-835)
    d.sgn = ((direction ? 1 : 0) == (k.orient ? 1 : 0)) ? -1 : 1;
    d.comp = k.orient ? 1 : 0;
    // in-bounds <=> lower <= start + sgn*t <= upper (the signed form of bnw_extend.c:843-868),
    // additionally clipped to the library so that no load can leave it
    const long long lo = std::max<long long>(k.lower_seq_bound, 0);
    const long long hi = std::min<long long>(k.upper_seq_bound, (long long) lib_len - 1);
    long long tmin, tmax;
    if (d.sgn > 0)
    {
      tmin = lo - d.start;
      tmax = hi - d.start;
    }
    else
    {
      tmin = d.start - hi;
      tmax = d.start - lo;
    }
    const long long lim = 1LL << 29;
    d.tmin = (int) std::max(-lim, std::min(lim, tmin));
    d.tmax = (int) std::max(-lim, std::min(lim, tmax));
    d.orig = i;
    d.pad = 0;
    inst.push_back(d);
  }
  // stable, descending tmax: an LSD radix sort of the indices (three 11-bit digits of the inverted key), then
  // one permutation -- the order std::stable_sort gives, at a fraction of its cost for the 10^4-instance families
  // (this runs once per pass on the critical path of every call)
  const size_t n = inst.size();
  if (n < 2)
    return;
  std::vector<uint32_t> key(n), idx(n), tmp(n);
  for (size_t i = 0; i < n; ++i)
  {
    key[i] = 0x7fffffffu - (uint32_t) (inst[i].tmax + (1 << 30));   // tmax in [-2^29, 2^29]
    idx[i] = (uint32_t) i;
  }
  for (int shift = 0; shift < 33; shift += 11)
  {
    uint32_t count[2049] = {0};
    for (size_t i = 0; i < n; ++i)
      ++count[((key[idx[i]] >> shift) & 2047u) + 1];
    for (int b = 0; b < 2048; ++b)
      count[b + 1] += count[b];
    for (size_t i = 0; i < n; ++i)
      tmp[count[(key[idx[i]] >> shift) & 2047u]++] = idx[i];
    idx.swap(tmp);
  }
  std::vector<InstDev> sorted(n);
  for (size_t i = 0; i < n; ++i)
    sorted[i] = inst[idx[i]];
  inst.swap(sorted);
}

// grid layout helpers shared by the host code and the CPU emulation driver
static inline int inst_per_warp(int G) { return 32 / G; }
// CTA size when n_ext instances are split evenly over n_ctas CTAs (whole warps of 32/G instances)
static inline int cta_threads(int n_ext, int n_ctas, int G)
{
  const int per_cta = (n_ext + n_ctas - 1) / n_ctas;
  const int ipw = inst_per_warp(G);
  return 32 * ((per_cta + ipw - 1) / ipw);
}

static inline void fill_family(FamDev &f, const ramcu_scoring_t *sc, const ramcu_params_t *pr,
                               int n_ext, int n_ctas)
{
  memset(&f, 0, sizeof(f));
  f.n_ext = n_ext;
  f.n_ctas = n_ctas;
  f.cta0 = 0;
  f.B = pr->bandwidth;
  f.W = 2 * pr->bandwidth + 1;
  f.go = sc->gapopen;
  f.ge = sc->gapextn;
  f.cap = pr->cappenalty;
  f.minimp = pr->minimprovement;
  f.L = pr->L;
  f.stop = pr->when_to_stop;
  f.mate = -1;
  f.one = 1;
  for (int code = 0; code < RAMCU_NCODES; ++code)
    for (int a = 0; a < 4; ++a)
      f.m[code][a] = sc->matrix[a][code];
}

// ram_extend.c:1234-1247 on the kernel's per-instance outputs
static inline void map_results(int n_cores, const std::vector<int> &orig, const int *thigh,
                               const int *tpos, int32_t *ext_len_out, int32_t *score_add_out)
{
  for (int i = 0; i < n_cores; ++i)
  {
    if (ext_len_out) ext_len_out[i] = 0;
    if (score_add_out) score_add_out[i] = 0;
  }
  for (size_t e = 0; e < orig.size(); ++e)
    if (thigh[e] > 0 && tpos[e] >= 0)
    {
      if (ext_len_out) ext_len_out[orig[e]] = tpos[e] + 1;
      if (score_add_out) score_add_out[orig[e]] = thigh[e];
    }
}

}  // namespace ramcu
