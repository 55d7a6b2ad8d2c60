// extend_kernel.cuh -- the banded-DP MSA extension kernel (sm_100a).
//
// Replaces, on the device, the column loop of extend_alignment() (reference
// ram_extend.c:970-1224) and the band row of compute_nw_row() (bnw_extend.c:800-1045).
//
// Mapping:
//   * an MSA instance's band of W = 2B+1 cells is split into G contiguous segments of C cells;
//     one lane owns one segment (G lanes per instance, any G in 1..32, G*C >= W).
//   * row state in registers, per cell, in the form the NEXT row consumes it:
//         P[k] = max(sub, gap)                  of the committed row      (bnw_extend.c:1012-1018)
//         D[k] = max(sub + go + ge, gap + ge)   of the cell to the RIGHT  (the deletion term of
//                                               bnw_extend.c:892-905 for cell k of the next row;
//                                               the same expression is the insertion term of
//                                               bnw_extend.c:972-985 for the cell right of it)
//   * sweep (all four candidate consensus bases, no dependency between cells): with non-positive
//     gap penalties the insertion term never exceeds the cell left of it, so the row best is
//         best_a = max_k max(P[k] + m[a][base_k], D[k])
//     exactly -- one fused add+max per cell and candidate, no insertion chain, no cross-lane scan.
//   * per-candidate bests are reduced over the G lanes, capped (max(best,0,high+CAPPENALTY)),
//     summed over the warp (__reduce_add_sync), the CTA (shared memory) and, for families
//     spanning several CTAs, over the grid with a data+flag exchange through L2 - no host
//     round trip and no cooperative launch per column.
//   * commit: only the chosen base's row is materialised (the reference's "5th pass",
//     ram_extend.c:1105-1150). The insertion chain H[k] = max(y[k], H[k-1] + ge) is serial inside
//     a lane; across the G lanes the lane aggregates are combined by a max-plus prefix scan
//     (__shfl_up_sync, decay C*ge per lane) and the true carry is folded in afterwards
//     (P[k] = max(P[k], carry + k*ge), H[k] = max(H[k], carry + (k+1)*ge)), together with the
//     first-max column index.
//   * integer max-plus only: no tensor cores. Sentinels (-1000000000 / -987654321) and the
//     edge rule are the reference's, so every cell value is bit-identical. Rows whose band hangs
//     over a sequence bound (or runtime-width bands, or positive gap penalties) take an exact
//     per-cell predicated path with an (aggregate, transfer) scan.
#pragma once
#ifdef RAMCU_EMUL
#include "cuda_emul.h"   // tests/emul: CPU lock-step emulation of the SIMT primitives (tests only)
#else
#include <cuda_runtime.h>
#define RAMCU_SHARED_STRUCT(type, name) __shared__ type name
#endif
#include <stdint.h>

namespace ramcu {

constexpr int kNeg = -1000000000;     // bnw_extend.c:760,802-804
constexpr int kOob = -987654321;      // bnw_extend.c:767
constexpr int kNegInf = -1050000000;  // "no carry": below every reachable cell value, and
                                      // 2*kNegInf still fits in int32
constexpr int kBig = 1 << 28;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxWarps = 32;         // CTA size <= 1024 threads

// One extendable instance of a family, in the direction being extended (device copy).
struct InstDev
{
  long long start;  // sequence[] index of band position t = 0 (bnw_extend.c:778-788)
  int sgn;          // +1 / -1 walk direction in sequence[] (bnw_extend.c:824-835)
  int comp;         // complement the base (orient == 1)
  int tmin, tmax;   // band positions t with lower_seq_bound <= pos(t) <= upper_seq_bound
  int orig;         // index of the core in the caller's array
  int pad;
};

// One family == one call of extend_alignment() for one direction (device copy).
struct FamDev
{
  const uint32_t *lib2;   // library, 2 bits/base (A0 C1 G2 T3), 16 bases per word
  const uint32_t *libn;   // library, N bitmask, 32 bases per word
  const InstDev *inst;    // n_ext instances
  unsigned char *master;  // out: chosen base per processed row
  int *thigh;             // out: trimmed_sequence_high_score      (per extendable instance)
  int *tpos;              // out: trimmed_sequence_high_score_pos
  int *result;            // out: [0] max_row [1] rows_processed [2] stop_row [3] max_ext
  unsigned long long *slots;    // column exchange area, zeroed before each launch:
                                // [2][n_ctas][4] (tag << 32 | sum) slots, or, for more than
                                // kSlotMaxCtas CTAs, [L][4] (count << 48 | sum) accumulators
  const unsigned char *forced;  // debug: forced consensus per row (or NULL)
  int *dbg_rows;          // debug: [rows][n_ext][W][2]
  int *dbg_best;          // debug: [rows][n_ext][4]
  int *dbg_sums;          // debug: [rows][4]
  unsigned char *dbg_path;   // -outmat: [rows][n_ext][W] state of the best path into each cell of the
                             // committed row: 0 substitution, 1 deletion, 2 insertion (bnw_extend.c:1027-1044)
  int *dbg_del;           // -outmat scratch: [n_ext][W] deletion term of the row being committed
  int *dbg_idx;           // -outmat: [rows][n_ext] band index of the row's first best cell
  long long *dbg_wait;    // diagnostics: per CTA [0] cycles spent in the grid exchange [1] total cycles
  int n_ext, n_ctas, cta0, B;
  int W, go, ge, cap;
  int minimp, L, stop, pad;
  int m[8][4];            // m[code][cand]; code 4 = N, 5 = out of bounds (value unused)
};

__device__ __forceinline__ int imax(int a, int b) { return max(a, b); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }
__device__ __forceinline__ int sel4(int a, int x0, int x1, int x2, int x3)
{
  const int lo = (a & 1) ? x1 : x0;
  const int hi = (a & 1) ? x3 : x2;
  return (a & 2) ? hi : lo;
}

// max / min over the G lanes of an instance, result in every lane of the group. Powers of two use
// an xor butterfly; other group sizes an upward inclusive scan + broadcast from the last lane.
template <int G>
__device__ __forceinline__ int group_max(int v, const int gl, const int lane)
{
  if (G == 1)
    return v;
  if ((G & (G - 1)) == 0)
  {
#pragma unroll
    for (int d = G >> 1; d >= 1; d >>= 1)
      v = max(v, __shfl_xor_sync(kFull, v, d));
    return v;
  }
  if (G <= 7)
  {
    // rotate through the group: G-1 shuffles from fixed source lanes (spare lanes above the last
    // whole group read wrapped lanes; their result is never used)
    const int base = lane - gl;
    int m = v;
#pragma unroll
    for (int d = 1; d < G; ++d)
      m = max(m, __shfl_sync(kFull, v, (base + ((gl + d >= G) ? gl + d - G : gl + d)) & 31));
    return m;
  }
#pragma unroll
  for (int d = 1; d < G; d <<= 1)
  {
    const int t = __shfl_up_sync(kFull, v, d);
    if (gl >= d)
      v = max(v, t);
  }
  return __shfl_sync(kFull, v, min(lane - gl + G - 1, 31));
}
template <int G>
__device__ __forceinline__ int group_min(int v, const int gl, const int lane)
{
  return -group_max<G>(-v, gl, lane);
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
#ifdef RAMCU_EMUL
  return emul_ld_relaxed_u64(p);
#else
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
#ifdef RAMCU_EMUL
  emul_st_relaxed_u64(p, v);
#else
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// Base code at band position t of an instance: 0..3 (complemented for reverse cores), 4 = N,
// 5 = outside [lower_seq_bound, upper_seq_bound].
__device__ __forceinline__ int fetch_code(const uint32_t *__restrict__ lib2,
                                          const uint32_t *__restrict__ libn, long long start,
                                          int sgn, int comp, int tmin, int tmax, int t)
{
  if (t < tmin || t > tmax)
    return 5;
  const long long p = start + (long long) sgn * t;
  const uint32_t w = __ldg(lib2 + (p >> 4));
  const uint32_t nm = __ldg(libn + (p >> 5));
  int c = (w >> (2 * ((int) p & 15))) & 3;
  if (comp)
    c = 3 - c;
  return ((nm >> ((int) p & 31)) & 1) ? 4 : c;
}

// byte j of a window register, zero-extended: one PRMT instead of shift + mask
__device__ __forceinline__ int win_byte(const uint32_t w, const int j)
{
#ifdef RAMCU_EMUL
  return (int) ((w >> (8 * j)) & 0xffu);
#else
  return (int) __byte_perm(w, 0u, 0x4440u | (unsigned) j);
#endif
}

// ---------------------------------------------------------------------------------------------
// flank feed: the base that enters the band at its high end, 16 band positions per load
// ---------------------------------------------------------------------------------------------
// eight 2-bit fields (low 16 bits) -> bits 0..1 of eight nibbles
__device__ __forceinline__ uint32_t spread2to4(uint32_t x)
{
  x &= 0xffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  return x;
}
// eight bits (low 8 bits) -> bit 0 of eight nibbles
__device__ __forceinline__ uint32_t spread1to4(uint32_t x)
{
  x &= 0xffu;
  x = (x | (x << 12)) & 0x000f000fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// Codes (0..3 base, complemented for reverse cores; 4 = N; 5 = outside the sequence bounds) of
// band positions t0 .. t0+15 of an instance as sixteen nibbles (nlo: t0..t0+7, nhi: t0+8..t0+15).
// Two 32-bit words of each library array are read per call instead of one word per row.
__device__ __forceinline__ void feed16(const uint32_t *__restrict__ lib2, const uint32_t *__restrict__ libn,
                                       const long long start, const int sgn, const int comp,
                                       const int tmin, const int tmax, const int t0, uint32_t &nlo,
                                       uint32_t &nhi)
{
  const int lo = min(max(tmin - t0, 0), 16), hi = min(max(tmax - t0 + 1, 0), 16);
  const uint32_t inb = (hi > lo) ? ((0xffffu >> (16 - hi)) & ~((1u << lo) - 1u)) : 0u;
  uint32_t codes = 0, nm = 0;
  if (inb)
  {
    // lowest sequence index of the group; every in-bounds position lies inside the library
    // (flatten_cores clips the bounds), the arrays are padded by four words on both sides
    const long long p = (sgn > 0) ? start + t0 : start - t0 - 15;
    const long long w2 = p >> 4, wn = p >> 5;
    codes = __funnelshift_r(__ldg(lib2 + w2), __ldg(lib2 + w2 + 1), 2 * (int) (p & 15));
    nm = __funnelshift_r(__ldg(libn + wn), __ldg(libn + wn + 1), (int) (p & 31)) & 0xffffu;
    if (sgn < 0)
    {
      codes = __brev(codes);
      codes = ((codes >> 1) & 0x55555555u) | ((codes & 0x55555555u) << 1);
      nm = __brev(nm) >> 16;
    }
    if (comp)
      codes = ~codes;
  }
  const uint32_t ob = ~inb;
  uint32_t c0 = spread2to4(codes), c1 = spread2to4(codes >> 16);
  const uint32_t n0 = spread1to4(nm), n1 = spread1to4(nm >> 8);
  const uint32_t o0 = spread1to4(ob), o1 = spread1to4(ob >> 8);
  c0 = (c0 & ~(n0 * 15u)) | (n0 << 2);
  c1 = (c1 & ~(n1 * 15u)) | (n1 << 2);
  nlo = (c0 & ~(o0 * 15u)) | (o0 * 5u);
  nhi = (c1 & ~(o1 * 15u)) | (o1 * 5u);
}

// ---------------------------------------------------------------------------------------------
// exact per-cell row pass (rows whose band hangs over a sequence bound, runtime band widths,
// positive gap penalties, debug output)
// ---------------------------------------------------------------------------------------------
struct RowDbg
{
  int *rows;             // [n_ext][W][2] of this row (or NULL)
  unsigned char *path;   // [n_ext][W] of this row (or NULL)
  int inst;
  bool valid;
};

// One pass over the lane's C cells for candidate base a, bnw_extend.c:800-1018 cell by cell:
//   in bounds : S = P + m[a][base], G = max(D, H of the cell to the left)      (D = NEG at offset +B,
//               H(left) = NEG at offset -B)
//   otherwise : S = G = edge value (offset < 0 && row < bandwidth) or the out-of-bounds sentinel
//   cell = max(S, G), H = max(S + go + ge, G + ge)
// hin is H of the cell left of the lane. FINAL: hin is the true value, `best` accumulates the row
// best over the real cells; COMMIT additionally writes the new state (P in place, H of cell k into
// the D slot of cell k-1, H of cell 0 to h0), the first cell equal to bestv and the debug output.
// Returns H of the lane's last cell.
template <int C, bool FINAL, bool COMMIT>
__device__ __forceinline__ int exact_pass(int (&Pp)[C], int (&Dd)[C], const uint32_t *win, const char *lutb,
                                          const int a, const int go, const int ge, const int hin,
                                          const int gl, const int W, const int B, const int r,
                                          const int tmin, const int tmax, int &best, const int bestv,
                                          int &idx, int &h0, const RowDbg &dbg)
{
  const int goe = go + ge;
  const int edge = go + (r + 1) * ge;   // bnw_extend.c:992
  const int tbase = r - B + gl * C;
  int hprev = hin;
  if (COMMIT)
    idx = kBig;
#pragma unroll
  for (int k = 0; k < C; ++k)
  {
    const int K = gl * C + k;
    const bool real = K < W;
    const int t = tbase + k;
    const bool inb = real && (t >= tmin) && (t <= tmax);
    const int E = (K < B && r < B) ? edge : kOob;   // bnw_extend.c:990-1002
    const int Dk = (K >= W - 1) ? kNeg : Dd[k];     // offset == bandwidth: no deletion (:892)
    const int m = *reinterpret_cast<const int *>(lutb + win_byte(win[k >> 2], k & 3) + 4 * a);
    const int S = inb ? Pp[k] + m : E;
    const int Gv = inb ? imax(Dk, hprev) : E;
    const int pn = imax(S, Gv);
    const int hn = imax(S + goe, Gv + ge);
    if (FINAL)
      best = real ? imax(best, pn) : best;
    if (COMMIT)
    {
      idx = min(idx, (real && pn == bestv) ? k : kBig);
      if (dbg.valid && real)
      {
        if (dbg.rows != nullptr)
        {
          int *o = dbg.rows + ((size_t) dbg.inst * W + K) * 2;
          o[0] = S;
          o[1] = Gv;
        }
        if (dbg.path != nullptr)   // sub == cell -> substitution, else del == cell -> deletion, else
                                   // insertion (bnw_extend.c:1036-1041); out of bounds all three are equal
          dbg.path[(size_t) dbg.inst * W + K] = (unsigned char) ((S == pn) ? 0 : (inb && Dk == pn) ? 1 : 2);
      }
      Pp[k] = real ? pn : kNegInf;
      if (k == 0)
        h0 = hn;
      else
        Dd[k - 1] = hn;
    }
    hprev = hn;
  }
  return hprev;
}

// True H of the cell left of each lane from the lanes' local aggregates: an (aggregate, transfer)
// max-plus scan over the G lanes of an instance. `agg` = H of the lane's last cell computed with no
// carry-in; `open` = every cell of the lane is an in-bounds band cell (an out-of-bounds cell restarts
// the chain, bnw_extend.c:988-1003, so nothing travels through such a lane).
template <int C, int G>
__device__ __forceinline__ int carry_scan_exact(int agg, const bool open, const int ge, const int gl)
{
  if (G == 1)
    return kNeg;
  int T = open ? C * ge : kNegInf;
#pragma unroll
  for (int d = 1; d < G; d <<= 1)
  {
    const int Tu = __shfl_up_sync(kFull, T, d);
    const int up = __shfl_up_sync(kFull, agg, d);
    if (gl >= d)
    {
      agg = imax(agg, imax(up, kNegInf) + T);
      T = imax(T + Tu, kNegInf);
    }
  }
  const int cin = __shfl_up_sync(kFull, agg, 1);
  return (gl == 0) ? kNeg : cin;   // offset == -bandwidth: no insertion (bnw_extend.c:972)
}

// Grid-level exchange of the four per-candidate column sums between the CTAs of a family.
// Every CTA publishes its four sums as (tag << 32 | sum) 64-bit words in its own 32-byte slot
// (plain stores: no atomics, no serialisation point in L2); warp 0 of every CTA then gathers all
// slots, kGatherU independent 32-byte loads per lane in flight, and re-reads only the slots whose
// tag (column + 1) has not arrived yet. The tag travels in the same word as the data, so no fence
// is needed and one L2 round trip suffices once the last CTA has published. Slots are double
// buffered by column parity: a CTA overwrites a slot two columns later, after everyone read it.
constexpr int kGatherU = 2;   // 64 slots per round covers the <= kSlotMaxCtas CTAs of this path

__device__ __forceinline__ void ld_slot(const unsigned long long *p, unsigned long long (&w)[4])
{
#ifdef RAMCU_EMUL
  for (int i = 0; i < 4; ++i)
    w[i] = emul_ld_relaxed_u64(p + i);
#else
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];"
               : "=l"(w[0]), "=l"(w[1]) : "l"(p) : "memory");
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];"
               : "=l"(w[2]), "=l"(w[3]) : "l"(p + 2) : "memory");
#endif
}

__device__ __forceinline__ void grid_exchange(unsigned long long *slots, const int n_ctas,
                                              const int cta_in_fam, const int par,
                                              const unsigned tag, unsigned (&sums)[4],
                                              const int lane)
{
  unsigned long long *base = slots + (size_t) par * n_ctas * 4;
  if (lane < 4)
  {
    const unsigned mine = (unsigned) sel4(lane, (int) sums[0], (int) sums[1], (int) sums[2], (int) sums[3]);
    st_relaxed_u64(base + (size_t) cta_in_fam * 4 + lane, ((unsigned long long) tag << 32) | mine);
  }
  unsigned acc[4] = {0, 0, 0, 0};
  for (int b0 = 0; b0 < n_ctas; b0 += 32 * kGatherU)
  {
    unsigned long long w[kGatherU][4];
    bool done[kGatherU];
#pragma unroll
    for (int u = 0; u < kGatherU; ++u)
      done[u] = (b0 + u * 32 + lane) >= n_ctas;
    bool all;
    do
    {
#pragma unroll
      for (int u = 0; u < kGatherU; ++u)
        if (!done[u])
          ld_slot(base + (size_t) (b0 + u * 32 + lane) * 4, w[u]);
      all = true;
#pragma unroll
      for (int u = 0; u < kGatherU; ++u)
        if (!done[u])
        {
          const bool ok = ((unsigned) (w[u][0] >> 32) == tag) && ((unsigned) (w[u][1] >> 32) == tag) &&
                          ((unsigned) (w[u][2] >> 32) == tag) && ((unsigned) (w[u][3] >> 32) == tag);
          if (ok)
          {
            done[u] = true;
            acc[0] += (unsigned) w[u][0];
            acc[1] += (unsigned) w[u][1];
            acc[2] += (unsigned) w[u][2];
            acc[3] += (unsigned) w[u][3];
          }
          else
            all = false;
        }
    } while (!all);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
    sums[a] = (unsigned) __reduce_add_sync(kFull, (int) acc[a]);
}

// Large families (more than kSlotMaxCtas CTAs): the all-to-all slot gather above would make every
// CTA read every slot; instead column r owns four 64-bit accumulators (zeroed before the launch)
// and every CTA adds (1 << 48 | partial) with one RED per candidate, so the arrival count travels
// in the same word as the data: a reader that sees count == n_ctas has the complete sum in the
// low 32 bits (mod 2^32, exactly the reference's int arithmetic). No fence, no reset.
constexpr int kSlotMaxCtas = 40;
constexpr int kRedStride = 1;      // 64-bit words between the accumulators of one column

__device__ __forceinline__ void grid_exchange_red(unsigned long long *acc_row, const int n_ctas,
                                                  unsigned (&sums)[4], const int lane)
{
  unsigned lo = 0;
  if (lane < 4)
  {
    const unsigned mine = (unsigned) sel4(lane, (int) sums[0], (int) sums[1], (int) sums[2], (int) sums[3]);
    // the four accumulators of a column sit in four different 128-byte lines (four L2 slices
    // serve the REDs and the polls in parallel)
    unsigned long long *word = acc_row + lane * kRedStride;
    atomicAdd(word, (1ull << 48) | (unsigned long long) mine);
    unsigned long long v;
    do
    {
      v = ld_relaxed_u64(word);
    } while ((int) (v >> 48) != n_ctas);
    lo = (unsigned) v;
  }
  sums[0] = __shfl_sync(kFull, lo, 0);
  sums[1] = __shfl_sync(kFull, lo, 1);
  sums[2] = __shfl_sync(kFull, lo, 2);
  sums[3] = __shfl_sync(kFull, lo, 3);
}

// per-CTA shared state: score LUT, per-warp partial column sums, broadcast sums, logical CTA id
struct SharedBlk
{
  int4 lut[8];
  int4 part[2][kMaxWarps];
  int4 bcast[2];
  int cta;
};

template <int C, int G, int WS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
extend_kernel(const FamDev *__restrict__ fams, const int *__restrict__ cta_fam, int *ticket)
{
  static_assert(G >= 1 && G <= 32, "1..32 lanes per instance");
  constexpr int IPW = 32 / G;   // instances per warp (lanes IPW*G..31 are spare when 32 % G != 0)
  static_assert(WS == 0 || ((G - 1) * C < WS && WS <= G * C), "band does not tile into G x C");
  constexpr int CW = (C + 3) / 4;
  constexpr int HW = (C + 31) / 32;
  // static band: index of the last real cell inside the last lane
  constexpr int kR = WS ? WS - 1 - (G - 1) * C : C - 1;

  RAMCU_SHARED_STRUCT(SharedBlk, sh);
  int4 (&s_lut)[8] = sh.lut;
  int4 (&s_part)[2][kMaxWarps] = sh.part;
  int4 (&s_bcast)[2] = sh.bcast;
  int &s_cta = sh.cta;

  // logical CTA id in start order: the CTAs of a multi-CTA family are co-resident or ahead of
  // every CTA that could wait for them.
  if (threadIdx.x == 0)
    s_cta = atomicAdd(ticket, 1);
  __syncthreads();
  const int cta = s_cta;
  const FamDev *__restrict__ F = fams + cta_fam[cta];
  const int cta_in_fam = cta - F->cta0;
  if (threadIdx.x < 8)
    s_lut[threadIdx.x] = make_int4(F->m[threadIdx.x][0], F->m[threadIdx.x][1],
                                   F->m[threadIdx.x][2], F->m[threadIdx.x][3]);
  __syncthreads();
  const char *lutb = reinterpret_cast<const char *>(s_lut);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int gl = lane % G;
  const int gw = lane / G;
  const bool is_last = (gl == G - 1);

  const int B = F->B, W = WS ? WS : F->W;
  const int go = F->go, ge = F->ge, goe = go + ge, cap = F->cap, minimp = F->minimp;
  const int L = F->L, stop = F->stop, n_ext = F->n_ext, n_ctas = F->n_ctas;
  // non-positive gap penalties: the insertion term of a cell never exceeds the cell left of it, so
  // the row best needs no insertion chain (sweep) and a carry can only raise a prefix of a lane
  const bool gaps_ok = (go <= 0) && (ge <= 0);
  const unsigned char *forced = F->forced;
  int *dbg_rows = F->dbg_rows;
  unsigned char *dbg_path = F->dbg_path;
  const bool dbg_on = (dbg_rows != nullptr) || (dbg_path != nullptr);

  // balanced partition of the family's instances over its CTAs (n_ctas <= n_ext)
  const int first = (int) ((long long) cta_in_fam * n_ext / n_ctas);
  const int next = (int) ((long long) (cta_in_fam + 1) * n_ext / n_ctas);
  const int grp = warp * IPW + gw;
  const bool valid = (gw < IPW) && (grp < next - first);
  const int inst = valid ? first + grp : next - 1;   // padding lanes shadow the last instance
  const InstDev I = F->inst[inst];
  const uint32_t *lib2 = F->lib2, *libn = F->libn;
  const int tmin = I.tmin, tmax = I.tmax;
  const int K0 = gl * C;

  // row state (see the header): boundary row S = G = (o == 0 ? 0 : |o| * ge + go), ram_extend.c:909-946
  int Pp[C], Dd[C];
  uint32_t win[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i)
    win[i] = 0;
  int maxD = kNeg;
#pragma unroll
  for (int k = 0; k < C; ++k)
  {
    const int K = K0 + k;
    const int o = K - B;
    const int v = (o == 0) ? 0 : abs(o) * ge + go;
    const int v1 = (o + 1 == 0) ? 0 : abs(o + 1) * ge + go;
    Pp[k] = (K < W) ? v : kNegInf;
    Dd[k] = (K < W - 1) ? imax(v1 + goe, v1 + ge) : (K == W - 1) ? kNeg : kNegInf;
    maxD = imax(maxD, Dd[k]);
    const int code = fetch_code(lib2, libn, I.start, I.sgn, I.comp, tmin, tmax, o);
    win[k >> 2] |= (uint32_t) (code << 4) << (8 * (k & 3));
  }

  const int nreal = min(max(W - K0, 0), C);   // real (in-band) cells of this lane
  // rows in which every lane of this warp has its whole segment inside the sequence bounds, and
  // the first row from which every band of the warp has left its window for good
  const int r_fast_lo = __reduce_max_sync(kFull, tmin + B - K0);
  const int r_fast_hi = __reduce_min_sync(kFull, tmax + B - K0 - nreal + 1);
  const int r_dead = max(__reduce_max_sync(kFull, tmax) + B + 1, B);
  const bool use_fast = (WS != 0) && gaps_ok && !dbg_on;

#ifndef RAMCU_EMUL
  const long long t_begin = F->dbg_wait ? clock64() : 0;
#endif
  int high = 0, pos = 0, thigh = 0, tpos = 0;   // ram_extend.c:900-914
  int max_row = -1, max_ext = 0, stop_row = -1;
  uint32_t nib_lo = 0, nib_hi = 0;   // codes of the next band positions entering at the high end
  int r = 0;

  for (r = 0; r < L; ++r)
  {
    const int par = r & 1;
    if ((r & 15) == 0 && is_last)
      feed16(lib2, libn, I.start, I.sgn, I.comp, tmin, tmax, r - B + G * C, nib_lo, nib_hi);

    const bool fast = use_fast && (r >= r_fast_lo) && (r <= r_fast_hi);
    const bool alldead = !dbg_on && (r >= r_dead);   // every cell out of bounds, for good

    int best[4] = {kOob, kOob, kOob, kOob};

    // ---- sweep: row best of every candidate base ---------------------------------------------
    if (!alldead)
    {
#pragma unroll
      for (int a = 0; a < 4; ++a)
        best[a] = kNeg;   // bnw_extend.c:760
      if (fast)
      {
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(win[k >> 2], k & 3));
          best[0] = imax(best[0], Pp[k] + m4.x);
          best[1] = imax(best[1], Pp[k] + m4.y);
          best[2] = imax(best[2], Pp[k] + m4.z);
          best[3] = imax(best[3], Pp[k] + m4.w);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
          best[a] = imax(best[a], maxD);
      }
      else if (gaps_ok)
      {
        const int edge = go + (r + 1) * ge;
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int K = K0 + k;
          const bool real = K < W;
          const int t = r - B + K;
          const bool inb = real && (t >= tmin) && (t <= tmax);
          const int E = (K < B && r < B) ? edge : kOob;
          const int Dk = (K >= W - 1) ? kNeg : Dd[k];
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(win[k >> 2], k & 3));
          const int v0 = inb ? imax(Pp[k] + m4.x, Dk) : E;
          const int v1 = inb ? imax(Pp[k] + m4.y, Dk) : E;
          const int v2 = inb ? imax(Pp[k] + m4.z, Dk) : E;
          const int v3 = inb ? imax(Pp[k] + m4.w, Dk) : E;
          best[0] = real ? imax(best[0], v0) : best[0];
          best[1] = real ? imax(best[1], v1) : best[1];
          best[2] = real ? imax(best[2], v2) : best[2];
          best[3] = real ? imax(best[3], v3) : best[3];
        }
      }
      else
      {
        // positive gap penalties: the full row of every candidate, without committing it
        const bool open = (r - B + K0 >= tmin) && (r - B + K0 + C - 1 <= tmax) && (nreal == C);
        const RowDbg nodbg = {nullptr, nullptr, 0, false};
#pragma unroll 1
        for (int a = 0; a < 4; ++a)
        {
          int nob = 0, noi = 0, noh = 0;
          const int agg = exact_pass<C, false, false>(Pp, Dd, win, lutb, a, go, ge, (gl == 0) ? kNeg : kNegInf, gl, W,
                                                      B, r, tmin, tmax, nob, 0, noi, noh, nodbg);
          const int hin = carry_scan_exact<C, G>(agg, open, ge, gl);
          int b = kNeg;
          exact_pass<C, true, false>(Pp, Dd, win, lutb, a, go, ge, hin, gl, W, B, r, tmin, tmax, b, 0, noi, noh,
                                     nodbg);
          best[a] = b;
        }
      }
      // row best over the instance's G lanes
#pragma unroll
      for (int a = 0; a < 4; ++a)
        best[a] = group_max<G>(best[a], gl, lane);
    }

    // ---- column sums over the family (ram_extend.c:1042-1062) -------------------------------
    unsigned sums[4];
    {
      const bool lead = valid && (gl == 0);
      int c[4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
      {
        c[a] = lead ? imax3(best[a], 0, high + cap) : 0;
        c[a] = __reduce_add_sync(kFull, c[a]);
      }
      if (lane == 0)
        s_part[par][warp] = make_int4(c[0], c[1], c[2], c[3]);
      __syncthreads();
      {
        const int4 p = (lane < nwarps) ? s_part[par][lane] : make_int4(0, 0, 0, 0);
        sums[0] = (unsigned) __reduce_add_sync(kFull, p.x);
        sums[1] = (unsigned) __reduce_add_sync(kFull, p.y);
        sums[2] = (unsigned) __reduce_add_sync(kFull, p.z);
        sums[3] = (unsigned) __reduce_add_sync(kFull, p.w);
      }
      if (n_ctas > 1)
      {
        if (warp == 0)
        {
#ifndef RAMCU_EMUL
          const long long t_ex0 = F->dbg_wait ? clock64() : 0;
#endif
          if (n_ctas > kSlotMaxCtas)
            grid_exchange_red(F->slots + (size_t) r * 4 * kRedStride, n_ctas, sums, lane);
          else
            grid_exchange(F->slots, n_ctas, cta_in_fam, par, (unsigned) (r + 1), sums, lane);
#ifndef RAMCU_EMUL
          if (F->dbg_wait && lane == 0)
            F->dbg_wait[2 * cta_in_fam] += clock64() - t_ex0;
#endif
          if (lane == 0)
            s_bcast[par] = make_int4((int) sums[0], (int) sums[1], (int) sums[2], (int) sums[3]);
        }
        __syncthreads();
        const int4 p = s_bcast[par];
        sums[0] = (unsigned) p.x;
        sums[1] = (unsigned) p.y;
        sums[2] = (unsigned) p.z;
        sums[3] = (unsigned) p.w;
      }
    }

    // ---- argmax over the four bases (ram_extend.c:973-974, 1081-1085) ------------------------
    int curr = 0, besta = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
      if ((int) sums[a] > curr)
      {
        curr = (int) sums[a];
        besta = a;
      }
    }
    if (forced != nullptr)
    {
      besta = forced[r] & 3;
      curr = (int) sums[besta];
    }

    if (F->dbg_best != nullptr)
    {
      if (valid && gl == 0)
      {
#pragma unroll
        for (int a = 0; a < 4; ++a)
          F->dbg_best[((size_t) r * n_ext + inst) * 4 + a] = best[a];
      }
      if (cta_in_fam == 0 && threadIdx.x < 4)
        F->dbg_sums[(size_t) r * 4 + threadIdx.x] =
          sel4(threadIdx.x, (int) sums[0], (int) sums[1], (int) sums[2], (int) sums[3]);
    }

    // ---- commit the chosen base's row (ram_extend.c:1105-1150) -------------------------------
    if (!alldead)
    {
      const int bestb = sel4(besta, best[0], best[1], best[2], best[3]);
      int idx = kBig;
      if (fast)
      {
        // local pass: S, the chain y/H and the cell value with no carry from the lane to the left
        uint32_t wa[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i)
          wa[i] = win[i] | (0x04040404u * (unsigned) besta);   // LUT offset of m[code][besta]
        unsigned hits[HW];
#pragma unroll
        for (int i = 0; i < HW; ++i)
          hits[i] = 0;
        int hprev = (gl == 0) ? kNeg : kNegInf;
        int h0 = 0;
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int m = *reinterpret_cast<const int *>(lutb + win_byte(wa[k >> 2], k & 3));
          const int S = Pp[k] + m;
          const int y = imax(S + goe, Dd[k] + ge);
          const int pn = imax3(S, Dd[k], hprev);
          const int hn = imax(hprev + ge, y);
          Pp[k] = pn;
          if (k == 0)
            h0 = hn;
          else
            Dd[k - 1] = hn;
          hprev = hn;
          if (G == 1 && pn == bestb && k <= kR)
            hits[k >> 5] |= 1u << (k & 31);
        }
        if (G > 1)
        {
          // carry: max-plus prefix scan of the lane aggregates, decay C * ge per lane. Lanes without
          // a left neighbour at distance d add kNegInf instead of being predicated: every value
          // that travels is a real cell score, so the sum stays above INT_MIN and below every score
          int agg = hprev;
#pragma unroll
          for (int d = 1; d < G; d <<= 1)
          {
            const int pen = (gl >= d) ? d * C * ge : kNegInf;
            agg = imax(agg, __shfl_up_sync(kFull, agg, d) + pen);
          }
          int c = __shfl_up_sync(kFull, agg, 1);
          c = (gl > 0) ? c : kNegInf;
          // fold the carry in: cell k sees carry + k * ge, its H carry + (k+1) * ge
#pragma unroll
          for (int k = 0; k < C; ++k)
          {
            const int pn = imax(Pp[k], c);
            Pp[k] = pn;
            if (pn == bestb && (k <= kR || !is_last))
              hits[k >> 5] |= 1u << (k & 31);
            c += ge;
            if (k == 0)
              h0 = imax(h0, c);
            else
              Dd[k - 1] = imax(Dd[k - 1], c);
          }
          Dd[C - 1] = __shfl_down_sync(kFull, h0, 1);
        }
        // cells right of the band: offset == bandwidth has no deletion, padding cells stay inert
#pragma unroll
        for (int k = kR; k < C; ++k)
        {
          if (G == 1)
          {
            Dd[k] = (k == kR) ? kNeg : kNegInf;
            if (k > kR)
              Pp[k] = kNegInf;
          }
          else
          {
            Dd[k] = is_last ? ((k == kR) ? kNeg : kNegInf) : Dd[k];
            if (k > kR)
              Pp[k] = is_last ? kNegInf : Pp[k];
          }
        }
        int md = kNeg;
#pragma unroll
        for (int k = 0; k + 1 < C; k += 2)
          md = imax3(md, Dd[k], Dd[k + 1]);
        if (C & 1)
          md = imax(md, Dd[C - 1]);
        maxD = md;
#pragma unroll
        for (int i = HW - 1; i >= 0; --i)
          if (hits[i])
            idx = 32 * i + __ffs((int) hits[i]) - 1;
      }
      else
      {
        const bool open = (r - B + K0 >= tmin) && (r - B + K0 + C - 1 <= tmax) && (nreal == C);
        const RowDbg nodbg = {nullptr, nullptr, 0, false};
        const RowDbg dbg = {dbg_rows ? dbg_rows + (size_t) r * n_ext * W * 2 : nullptr,
                            dbg_path ? dbg_path + (size_t) r * n_ext * W : nullptr, inst, valid};
        int nob = 0, noi = 0, h0 = 0;
        const int agg = exact_pass<C, false, false>(Pp, Dd, win, lutb, besta, go, ge, (gl == 0) ? kNeg : kNegInf, gl,
                                                    W, B, r, tmin, tmax, nob, 0, noi, h0, nodbg);
        const int hin = carry_scan_exact<C, G>(agg, open, ge, gl);
        exact_pass<C, true, true>(Pp, Dd, win, lutb, besta, go, ge, hin, gl, W, B, r, tmin, tmax, nob, bestb, idx,
                                  h0, dbg);
        Dd[C - 1] = (G > 1) ? __shfl_down_sync(kFull, h0, 1) : kNeg;
        int md = kNeg;
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int K = K0 + k;
          Dd[k] = (K < W - 1) ? Dd[k] : (K == W - 1) ? kNeg : kNegInf;
          md = imax(md, Dd[k]);
        }
        maxD = md;
      }
      int kidx = (idx < kBig) ? K0 + idx : kBig;
      kidx = group_min<G>(kidx, gl, lane);
      if (bestb > high)   // record setter, ram_extend.c:1140-1150
      {
        high = bestb;
        pos = r - B + kidx;
      }
      if (dbg_path != nullptr && valid && gl == 0)
        F->dbg_idx[(size_t) r * n_ext + inst] = kidx;
    }

    // ---- improvement / stop rule (ram_extend.c:1194-1223) ------------------------------------
    if (curr >= max_ext + abs(max_row - r) * minimp)
    {
      max_row = r;
      max_ext = curr;
      thigh = high;
      tpos = pos;
    }
    if (cta_in_fam == 0 && threadIdx.x == 0)
      F->master[r] = (unsigned char) besta;
    if (r - max_row >= stop)
    {
      stop_row = r;
      break;
    }

    // ---- slide the base window by one band position ------------------------------------------
    {
      uint32_t in = (nib_lo << 4) & 0xf0u;
      nib_lo = __funnelshift_r(nib_lo, nib_hi, 4);
      nib_hi >>= 4;
      if (G > 1)
      {
        const uint32_t nb = __shfl_down_sync(kFull, win[0] & 0xffu, 1);
        in = is_last ? in : nb;
      }
#pragma unroll
      for (int i = 0; i + 1 < CW; ++i)
        win[i] = __funnelshift_r(win[i], win[i + 1], 8);
      win[CW - 1] >>= 8;
      win[(C - 1) >> 2] |= in << (8 * ((C - 1) & 3));
    }
  }

#ifndef RAMCU_EMUL
  if (F->dbg_wait && threadIdx.x == 0)
    F->dbg_wait[2 * cta_in_fam + 1] = clock64() - t_begin;
#endif
  // ---- write back (ram_extend.c:1234-1257) ---------------------------------------------------
  if (valid && gl == 0)
  {
    F->thigh[inst] = thigh;
    F->tpos[inst] = tpos;
  }
  if (cta_in_fam == 0 && threadIdx.x == 0)
  {
    F->result[0] = max_row;
    F->result[1] = (stop_row >= 0) ? stop_row + 1 : L;
    F->result[2] = stop_row;
    F->result[3] = max_ext;
  }
}

}  // namespace ramcu
