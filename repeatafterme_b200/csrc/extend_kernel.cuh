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
// -DRAMCU_CHECK: a build with in-kernel index assertions (library word indices, instance / CTA indices, row
// indices of every per-row output); tools/build_variant.sh check "-DRAMCU_CHECK", then the GPU suite under
// RAMCU_LIB=tools/variants/libramcu_check.so. A failed assertion traps the kernel and the host reports it.
// The adds of the sweep go to the FMA pipe (add_fma) unless -DRAMCU_NO_SWEEP_FMA (A/B builds): config 4 16.7 -> 16.4 ms,
// a 96-family batch 317 -> 315 ms.
#ifndef RAMCU_NO_SWEEP_FMA
#define RAMCU_SWEEP_FMA 1
#endif
// The sweep reads the four candidate scores of a cell's base as int16 in one 64-bit load and adds the selected
// half with IDP.2A (__dp2a_lo) unless -DRAMCU_NO_SWEEP_DP2A: half the registers per load in flight, so the lone
// warp of a scheduler waits less for its score loads (config 4: 16.8 -> 16.15 ms). Matrices beyond int16 take the
// exact path.
#ifndef RAMCU_NO_SWEEP_DP2A
#define RAMCU_SWEEP_DP2A 1
#endif
#ifdef RAMCU_CHECK
#include <assert.h>
#define RAMCU_ASSERT(cond) assert(cond)
#else
#define RAMCU_ASSERT(cond) ((void) 0)
#endif

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
  unsigned long long *slots;    // column exchange area of a multi-CTA family, zeroed before each
                                // launch: [L][4] (count << 48 | sum) accumulators
  const unsigned char *forced;  // debug: forced consensus per row (or NULL)
  int *dbg_rows;          // debug: [rows][n_ext][W][2]
  int *dbg_best;          // debug: [rows][n_ext][4]
  int *dbg_sums;          // debug: [rows][4]
  unsigned char *dbg_path;   // -outmat: [rows][n_ext][W] state of the best path into each cell of the
                             // committed row: 0 substitution, 1 deletion, 2 insertion (bnw_extend.c:1027-1044)
  int *dbg_idx;           // -outmat: [rows][n_ext] band index of the row's first best cell
  int *seed_rows;         // seed extension: [rows][n_ext][2] = compute_nw_row()'s return value for the given
                          // consensus base and *max_score_sequence_idx (bnw_extend.c:1020-1024)
  int *dbg_cand;          // verbose dumps (VERBOSE >= 10, ram_extend.c:992-1075): [dbg_nrows][n_ext][4][3] per candidate
                          // base: *max_score_sequence_idx of its row, the gap value of band cell 0 and of band
                          // cell 2B (the reference's **OUT_OF_SEQ** tests, ram_extend.c:1037-1039, 1139-1141)
  int *dbg_cells;         // VERBOSE >= 12 (ram_extend.c:1013-1024): [dbg_nrows][n_ext][4][W][2] (sub, gap) of every
                          // candidate's row (or NULL)
  long long *dbg_wait;    // diagnostics: per CTA [0] cycles spent in the grid exchange [1] total cycles
  int n_ext, n_ctas, cta0, B;
  int W, go, ge, cap;
  int minimp, L, stop;
  int mate;               // index of the family's other pass in the same launch (two-pass CTAs), or -1
  int dbg_nrows;          // rows the dbg_* arrays hold (rows beyond are not reported)
  long long lib_len;      // bases in the library (index assertions of the RAMCU_CHECK build)
  int one;                // the constant 1, opaque to the compiler (add_fma)
  int m[8][4];            // m[code][cand]; code 4 = N, 5 = out of bounds (value unused)
  int mN[8];              // mN[code]: score of an N consensus base (forced consensus 4, seed extension)
};

__device__ __forceinline__ int imax(int a, int b) { return max(a, b); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(max(a, b), c); }
// a + b issued on the FMA pipe: a * one + b with a multiplier the compiler cannot see through (ptxas turns a
// literal 1 back into an add and fuses it with the max). With one warp per scheduler the INT32 ALU pipe (two
// cycles per warp instruction) is what a row waits for, and the FMA pipe is idle.
__device__ __forceinline__ int add_fma(int a, int b, int one)
{
  return a * one + b;
}
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
                                          int sgn, int comp, int tmin, int tmax, int t, const long long lib_len)
{
  if (t < tmin || t > tmax)
    return 5;
  const long long p = start + (long long) sgn * t;
  RAMCU_ASSERT(p >= 0 && p < lib_len);
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
                                       uint32_t &nhi, const long long lib_len)
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
    // two words of each array; the arrays are padded by four words on both sides
    RAMCU_ASSERT(w2 >= -4 && w2 + 1 < ((lib_len + 31) / 32) * 2 + 4);
    RAMCU_ASSERT(wn >= -4 && wn + 1 < (lib_len + 31) / 32 + 4);
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

// Row of candidate base a WITHOUT committing it, for the verbose dumps: the same cell loop as exact_pass
// with the true hin; reports the lane's first cell equal to bestv, the gap value of band cell 0 / W-1 when
// the lane owns it, and (cells != NULL) the (sub, gap) pair of every real cell.
template <int C>
__device__ __forceinline__ void probe_pass(const int (&Pp)[C], const int (&Dd)[C], const uint32_t *win,
                                           const char *lutb, const int a, const int go, const int ge,
                                           const int hin, const int gl, const int W, const int B, const int r,
                                           const int tmin, const int tmax, const int bestv, int &idx,
                                           int &g_first, int &g_last, int *cells)
{
  const int goe = go + ge;
  const int edge = go + (r + 1) * ge;
  const int tbase = r - B + gl * C;
  int hprev = hin;
  idx = kBig;
#pragma unroll
  for (int k = 0; k < C; ++k)
  {
    const int K = gl * C + k;
    const bool real = K < W;
    const int t = tbase + k;
    const bool inb = real && (t >= tmin) && (t <= tmax);
    const int E = (K < B && r < B) ? edge : kOob;
    const int Dk = (K >= W - 1) ? kNeg : Dd[k];
    const int m = *reinterpret_cast<const int *>(lutb + win_byte(win[k >> 2], k & 3) + 4 * a);
    const int S = inb ? Pp[k] + m : E;
    const int Gv = inb ? imax(Dk, hprev) : E;
    const int pn = imax(S, Gv);
    idx = min(idx, (real && pn == bestv) ? k : kBig);
    if (K == 0)
      g_first = Gv;
    if (K == W - 1)
      g_last = Gv;
    if (cells != nullptr && real)
    {
      cells[2 * K] = S;
      cells[2 * K + 1] = Gv;
    }
    hprev = imax(S + goe, Gv + ge);
  }
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

// Column exchange between the CTAs of a family: column r owns four 64-bit accumulators (zeroed before
// the launch) and every CTA adds (1 << 48 | partial) with one RED per candidate, so the arrival count
// travels in the same word as the data: a reader that sees count == n_ctas has the complete sum in
// the low 32 bits (mod 2^32, exactly the reference's int arithmetic). No fence, no reset.
constexpr int kRedStride = 1;      // 64-bit words between the accumulators of one column

// ---------------------------------------------------------------------------------------------
// per-column handshake of one pass: CTA-level sum of the warps' contributions, grid-level sum of
// the CTAs' contributions, delivery of the column sums to every warp -- without a CTA barrier
// ---------------------------------------------------------------------------------------------
// Column c uses slot c & 1. Contribution: lane 0 of every warp adds its four partial sums to
// acc[slot] (shared-memory atomics); after the one CTA barrier of the row and pass, warp 0 owns the CTA's sums:
// it clears the slot and either publishes them (family on one CTA) or adds them to the column's
// global accumulators (one RED per candidate, the arrival count in the upper bits).
// Delivery: warp 0 polls the global accumulators and publishes the complete sums; every warp waits
// for res_tag[slot] == c + 1 when it next needs the column, i.e. after it has done the other pass's
// row in between. A slot is reused two columns later, which every warp of the CTA can only reach
// after it has consumed this column.
struct PassShared
{
  int acc[2][4];
  int res[2][4];
  int res_tag[2];
};

__device__ __forceinline__ int ld_volatile_shared(const int *p)
{
#ifdef RAMCU_EMUL
  sched_yield();
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  return *reinterpret_cast<const volatile int *>(p);
#endif
}
__device__ __forceinline__ void st_volatile_shared(int *p, int v)
{
#ifdef RAMCU_EMUL
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  *reinterpret_cast<volatile int *>(p) = v;
#endif
}

// Waiting is bounded: the exchange relies on every CTA of a family being resident (launch layout,
// ramcu.cu). If another process or kernel holds SMs of the device that can fail; the kernel then
// traps (tens of seconds) and the host reports an error instead of hanging for ever.
constexpr unsigned kMaxGlobalSpins = 1u << 27;
constexpr unsigned kMaxSharedSpins = 1u << 30;
__device__ __forceinline__ void spin_timeout()
{
#ifdef RAMCU_EMUL
  abort();
#else
  __trap();
#endif
}

// RAMCU_STATS: time stamps of the exchange of kStatCols columns of the first pass (globaltimer, ns)
constexpr int kStatCol0 = 1024, kStatCols = 64;
__device__ __forceinline__ long long global_ns()
{
#ifdef RAMCU_EMUL
  return 0;
#else
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}

// per-CTA shared state: score LUT, the handshake areas of the passes, logical CTA id
template <int NP>
struct SharedBlk
{
  int4 lut[16];   // [code][A,C,G,T | N-consensus,-,-,-]: 32 bytes per base code
  PassShared ps[NP];
  int cta;
};

template <int P>
struct PassTag
{
  static constexpr int value = P;
};

// registers of one pass (one direction of one family) in one thread
template <int C>
struct PassRegs
{
  int Pp[C], Dd[C];             // row state, see the header
  uint32_t win[(C + 3) / 4];    // base window: byte k = LUT offset (code * 32) of cell k's base
  uint32_t nib_lo, nib_hi;      // codes of the next band positions entering at the high end
  int maxD;                     // max of Dd over the lane's real cells
  int best[4];                  // row bests of the pending column (instance-wide)
  int high, pos, thigh, tpos;   // ram_extend.c:900-914
  int max_row, max_ext, stop_row;
  long long start;              // instance geometry (InstDev)
  int sgn, comp, tmin, tmax;
  int r_fast_lo, r_fast_hi, r_dead;
  int inst;
  bool valid, active;
};

// EX: the instantiation also carries the code of the diagnostic and per-row outputs (forced consensus, debug
// rows, -outmat path states, per-row seed values, verbose dumps, RAMCU_STATS). The lean instantiations
// (EX = false) are what plain extensions run: none of that code, fewer registers (e.g. <27,3,81,384,1>: 136
// instead of 204 bytes of spill stores).
template <int C, int G, int WS, int MAXT, int NP, bool EX>
__global__ void __launch_bounds__(MAXT, 1)
extend_kernel(const FamDev *__restrict__ fams, const int *__restrict__ cta_fam, int *ticket)
{
  static_assert(G >= 1 && G <= 32, "1..32 lanes per instance");
  static_assert(NP == 1 || NP == 2, "one pass, or the right and the left pass of a family");
  constexpr int IPW = 32 / G;   // instances per warp (lanes IPW*G..31 are spare when 32 % G != 0)
  static_assert(WS == 0 || ((G - 1) * C < WS && WS <= G * C), "band does not tile into G x C");
  constexpr int CW = (C + 3) / 4;
  constexpr int HW = (C + 31) / 32;
  // static band: index of the last real cell inside the last lane
  constexpr int kR = WS ? WS - 1 - (G - 1) * C : C - 1;
  static_assert(!EX || NP == 1, "diagnostic outputs exist for single passes only");
  constexpr bool kExtras = EX;

  RAMCU_SHARED_STRUCT(SharedBlk<NP>, sh);

  // logical CTA id in start order: the CTAs of a multi-CTA family are co-resident or ahead of
  // every CTA that could wait for them.
  if (threadIdx.x == 0)
    sh.cta = atomicAdd(ticket, 1);
  for (int i = threadIdx.x; i < NP * (int) (sizeof(PassShared) / sizeof(int)); i += blockDim.x)
    reinterpret_cast<int *>(sh.ps)[i] = 0;
  __syncthreads();
  const int cta = sh.cta;
  const FamDev *__restrict__ F0 = fams + cta_fam[cta];
  const int cta_in_fam = cta - F0->cta0;
  RAMCU_ASSERT(cta >= 0 && cta < (int) gridDim.x && cta_in_fam >= 0 && cta_in_fam < F0->n_ctas);
  if (threadIdx.x < 8)
  {
    sh.lut[2 * threadIdx.x] = make_int4(F0->m[threadIdx.x][0], F0->m[threadIdx.x][1],
                                        F0->m[threadIdx.x][2], F0->m[threadIdx.x][3]);
#ifdef RAMCU_SWEEP_DP2A
    // bytes 24..31 of the entry: the four candidate scores as int16 (A | C << 16, G | T << 16) for the sweep
    sh.lut[2 * threadIdx.x + 1] =
        make_int4(F0->mN[threadIdx.x], 0,
                  (int) (((unsigned) F0->m[threadIdx.x][0] & 0xffffu) | ((unsigned) F0->m[threadIdx.x][1] << 16)),
                  (int) (((unsigned) F0->m[threadIdx.x][2] & 0xffffu) | ((unsigned) F0->m[threadIdx.x][3] << 16)));
#else
    sh.lut[2 * threadIdx.x + 1] = make_int4(F0->mN[threadIdx.x], 0, 0, 0);
#endif
  }
  __syncthreads();
  const char *lutb = reinterpret_cast<const char *>(sh.lut);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int gl = lane % G;
  const int gw = lane / G;
  const bool is_last = (gl == G - 1);
  const int K0 = gl * C;

  // scalar parameters: the passes of a launch unit are the two directions of one family
  const int B = F0->B, W = WS ? WS : F0->W;
  const int go = F0->go, ge = F0->ge, goe = go + ge, cap = F0->cap, minimp = F0->minimp;
  const int L = F0->L, stop = F0->stop, n_ctas = F0->n_ctas;
  // non-positive gap penalties: the insertion term of a cell never exceeds the cell left of it, so
  // the row best needs no insertion chain (sweep) and a carry can only raise a prefix of a lane
  const bool gaps_ok = (go <= 0) && (ge <= 0);
#ifdef RAMCU_SWEEP_DP2A
  // the sweep reads the scores as int16: a matrix beyond that range takes the exact path on every row
  bool m16 = true;
  for (int i = 0; i < 8; ++i)
    for (int a = 0; a < 4; ++a)
      m16 = m16 && (F0->m[i][a] >= -32768) && (F0->m[i][a] <= 32767);
#else
  const bool m16 = true;
#endif
  const int nreal = min(max(W - K0, 0), C);   // real (in-band) cells of this lane

  const FamDev *Fp[NP];
  PassRegs<C> ps[NP];
  bool dbg_on[NP], use_fast[NP], has_forced[NP], has_dbg_best[NP], has_dbg_path[NP], has_seed_rows[NP], has_dbg_cand[NP];

  // ---- set-up of every pass: instance, boundary row, first base window ------------------------
#pragma unroll
  for (int p = 0; p < NP; ++p)
  {
    const FamDev *__restrict__ F = (p == 0) ? F0 : ((F0->mate >= 0) ? fams + F0->mate : F0);
    Fp[p] = F;
    PassRegs<C> &S = ps[p];
    S.active = (p == 0) || (F0->mate >= 0);
    const int n_ext = F->n_ext;
    // balanced partition of the pass's instances over the CTAs of the launch unit
    const int first = (int) ((long long) cta_in_fam * n_ext / n_ctas);
    const int next = (int) ((long long) (cta_in_fam + 1) * n_ext / n_ctas);
    const int grp = warp * IPW + gw;
    S.valid = S.active && (gw < IPW) && (grp < next - first);
    S.inst = S.valid ? first + grp : max(min(next, n_ext) - 1, 0);   // other lanes shadow an instance
    RAMCU_ASSERT(S.inst >= 0 && S.inst < n_ext);
    const InstDev I = F->inst[S.inst];
    S.start = I.start;
    S.sgn = I.sgn;
    S.comp = I.comp;
    S.tmin = I.tmin;
    S.tmax = I.tmax;
    // forced consensus, debug rows, -outmat states and per-row seed values exist for single passes only
    // (ramcu_job_create_pair never asks for them): the two-pass kernel carries none of that code
    dbg_on[p] = kExtras && ((F->dbg_rows != nullptr) || (F->dbg_path != nullptr) || (F->dbg_cand != nullptr));
    has_forced[p] = kExtras && F->forced != nullptr;
    has_dbg_best[p] = kExtras && F->dbg_best != nullptr;
    has_dbg_path[p] = kExtras && F->dbg_path != nullptr;
    has_seed_rows[p] = kExtras && F->seed_rows != nullptr;
    has_dbg_cand[p] = kExtras && F->dbg_cand != nullptr;
    use_fast[p] = (WS != 0) && gaps_ok && m16 && !dbg_on[p];
#pragma unroll
    for (int i = 0; i < CW; ++i)
      S.win[i] = 0;
    S.maxD = kNeg;
    // boundary row S = G = (o == 0 ? 0 : |o| * ge + go), ram_extend.c:909-946
#pragma unroll
    for (int k = 0; k < C; ++k)
    {
      const int K = K0 + k;
      const int o = K - B;
      const int v = (o == 0) ? 0 : abs(o) * ge + go;
      const int v1 = (o + 1 == 0) ? 0 : abs(o + 1) * ge + go;
      S.Pp[k] = (K < W) ? v : kNegInf;
      S.Dd[k] = (K < W - 1) ? imax(v1 + goe, v1 + ge) : (K == W - 1) ? kNeg : kNegInf;
      S.maxD = imax(S.maxD, S.Dd[k]);
    }
#pragma unroll 1
    for (int k = 0; k < C; ++k)
    {
      const int code = fetch_code(F->lib2, F->libn, I.start, I.sgn, I.comp, S.tmin, S.tmax, K0 + k - B, F->lib_len);
      const uint32_t b = (uint32_t) (code << 5) << (8 * (k & 3));
#pragma unroll
      for (int i = 0; i < CW; ++i)
        S.win[i] |= (i == (k >> 2)) ? b : 0u;
    }
    // rows in which every lane of this warp has its whole segment inside the sequence bounds, and
    // the first row from which every band of the warp has left its window for good
    S.r_fast_lo = __reduce_max_sync(kFull, S.tmin + B - K0);
    S.r_fast_hi = __reduce_min_sync(kFull, S.tmax + B - K0 - nreal + 1);
    S.r_dead = max(__reduce_max_sync(kFull, S.tmax) + B + 1, B);
    S.high = S.pos = S.thigh = S.tpos = 0;
    S.max_row = -1;
    S.max_ext = 0;
    S.stop_row = -1;
    S.nib_lo = S.nib_hi = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
      S.best[a] = kOob;
    if (L <= 0)
      S.active = false;
  }

#ifndef RAMCU_EMUL
  const bool stats = kExtras && (F0->dbg_wait != nullptr);   // RAMCU_STATS diagnostics: single passes only
  const long long t_begin = stats ? clock64() : 0;
  long long t_wait = 0;
#endif

  // ---- sweep of row r: row best of every candidate base, then the pass's contribution to the
  //      column sums (ram_extend.c:1042-1062) goes on its way ---------------------------------------
  auto sweep_and_send = [&](auto pc, const int r) {
    constexpr int p = decltype(pc)::value;
    PassRegs<C> &S = ps[p];
    const FamDev *__restrict__ F = Fp[p];
    const bool fast = use_fast[p] && (r >= S.r_fast_lo) && (r <= S.r_fast_hi);
    const bool alldead = !dbg_on[p] && !has_seed_rows[p] && (r >= S.r_dead);   // every cell out of bounds, for good
    int best[4] = {kOob, kOob, kOob, kOob};
    if (!alldead)
    {
#pragma unroll
      for (int a = 0; a < 4; ++a)
        best[a] = kNeg;   // bnw_extend.c:760
      if (fast)
      {
#if defined(RAMCU_SWEEP_DP2A)
        // two cells per step; the four candidate scores of a cell's base travel as int16 in ONE 64-bit load (half
        // the registers per load in flight of the int4 form) and the add selects its half: IDP.2A, c + lo * 1 + hi * 0
        const int slo = F->one, shi = F->one << 8;   // (1, 0) and (0, 1) as bytes, opaque to the compiler
#pragma unroll
        for (int k = 0; k + 1 < C; k += 2)
        {
          const int2 m2 = *reinterpret_cast<const int2 *>(lutb + 24 + win_byte(S.win[k >> 2], k & 3));
          const int2 n2 = *reinterpret_cast<const int2 *>(lutb + 24 + win_byte(S.win[(k + 1) >> 2], (k + 1) & 3));
          best[0] = imax3(best[0], __dp2a_lo(m2.x, slo, S.Pp[k]), __dp2a_lo(n2.x, slo, S.Pp[k + 1]));
          best[1] = imax3(best[1], __dp2a_lo(m2.x, shi, S.Pp[k]), __dp2a_lo(n2.x, shi, S.Pp[k + 1]));
          best[2] = imax3(best[2], __dp2a_lo(m2.y, slo, S.Pp[k]), __dp2a_lo(n2.y, slo, S.Pp[k + 1]));
          best[3] = imax3(best[3], __dp2a_lo(m2.y, shi, S.Pp[k]), __dp2a_lo(n2.y, shi, S.Pp[k + 1]));
        }
        if (C & 1)
        {
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(S.win[(C - 1) >> 2], (C - 1) & 3));
          best[0] = imax(best[0], S.Pp[C - 1] + m4.x);
          best[1] = imax(best[1], S.Pp[C - 1] + m4.y);
          best[2] = imax(best[2], S.Pp[C - 1] + m4.z);
          best[3] = imax(best[3], S.Pp[C - 1] + m4.w);
        }
#elif defined(RAMCU_SWEEP_FMA)
        // two cells per step: the adds go to the FMA pipe, one three-input max per candidate to the ALU pipe
        const int one = F->one;   // 1, but not to the compiler
#pragma unroll
        for (int k = 0; k + 1 < C; k += 2)
        {
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(S.win[k >> 2], k & 3));
          const int4 n4 = *reinterpret_cast<const int4 *>(lutb + win_byte(S.win[(k + 1) >> 2], (k + 1) & 3));
          best[0] = imax3(best[0], add_fma(S.Pp[k], m4.x, one), add_fma(S.Pp[k + 1], n4.x, one));
          best[1] = imax3(best[1], add_fma(S.Pp[k], m4.y, one), add_fma(S.Pp[k + 1], n4.y, one));
          best[2] = imax3(best[2], add_fma(S.Pp[k], m4.z, one), add_fma(S.Pp[k + 1], n4.z, one));
          best[3] = imax3(best[3], add_fma(S.Pp[k], m4.w, one), add_fma(S.Pp[k + 1], n4.w, one));
        }
        if (C & 1)
        {
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(S.win[(C - 1) >> 2], (C - 1) & 3));
          best[0] = imax(best[0], S.Pp[C - 1] + m4.x);
          best[1] = imax(best[1], S.Pp[C - 1] + m4.y);
          best[2] = imax(best[2], S.Pp[C - 1] + m4.z);
          best[3] = imax(best[3], S.Pp[C - 1] + m4.w);
        }
#else
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(S.win[k >> 2], k & 3));
          best[0] = imax(best[0], S.Pp[k] + m4.x);
          best[1] = imax(best[1], S.Pp[k] + m4.y);
          best[2] = imax(best[2], S.Pp[k] + m4.z);
          best[3] = imax(best[3], S.Pp[k] + m4.w);
        }
#endif
#pragma unroll
        for (int a = 0; a < 4; ++a)
          best[a] = imax(best[a], S.maxD);
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
          const bool inb = real && (t >= S.tmin) && (t <= S.tmax);
          const int E = (K < B && r < B) ? edge : kOob;
          const int Dk = (K >= W - 1) ? kNeg : S.Dd[k];
          const int4 m4 = *reinterpret_cast<const int4 *>(lutb + win_byte(S.win[k >> 2], k & 3));
          const int v0 = inb ? imax(S.Pp[k] + m4.x, Dk) : E;
          const int v1 = inb ? imax(S.Pp[k] + m4.y, Dk) : E;
          const int v2 = inb ? imax(S.Pp[k] + m4.z, Dk) : E;
          const int v3 = inb ? imax(S.Pp[k] + m4.w, Dk) : E;
          best[0] = real ? imax(best[0], v0) : best[0];
          best[1] = real ? imax(best[1], v1) : best[1];
          best[2] = real ? imax(best[2], v2) : best[2];
          best[3] = real ? imax(best[3], v3) : best[3];
        }
      }
      else
      {
        // positive gap penalties: the full row of every candidate, without committing it
        const bool open = (r - B + K0 >= S.tmin) && (r - B + K0 + C - 1 <= S.tmax) && (nreal == C);
        const RowDbg nodbg = {nullptr, nullptr, 0, false};
#pragma unroll 1
        for (int a = 0; a < 4; ++a)
        {
          int nob = 0, noi = 0, noh = 0;
          const int agg = exact_pass<C, false, false>(S.Pp, S.Dd, S.win, lutb, a, go, ge, (gl == 0) ? kNeg : kNegInf,
                                                      gl, W, B, r, S.tmin, S.tmax, nob, 0, noi, noh, nodbg);
          const int hin = carry_scan_exact<C, G>(agg, open, ge, gl);
          int b = kNeg;
          exact_pass<C, true, false>(S.Pp, S.Dd, S.win, lutb, a, go, ge, hin, gl, W, B, r, S.tmin, S.tmax, b, 0, noi,
                                     noh, nodbg);
          best[0] = (a == 0) ? b : best[0];
          best[1] = (a == 1) ? b : best[1];
          best[2] = (a == 2) ? b : best[2];
          best[3] = (a == 3) ? b : best[3];
        }
      }
      // row best over the instance's G lanes
#pragma unroll
      for (int a = 0; a < 4; ++a)
        best[a] = group_max<G>(best[a], gl, lane);
      if (kExtras && has_dbg_cand[p] && r < F->dbg_nrows)
      {
        // verbose dumps: what compute_nw_row() leaves behind for EVERY candidate base (ram_extend.c:1006-1039)
        const bool open = (r - B + K0 >= S.tmin) && (r - B + K0 + C - 1 <= S.tmax) && (nreal == C);
        const RowDbg nodbg = {nullptr, nullptr, 0, false};
        const int n_ext = F->n_ext;
#pragma unroll 1
        for (int a = 0; a < 4; ++a)
        {
          int nob = 0, noi = 0, noh = 0;
          const int agg = exact_pass<C, false, false>(S.Pp, S.Dd, S.win, lutb, a, go, ge, (gl == 0) ? kNeg : kNegInf,
                                                      gl, W, B, r, S.tmin, S.tmax, nob, 0, noi, noh, nodbg);
          const int hin = carry_scan_exact<C, G>(agg, open, ge, gl);
          const int bv = sel4(a, best[0], best[1], best[2], best[3]);
          int idx = kBig, gf = -2147483647, gla = -2147483647;
          int *cells = (F->dbg_cells != nullptr && S.valid)
                           ? F->dbg_cells + (((size_t) r * n_ext + S.inst) * 4 + a) * (size_t) W * 2
                           : nullptr;
          probe_pass<C>(S.Pp, S.Dd, S.win, lutb, a, go, ge, hin, gl, W, B, r, S.tmin, S.tmax, bv, idx, gf, gla, cells);
          int kidx = (idx < kBig) ? K0 + idx : kBig;
          kidx = group_min<G>(kidx, gl, lane);
          gf = group_max<G>(gf, gl, lane);
          gla = group_max<G>(gla, gl, lane);
          if (S.valid && gl == 0)
          {
            int *o = F->dbg_cand + (((size_t) r * n_ext + S.inst) * 4 + a) * 3;
            o[0] = r - B + kidx;
            o[1] = gf;
            o[2] = gla;
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      S.best[a] = best[a];

    // contribution of this warp, then of this CTA
    const bool lead = S.valid && (gl == 0);
    int c[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
      c[a] = lead ? imax3(best[a], 0, S.high + cap) : 0;
      c[a] = __reduce_add_sync(kFull, c[a]);
    }
    PassShared &H = sh.ps[p];
    const int slot = r & 1;
    const bool sender = (warp == 0);
    RAMCU_ASSERT(r >= 0 && r <= L);   // the exchange area holds L + 1 columns
    if (nwarps > 1)
    {
      if (lane == 0)
      {
#pragma unroll
        for (int a = 0; a < 4; ++a)
          atomicAdd(&H.acc[slot][a], c[a]);
      }
      // One CTA barrier per row and pass; warp 0 then owns the CTA's sums. (Letting the warp that
      // arrives last send instead, with no barrier at all, measured the same for two passes per thread
      // and 5 % slower for one.)
      __syncthreads();
    }
    if (sender)
    {
      int mine = 0;
      if (nwarps > 1)
      {
        if (lane < 4)
        {
          mine = ld_volatile_shared(&H.acc[slot][lane]);
          st_volatile_shared(&H.acc[slot][lane], 0);   // next used two columns (two barriers) later
        }
      }
      else
        mine = sel4(lane & 3, c[0], c[1], c[2], c[3]);
      if (n_ctas > 1)
      {
        if (lane < 4)   // column r owns four accumulators: (arrivals << 48 | sum), zeroed before the launch
          atomicAdd(F->slots + (size_t) r * 4 * kRedStride + lane * kRedStride,
                    (1ull << 48) | (unsigned long long) (unsigned) mine);
#ifndef RAMCU_EMUL
        if (stats && p == 0 && lane == 0 && r >= kStatCol0 && r < kStatCol0 + kStatCols)
          F0->dbg_wait[2 * n_ctas + ((size_t) (r - kStatCol0) * n_ctas + cta_in_fam) * 3] = global_ns();
#endif
      }
      else
      {
        if (lane < 4)
          st_volatile_shared(&H.res[slot][lane], mine);
        __syncwarp();
        __threadfence_block();
        if (lane == 0)
          st_volatile_shared(&H.res_tag[slot], r + 1);
      }
    }
  };

  // ---- warp 0 asks for the accumulators of a column ahead of time: the load is in flight while the
  //      other pass's row is computed, so that the receive below rarely waits for an L2 round trip -------
  unsigned long long pre[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p)
    pre[p] = 0;
  auto prefetch = [&](auto pc, const int r) {
    constexpr int p = decltype(pc)::value;
    if (n_ctas > 1 && warp == 0 && lane < 4 && ps[p].active)
      pre[p] = ld_relaxed_u64(Fp[p]->slots + (size_t) r * 4 * kRedStride + lane * kRedStride);
  };

  // ---- the column sums of row r arrive --------------------------------------------------------------
  auto receive = [&](auto pc, const int r, int (&sums)[4]) {
    constexpr int p = decltype(pc)::value;
    const FamDev *__restrict__ F = Fp[p];
    PassShared &H = sh.ps[p];
    const int slot = r & 1;
    if (n_ctas > 1 && warp == 0)
    {
#ifndef RAMCU_EMUL
      const long long t0 = stats ? clock64() : 0;
      const bool stamp = stats && p == 0 && lane == 0 && r >= kStatCol0 && r < kStatCol0 + kStatCols;
      if (stamp)
        F0->dbg_wait[2 * n_ctas + ((size_t) (r - kStatCol0) * n_ctas + cta_in_fam) * 3 + 1] = global_ns();
#endif
      unsigned lo = 0;
      if (lane < 4)
      {
        const unsigned long long *word = F->slots + (size_t) r * 4 * kRedStride + lane * kRedStride;
        unsigned long long v = pre[p];   // asked for during the other pass's row (0 if not)
        pre[p] = 0;
        unsigned spins = 0;
        while ((int) (v >> 48) != n_ctas)
        {
          v = ld_relaxed_u64(word);
          if (++spins > kMaxGlobalSpins)
            spin_timeout();   // a CTA of this family never became resident (something else holds the SMs)
        }
        lo = (unsigned) v;
        st_volatile_shared(&H.res[slot][lane], (int) lo);
      }
      __syncwarp();
      __threadfence_block();
      if (lane == 0)
        st_volatile_shared(&H.res_tag[slot], r + 1);
#ifndef RAMCU_EMUL
      if (stats)
        t_wait += clock64() - t0;
      if (stamp)
        F0->dbg_wait[2 * n_ctas + ((size_t) (r - kStatCol0) * n_ctas + cta_in_fam) * 3 + 2] = global_ns();
#endif
    }
    {
      unsigned spins = 0;
      while (ld_volatile_shared(&H.res_tag[slot]) != r + 1)
        if (++spins > kMaxSharedSpins)
          spin_timeout();
    }
    __threadfence_block();
#pragma unroll
    for (int a = 0; a < 4; ++a)
      sums[a] = ld_volatile_shared(&H.res[slot][a]);
  };

  // ---- argmax, commit of row r, record / improvement / stop rule, window slide -----------------------
  auto commit = [&](auto pc, const int r, const int (&sums)[4]) {
    constexpr int p = decltype(pc)::value;
    PassRegs<C> &S = ps[p];
    const FamDev *__restrict__ F = Fp[p];
    const int n_ext = F->n_ext;
    if ((r & 15) == 0 && is_last)
      feed16(F->lib2, F->libn, S.start, S.sgn, S.comp, S.tmin, S.tmax, r - B + G * C, S.nib_lo, S.nib_hi, F->lib_len);
    const bool fast_row = use_fast[p] && (r >= S.r_fast_lo) && (r <= S.r_fast_hi);
    const bool alldead = !dbg_on[p] && !has_seed_rows[p] && (r >= S.r_dead);
    RAMCU_ASSERT(r >= 0 && r < L);    // master[] and every per-row output hold L rows

    // argmax over the four bases (ram_extend.c:973-974, 1081-1085)
    int curr = 0, besta = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
      if (sums[a] > curr)
      {
        curr = sums[a];
        besta = a;
      }
    }
    if (has_forced[p])
    {
      besta = min((int) F->forced[r], 4);   // 4: an N base of a given query (seed extension)
      curr = sel4(besta & 3, sums[0], sums[1], sums[2], sums[3]);
    }
    if (has_dbg_best[p] && r < F->dbg_nrows)
    {
      if (S.valid && gl == 0)
      {
#pragma unroll
        for (int a = 0; a < 4; ++a)
          F->dbg_best[((size_t) r * n_ext + S.inst) * 4 + a] = S.best[a];
      }
      if (cta_in_fam == 0 && threadIdx.x < 4)
        F->dbg_sums[(size_t) r * 4 + threadIdx.x] = sel4(threadIdx.x, sums[0], sums[1], sums[2], sums[3]);
    }

    // commit the chosen base's row (ram_extend.c:1105-1150)
    if (!alldead)
    {
      // an N base of a given query is not one of the four candidates of the sweep: its row takes the exact path
      const bool fast = fast_row && (besta < 4);
      int bestb = sel4(besta, S.best[0], S.best[1], S.best[2], S.best[3]);
      int idx = kBig;
      if (fast)
      {
        // local pass: S, the chain y/H and the cell value with no carry from the lane to the left
        uint32_t wa[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i)
          wa[i] = S.win[i] | (0x04040404u * (unsigned) besta);   // LUT offset of m[code][besta]
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
          const int Sv = S.Pp[k] + m;
          const int y = imax(Sv + goe, S.Dd[k] + ge);   // (these adds on the FMA pipe: 16.4 -> 17.5 ms, not kept)
          const int pn = imax3(Sv, S.Dd[k], hprev);
          const int hn = imax(hprev + ge, y);
          S.Pp[k] = pn;
          if (k == 0)
            h0 = hn;
          else
            S.Dd[k - 1] = hn;
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
          int cy = __shfl_up_sync(kFull, agg, 1);
          cy = (gl > 0) ? cy : kNegInf;
          // fold the carry in: cell k sees carry + k * ge, its H carry + (k+1) * ge
#pragma unroll
          for (int k = 0; k < C; ++k)
          {
            const int pn = imax(S.Pp[k], cy);
            S.Pp[k] = pn;
            if (pn == bestb && (k <= kR || !is_last))
              hits[k >> 5] |= 1u << (k & 31);
            cy += ge;
            if (k == 0)
              h0 = imax(h0, cy);
            else
              S.Dd[k - 1] = imax(S.Dd[k - 1], cy);
          }
          S.Dd[C - 1] = __shfl_down_sync(kFull, h0, 1);
        }
        // cells right of the band: offset == bandwidth has no deletion, padding cells stay inert
#pragma unroll
        for (int k = kR; k < C; ++k)
        {
          if (G == 1)
          {
            S.Dd[k] = (k == kR) ? kNeg : kNegInf;
            if (k > kR)
              S.Pp[k] = kNegInf;
          }
          else
          {
            S.Dd[k] = is_last ? ((k == kR) ? kNeg : kNegInf) : S.Dd[k];
            if (k > kR)
              S.Pp[k] = is_last ? kNegInf : S.Pp[k];
          }
        }
        int md = kNeg;
#pragma unroll
        for (int k = 0; k + 1 < C; k += 2)
          md = imax3(md, S.Dd[k], S.Dd[k + 1]);
        if (C & 1)
          md = imax(md, S.Dd[C - 1]);
        S.maxD = md;
#pragma unroll
        for (int i = HW - 1; i >= 0; --i)
          if (hits[i])
            idx = 32 * i + __ffs((int) hits[i]) - 1;
      }
      else
      {
        const bool open = (r - B + K0 >= S.tmin) && (r - B + K0 + C - 1 <= S.tmax) && (nreal == C);
        const RowDbg nodbg = {nullptr, nullptr, 0, false};
        const bool drow = r < F->dbg_nrows;   // rows the per-row outputs hold
        const RowDbg dbg = {(F->dbg_rows && drow) ? F->dbg_rows + (size_t) r * n_ext * W * 2 : nullptr,
                            (F->dbg_path && drow) ? F->dbg_path + (size_t) r * n_ext * W : nullptr, S.inst, S.valid};
        int nob = 0, noi = 0, h0 = 0;
        const int agg = exact_pass<C, false, false>(S.Pp, S.Dd, S.win, lutb, besta, go, ge, (gl == 0) ? kNeg : kNegInf,
                                                    gl, W, B, r, S.tmin, S.tmax, nob, 0, noi, h0, nodbg);
        const int hin = carry_scan_exact<C, G>(agg, open, ge, gl);
        if (besta == 4)
        {
          int b = kNeg;
          exact_pass<C, true, false>(S.Pp, S.Dd, S.win, lutb, besta, go, ge, hin, gl, W, B, r, S.tmin, S.tmax, b, 0, noi,
                                     h0, nodbg);
          bestb = group_max<G>(b, gl, lane);
        }
        exact_pass<C, true, true>(S.Pp, S.Dd, S.win, lutb, besta, go, ge, hin, gl, W, B, r, S.tmin, S.tmax, nob, bestb,
                                  idx, h0, dbg);
        S.Dd[C - 1] = (G > 1) ? __shfl_down_sync(kFull, h0, 1) : kNeg;
        int md = kNeg;
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int K = K0 + k;
          S.Dd[k] = (K < W - 1) ? S.Dd[k] : (K == W - 1) ? kNeg : kNegInf;
          md = imax(md, S.Dd[k]);
        }
        S.maxD = md;
      }
      int kidx = (idx < kBig) ? K0 + idx : kBig;
      kidx = group_min<G>(kidx, gl, lane);
      if (bestb > S.high)   // record setter, ram_extend.c:1140-1150
      {
        S.high = bestb;
        S.pos = r - B + kidx;
      }
      if (has_dbg_path[p] && S.valid && gl == 0 && r < F->dbg_nrows)
        F->dbg_idx[(size_t) r * n_ext + S.inst] = kidx;
      if (has_seed_rows[p] && S.valid && gl == 0)
      {
        F->seed_rows[((size_t) r * n_ext + S.inst) * 2] = bestb;
        F->seed_rows[((size_t) r * n_ext + S.inst) * 2 + 1] = r - B + kidx;
      }
    }

    // improvement / stop rule (ram_extend.c:1194-1223)
    if (curr >= S.max_ext + abs(S.max_row - r) * minimp)
    {
      S.max_row = r;
      S.max_ext = curr;
      S.thigh = S.high;
      S.tpos = S.pos;
    }
    if (cta_in_fam == 0 && threadIdx.x == 0)
      F->master[r] = (unsigned char) besta;
    if (r - S.max_row >= stop)
    {
      S.stop_row = r;
      S.active = false;
    }
    else if (r + 1 >= L)
      S.active = false;
    if (S.active)
    {
      // slide the base window by one band position
      uint32_t in = (S.nib_lo << 5) & 0x1e0u;
      S.nib_lo = __funnelshift_r(S.nib_lo, S.nib_hi, 4);
      S.nib_hi >>= 4;
      if (G > 1)
      {
        const uint32_t nb = __shfl_down_sync(kFull, S.win[0] & 0xffu, 1);
        in = is_last ? in : nb;
      }
#pragma unroll
      for (int i = 0; i + 1 < CW; ++i)
        S.win[i] = __funnelshift_r(S.win[i], S.win[i + 1], 8);
      S.win[CW - 1] >>= 8;
      S.win[(C - 1) >> 2] |= in << (8 * ((C - 1) & 3));
    }
  };

  // ---- column loop: the passes take turns, so that the exchange of one pass's column sums is in
  //      flight while the other pass's row is computed --------------------------------------------------
  if (ps[0].active)
    sweep_and_send(PassTag<0>(), 0);
  if (NP > 1 && ps[NP - 1].active)
    sweep_and_send(PassTag<NP - 1>(), 0);
  for (int r = 0; r < L; ++r)
  {
    bool any = false;
    if (ps[0].active)
    {
      int sums[4];
      receive(PassTag<0>(), r, sums);
      commit(PassTag<0>(), r, sums);
      if (NP > 1)
        prefetch(PassTag<NP - 1>(), r);
      if (ps[0].active)
        sweep_and_send(PassTag<0>(), r + 1);
      any = true;
    }
    if (NP > 1 && ps[NP - 1].active)
    {
      int sums[4];
      receive(PassTag<NP - 1>(), r, sums);
      commit(PassTag<NP - 1>(), r, sums);
      prefetch(PassTag<0>(), r + 1);
      if (ps[NP - 1].active)
        sweep_and_send(PassTag<NP - 1>(), r + 1);
      any = true;
    }
    if (!any)
      break;
  }

  // ---- write back (ram_extend.c:1234-1257) ---------------------------------------------------
#pragma unroll
  for (int p = 0; p < NP; ++p)
  {
    if (p > 0 && F0->mate < 0)
      continue;
    const FamDev *__restrict__ F = Fp[p];
    const PassRegs<C> &S = ps[p];
    if (S.valid && gl == 0)
    {
      F->thigh[S.inst] = S.thigh;
      F->tpos[S.inst] = S.tpos;
    }
    if (cta_in_fam == 0 && threadIdx.x == 0)
    {
      F->result[0] = S.max_row;
      F->result[1] = (S.stop_row >= 0) ? S.stop_row + 1 : L;
      F->result[2] = S.stop_row;
      F->result[3] = S.max_ext;
    }
  }
#ifndef RAMCU_EMUL
  if (stats && threadIdx.x == 0)
  {
    F0->dbg_wait[2 * cta_in_fam] = t_wait;
    F0->dbg_wait[2 * cta_in_fam + 1] = clock64() - t_begin;
  }
#endif
}

}  // namespace ramcu
