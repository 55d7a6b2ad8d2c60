// extend_kernel.cuh -- the banded-DP MSA extension kernel (sm_100a).
//
// Replaces, on the device, the column loop of extend_alignment() (reference
// ram_extend.c:970-1224) and the band row of compute_nw_row() (bnw_extend.c:800-1045).
//
// Mapping:
//   * an MSA instance's band of W = 2B+1 cells is split into G contiguous segments of C cells;
//     one lane owns one segment (G lanes per instance, G in {1,2,4,..,32}, G*C >= W). The two
//     DP states (sub/gap) of the committed previous row live in registers: Sp[C], Gp[C].
//   * all four candidate consensus bases are evaluated in one sweep over the segment; per
//     candidate only the running (S,G) pair and the row best are kept.
//   * the in-row insertion chain G[o] = max(X[o], G[o-1]+ge) is a max-plus prefix scan: serial
//     inside a lane, Hillis-Steele over the G lanes with __shfl_up_sync (decay C*ge per lane).
//   * per-candidate best scores are butterfly-reduced over the G lanes (__shfl_xor_sync), capped
//     (max(best,0,high+CAPPENALTY)), summed over the warp (__reduce_add_sync), the CTA (shared
//     memory) and, for families spanning several CTAs, over the grid with a data+flag ("LL")
//     exchange through L2 - no host round trip and no cooperative launch per column.
//   * the chosen base's row is then re-materialised in place (the reference's "5th pass",
//     ram_extend.c:1105-1150) with the true carries, together with the first-max column index.
//   * integer max-plus only: no tensor cores. Sentinels (-1000000000 / -987654321) and the
//     edge rule are the reference's, so every cell value is bit-identical.
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

// One pass over the lane's C cells for NC candidates (NC = 4: all bases; NC = 1: base a0).
//   SLOW   : per-cell bounds predicates (exact out-of-bounds / edge / dead-cell semantics)
//   COMMIT : write the new row into Sp/Gp and track the first cell whose score == bestv
//   BEST   : accumulate the row best per candidate
//   PDOUT  : leave P = max(Sp,Gp) and D (the two candidate-independent terms of the row) in Sp/Gp
//            in place of the previous row, which this pass has finished reading at that point
//   PDIN   : Sp/Gp already hold P and D (the commit after a PDOUT sweep recomputes neither)
// sin/gin  : S and G of the cell left of this lane's first cell (current row)
// gout     : G of this lane's last cell
template <int C, int G, int WS, bool SLOW, int NC, bool COMMIT, bool BEST, bool PDIN, bool PDOUT>
__device__ __forceinline__ void band_pass(int (&Sp)[C], int (&Gp)[C], const int SpN,
                                          const int GpN, const uint32_t *win, const char *lutb,
                                          const int a0, const int go, const int ge,
                                          const int (&sin)[NC], const int (&gin)[NC],
                                          int (&gout)[NC], int (&best)[NC], const int bestv,
                                          int &idx, const int gl, const int W, const int B,
                                          const int r, const int tmin, const int tmax)
{
  const int goe = go + ge;
  const bool is_last = (gl == G - 1);
  const int Wc = WS ? WS : W;
  const int kR = Wc - 1 - (G - 1) * C;   // last real cell of the last lane (static if WS != 0)
  const int edge = go + (r + 1) * ge;    // bnw_extend.c:992
  const int tbase = r - B + gl * C;      // band position t of cell 0
  int sprev[NC], gprev[NC];
#pragma unroll
  for (int a = 0; a < NC; ++a)
  {
    sprev[a] = sin[a];
    gprev[a] = gin[a];
  }
  static_assert(!(PDIN && PDOUT) && !(PDOUT && COMMIT), "P/D hand-over is sweep -> commit");
  if (COMMIT)
    idx = kBig;
  unsigned hits = 0;   // COMMIT, C <= 32: bit k set when cell k holds the row best

#pragma unroll
  for (int k = 0; k < C; ++k)
  {
    int P, D;
    if (PDIN)
    {
      P = Sp[k];
      D = Gp[k];
    }
    else
    {
      const int sN = (k + 1 < C) ? Sp[(k + 1 < C) ? k + 1 : 0] : SpN;
      const int gN = (k + 1 < C) ? Gp[(k + 1 < C) ? k + 1 : 0] : GpN;
      P = imax(Sp[k], Gp[k]);
      D = imax(sN + goe, gN + ge);
    }
    const int byte = win_byte(win[k >> 2], k & 3);
    bool real = true, inb = true;
    int E = 0;
    if (SLOW)
    {
      const int K = gl * C + k;
      real = K < Wc;
      if (K >= Wc - 1)
        D = kNeg;                          // offset == bandwidth: no deletion (bnw_extend.c:892)
      const int t = tbase + k;
      inb = real && (t >= tmin) && (t <= tmax);
      E = (K < B && r < B) ? edge : kOob;  // bnw_extend.c:990-1002
    }
    else
    {
      if (k == kR)
        D = is_last ? kNeg : D;
      if (k > kR)
        real = !is_last;
    }
    if (PDOUT)
    {
      Sp[k] = P;
      Gp[k] = D;
    }
    int mv[NC];
    if constexpr (NC == 4)
    {
      const int4 m4 = *reinterpret_cast<const int4 *>(lutb + byte);
      mv[0] = m4.x;
      mv[1] = m4.y;
      mv[2] = m4.z;
      mv[3] = m4.w;
    }
    else
    {
      mv[0] = *reinterpret_cast<const int *>(lutb + byte + 4 * a0);
    }
    int s0 = 0, g0 = 0;
#pragma unroll
    for (int a = 0; a < NC; ++a)
    {
      int S = P + mv[a];
      int Gv = imax3(D, sprev[a] + goe, gprev[a] + ge);
      if (SLOW)
      {
        S = inb ? S : E;
        Gv = inb ? Gv : E;
      }
      if (BEST)
      {
        const int nb = imax3(best[a], S, Gv);
        if (SLOW || k > kR)
          best[a] = real ? nb : best[a];
        else
          best[a] = nb;
      }
      sprev[a] = S;
      gprev[a] = Gv;
      if (a == 0)
      {
        s0 = S;
        g0 = Gv;
      }
    }
    if (COMMIT)
    {
      Sp[k] = s0;
      Gp[k] = g0;
      const bool hit = (imax(s0, g0) == bestv) && real;
      if constexpr (C <= 32)
      {
        if (hit)
          hits |= 1u << k;
      }
      else
        idx = min(idx, hit ? k : kBig);
    }
  }
  if constexpr (COMMIT && C <= 32)
    idx = hits ? (__ffs((int) hits) - 1) : kBig;
#pragma unroll
  for (int a = 0; a < NC; ++a)
    gout[a] = gprev[a];
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
  // fast path: non-positive gaps (the carry peaks at a lane's first cell) of modest size (the scan
  // adds kNegInf to travelling scores, see below)
  const bool gaps_ok = (go <= 0) && (ge <= 0) && (go >= -100000) && ((long long) ge * (2 * F->B + 1) >= -(1 << 20));
  const unsigned char *forced = F->forced;
  int *dbg_rows = F->dbg_rows;
  const bool dbg_any = (F->dbg_rows != nullptr) || (F->dbg_best != nullptr);
  unsigned char *dbg_path = F->dbg_path;

  // balanced partition of the family's instances over its CTAs (n_ctas <= n_ext)
  const int first = (int) ((long long) cta_in_fam * n_ext / n_ctas);
  const int next = (int) ((long long) (cta_in_fam + 1) * n_ext / n_ctas);
  const int grp = warp * IPW + gw;
  const bool valid = (gw < IPW) && (grp < next - first);
  const int inst = valid ? first + grp : next - 1;   // padding lanes shadow the last instance
  const InstDev I = F->inst[inst];
  const uint32_t *lib2 = F->lib2, *libn = F->libn;
  const int tmin = I.tmin, tmax = I.tmax;

  int Sp[C], Gp[C];
  uint32_t win[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i)
    win[i] = 0;
#pragma unroll
  for (int k = 0; k < C; ++k)
  {
    const int o = gl * C + k - B;
    const int v = (o == 0) ? 0 : abs(o) * ge + go;   // boundary row, ram_extend.c:909-946
    Sp[k] = v;
    Gp[k] = v;
    const int code = fetch_code(lib2, libn, I.start, I.sgn, I.comp, tmin, tmax, o);
    win[k >> 2] |= (uint32_t) (code << 4) << (8 * (k & 3));
  }

  const int nreal = min(max(W - gl * C, 0), C);   // real (in-band) cells of this lane
#ifndef RAMCU_EMUL
  const long long t_begin = F->dbg_wait ? clock64() : 0;
#endif
  int high = 0, pos = 0, thigh = 0, tpos = 0;   // ram_extend.c:900-914
  int max_row = -1, max_ext = 0, stop_row = -1;
  int r = 0;

  for (r = 0; r < L; ++r)
  {
    const int par = r & 1;
    // base entering the band at its high end for the next row
    int newc = 0;
    if (is_last)
      newc = fetch_code(lib2, libn, I.start, I.sgn, I.comp, tmin, tmax, r - B + G * C);

    const int tfirst = r - B + gl * C;
    const bool lane_ok = (tfirst >= tmin) && (tfirst + nreal - 1 <= tmax);
    const bool dead = (r - B > tmax) && (r >= B);   // every cell out of bounds, for good
    const bool alldead = (dbg_rows == nullptr) && (dbg_path == nullptr) && __all_sync(kFull, dead);
    const bool fast = (WS != 0) && gaps_ok && __all_sync(kFull, lane_ok);

    int best[4] = {kOob, kOob, kOob, kOob};
    int sin4[4] = {0, 0, 0, 0}, carry[4] = {0, 0, 0, 0};
    int SpN = 0, GpN = 0;
    int idx_dummy = 0;

    if (!alldead)
    {
      if (G > 1)
      {
        SpN = __shfl_down_sync(kFull, Sp[0], 1);
        GpN = __shfl_down_sync(kFull, Gp[0], 1);
      }
      // S of this lane's last cell for the four candidates -> left input of the next lane
      {
        const int byte = win_byte(win[(C - 1) >> 2], (C - 1) & 3);
        const int4 m4 = *reinterpret_cast<const int4 *>(lutb + byte);
        const int P = imax(Sp[C - 1], Gp[C - 1]);
        int sl[4] = {P + m4.x, P + m4.y, P + m4.z, P + m4.w};
        if (!fast)
        {
          const int K = gl * C + C - 1;
          const int t = tfirst + C - 1;
          const bool inb = (K < W) && (t >= tmin) && (t <= tmax);
          const int E = (K < B && r < B) ? (go + (r + 1) * ge) : kOob;
#pragma unroll
          for (int a = 0; a < 4; ++a)
            sl[a] = inb ? sl[a] : E;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
        {
          const int up = (G > 1) ? __shfl_up_sync(kFull, sl[a], 1) : 0;
          sin4[a] = (gl == 0) ? (kNeg - goe) : up;   // offset == -bandwidth: no insertion (972)
        }
      }
      int gin[4], agg[4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
      {
        gin[a] = (gl == 0) ? (kNeg - ge) : kNegInf;
        best[a] = kNeg;
      }
      if (fast)
      {
        band_pass<C, G, WS, false, 4, false, true, false, true>(Sp, Gp, SpN, GpN, win, lutb, 0, go, ge, sin4,
                                                   gin, agg, best, 0, idx_dummy, gl, W, B, r,
                                                   tmin, tmax);
        // lanes without a left neighbour at distance d add kNegInf instead of being predicated:
        // every value that travels is a real cell score here (>= kNeg - |small|), so the sum stays
        // above INT_MIN and below every real score
#pragma unroll
        for (int d = 1; d < G; d <<= 1)
        {
          const int pen = (gl >= d) ? d * C * ge : kNegInf;
#pragma unroll
          for (int a = 0; a < 4; ++a)
            agg[a] = imax(agg[a], __shfl_up_sync(kFull, agg[a], d) + pen);
        }
        {
          const int pen = (gl > 0) ? ge : kNegInf;
#pragma unroll
          for (int a = 0; a < 4; ++a)
          {
            carry[a] = (G > 1) ? __shfl_up_sync(kFull, agg[a], 1) : 0;
            best[a] = imax(best[a], carry[a] + pen);   // ge <= 0: the carry peaks at cell 0
          }
        }
      }
      else
      {
        int nobest[4] = {0, 0, 0, 0};
        band_pass<C, G, WS, true, 4, false, false, false, false>(Sp, Gp, SpN, GpN, win, lutb, 0, go, ge, sin4,
                                                   gin, agg, nobest, 0, idx_dummy, gl, W, B, r,
                                                   tmin, tmax);
        // (aggregate, transfer) scan: an out-of-bounds cell anywhere in a lane blocks the carry
        int T = (lane_ok && nreal == C) ? C * ge : kNegInf;
#pragma unroll
        for (int d = 1; d < G; d <<= 1)
        {
          const int Tu = __shfl_up_sync(kFull, T, d);
#pragma unroll
          for (int a = 0; a < 4; ++a)
          {
            const int up = __shfl_up_sync(kFull, agg[a], d);
            if (gl >= d)
              agg[a] = imax(agg[a], imax(up, kNegInf) + T);
          }
          if (gl >= d)
            T = imax(T + Tu, kNegInf);
        }
        int gin2[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
        {
          carry[a] = (G > 1) ? __shfl_up_sync(kFull, agg[a], 1) : 0;
          gin2[a] = (gl == 0) ? (kNeg - ge) : carry[a];
        }
        int agg2[4];
        band_pass<C, G, WS, true, 4, false, true, false, true>(Sp, Gp, SpN, GpN, win, lutb, 0, go, ge, sin4,
                                                  gin2, agg2, best, 0, idx_dummy, gl, W, B, r,
                                                  tmin, tmax);
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

    if (dbg_any && F->dbg_best != nullptr)
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
      int cs[1] = {sel4(besta, sin4[0], sin4[1], sin4[2], sin4[3])};
      int cg[1] = {(gl == 0) ? (kNeg - ge) : sel4(besta, carry[0], carry[1], carry[2], carry[3])};
      int gout1[1], nob[1] = {0};
      int idx = kBig;
      if (dbg_path != nullptr && valid)
      {
        // Gp holds the deletion term D of every cell until the commit overwrites it
#pragma unroll
        for (int k = 0; k < C; ++k)
          if (gl * C + k < W)
            F->dbg_del[(size_t) inst * W + gl * C + k] = Gp[k];
      }
      if (fast)
        band_pass<C, G, WS, false, 1, true, false, true, false>(Sp, Gp, SpN, GpN, win, lutb, besta, go, ge, cs,
                                                   cg, gout1, nob, bestb, idx, gl, W, B, r, tmin,
                                                   tmax);
      else
        band_pass<C, G, WS, true, 1, true, false, true, false>(Sp, Gp, SpN, GpN, win, lutb, besta, go, ge, cs,
                                                  cg, gout1, nob, bestb, idx, gl, W, B, r, tmin,
                                                  tmax);
      int kidx = (idx < kBig) ? gl * C + idx : kBig;
      kidx = group_min<G>(kidx, gl, lane);
      if (bestb > high)   // record setter, ram_extend.c:1140-1150
      {
        high = bestb;
        pos = r - B + kidx;
      }
      if (dbg_rows != nullptr && valid)
      {
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int K = gl * C + k;
          if (K < W)
          {
            int *o = dbg_rows + (((size_t) r * n_ext + inst) * W + K) * 2;
            o[0] = Sp[k];
            o[1] = Gp[k];
          }
        }
      }
      if (dbg_path != nullptr && valid)
      {
        // sub == cell -> substitution, else del == cell -> deletion, else insertion
        // (bnw_extend.c:1036-1041; each thread reads back the D it stored above)
#pragma unroll
        for (int k = 0; k < C; ++k)
        {
          const int K = gl * C + k;
          if (K < W)
          {
            const int cell = imax(Sp[k], Gp[k]);
            const int del = F->dbg_del[(size_t) inst * W + K];
            dbg_path[((size_t) r * n_ext + inst) * W + K] =
              (unsigned char) ((Sp[k] == cell) ? 0 : (del == cell) ? 1 : 2);
          }
        }
        if (gl == 0)
          F->dbg_idx[(size_t) r * n_ext + inst] = kidx;
      }
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
      uint32_t in = (uint32_t) (newc << 4);
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
