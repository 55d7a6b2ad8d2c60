// kernels.h -- registry of the band tilings (template instantiations of extend_kernel) that
// libramcu.so carries. Each kernels_*.cu instantiates the tilings of one bandwidth so that the
// translation units compile in parallel.
#pragma once
#include "extend_kernel.cuh"

namespace ramcu {

typedef void (*extend_fn)(const FamDev *, const int *, int *);

struct KernelCfg
{
  int B;    // bandwidth this instantiation is specialised for (-1 = any, runtime W)
  int C;    // cells per lane
  int G;    // lanes per instance
  // [0]: one pass per CTA; [1]: the right and the left pass of a family in the same threads (NULL
  // when the two row states do not fit the register file)
  int maxt[2];      // largest CTA size (threads); caps registers per thread at 65536 / maxt
  extend_fn fn[2];
  // single pass with the diagnostic / per-row outputs compiled in (forced consensus, debug rows, -outmat path
  // states, seed-extension rows, verbose dumps, RAMCU_STATS); same CTA limit as fn[0]
  extend_fn fnx;
};

// MAXT: wide lanes keep the whole 255-register budget (no spills); narrow lanes trade registers
// for more resident warps. Two passes per thread hold two row states.
// Lanes of 21..41 cells, one pass: 170 registers (12 resident warps per SM) rather than 255 (8). In a crowded
// batch that is worth 9 % (config-5 mix: 1616 against 1486 GCUPS); a family alone on the device loses 8 %
// (two single-pass launches of config 4: 25.4 against 23.4 ms) -- but a family alone runs both passes in
// the same threads (the two-pass variants keep 255 registers).
#ifndef RAMCU_MAXT1_WIDE
#define RAMCU_MAXT1_WIDE 384
#endif
constexpr int maxt_for(int C)
{
  return C <= 8 ? 768 : C <= 13 ? 512 : C <= 20 ? 384 : C <= 41 ? RAMCU_MAXT1_WIDE : 256;
}
#ifndef RAMCU_MAXT2_MID
#define RAMCU_MAXT2_MID 19
#endif
constexpr int maxt2_for(int C)
{
  return C <= 6 ? 768 : C <= 11 ? 512 : C <= RAMCU_MAXT2_MID ? 384 : 256;
}
#define RAMCU_CFG(B_, C_, G_)                                                                              \
  {B_, C_, G_, {maxt_for(C_), maxt2_for(C_)},                                                              \
   {extend_kernel<C_, G_, 2 * B_ + 1, maxt_for(C_), 1, false>, extend_kernel<C_, G_, 2 * B_ + 1, maxt2_for(C_), 2, false>}, \
   extend_kernel<C_, G_, 2 * B_ + 1, maxt_for(C_), 1, true>}
// wide lanes: one pass per CTA only
#define RAMCU_CFG1(B_, C_, G_) \
  {B_, C_, G_, {maxt_for(C_), 0}, {extend_kernel<C_, G_, 2 * B_ + 1, maxt_for(C_), 1, false>, nullptr}, \
   extend_kernel<C_, G_, 2 * B_ + 1, maxt_for(C_), 1, true>}
// same tiling with an explicit (larger) CTA limit: fewer registers per thread, more resident
// instances -- the capacity fallback for very large families with wide bands
#define RAMCU_CFG_T(B_, C_, G_, T_) \
  {B_, C_, G_, {T_, 0}, {extend_kernel<C_, G_, 2 * B_ + 1, T_, 1, false>, nullptr}, extend_kernel<C_, G_, 2 * B_ + 1, T_, 1, true>}
#define RAMCU_GEN(C_, G_)                                                                \
  {-1, C_, G_, {maxt_for(C_), maxt2_for(C_)},                                            \
   {extend_kernel<C_, G_, 0, maxt_for(C_), 1, false>, extend_kernel<C_, G_, 0, maxt2_for(C_), 2, false>}, \
   extend_kernel<C_, G_, 0, maxt_for(C_), 1, true>}
#define RAMCU_GEN1(C_, G_) \
  {-1, C_, G_, {maxt_for(C_), 0}, {extend_kernel<C_, G_, 0, maxt_for(C_), 1, false>, nullptr}, \
   extend_kernel<C_, G_, 0, maxt_for(C_), 1, true>}

struct CfgList
{
  KernelCfg *cfgs;
  int n;
};
CfgList cfgs_b14();
CfgList cfgs_b20();
CfgList cfgs_b40();
CfgList cfgs_b60();
CfgList cfgs_b80();
CfgList cfgs_b100();
CfgList cfgs_generic();

}  // namespace ramcu
