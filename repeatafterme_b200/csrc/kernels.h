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
  int maxt; // largest CTA size (threads) supported; caps registers per thread at 65536 / maxt
  extend_fn fn;
  int regs; // filled at init
};

// MAXT: wide lanes keep the whole 255-register budget (no spills); narrow lanes trade registers
// for more resident warps.
constexpr int maxt_for(int C)
{
  return C <= 8 ? 768 : C <= 13 ? 512 : C <= 20 ? 384 : 256;
}
#define RAMCU_CFG(B_, C_, G_) {B_, C_, G_, maxt_for(C_), extend_kernel<C_, G_, 2 * B_ + 1, maxt_for(C_)>, 0}
// same tiling with an explicit (larger) CTA limit: fewer registers per thread, more resident
// instances -- the capacity fallback for very large families with wide bands
#define RAMCU_CFG_T(B_, C_, G_, T_) {B_, C_, G_, T_, extend_kernel<C_, G_, 2 * B_ + 1, T_>, 0}
#define RAMCU_GEN(C_, G_) {-1, C_, G_, maxt_for(C_), extend_kernel<C_, G_, 0, maxt_for(C_)>, 0}

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
