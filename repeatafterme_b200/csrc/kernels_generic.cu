// kernels_generic.cu -- band tilings instantiated for any bandwidth (runtime band width)
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_GEN(4, 4), RAMCU_GEN(8, 4), RAMCU_GEN(16, 4), RAMCU_GEN(32, 4), RAMCU_GEN(16, 16), RAMCU_GEN1(32, 16), RAMCU_GEN(16, 32), RAMCU_GEN1(32, 32), RAMCU_GEN1(64, 32),
};

CfgList cfgs_generic() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
