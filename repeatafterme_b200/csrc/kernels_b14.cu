// kernels_b14.cu -- band tilings instantiated for bandwidth 14
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_CFG(14, 29, 1), RAMCU_CFG(14, 15, 2), RAMCU_CFG(14, 10, 3), RAMCU_CFG(14, 8, 4), RAMCU_CFG(14, 4, 8),
};

CfgList cfgs_b14() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
