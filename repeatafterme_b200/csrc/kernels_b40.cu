// kernels_b40.cu -- band tilings instantiated for bandwidth 40
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_CFG1(40, 81, 1), RAMCU_CFG1(40, 41, 2), RAMCU_CFG(40, 27, 3), RAMCU_CFG(40, 21, 4), RAMCU_CFG(40, 17, 5), RAMCU_CFG(40, 14, 6), RAMCU_CFG(40, 11, 8),
};

CfgList cfgs_b40() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
