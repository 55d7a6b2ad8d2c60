// kernels_b100.cu -- band tilings instantiated for bandwidth 100
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_CFG1(100, 41, 5), RAMCU_CFG(100, 34, 6), RAMCU_CFG(100, 29, 7), RAMCU_CFG(100, 26, 8), RAMCU_CFG(100, 13, 16), RAMCU_CFG_T(100, 26, 8, 640),
};

CfgList cfgs_b100() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
