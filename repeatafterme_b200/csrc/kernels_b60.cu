// kernels_b60.cu -- band tilings instantiated for bandwidth 60
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_CFG1(60, 41, 3), RAMCU_CFG(60, 31, 4), RAMCU_CFG(60, 25, 5), RAMCU_CFG(60, 21, 6), RAMCU_CFG(60, 16, 8), RAMCU_CFG(60, 8, 16), RAMCU_CFG_T(60, 16, 8, 640),
};

CfgList cfgs_b60() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
