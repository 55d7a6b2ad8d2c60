// kernels_b20.cu -- band tilings instantiated for bandwidth 20
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_CFG1(20, 41, 1), RAMCU_CFG(20, 21, 2), RAMCU_CFG(20, 14, 3), RAMCU_CFG(20, 11, 4), RAMCU_CFG(20, 7, 6),
};

CfgList cfgs_b20() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
