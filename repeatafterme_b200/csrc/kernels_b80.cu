// kernels_b80.cu -- band tilings instantiated for bandwidth 80
#include "kernels.h"

namespace ramcu {

static KernelCfg g_list[] = {
  RAMCU_CFG1(80, 41, 4), RAMCU_CFG(80, 33, 5), RAMCU_CFG(80, 27, 6), RAMCU_CFG(80, 23, 7), RAMCU_CFG(80, 21, 8), RAMCU_CFG_T(80, 21, 8, 640),
};

CfgList cfgs_b80() { return CfgList{g_list, (int) (sizeof(g_list) / sizeof(g_list[0]))}; }

}  // namespace ramcu
