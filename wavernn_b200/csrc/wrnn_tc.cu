// wrnn_tc.cu -- tcgen05 engine (placeholder until the tensor-core kernel lands).
#include "wrnn_engine.h"
namespace wrnn {
int make_tc_engine(const wrnn_cfg&, const HostWeights&, int, Engine**) {
  set_error("tcgen05 engine not available in this build");
  return WRNN_E_INVALID;
}
}  // namespace wrnn
