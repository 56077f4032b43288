#include "common.cuh"
namespace cmcd {
int launch_wide_fwd(const BridgeArgs& a, const cmcd_target* tg, int D, cudaStream_t st, int num_sms) { set_error("lgcp wide path: not built yet"); return 2; }
}
