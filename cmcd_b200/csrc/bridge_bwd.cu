#include "net.cuh"
namespace cmcd {
size_t bridge_bwd_workspace_bytes(int D, int K, int HP, int arch, int num_sms) { return 16; }
int launch_bridge_bwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                      float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                      const cmcd_net_grad* g_net, void* ws, size_t ws_bytes) { set_error("bwd: not built yet"); return 2; }
}
