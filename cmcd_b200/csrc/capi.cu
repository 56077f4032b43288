// C ABI of libcmcd_b200.so (declared in include/cmcd_b200.h): argument validation and
// marshalling into the kernel launchers.  No torch types, no allocation, enqueue-only.
#include <atomic>
#include <cstdarg>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cmcd {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int launch_bridge_fwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms);
bool fwd_tc_supported(const BridgeArgs& a, int D);
int launch_bridge_fwd_tc(const BridgeArgs& a, int D, cudaStream_t st, int num_sms);
bool fwd_tcw_supported(const BridgeArgs& a, int D, int num_sms);
int launch_bridge_fwd_tcw(const BridgeArgs& a, int D, cudaStream_t st, int num_sms);
int launch_bridge_bwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                      float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                      const cmcd_net_grad* g_net, void* ws, size_t ws_bytes);
size_t bridge_bwd_workspace_bytes(int D, int K, int HP, int arch, int num_sms);
bool bwd_tc_supported(const BridgeArgs& a, int D);
size_t bridge_bwd_tc_workspace_bytes(int D, int K, int num_sms);
int launch_bridge_bwd_tc(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                         float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                         const cmcd_net_grad* g_net, void* ws, size_t ws_bytes);
bool blk_supported(const BridgeArgs& a, int D, int num_sms);
int launch_bridge_fwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms);
int launch_bridge_bwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                          float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                          const cmcd_net_grad* g_net, void* ws, size_t ws_bytes);
bool blk_ud_supported(const BridgeArgs& a, int D, int num_sms);
int launch_bridge_ud_fwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms);
int launch_bridge_ud_bwd_blk(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                             float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                             const cmcd_net_grad* g_net, void* ws, size_t ws_bytes);
int launch_bridge_ud_fwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms);
int launch_bridge_ud_bwd(const BridgeArgs& a, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                         float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                         const cmcd_net_grad* g_net, void* ws, size_t ws_bytes);
size_t bridge_ud_bwd_workspace_bytes(int mode, int D, int K, int HP, int arch, int num_sms);
int launch_bridge_uha_fwd(const BridgeArgs& a, int D, int lfsteps, cudaStream_t st, int num_sms);
int launch_bridge_uha_bwd(const BridgeArgs& a, int D, int lfsteps, cudaStream_t st, int num_sms, const float* cot_negw,
                          float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps, void* ws, size_t ws_bytes);
size_t bridge_uha_bwd_workspace_bytes(int D, int K, int num_sms);
int launch_loss_stats(cudaStream_t st, const float* negw, long long n, float* out4);
int launch_batched_elbo_lnz(cudaStream_t st, const float* losses, int batches, int n, float* elbo, float* lnz);
int launch_threefry(cudaStream_t st, const uint32_t* key2, const uint32_t* x0, const uint32_t* x1, long long n, uint32_t* y0, uint32_t* y1);
int launch_particle_noise(cudaStream_t st, const int32_t* seeds, long long n, int d, int K, float* xi0, float* xi);
int launch_target_eval(cudaStream_t st, const TargetDesc& t, int D, const float* x, long long n, const float* v,
                       float* lp, float* score, float* hvp);
int launch_ffma_peak(cudaStream_t st, float* scratch, int blocks, int iters);
int launch_wide_fwd(const BridgeArgs& a, const cmcd_target* tg, int D, cudaStream_t st, int num_sms, void* ws, size_t ws_bytes);
size_t wide_fwd_workspace_bytes(long long N, int d, int HP);
size_t wide_bwd_workspace_bytes(long long N, int d, int HP, int K);
int launch_wide_bwd(const BridgeArgs& a, const cmcd_target* tg, int D, cudaStream_t st, int num_sms, const float* cot_negw,
                    float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps, const cmcd_net_grad* g,
                    void* ws, size_t ws_bytes);

int launch_adam_project(cudaStream_t st, float* p, const float* g, float* m, float* v, const float* lo, const float* hi, long long n,
                        float lr, float b1, float b2, float eps, float clip, int step, float* ema, float ema_step,
                        const int32_t* skip_flag);
int launch_chain_fwd(const cmcd_chain* c, cudaStream_t st, const float* params_flat, float* betas, float* eps, float* c1, float* c2, float* c3,
                     float* U1p, float* U2p, float* W2p, float* W3p);
size_t chain_bwd_scratch_floats(const cmcd_chain* c);
int launch_chain_bwd(const cmcd_chain* c, cudaStream_t st, const float* params_flat, const float* g_betas, const float* g_eps,
                     const float* g_vd_mean, const float* g_vd_logdiag, const cmcd_net_grad* gn, float* scratch, size_t scratch_floats,
                     float* grad_flat);
int launch_randint(cudaStream_t st, uint32_t key0, uint32_t key1, long long n, int32_t minval, int32_t maxval, int32_t* out);

// SM count of the CURRENT device, cached per device id (one process may drive several GPUs; relaxed atomics: the value is
// idempotent, so a race only repeats the query)
static int num_sms() {
    constexpr int MAX_DEV = 64;
    static std::atomic<int> cache[MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < MAX_DEV) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < MAX_DEV) cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

static bool is_ud(int mode) { return (mode >= CMCD_MODE_UD_NONE && mode <= CMCD_MODE_UD_NET_ZRHO) || mode == CMCD_MODE_UD_CAIS; }
static bool mode_uses_net(int mode) { return mode != CMCD_MODE_ULA && mode != CMCD_MODE_UD_NONE && mode != CMCD_MODE_UHA; }

static int build_args(const cmcd_bridge_desc* d, const int32_t* seeds, const float* vd_mean, const float* vd_logdiag,
                      const float* betas, const float* eps, const cmcd_net* net, const cmcd_target* tg, BridgeArgs& a) {
    if (!d || !tg) { set_error("null descriptor"); return 2; }
    if (d->mode < CMCD_MODE_ULA || d->mode > CMCD_MODE_UD_CAIS) { set_error("Mode not implemented."); return 2; }
    if (d->nbridges < 0 || d->n_particles < 0 || d->dim < 1) { set_error("bad sizes N=%d K=%d d=%d", d->n_particles, d->nbridges, d->dim); return 2; }
    std::memset(&a, 0, sizeof(a));
    a.mode = d->mode; a.K = d->nbridges; a.N = d->n_particles;
    a.clip_t = d->clip_target; a.clip_q = d->clip_q;
    a.seeds = seeds; a.vd_mean = vd_mean; a.vd_logdiag = vd_logdiag; a.betas = betas; a.eps = eps;
    const bool needs_net = mode_uses_net(d->mode) && d->nbridges >= 1;
    if (needs_net && (!net || net->arch == CMCD_ARCH_NONE)) { set_error("mode %d needs a drift network", d->mode); return 2; }
    if (net && net->arch != CMCD_ARCH_NONE && needs_net) {
        if (net->arch != CMCD_ARCH_GEFFNER && net->arch != CMCD_ARCH_DDS) { set_error("nn_arch %d not implemented", net->arch); return 2; }
        if (net->hidden_pad % 8 || net->hidden_pad < net->hidden || net->hidden < 1) { set_error("hidden_pad must be a multiple of 8 and >= hidden"); return 2; }
        if (net->n_rows < d->nbridges + 1) { set_error("net tables need nbridges+1 rows"); return 2; }
        a.net.arch = net->arch; a.net.H = net->hidden; a.net.HP = net->hidden_pad; a.net.T = net->n_rows;
        a.net.U1 = net->U1; a.net.U2 = net->U2; a.net.U3 = net->U3; a.net.W2 = net->W2; a.net.W3 = net->W3;
        a.net.c1 = net->c1; a.net.c2 = net->c2; a.net.c3 = net->c3;
        a.net.out_scale = net->out_scale; a.net.out_clip = net->out_clip; a.net.out_scale_dev = net->out_scale_dev;
    } else {
        a.net.arch = CMCD_ARCH_NONE;
    }
    a.tgt.kind = tg->kind; a.tgt.ncomp = tg->ncomp; a.tgt.mix = tg->mix;
    switch (tg->kind) {
        case CMCD_TARGET_GMM:
        case CMCD_TARGET_MANY_GMM:
            if (d->dim != 2) { set_error("mixture targets are 2-D (model_handler.py:245-249), got dim=%d", d->dim); return 2; }
            if (tg->ncomp < 1 || tg->ncomp > MIX_MAX || !tg->mix) { set_error("mixture needs 1..%d components", MIX_MAX); return 2; }
            a.tgt.scale = tg->scale; a.tgt.inv_var = 1.0f / (tg->scale * tg->scale);
            a.tgt.comp_norm = 0.9189385332046727f + logf(tg->scale);
            a.tgt.log_mix = -logf((float)tg->ncomp);
            a.tgt.invalid_below = tg->invalid_below;
            break;
        case CMCD_TARGET_FUNNEL:
            if (d->dim < 2) { set_error("funnel needs dim >= 2"); return 2; }
            break;
        case CMCD_TARGET_LGCP:
            if (!tg->lgcp_kinv || !tg->lgcp_counts) { set_error("lgcp needs kinv and counts"); return 2; }
            if (is_ud(d->mode) || d->mode == CMCD_MODE_UHA) { set_error("the underdamped modes have no wide (lgcp) path"); return 2; }
            break;
        case CMCD_TARGET_CALLBACK:
            if (!tg->eval) { set_error("callback target needs an eval function"); return 2; }
            if (d->mode > CMCD_MODE_CAIS_VAR_SN) { set_error("callback targets are served for the overdamped modes (0..3)"); return 2; }
            break;
        default: set_error("target kind %d not in the registry", tg->kind); return 2;
    }
    return 0;
}

}  // namespace cmcd

using namespace cmcd;

extern "C" {

const char* cmcd_last_error(void) { return g_err; }
int cmcd_version(void) { return 100; }
int cmcd_num_sms(void) { return num_sms(); }

size_t cmcd_bridge_fwd_workspace_bytes(const cmcd_bridge_desc* desc, const cmcd_net* net, const cmcd_target* target) {
    if (!desc || !target || (target->kind != CMCD_TARGET_LGCP && target->kind != CMCD_TARGET_CALLBACK)) return 0;
    const bool uses_net = net && net->arch != CMCD_ARCH_NONE && mode_uses_net(desc->mode) && desc->nbridges >= 1;
    return wide_fwd_workspace_bytes(desc->n_particles, desc->dim, uses_net ? net->hidden_pad : 0);
}

int cmcd_bridge_fwd(const cmcd_bridge_desc* desc, void* stream, const int32_t* seeds, const float* vd_mean,
                    const float* vd_logdiag, const float* betas, const float* eps, const cmcd_net* net,
                    const cmcd_target* target, float* out_negw, float* out_z, float* traj,
                    void* workspace, size_t workspace_bytes) {
    BridgeArgs a;
    if (int rc = build_args(desc, seeds, vd_mean, vd_logdiag, betas, eps, net, target, a)) return rc;
    a.out_negw = out_negw; a.out_z = out_z; a.traj = traj;
    if (a.N == 0) return 0;
    const int sms = num_sms();
    if (sms <= 0) { set_error("no CUDA device"); return 1; }
    if (target->kind == CMCD_TARGET_LGCP || target->kind == CMCD_TARGET_CALLBACK)
        return launch_wide_fwd(a, target, desc->dim, (cudaStream_t)stream, sms, workspace, workspace_bytes);
    if (is_ud(desc->mode)) {
        if (blk_ud_supported(a, desc->dim, sms) && !std::getenv("CMCD_DISABLE_BLK"))
            return launch_bridge_ud_fwd_blk(a, desc->dim, (cudaStream_t)stream, sms);
        return launch_bridge_ud_fwd(a, desc->dim, (cudaStream_t)stream, sms);
    }
    if (desc->mode == CMCD_MODE_UHA)
        return launch_bridge_uha_fwd(a, desc->dim, desc->lfsteps > 0 ? desc->lfsteps : 1, (cudaStream_t)stream, sms);
    // hidden width 64 and 129..144: tcgen05 tiles; other widths: FP32 FMA kernels.  CMCD_DISABLE_TC=1 forces the FP32 kernels (A/B runs).
    if (fwd_tcw_supported(a, desc->dim, sms) && !std::getenv("CMCD_DISABLE_TC"))
        return launch_bridge_fwd_tcw(a, desc->dim, (cudaStream_t)stream, sms);
    if (fwd_tc_supported(a, desc->dim) && !std::getenv("CMCD_DISABLE_TC"))
        return launch_bridge_fwd_tc(a, desc->dim, (cudaStream_t)stream, sms);
    // few particles x wide network: block-cooperative mapping (bridge_blk.cu).  CMCD_DISABLE_BLK=1 keeps one thread per particle.
    if (blk_supported(a, desc->dim, sms) && !std::getenv("CMCD_DISABLE_BLK"))
        return launch_bridge_fwd_blk(a, desc->dim, (cudaStream_t)stream, sms);
    return launch_bridge_fwd(a, desc->dim, (cudaStream_t)stream, sms);
}

size_t cmcd_bridge_bwd_workspace_bytes(const cmcd_bridge_desc* desc, const cmcd_net* net) {
    const int sms = num_sms() > 0 ? num_sms() : 148;
    const int arch = (net && mode_uses_net(desc->mode)) ? net->arch : CMCD_ARCH_NONE;
    if (desc->mode == CMCD_MODE_UHA) return bridge_uha_bwd_workspace_bytes(desc->dim, desc->nbridges, sms);
    if (is_ud(desc->mode))
        return bridge_ud_bwd_workspace_bytes(desc->mode, desc->dim, desc->nbridges, net ? net->hidden_pad : 0, arch, sms);
    if (desc->dim > 64)   // wide path (lgcp, d = 1600): [N x d] / [N x hidden] state + split-K partials
        return wide_bwd_workspace_bytes(desc->n_particles, desc->dim, arch != CMCD_ARCH_NONE ? net->hidden_pad : 0, desc->nbridges);
    size_t need = bridge_bwd_workspace_bytes(desc->dim, desc->nbridges, net ? net->hidden_pad : 0, arch, sms);
    if (arch == CMCD_ARCH_DDS && net->hidden_pad == 64 && desc->dim == 2) {
        const size_t tc = bridge_bwd_tc_workspace_bytes(desc->dim, desc->nbridges, sms);
        if (tc > need) need = tc;
    }
    return need;
}

size_t cmcd_bridge_bwd_workspace_bytes_for_target(const cmcd_bridge_desc* desc, const cmcd_net* net, const cmcd_target* target) {
    if (desc && target && target->kind == CMCD_TARGET_CALLBACK) {   // step-wise path at any dim
        const int arch = (net && mode_uses_net(desc->mode)) ? net->arch : CMCD_ARCH_NONE;
        return wide_bwd_workspace_bytes(desc->n_particles, desc->dim, arch != CMCD_ARCH_NONE ? net->hidden_pad : 0, desc->nbridges);
    }
    return cmcd_bridge_bwd_workspace_bytes(desc, net);
}

int cmcd_bridge_bwd(const cmcd_bridge_desc* desc, void* stream, const int32_t* seeds, const float* vd_mean,
                    const float* vd_logdiag, const float* betas, const float* eps, const cmcd_net* net,
                    const cmcd_target* target, const float* traj, const float* cot_negw, float* g_vd_mean,
                    float* g_vd_logdiag, float* g_betas, float* g_eps, const cmcd_net_grad* g_net,
                    void* workspace, size_t workspace_bytes) {
    BridgeArgs a;
    if (int rc = build_args(desc, seeds, vd_mean, vd_logdiag, betas, eps, net, target, a)) return rc;
    a.traj = const_cast<float*>(traj);
    const int sms = num_sms();
    if (sms <= 0) { set_error("no CUDA device"); return 1; }
    if (!traj || !cot_negw) { set_error("bridge_bwd needs traj and cot_negw"); return 2; }
    if (a.N == 0) { set_error("bridge_bwd: empty particle batch"); return 2; }
    if (target->kind == CMCD_TARGET_LGCP || target->kind == CMCD_TARGET_CALLBACK)
        return launch_wide_bwd(a, target, desc->dim, (cudaStream_t)stream, sms, cot_negw, g_vd_mean, g_vd_logdiag, g_betas,
                               g_eps, g_net, workspace, workspace_bytes);
    if (desc->mode == CMCD_MODE_UHA)
        return launch_bridge_uha_bwd(a, desc->dim, desc->lfsteps > 0 ? desc->lfsteps : 1, (cudaStream_t)stream, sms, cot_negw,
                                     g_vd_mean, g_vd_logdiag, g_betas, g_eps, workspace, workspace_bytes);
    if (is_ud(desc->mode) && blk_ud_supported(a, desc->dim, sms) && !std::getenv("CMCD_DISABLE_BLK"))
        return launch_bridge_ud_bwd_blk(a, desc->dim, (cudaStream_t)stream, sms, cot_negw, g_vd_mean, g_vd_logdiag, g_betas,
                                        g_eps, g_net, workspace, workspace_bytes);
    if (is_ud(desc->mode))
        return launch_bridge_ud_bwd(a, desc->dim, (cudaStream_t)stream, sms, cot_negw, g_vd_mean, g_vd_logdiag, g_betas,
                                    g_eps, g_net, workspace, workspace_bytes);
    if (bwd_tc_supported(a, desc->dim) && !std::getenv("CMCD_DISABLE_TC"))
        return launch_bridge_bwd_tc(a, desc->dim, (cudaStream_t)stream, sms, cot_negw, g_vd_mean, g_vd_logdiag, g_betas,
                                    g_eps, g_net, workspace, workspace_bytes);
    if (blk_supported(a, desc->dim, sms) && !std::getenv("CMCD_DISABLE_BLK"))
        return launch_bridge_bwd_blk(a, desc->dim, (cudaStream_t)stream, sms, cot_negw, g_vd_mean, g_vd_logdiag, g_betas,
                                     g_eps, g_net, workspace, workspace_bytes);
    return launch_bridge_bwd(a, desc->dim, (cudaStream_t)stream, sms, cot_negw, g_vd_mean, g_vd_logdiag, g_betas,
                             g_eps, g_net, workspace, workspace_bytes);
}

int cmcd_loss_stats(void* stream, const float* negw, int64_t n, float* out4) {
    return launch_loss_stats((cudaStream_t)stream, negw, n, out4);
}
int cmcd_batched_elbo_lnz(void* stream, const float* losses, int32_t batches, int32_t n, float* elbo, float* lnz) {
    if (batches < 1 || n < 1) { set_error("empty batch"); return 2; }
    return launch_batched_elbo_lnz((cudaStream_t)stream, losses, batches, n, elbo, lnz);
}

int cmcd_chain_fwd(const cmcd_chain* chain, void* stream, const float* params_flat, float* betas, float* eps, float* c1, float* c2,
                   float* c3, float* U1_pad, float* U2_pad, float* W2_pad, float* W3_pad) {
    return launch_chain_fwd(chain, (cudaStream_t)stream, params_flat, betas, eps, c1, c2, c3, U1_pad, U2_pad, W2_pad, W3_pad);
}
size_t cmcd_chain_bwd_scratch_floats(const cmcd_chain* chain) { return chain_bwd_scratch_floats(chain); }
int cmcd_chain_bwd(const cmcd_chain* chain, void* stream, const float* params_flat, const float* g_betas, const float* g_eps,
                   const float* g_vd_mean, const float* g_vd_logdiag, const cmcd_net_grad* g_net, float* scratch,
                   size_t scratch_floats, float* grad_flat) {
    return launch_chain_bwd(chain, (cudaStream_t)stream, params_flat, g_betas, g_eps, g_vd_mean, g_vd_logdiag, g_net, scratch,
                            scratch_floats, grad_flat);
}

int cmcd_bridge_evolve(const cmcd_bridge_desc* desc, void* stream, const float* z0, const uint32_t* keys, const float* vd_mean,
                       const float* vd_logdiag, const float* betas, const float* eps, const cmcd_net* net,
                       const cmcd_target* target, float* out_z, float* out_w) {
    BridgeArgs a;
    if (int rc = build_args(desc, nullptr, vd_mean, vd_logdiag, betas, eps, net, target, a)) return rc;
    if (!z0 || !keys || !out_z || !out_w) { set_error("bridge_evolve needs z0, keys, out_z and out_w"); return 2; }
    if (desc->mode > CMCD_MODE_CAIS_VAR_SN) { set_error("Mode not implemented. (bridge_evolve serves the overdamped modes 0..3)"); return 2; }
    if (target->kind == CMCD_TARGET_LGCP || target->kind == CMCD_TARGET_CALLBACK) { set_error("bridge_evolve: target kind %d not in the registry of the fused small-d kernels", target->kind); return 2; }
    if (desc->nbridges < 1) { set_error("bridge_evolve needs nbridges >= 1"); return 2; }
    a.z0 = z0; a.keys = keys; a.out_z = out_z; a.out_negw = out_w; a.traj = nullptr;
    if (a.N == 0) return 0;
    const int sms = num_sms();
    if (sms <= 0) { set_error("no CUDA device"); return 1; }
    // same kernels as cmcd_bridge_fwd (the block-cooperative mapping has no evolve entry: one thread per particle instead)
    if (fwd_tcw_supported(a, desc->dim, sms) && !std::getenv("CMCD_DISABLE_TC"))
        return launch_bridge_fwd_tcw(a, desc->dim, (cudaStream_t)stream, sms);
    if (fwd_tc_supported(a, desc->dim) && !std::getenv("CMCD_DISABLE_TC"))
        return launch_bridge_fwd_tc(a, desc->dim, (cudaStream_t)stream, sms);
    return launch_bridge_fwd(a, desc->dim, (cudaStream_t)stream, sms);
}

int cmcd_bridge_fwd_host(const cmcd_bridge_desc* desc, void* stream, const int32_t* seeds_host, const float* vd_mean,
                         const float* vd_logdiag, const float* betas, const float* eps, const cmcd_net* net,
                         const cmcd_target* target, int32_t* seeds_dev, float* negw_dev, float* z_dev,
                         float* out_negw_host, float* out_z_host) {
    if (!desc || !target) { set_error("null descriptor"); return 2; }
    if (!seeds_host || !seeds_dev || !negw_dev || !z_dev || !out_negw_host) { set_error("bridge_fwd_host: null buffer"); return 2; }
    if (target->kind == CMCD_TARGET_LGCP || target->kind == CMCD_TARGET_CALLBACK) { set_error("bridge_fwd_host: lgcp / callback targets need the workspace entry point"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = desc->n_particles;
    CMCD_CUDA_OK(cudaMemcpyAsync(seeds_dev, seeds_host, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (int rc = cmcd_bridge_fwd(desc, stream, seeds_dev, vd_mean, vd_logdiag, betas, eps, net, target, negw_dev, z_dev, nullptr, nullptr, 0)) return rc;
    CMCD_CUDA_OK(cudaMemcpyAsync(out_negw_host, negw_dev, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out_z_host) CMCD_CUDA_OK(cudaMemcpyAsync(out_z_host, z_dev, n * desc->dim * sizeof(float), cudaMemcpyDeviceToHost, st));
    CMCD_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int cmcd_target_eval(const cmcd_target* target, int32_t dim, void* stream, const float* x, int64_t n, const float* v,
                     float* out_logp, float* out_score, float* out_hvp) {
    cmcd_bridge_desc d;
    d.mode = CMCD_MODE_ULA; d.dim = dim; d.nbridges = 0; d.n_particles = (int32_t)n;
    d.clip_target = d.clip_q = INFINITY;
    d.lfsteps = 0;
    BridgeArgs a;
    if (int rc = build_args(&d, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, target, a)) return rc;
    if (target->kind == CMCD_TARGET_LGCP || target->kind == CMCD_TARGET_CALLBACK) { set_error("target_eval: lgcp / callback targets are served by the wide path"); return 2; }
    if (n == 0) return 0;
    return launch_target_eval((cudaStream_t)stream, a.tgt, dim, x, n, v, out_logp, out_score, out_hvp);
}

int cmcd_adam_project_step(void* stream, float* params, const float* grad, float* m, float* v, const float* lo, const float* hi,
                           int64_t n, float lr, float b1, float b2, float eps, float clip, int32_t step, float* ema,
                           float ema_step, const int32_t* skip_flag) {
    if (!params || !grad || !m || !v || n < 0 || step < 1) { set_error("adam_project_step: bad arguments"); return 2; }
    if (n == 0) return 0;
    return launch_adam_project((cudaStream_t)stream, params, grad, m, v, lo, hi, n, lr, b1, b2, eps, clip, step, ema, ema_step, skip_flag);
}
int cmcd_randint(void* stream, uint32_t key0, uint32_t key1, int64_t n, int32_t minval, int32_t maxval, int32_t* out) {
    if (!out || n < 0) { set_error("randint: bad arguments"); return 2; }
    if (n == 0) return 0;
    return launch_randint((cudaStream_t)stream, key0, key1, n, minval, maxval, out);
}

int cmcd_ffma_probe(void* stream, float* scratch, int32_t blocks, int32_t iters) {
    if (blocks < 1 || iters < 1 || !scratch) { set_error("ffma_probe: bad arguments"); return 2; }
    return launch_ffma_peak((cudaStream_t)stream, scratch, blocks, iters);
}

int cmcd_threefry2x32(void* stream, const uint32_t* key2, const uint32_t* x0, const uint32_t* x1, int64_t n, uint32_t* y0, uint32_t* y1) {
    return launch_threefry((cudaStream_t)stream, key2, x0, x1, n, y0, y1);
}
int cmcd_particle_noise(void* stream, const int32_t* seeds, int64_t n, int32_t dim, int32_t nbridges, float* xi0, float* xi) {
    return launch_particle_noise((cudaStream_t)stream, seeds, n, dim, nbridges, xi0, xi);
}

}  // extern "C"
