/*
 * cmcd_b200 -- C ABI of the B200-native CMCD bridge hot path (libcmcd_b200.so).
 *
 * Drop-in boundary for the data-parallel hot path of shreyaspadhy/CMCD: the per-particle
 * annealed-Langevin bridge loop behind the boundmodes MCD_ULA, MCD_ULA_sn, MCD_CAIS_sn,
 * MCD_CAIS_var_sn -- and, behind the same entry points, the momentum-augmented operators of the
 * dispatcher (LDVI family, 2nd-order CMCD, UHA) and targets given as a score callback.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference's src/).  All pointers are DEVICE pointers unless the
 * name ends in _host; arrays are dense row-major float32 / int32; the callee only enqueues
 * work on `stream` (no allocation, no synchronisation) so the calls are CUDA-graph safe and
 * usable from an XLA-FFI handler.  Return value: 0 = OK, non-zero =
 * error, message via cmcd_last_error().  Unsupported (mode, target, arch, dim) combinations
 * fail loudly -- there is no CPU fallback.
 */
#ifndef CMCD_B200_H_
#define CMCD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* boundmode -- the `mode` string of mcd_utils.evolve (mcd_utils.py:34-190).
 * 4..6: the underdamped (momentum-augmented) operators evolve_underdamped_lp_a / _lp_e / _lp_ea
 *   (mcd_under_lp_a.py:6-87 "LDVI", mcd_under_lp_e.py:6-74, mcd_under_lp_ea.py:6-104; dispatch mcd_utils.py:59-133).
 *   The three scan bodies are one step with different coefficients, so the kernel mode only says what the score
 *   network sees: nothing (MCD_U_a-lp, MCD_U_e-lp), z (MCD_U_a-lp-sna, MCD_U_e-lp-sna) or (z, rho') (MCD_U_a-lp-sn,
 *   MCD_U_ea-lp-sn; rho_dim = dim, mcdboundingmachine.py:84-102).  For these modes:
 *     eps / g_eps = [7][K] = rows (eps, a_f, s_f, a_b, c_n, s_b, c_f): forward-kernel mean a_f rho (+ c_f NN, mode 8)
 *                   and scale s_f, backward-kernel mean a_b rho' + c_n NN and scale s_b (formulas per operator in
 *                   csrc/bridge_ud.cu) -- the host
 *                   forms the rows from (eps, gamma, eta) and chains the cotangents back;
 *     traj        = [K+1][3 dim][N] = (z_j, rho_j, rho'_j) per node;
 *     cmcd_net    U1 / U2 = [in][HP], U3 = [in][dim] with in = dim (NET_Z) or 2 dim (NET_ZRHO);
 *     clip_target / clip_q are ignored in modes 4..6 (these operators take no grad_clipping).
 * 8: MCD_CAIS_UHA_sn ("2nd order CMCD", README.md:16; evolve_underdamped_lp_a_cais, mcd_under_lp_a_cais.py:6-115): as mode 6
 *   with the network also in the forward-kernel mean (row c_f), eps_i on the cosine schedule and the target score clipped at
 *   clip_target (1e2 in the reference body).  The reference dispatcher cannot call it at HEAD (keyword mismatch,
 *   mcd_utils.py:176-188); built to the function body as written.
 * 7: UHA -- boundmode "UHA" (main.py:115-133): boundingmachine.compute_log_elbo (boundingmachine.py:73-111) over
 *   ais_utils.evolve (ais_utils.py:7-69) with the diagonal momentum distribution of momdist.py; no score network.
 *     eps / g_eps = [3][K] = rows (eps, a, s): momentum refresh rho_r = a rho + s exp(md) xi (a = eta, s = sqrt(1 - eta^2));
 *     vd_logdiag / g_vd_logdiag = [2][dim] = (q log-scales, md = momentum log-scales, momdist.py:9-11);
 *     desc.lfsteps = leapfrog steps per bridge (1..8, configs/base.py:83); traj = [K+1][3 dim][N] = (z_j, rho_j, rho_r_j). */
enum { CMCD_MODE_ULA = 0, CMCD_MODE_ULA_SN = 1, CMCD_MODE_CAIS_SN = 2, CMCD_MODE_CAIS_VAR_SN = 3,
       CMCD_MODE_UD_NONE = 4, CMCD_MODE_UD_NET_Z = 5, CMCD_MODE_UD_NET_ZRHO = 6, CMCD_MODE_UHA = 7, CMCD_MODE_UD_CAIS = 8 };
/* target registry -- model_handler.load_model (model_handler.py:30-43) */
enum { CMCD_TARGET_GMM = 0, CMCD_TARGET_MANY_GMM = 1, CMCD_TARGET_FUNNEL = 2, CMCD_TARGET_LGCP = 3,
       CMCD_TARGET_CALLBACK = 4 };

/*
 * Generic target (CMCD_TARGET_CALLBACK): a batched score callback for densities outside the registry -- what the
 * reference gets from jax.grad of an arbitrary log_prob_model (inference-gym and numpyro models, model_handler.py:46-86).
 * Called between the half-steps of every bridge step (once per trajectory point; in the reverse pass once more with v for
 * the Hessian-vector product).  It must ENQUEUE its work on `stream` and return 0:
 *   x[n][dim] -> out_logp[n] (if non-NULL), out_score[n][dim] = grad log p (if non-NULL),
 *   out_hvp[n][dim] = Hessian(log p)(x) v (if v and out_hvp non-NULL).  All pointers are device pointers.
 * Served by the step-wise ("wide") path: any dim, modes 0..3, nn_arch none / geffner.
 */
typedef int (*cmcd_target_fn)(void* user, void* stream, const float* x, int64_t n, int32_t dim, const float* v,
                              float* out_logp, float* out_score, float* out_hvp);
/* drift network -- nn.initialize_network (nn.py:21-39) */
enum { CMCD_ARCH_NONE = 0, CMCD_ARCH_GEFFNER = 1, CMCD_ARCH_DDS = 2 };

#define CMCD_MIX_STRIDE 6

/*
 * Drift network in "per-step table" form.  Both reference networks evaluate, per particle x
 * and step index t (apply_fun_sn(params["sn"], x, t); nn.py:66-70, nn_dds.py:145-164):
 *     a1  = act(U1^T x + c1[t])                       act = softplus (geffner) | exact-erf gelu (dds)
 *     a2  = act(W2^T a1 + U2^T x + c2[t])
 *     o   = W3^T (a2 + skip*a1) + U3^T x + c3[t]       skip = 1 (geffner) | 0 (dds)
 *     out = out_scale * clamp(o, -out_clip, out_clip)
 * where everything that depends only on the step (embedding row / time-coder output pushed
 * through the first-layer weights, biases) is folded into the tables c1,c2,c3 by the host
 * wrapper (O(K) work, cmcd_b200/nn.py).  Leading dimension of every [.,hidden] array is
 * hidden_pad (multiple of 8, zero padded).
 */
typedef struct cmcd_net {
    int32_t arch;        /* CMCD_ARCH_* */
    int32_t hidden;      /* H */
    int32_t hidden_pad;  /* HP >= H, multiple of 8 */
    int32_t n_rows;      /* rows of c1/c2/c3 (nbridges + 1) */
    const float* U1;     /* [dim][HP] */
    const float* U2;     /* [dim][HP] or NULL (treated as zero) */
    const float* U3;     /* [dim][dim] or NULL */
    const float* W2;     /* [HP][HP]  (input-major: W2[i][j] multiplies a1[i] into unit j) */
    const float* W3;     /* [HP][dim] */
    const float* c1;     /* [n_rows][HP] */
    const float* c2;     /* [n_rows][HP] */
    const float* c3;     /* [n_rows][dim] */
    float out_scale;     /* factor_sn (geffner) | 1 (dds) */
    float out_clip;      /* +inf (geffner) | 1e4 (dds) */
    const float* out_scale_dev; /* optional device scalar: if non-NULL the kernels read out_scale from here instead (keeps the
                                   trainable factor_sn, nn.py:63, on the device: no host read-back per iteration, CUDA-graph safe) */
} cmcd_net;

/* Cotangents of every differentiable cmcd_net field (same shapes); NULL members are skipped.
 * arch = CMCD_ARCH_DDS: PISNet's c2 / c3 tables repeat one bias row for every step (nn_dds.py:159-164: st2.b, LinearZero bias),
 * so only the SUM over the rows of their cotangents is meaningful; the tensor-core adjoint reports that sum in row 0 and zeros
 * elsewhere (the other kernels report per-row values whose sum is the same). */
typedef struct cmcd_net_grad {
    float* U1; float* U2; float* U3; float* W2; float* W3; float* c1; float* c2; float* c3;
    float* out_scale;    /* [1] */
} cmcd_net_grad;

/* Target density descriptor (closed registry; model_handler.py:124-409). */
typedef struct cmcd_target {
    int32_t kind;        /* CMCD_TARGET_* */
    int32_t ncomp;       /* mixtures: number of components (<= 64) */
    float scale;         /* many_gmm: shared component scale (softplus(0.1), model_handler.py:262-267) */
    float invalid_below; /* many_gmm: log p <= this -> -inf with zero gradient (-1e4, :279-280) */
    const float* mix;    /* [ncomp][CMCD_MIX_STRIDE]: many_gmm (mu0,mu1,-,-,-,-); gmm (m0,m1,p00,p01,p11,logc) */
    const float* lgcp_kinv;   /* lgcp: [dim][dim] dense K^-1 (symmetric) */
    const float* lgcp_linv;   /* lgcp: reserved (the kernels use the K^-1 form), may be NULL */
    const float* lgcp_counts; /* lgcp: [dim] */
    float lgcp_mu0;           /* lgcp: constant prior mean */
    float lgcp_log_norm;      /* lgcp: -d/2 log(2 pi) - sum log diag L */
    float lgcp_bin_area;      /* lgcp: 1/dim */
    cmcd_target_fn eval;      /* CMCD_TARGET_CALLBACK: the batched score callback */
    void* user;               /* CMCD_TARGET_CALLBACK: passed back to eval */
} cmcd_target;

/* Static description of one bridge problem (params_fixed + the flags of main.py:162-172). */
typedef struct cmcd_bridge_desc {
    int32_t mode;        /* CMCD_MODE_* */
    int32_t dim;         /* d */
    int32_t nbridges;    /* K >= 0 (K = 0: MFVI bound, boundingmachine.py:73-111) */
    int32_t n_particles; /* N (this rank's shard) */
    float clip_target;   /* grad_clipping: clip on the target score (1e3 KL, 1e2 log-var; +inf = off) */
    float clip_q;        /* clip on the q score (log-var mode only, mcd_cais_var.py:33-40; +inf = off) */
    int32_t lfsteps;     /* CMCD_MODE_UHA only: leapfrog steps per bridge (params_fixed[2], boundingmachine.py:66); 0 reads as 1 */
} cmcd_bridge_desc;

const char* cmcd_last_error(void);
int cmcd_version(void);
int cmcd_num_sms(void);
/* 1 if the library was built with the XLA-FFI handlers CmcdBridgeFwd / CmcdBridgeBwd (csrc/xla_ffi.cc; needs jaxlib's
 * xla/ffi/api/ffi.h at build time), 0 otherwise.  The handlers wrap cmcd_bridge_fwd / cmcd_bridge_bwd for
 * jax.ffi.ffi_call under the reference's jax.jit(jax.grad(compute_bound_fn)) (main.py:162-177); see INTEGRATION.md. */
int cmcd_xla_ffi_available(void);

/*
 * Forward bridge: replaces the jitted vmap(compute_log_elbo) of
 * mcdboundingmachine.compute_bound (mcdboundingmachine.py:126-205) = mcd_utils.evolve
 * (mcd_utils.py:24-33) over mcd_cais.py:46-96 / mcd_cais_var.py:56-107 / mcd_over_orig.py:18-60.
 *   seeds[N] int32; vd_mean[d], vd_logdiag[d]; betas[K]; eps[K] (per-step step size after
 *   the eps schedule, mcd_cais.py:34-44,54-59); out_negw[N] = -w (the per-particle loss);
 *   out_z[N][d] = z_K; traj = NULL or [K+1][d][N] (z_k for the reverse pass).
 *   Modes 4..8 widen eps / vd_logdiag / traj as described at the mode enum (coefficient rows, momentum scales, (z, rho, rho')).
 *   workspace: scratch of cmcd_bridge_fwd_workspace_bytes() bytes (0 for the small-d targets; the lgcp wide
 *   path keeps its [N][1600] state, split-K partials and key chain there).
 */
size_t cmcd_bridge_fwd_workspace_bytes(const cmcd_bridge_desc* desc, const cmcd_net* net, const cmcd_target* target);
int cmcd_bridge_fwd(const cmcd_bridge_desc* desc, void* stream, const int32_t* seeds,
                    const float* vd_mean, const float* vd_logdiag, const float* betas, const float* eps,
                    const cmcd_net* net, const cmcd_target* target,
                    float* out_negw, float* out_z, float* traj, void* workspace, size_t workspace_bytes);

/*
 * Reverse bridge: replaces jax.grad(compute_bound, argnum=1) (main.py:174-176) for the part
 * of the graph that is O(N*K): given traj from cmcd_bridge_fwd and cot_negw[N] = dL/d(-w_n)
 * (1/N for the KL mean loss, 2(l_n - mean l)/N for the log-variance loss,
 * mcdboundingmachine.py:205,231), accumulates cotangents of vd_mean[d], vd_logdiag[d],
 * betas[K], eps[K] and the network tables.  `workspace` of cmcd_bridge_bwd_workspace_bytes()
 * bytes is scratch.  Outputs are overwritten (not accumulated into).
 */
size_t cmcd_bridge_bwd_workspace_bytes(const cmcd_bridge_desc* desc, const cmcd_net* net);
/* same, given the target too: callback targets take the step-wise path at any dim and need its (larger) scratch */
size_t cmcd_bridge_bwd_workspace_bytes_for_target(const cmcd_bridge_desc* desc, const cmcd_net* net, const cmcd_target* target);
int cmcd_bridge_bwd(const cmcd_bridge_desc* desc, void* stream, const int32_t* seeds,
                    const float* vd_mean, const float* vd_logdiag, const float* betas, const float* eps,
                    const cmcd_net* net, const cmcd_target* target,
                    const float* traj, const float* cot_negw,
                    float* g_vd_mean, float* g_vd_logdiag, float* g_betas, float* g_eps,
                    const cmcd_net_grad* g_net, void* workspace, size_t workspace_bytes);

/*
 * Reductions over the per-particle losses (mcdboundingmachine.py:205,231; utils.py:227-237).
 * out[0]=sum l, out[1]=sum l^2, out[2]=max(-l), out[3]=sum exp(-l - max(-l)); callers combine
 * across ranks (sum / max+sum allreduce) and finish: mean, var(ddof=0), ln Z = log(out3)+out2-log n.
 */
int cmcd_loss_stats(void* stream, const float* negw, int64_t n, float* out4);
/* per-batch estimators for the eval driver: losses[batches][n] -> elbo[batches], lnz[batches] */
int cmcd_batched_elbo_lnz(void* stream, const float* losses, int32_t batches, int32_t n, float* elbo, float* lnz);

/*
 * Host-buffer convenience entry (what the reference's host loop hands over per call,
 * opt.py:93-97): seeds on the host, loss and z back on the host; parameters/tables resident
 * on the device.  Copies run on `stream`; the call returns after the D2H copy completed.
 */
int cmcd_bridge_fwd_host(const cmcd_bridge_desc* desc, void* stream, const int32_t* seeds_host,
                         const float* vd_mean, const float* vd_logdiag, const float* betas, const float* eps,
                         const cmcd_net* net, const cmcd_target* target,
                         int32_t* seeds_dev_scratch, float* negw_dev_scratch, float* z_dev_scratch,
                         float* out_negw_host, float* out_z_host);

/*
 * The O(K) chain around the bridge, fused ("prologue" / "epilogue" of one compute_log_elbo call and of its gradient).
 * compute_log_elbo derives, from the flat parameter vector alone: betas (mcdboundingmachine.py:146-149: cumsum(mgridref_y) /
 * sum, interp at target_x), the per-step step sizes (mcd_cais.py:34-44,54-59) and the step-dependent part of the drift
 * network -- PISNet's time coder pushed through the first state layer (nn_dds.py:130-143,156-161) or the geffner net's
 * clamped embedding gather pushed through the three Dense layers (nn.py:62-70); jax.grad then transposes all of it.  These
 * two entry points do that directly on params_flat = ravel_pytree((params_train, params_notrain)) (mcdboundingmachine.py:122):
 * cmcd_chain_fwd writes betas[K], eps[K] and the tables c1/c2/c3 of cmcd_net (plus zero-padded copies of the geffner
 * weights when hidden != hidden_pad; for dds, U1 / W2 / W3 are views of params_flat); cmcd_chain_bwd takes the cotangents
 * cmcd_bridge_bwd produced and writes the flat gradient (zero for leaves outside train_mask -- stop_gradient(params_notrain),
 * mcdboundingmachine.py:142).  off[] holds the offset (in floats) of each pytree leaf inside params_flat, -1 if absent.
 */
enum { CMCD_EPS_CONST = 0, CMCD_EPS_LINEAR = 1, CMCD_EPS_COS_SQ = 2 };
enum {
    CMCD_LEAF_VD_MEAN = 0, CMCD_LEAF_VD_LOGDIAG, CMCD_LEAF_EPS, CMCD_LEAF_MGRID_Y, CMCD_LEAF_GRID_X, CMCD_LEAF_TARGET_X,
    CMCD_LEAF_DDS_PHASE, CMCD_LEAF_DDS_TC1_W, CMCD_LEAF_DDS_TC1_B, CMCD_LEAF_DDS_TC2_W, CMCD_LEAF_DDS_TC2_B,
    CMCD_LEAF_DDS_ST1_W, CMCD_LEAF_DDS_ST1_B, CMCD_LEAF_DDS_ST2_W, CMCD_LEAF_DDS_ST2_B, CMCD_LEAF_DDS_OUT_W, CMCD_LEAF_DDS_OUT_B,
    CMCD_LEAF_GEF_EMB, CMCD_LEAF_GEF_FACTOR, CMCD_LEAF_GEF_W1, CMCD_LEAF_GEF_B1, CMCD_LEAF_GEF_W2, CMCD_LEAF_GEF_B2,
    CMCD_LEAF_GEF_W3, CMCD_LEAF_GEF_B3, CMCD_LEAF_COUNT
};
typedef struct cmcd_chain {
    int32_t arch;          /* CMCD_ARCH_* */
    int32_t dim;           /* d: output width of the network = dimension of z */
    int32_t in_dim;        /* rows of the first-layer weights that see the particle (d) */
    int32_t nbridges;      /* K >= 1 */
    int32_t emb_dim;       /* geffner: E */
    int32_t hidden;        /* H (dds: 64; geffner: in_dim + E) */
    int32_t hidden_pad;    /* HP */
    int32_t eps_schedule;  /* CMCD_EPS_* */
    int32_t ngrid;         /* len(mgridref_y) = ngridb + 1 <= 39 */
    uint32_t train_mask;   /* bit CMCD_LEAF_x set: the leaf is in params_train */
    int64_t n_params;      /* len(params_flat) */
    int64_t off[CMCD_LEAF_COUNT];
    const float* dds_coeff; /* device, 64 floats: linspace(0.1, 100, 64) (nn_dds.py:108) */
} cmcd_chain;
int cmcd_chain_fwd(const cmcd_chain* chain, void* stream, const float* params_flat, float* betas, float* eps,
                   float* c1, float* c2, float* c3, float* U1_pad, float* U2_pad, float* W2_pad, float* W3_pad);
size_t cmcd_chain_bwd_scratch_floats(const cmcd_chain* chain);
int cmcd_chain_bwd(const cmcd_chain* chain, void* stream, const float* params_flat, const float* g_betas, const float* g_eps,
                   const float* g_vd_mean, const float* g_vd_logdiag, const cmcd_net_grad* g_net,
                   float* scratch, size_t scratch_floats, float* grad_flat);

/*
 * mcd_utils.evolve(z, betas, params, rng_key_gen, params_fixed, log_prob_model, eps_schedule, grad_clipping) -> (z, w, None)
 * (src/mcd_utils.py:24-33; bodies src/mcd_cais.py:6-99, src/mcd_cais_var.py:7-112, src/mcd_over_orig.py:6-65), batched over
 * particles: the K bridge steps started from caller-supplied states z0[N][dim] and per-particle PRNG keys keys[N][2] (uint32,
 * the `rng_key_gen` argument: the body consumes it exactly like the reference -- split once, then split / normal / split per
 * step).  out_z[N][dim] = z_K, out_w[N] = sum over the steps of log B_k - log F_k (NOT negated; no -log q(z0) and no
 * log p(z_K): compute_log_elbo adds those, src/mcdboundingmachine.py:157,178).  Forward only (the differentiable entry is
 * cmcd_bridge_fwd / cmcd_bridge_bwd); overdamped modes (0..3), dim 2 or 10, registry targets.  desc, vd_*, betas, eps, net,
 * target as in cmcd_bridge_fwd.
 */
int cmcd_bridge_evolve(const cmcd_bridge_desc* desc, void* stream, const float* z0, const uint32_t* keys,
                       const float* vd_mean, const float* vd_logdiag, const float* betas, const float* eps,
                       const cmcd_net* net, const cmcd_target* target, float* out_z, float* out_w);

/*
 * log p(x), grad log p(x) and (if v != NULL) Hessian(log p)(x) v for x[n][dim] -- replaces calling
 * log_prob_model / jax.grad(log_prob_model) on a batch (model_handler.py:124-284; used by
 * utils.py:54 for plotting and by the tests).  Any output pointer may be NULL.
 */
int cmcd_target_eval(const cmcd_target* target, int32_t dim, void* stream, const float* x, int64_t n,
                     const float* v, float* out_logp, float* out_score, float* out_hvp);

/*
 * Training-loop step of opt.run (opt.py:92-132), device side ("next" rows of the scope table):
 * cmcd_adam_project_step = optimizer.update + optax.apply_updates + project (opt.py:126-128, :14-24) for
 * optimizer = optax.chain(optax.clip(clip), optax.adam(lr, b1, b2, eps)) (opt.py:26-35), fused in one launch.
 *   params, m, v [n] are updated in place (m, v: Adam first / second moments, zero-initialised by the caller);
 *   step = 1-based update count (bias correction); lo / hi [n] or NULL: per-element projection bounds
 *   (eps in [1e-7, 0.5], eta in [0, 0.99], gamma >= 1e-3, mgridref_y >= 1e-3; -inf / +inf elsewhere);
 *   ema [n] or NULL: ema = ema_step * params_new + (1 - ema_step) * ema (optax.incremental_update, opt.py:129-132);
 *   skip_flag (device int32*) or NULL: if *skip_flag != 0 the launch changes nothing -- the NaN "Diverged" guard of
 *   opt.py:122-124 without a host synchronisation.
 * cmcd_randint = jax.random.randint(key, (n,), minval, maxval) int32 (opt.py:93-94, :182-184), key = (key0, key1).
 */
int cmcd_adam_project_step(void* stream, float* params, const float* grad, float* m, float* v, const float* lo,
                           const float* hi, int64_t n, float lr, float b1, float b2, float eps, float clip, int32_t step,
                           float* ema, float ema_step, const int32_t* skip_flag);
int cmcd_randint(void* stream, uint32_t key0, uint32_t key1, int64_t n, int32_t minval, int32_t maxval, int32_t* out);

/*
 * FP32 FMA-pipe probe: `blocks` x 256 threads each run iters*16 dependent-chain FFMAs (2*16*iters*256*blocks
 * flops); scratch needs blocks*256 floats.  bench.py times it with CUDA events to get the measured FP32 roofline
 * denominator (no reference counterpart).
 */
int cmcd_ffma_probe(void* stream, float* scratch, int32_t blocks, int32_t iters);

/* Test hooks for the bit-exact PRNG (jax.random.* call sites listed in csrc/prng.cuh). */
int cmcd_threefry2x32(void* stream, const uint32_t* key2, const uint32_t* x0, const uint32_t* x1, int64_t n,
                      uint32_t* y0, uint32_t* y1);
/* xi0[N][d], xi[K][N][d]: every Gaussian particle n consumes, in order. */
int cmcd_particle_noise(void* stream, const int32_t* seeds, int64_t n, int32_t dim, int32_t nbridges,
                        float* xi0, float* xi);

#ifdef __cplusplus
}
#endif
#endif /* CMCD_B200_H_ */
