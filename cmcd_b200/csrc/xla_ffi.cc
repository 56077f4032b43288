// XLA-FFI (jax.ffi custom-call) shim over the C ABI of include/cmcd_b200.h.
//
// The reference calls its hot path from jitted JAX (src/main.py:162-177: jax.jit(jax.grad(compute_bound_fn, 1,
// has_aux=True), static_argnums=(2,3,4))); a drop-in therefore has to be an XLA custom call.  This translation unit
// binds the two bridge entry points as typed FFI handlers.  It is compiled only where jaxlib's headers exist
// (`python -m cmcd_b200.build --xla-ffi $(python -c "import jax.ffi; print(jax.ffi.include_dir())")`): neither the build
// container nor the GPU box of this project has jax / jaxlib / xla/ffi/api/ffi.h, so this file
// is exercised nowhere in CI -- the tests drive the very same C entry points through ctypes.  INTEGRATION.md shows
// the Python side (jax.ffi.register_ffi_target + jax.custom_vjp).
//
// Buffer order (all device buffers, XLA-owned, dense row-major):
//   fwd  args : seeds s32[N], vd_mean f32[d], vd_logdiag f32[d], betas f32[K], eps f32[K],
//               U1 f32[d,HP], U2 f32[d,HP], U3 f32[d,d], W2 f32[HP,HP], W3 f32[HP,d], c1 f32[T,HP], c2 f32[T,HP], c3 f32[T,d],
//               mix f32[ncomp,6]
//        rets : negw f32[N], z f32[N,d], traj f32[K+1,d,N]
//        attrs: mode, target, arch, hidden, ncomp (i32); clip_target, clip_q, out_scale, out_clip, scale, invalid_below (f32)
//   bwd  args : the fwd args + traj + cot_negw f32[N]
//        rets : g_vd_mean, g_vd_logdiag, g_betas, g_eps, g_U1, g_U2, g_U3, g_W2, g_W3, g_c1, g_c2, g_c3, g_out_scale f32[1],
//               workspace u8[cmcd_bridge_bwd_workspace_bytes]
// (lgcp passes its K^-1 / counts buffers in place of `mix`; omitted here for brevity -- same pattern.)
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define CMCD_HAVE_XLA_FFI 1
#endif
#endif

#ifdef CMCD_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/cmcd_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

namespace {

struct Static {
    int32_t mode, target, arch, hidden, ncomp;
    float clip_target, clip_q, out_scale, out_clip, scale, invalid_below;
};

void fill(const Static& s, const S32& seeds, const F32& vd_mean, const F32& betas, const F32& U1, const F32& U2, const F32& U3,
          const F32& W2, const F32& W3, const F32& c1, const F32& c2, const F32& c3, const F32& mix, cmcd_bridge_desc& d,
          cmcd_net& net, cmcd_target& tg) {
    d.mode = s.mode;
    d.dim = (int32_t)vd_mean.element_count();
    d.nbridges = (int32_t)betas.element_count();
    d.n_particles = (int32_t)seeds.element_count();
    d.clip_target = s.clip_target;
    d.clip_q = s.clip_q;
    d.lfsteps = 0;
    net = cmcd_net{};
    net.arch = s.arch;
    net.hidden = s.hidden;
    net.hidden_pad = (int32_t)W2.dimensions()[0];
    net.n_rows = (int32_t)c1.dimensions()[0];
    net.U1 = U1.typed_data(); net.U2 = U2.element_count() ? U2.typed_data() : nullptr;
    net.U3 = U3.element_count() ? U3.typed_data() : nullptr;
    net.W2 = W2.typed_data(); net.W3 = W3.typed_data();
    net.c1 = c1.typed_data(); net.c2 = c2.typed_data(); net.c3 = c3.typed_data();
    net.out_scale = s.out_scale; net.out_clip = s.out_clip;
    tg = cmcd_target{};
    tg.kind = s.target; tg.ncomp = s.ncomp; tg.scale = s.scale; tg.invalid_below = s.invalid_below;
    tg.mix = mix.typed_data();
}

ffi::Error status(int rc) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(rc == 2 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, cmcd_last_error());
}

ffi::Error BridgeFwdImpl(cudaStream_t stream, S32 seeds, F32 vd_mean, F32 vd_logdiag, F32 betas, F32 eps, F32 U1, F32 U2, F32 U3,
                         F32 W2, F32 W3, F32 c1, F32 c2, F32 c3, F32 mix, RF32 negw, RF32 z, RF32 traj, int32_t mode,
                         int32_t target, int32_t arch, int32_t hidden, int32_t ncomp, float clip_target, float clip_q,
                         float out_scale, float out_clip, float scale, float invalid_below) {
    const Static s{mode, target, arch, hidden, ncomp, clip_target, clip_q, out_scale, out_clip, scale, invalid_below};
    cmcd_bridge_desc d; cmcd_net net; cmcd_target tg;
    fill(s, seeds, vd_mean, betas, U1, U2, U3, W2, W3, c1, c2, c3, mix, d, net, tg);
    return status(cmcd_bridge_fwd(&d, stream, seeds.typed_data(), vd_mean.typed_data(), vd_logdiag.typed_data(), betas.typed_data(),
                                  eps.typed_data(), &net, &tg, negw->typed_data(), z->typed_data(), traj->typed_data(), nullptr, 0));
}

ffi::Error BridgeBwdImpl(cudaStream_t stream, S32 seeds, F32 vd_mean, F32 vd_logdiag, F32 betas, F32 eps, F32 U1, F32 U2, F32 U3,
                         F32 W2, F32 W3, F32 c1, F32 c2, F32 c3, F32 mix, F32 traj, F32 cot_negw, RF32 g_mean, RF32 g_logdiag,
                         RF32 g_betas, RF32 g_eps, RF32 g_U1, RF32 g_U2, RF32 g_U3, RF32 g_W2, RF32 g_W3, RF32 g_c1, RF32 g_c2,
                         RF32 g_c3, RF32 g_os, RU8 ws, int32_t mode, int32_t target, int32_t arch, int32_t hidden, int32_t ncomp,
                         float clip_target, float clip_q, float out_scale, float out_clip, float scale, float invalid_below) {
    const Static s{mode, target, arch, hidden, ncomp, clip_target, clip_q, out_scale, out_clip, scale, invalid_below};
    cmcd_bridge_desc d; cmcd_net net; cmcd_target tg;
    fill(s, seeds, vd_mean, betas, U1, U2, U3, W2, W3, c1, c2, c3, mix, d, net, tg);
    cmcd_net_grad g{};
    g.U1 = g_U1->typed_data(); g.U2 = g_U2->element_count() ? g_U2->typed_data() : nullptr;
    g.U3 = g_U3->element_count() ? g_U3->typed_data() : nullptr;
    g.W2 = g_W2->typed_data(); g.W3 = g_W3->typed_data(); g.c1 = g_c1->typed_data(); g.c2 = g_c2->typed_data();
    g.c3 = g_c3->typed_data(); g.out_scale = g_os->typed_data();
    return status(cmcd_bridge_bwd(&d, stream, seeds.typed_data(), vd_mean.typed_data(), vd_logdiag.typed_data(), betas.typed_data(),
                                  eps.typed_data(), &net, &tg, traj.typed_data(), cot_negw.typed_data(), g_mean->typed_data(),
                                  g_logdiag->typed_data(), g_betas->typed_data(), g_eps->typed_data(), &g, ws->typed_data(),
                                  ws->element_count()));
}

}  // namespace

#define CMCD_STATIC_ATTRS                                                                                             \
    .Attr<int32_t>("mode").Attr<int32_t>("target").Attr<int32_t>("arch").Attr<int32_t>("hidden").Attr<int32_t>("ncomp") \
    .Attr<float>("clip_target").Attr<float>("clip_q").Attr<float>("out_scale").Attr<float>("out_clip")                  \
    .Attr<float>("scale").Attr<float>("invalid_below")

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmcdBridgeFwd, BridgeFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()                       // seeds, vd, betas, eps
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()  // net
                                  .Arg<F32>()                                                                 // mix
                                  .Ret<F32>().Ret<F32>().Ret<F32>() CMCD_STATIC_ATTRS);

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmcdBridgeBwd, BridgeBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>()                                           // mix, traj, cot_negw
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()                                // vd, betas, eps
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<ffi::Buffer<ffi::U8>>() CMCD_STATIC_ATTRS);

extern "C" int cmcd_xla_ffi_available(void) { return 1; }
#else
// No XLA FFI headers on this machine: the shim is intentionally empty (the C ABI is complete without it).
extern "C" int cmcd_xla_ffi_available(void) { return 0; }
#endif
