// Guided partial reverse process driver: the whole iteration structure of the reference's
// GaussianDiffusion.guided_reverse_process (Model/model.py:517-642, explicit t_start branches) is
// enqueued on ONE stream with no host round trip: re-noising, lambda selection (cosine scalar /
// per-pixel map / constant), UNet forward, reduce+apply sampler step, post-iteration clamp, the
// first-iteration delta-map (median by radix select, 4x4 average pool, exp, polynomial curve), guidance
// blends, restart-from-input, and the final mean of the last two iterates.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace ipdm { int schedule_at(int T, double p, int t, double out[10]); }
using namespace ipdm;

extern "C" int ipdm_guided_noise_count(const ipdm_guided_params* p) {
    if (!p || p->n_iters < 1 || p->n_iters > 8) return -1;
    int n = p->n_iters;
    for (int i = 0; i < p->n_iters; ++i) n += p->t_start[i];
    return n;
}

namespace {
size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
struct GuidedWs { float *x, *guide, *eps, *lam_exp, *lam_map; void* sampler; };
size_t carve(const ipdm_guided_params* p, int batch, int h, int w, void* base, GuidedWs* out) {
    const size_t n = (size_t)batch * h * w * sizeof(float);
    const int ks = p->kernel_size > 0 ? p->kernel_size : 1;
    const size_t nl = (size_t)batch * (h / ks) * (w / ks) * sizeof(float);
    char* c = (char*)base; size_t off = 0;
    auto take = [&](size_t b) { char* r = c ? c + off : nullptr; off += a256(b); return r; };
    GuidedWs g;
    g.x = (float*)take(n); g.guide = (float*)take(n); g.eps = (float*)take(n);
    g.lam_exp = (float*)take(nl); g.lam_map = (float*)take(nl);
    g.sampler = take(ipdm_sampler_workspace_bytes(batch, h, w));
    if (out) *out = g;
    return off;
}
}  // namespace

extern "C" size_t ipdm_guided_workspace_bytes(const ipdm_guided_params* p, int batch, int h, int w) {
    if (!p) return 0;
    return carve(p, batch, h, w, nullptr, nullptr);
}

// One body for the whole process (first_it = 0) and for the continuation of the adaptive-schedule branch (first_it = 1: the probing
// iteration has been run by a separate call, its lambda-exponent map is handed in, x restarts from the input, :629-630).
static int guided_impl(ipdm_unet* net, const ipdm_guided_params* p, const float* img, const float* ldct, const float* noise,
                       const float* lam_exp_in, uint64_t call_base, float* iters_out, int batch, int h, int w, void* workspace, void* stream) {
    IPDM_REQUIRE(net && p && img && iters_out && workspace && batch > 0, "ipdm_guided_process: bad arguments");
    IPDM_REQUIRE(p->n_iters >= 1 && p->n_iters <= 8, "ipdm_guided_process: n_iters must be in [1, 8] (the adaptive t_start=None schedule is chosen by the host: ipdm_delta_exp_max + ipdm_guided_process_resume)");
    IPDM_REQUIRE(p->mode == 0 || p->mode == 1, "ipdm_guided_process: mode must be 0 (proj) or 1 (img)");
    IPDM_REQUIRE(p->mode == 0 || ldct != nullptr, "ipdm_guided_process: img mode needs ldct");
    const bool adaptive = !p->constant_guidance_set;
    const int first_it = lam_exp_in ? 1 : 0;
    IPDM_REQUIRE(!lam_exp_in || adaptive, "ipdm_guided_process_resume: only the adaptive-lambda process has a lambda map to resume from");
    if (adaptive) IPDM_REQUIRE(p->kernel_size > 0 && h % p->kernel_size == 0 && w % p->kernel_size == 0,
                               "ipdm_guided_process: H and W must be multiples of kernel_size for the per-pixel lambda map");
    cudaStream_t st = (cudaStream_t)stream;
    GuidedWs ws; carve(p, batch, h, w, workspace, &ws);
    const size_t n1 = (size_t)h * w, n = n1 * batch;
    const float INF = INFINITY;
    uint64_t call = call_base;
    auto tape = [&](uint64_t k) -> const float* { return noise ? noise + (k - call_base) * n : nullptr; };
    const float* lam_exp = lam_exp_in ? lam_exp_in : ws.lam_exp;

    IPDM_CHECK_CUDA(cudaMemcpyAsync(ws.x, img, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const float* guide = img;                                            // imgs = img.clone()  (:538)
    for (int k = 0; k < p->n_iters; ++k) {
        const int it = k + first_it;                                     // iteration number in the reference's `iters` counter
        const int ts = p->t_start[k];
        IPDM_REQUIRE(ts >= 1 && ts < p->timesteps, "ipdm_guided_process: t_start[%d] = %d out of range", k, ts);
        double tab[10];
        IPDM_CHECK(schedule_at(p->timesteps, p->schedule_power, ts, tab));
        IPDM_CHECK(ipdm_q_sample(ws.x, tape(call), ws.x, (float)tab[2], (float)tab[3], n1, batch, p->seed, call, st));   // :545
        ++call;
        std::vector<double> lam_cos(ts);
        if (adaptive && it == 0) IPDM_CHECK(ipdm_cosine_beta_schedule(ts, p->lambda_ratio, lam_cos.data()));          // :546
        for (int i = ts - 1; i >= 0; --i) {
            float lam_scalar = 0.f; const float* lam_map = nullptr;
            if (adaptive) {
                if (it == 0) lam_scalar = (float)lam_cos[i];
                else { IPDM_CHECK(ipdm_lambda_step_map(lam_exp, ws.lam_map, (size_t)batch * (h / p->kernel_size) * (w / p->kernel_size), i, ts, st)); lam_map = ws.lam_map; }
            } else lam_scalar = (float)p->constant_guidance;
            IPDM_CHECK(ipdm_unet_forward(net, ws.x, i, ws.eps, batch, h, w, st));
            IPDM_CHECK(schedule_at(p->timesteps, p->schedule_power, i, tab));
            const float coef[7] = {(float)tab[2], (float)tab[3], (float)tab[4], (float)tab[5], (float)tab[8], (float)tab[9],
                                   expf(0.5f * (float)tab[7])};
            IPDM_CHECK(ipdm_sampler_step(ws.x, guide, ws.eps, tape(call), ws.x, batch, h, w, coef, lam_scalar, lam_map, p->kernel_size,
                                         p->clip, i != 0, p->seed, call, ws.sampler, st));
            ++call;
        }
        if (p->clip) IPDM_CHECK(ipdm_clamp(ws.x, 0.f, p->mode == 1 ? 1.f : INF, n, st));                                   // :569-573
        float* out_it = iters_out + (size_t)k * n;
        IPDM_CHECK_CUDA(cudaMemcpyAsync(out_it, ws.x, n * sizeof(float), cudaMemcpyDeviceToDevice, st));                  // :619
        if (adaptive) {
            if (it == 0) {
                if (p->mode == 0)
                    IPDM_CHECK(ipdm_delta_lambda_map(ws.x, img, ws.lam_exp, nullptr, batch, h, w, p->kernel_size, (float)p->amplitude,
                                                     p->curve_kind, ws.sampler, st));                                   // :596-600, :614
                else                                                                                                     // :591-595 (lam_map is free scratch here)
                    IPDM_CHECK(ipdm_delta_lambda_map_img(ws.x, img, ws.lam_exp, nullptr, ws.lam_map, batch, h, w, p->kernel_size,
                                                         (float)p->amplitude, p->curve_kind, ws.sampler, st));
                IPDM_CHECK_CUDA(cudaMemcpyAsync(ws.x, img, n * sizeof(float), cudaMemcpyDeviceToDevice, st));             // :630
            } else {
                if (p->mode == 0) IPDM_CHECK(ipdm_lincomb(ws.guide, (float)p->eta, out_it, (float)(1 - p->eta), img, 0.f, nullptr, n, st));   // :626
                else IPDM_CHECK(ipdm_lincomb(ws.guide, (float)p->eta, out_it, (float)(0.95 - p->eta), img, 0.05f, ldct, n, st));                // :628
                guide = ws.guide;
            }
        } else {
            if (p->mode == 0) IPDM_CHECK(ipdm_lincomb(ws.guide, (float)p->eta, out_it, (float)(1 - p->eta), img, 0.f, nullptr, n, st));      // :633
            else IPDM_CHECK(ipdm_lincomb(ws.guide, (float)p->eta, out_it, (float)(0.95 - p->eta), img, 0.05f, ldct, n, st));                   // :635
            guide = ws.guide;
        }
    }
    if (p->n_iters + first_it > 1 && p->n_iters > 1) {                                                                    // :637-638
        const float* a = iters_out + (size_t)(p->n_iters - 1) * n;
        const float* b = iters_out + (size_t)(p->n_iters - 2) * n;
        float* o = iters_out + (size_t)p->n_iters * n;
        IPDM_CHECK(ipdm_lincomb(o, 1.f, a, 1.f, b, 0.f, nullptr, n, st));
        IPDM_CHECK(ipdm_lincomb(o, 0.5f, o, 0.f, o, 0.f, nullptr, n, st));
    }
    return IPDM_OK;
}

extern "C" int ipdm_guided_process(ipdm_unet* net, const ipdm_guided_params* p, const float* img, const float* ldct,
                                   const float* noise, float* iters_out, int batch, int h, int w, void* workspace, void* stream) {
    return guided_impl(net, p, img, ldct, noise, nullptr, 0, iters_out, batch, h, w, workspace, stream);
}

extern "C" int ipdm_guided_process_resume(ipdm_unet* net, const ipdm_guided_params* p, const float* img, const float* ldct,
                                          const float* noise, const float* lam_exp, uint64_t call_base, float* iters_out,
                                          int batch, int h, int w, void* workspace, void* stream) {
    IPDM_REQUIRE(lam_exp != nullptr, "ipdm_guided_process_resume: lam_exp is required");
    IPDM_REQUIRE(p && p->n_iters >= 2, "ipdm_guided_process_resume: the continuation holds at least two iterations (its output ends with their mean)");
    return guided_impl(net, p, img, ldct, noise, lam_exp, call_base, iters_out, batch, h, w, workspace, stream);
}
