// UNet noise predictor: architecture builder, weight repacking and the per-shape execution plan.
// Mirrors the reference's Model/model.py UNetModel.__init__ :190-281 (block structure, channel
// plan, attention placement) and UNetModel.forward :283-310 (skip stack, concat, upsample-to-size).
//
// A plan is a flat list of kernel launches over NHWC fp32 tensors carved from one arena with
// liveness-based reuse; it is built once per (batch, H, W) and replayed on the caller's stream with
// no allocation, no host synchronisation and no timestep-dependent host work (the timestep is a
// device scalar that selects the row of each ResBlock's precomputed bias+time-embedding table).
#include "unet_ops.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <tuple>
#include <vector>

namespace ipdm {

static int round_up(int a, int b) { return (a + b - 1) / b * b; }
// Channel stride of an activation tensor.  Thin tensors (<= 16 channels: the 2000x912 and 1000x456 levels, where a pad channel is
// a full HBM stream) are stored dense; wider ones are padded to 32 floats = one 128-byte K chunk of the tf32 tensor-core path.
// A 16-channel tensor that is read RAW by a tensor-core conv (1x1 shortcut over a concat) is widened to 32 by build_plan.
static int alloc_cs(int c) { return c > 16 ? round_up(c, 32) : c; }
static int tc_src_cs(int c) { return c >= 16 ? round_up(c, 32) : c; }      // stride a raw tcgen05 source must have

static int gn_groups(int c) {           // norm_layer, model.py:82-90
    if (c % 32 == 0) return 32;
    if (c < 32) return c;
    int best = 1; long long bd = -1;
    for (int i = 1; i * i <= c; ++i)
        if (c % i == 0)
            for (int f : {i, c / i}) {
                const long long dd = (long long)(f - 32) * (f - 32);
                if (bd < 0 || dd < bd) { bd = dd; best = f; }   // first minimum in the reference's factor order
            }
    return best;
}

struct GNW { float* gamma = nullptr; float* beta = nullptr; int C = 0, groups = 0; };
struct ConvW {
    int cin = 0, cout = 0, k = 0;
    std::vector<float> w_host, b_host;       // PyTorch layout [cout][cin][k][k], [cout]
    float* w_dev = nullptr; float* w_dev_lo = nullptr; float* b_dev = nullptr;
    bool tc = false; int c0 = 0, cs0 = 0, c1 = 0, cs1 = 0, kpad = 0;
    bool bf16 = false;                       // operands (activation tile and packed weights) are bf16
    float* w_dev_direct = nullptr;           // thin layers: the [tap][cin][cout] copy for the direct-kernel fallback
    bool thin = false; int thin_cs = 0;      // thin tensor-core path (conv_thin.cu): operand channel stride 8 / 16 / 32
    // width-folded tensor-core path (pack_fold): `fold` pixels of a row are one pixel of fold*C channels
    float* w_dev_phase = nullptr; int phase_k = 0;   // Upsample conv as four output-parity phases (pack_phase)
    int fold = 0, fold_c0 = 0, fold_c1 = 0, fold_c2 = 0, fold_k = 0; float* w_dev_fold = nullptr; uint64_t fold_mask[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};
struct ResW { GNW gn1, gn2; ConvW conv1, conv2, shortcut, conv2sc /* conv2 with the shortcut folded in (pack_conv2_shortcut) */; bool has_shortcut = false; std::vector<float> temb_w, temb_b; float* bias1_t = nullptr; int cin = 0, cout = 0; };
struct AttnW { GNW norm; ConvW qkv, proj; int C = 0; };
struct LayerRef { enum Kind { CONV_IN, RES, ATTN, DOWN, UP } kind; int idx; };
typedef std::vector<LayerRef> Block;

struct VTensor { int n = 0, h = 0, w = 0, c = 0, cs = 0; int def = -1, last = -1; size_t off = 0; bool external = false; int bf16 = 0;
                 int stats_t = -1, stats_rows = 0, stats_fold = 0; };   // companion tensor with the producer's GroupNorm partials (conv_tc epilogue), rows per slice

struct Op {
    enum Kind { GN_STATS, GN_APPLY, CONV_TC, CONV_DIRECT, CONV_THIN, UPSAMPLE, ATTN } kind;
    int src[3] = {-1, -1, -1}; int nsrc = 0; int n_ident = 0; int dst = -1, res = -1, aux = -1, aux2 = -1, aux3 = -1;
    const GNW* gn = nullptr; int norm_slot = -1; int act = 1;
    const ConvW* cw = nullptr; const float* bias = nullptr; int bias_t_stride = 0; bool use_t = false;
    int stride = 1, upsample = 0, qkv = 0;
    int stats = -1;                            // CONV_TC: tensor that receives the GroupNorm partials of the output
    int fold = 0;                              // CONV_TC: width-fold factor (0: plain)
    int phase = 0;                             // CONV_TC: Upsample conv as four output-parity phases on the low-res source
    // materialised
    ConvTcParams tcp; ConvThinParams thp; ConvDirectDesc cd; GroupNormDesc gd; TensorNHWC out_t, src_t; AttentionParams ap; AttentionDesc ad;
    double flops = 0;
};

struct Plan {
    int B = 0, H = 0, W = 0;
    std::vector<VTensor> vt;
    std::vector<Op> ops;
    float* arena = nullptr; size_t arena_bytes = 0;
    float* norm_buf = nullptr; double* gn_partials = nullptr;
    int x_id = -1, eps_id = -1, first_op = -1, last_op = -1;
    double flops = 0;
    ~Plan() { cudaFree(arena); cudaFree(norm_buf); cudaFree(gn_partials); }
};

}  // namespace ipdm

using namespace ipdm;

struct ipdm_unet {
    ipdm_unet_config cfg;
    std::vector<ConvW> convs;           // conv_in, downs, ups, out conv
    std::vector<ResW> res;
    std::vector<AttnW> attn;
    std::vector<Block> down_blocks, up_blocks;
    Block middle;
    GNW out_gn; int out_conv = -1;
    std::vector<float*> dev_allocs;
    int* t_dev = nullptr;
    int heads = 4;
    int precision = IPDM_PREC_TF32;
    bool force_thin = false;                 // tests: thin tensor-core path regardless of the precision mode
    bool force_fold = false;                 // tests: width-folded path regardless of the precision mode
    std::map<std::tuple<int, int, int>, std::unique_ptr<Plan>> plans;
    ~ipdm_unet() { for (float* p : dev_allocs) cudaFree(p); cudaFree(t_dev); }
};

namespace ipdm {

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
struct Reader {
    const float* p; size_t n, pos = 0; bool ok = true;
    std::vector<float> take(size_t k) {
        if (pos + k > n) { ok = false; return std::vector<float>(k, 0.f); }
        std::vector<float> v(p + pos, p + pos + k); pos += k; return v;
    }
};

static int upload(ipdm_unet* net, const std::vector<float>& h, float** out) {
    float* d = nullptr;
    IPDM_CHECK_CUDA(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(float)));
    IPDM_CHECK_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    net->dev_allocs.push_back(d);
    *out = d;
    return IPDM_OK;
}

static void read_conv(Reader& r, ConvW& c, int cin, int cout, int k, bool bias) {
    c.cin = cin; c.cout = cout; c.k = k;
    c.w_host = r.take((size_t)cout * cin * k * k);
    if (bias) c.b_host = r.take(cout);
}
static int read_gn(ipdm_unet* net, Reader& r, GNW& g, int C) {
    g.C = C; g.groups = gn_groups(C);
    IPDM_CHECK(upload(net, r.take(C), &g.gamma));
    IPDM_CHECK(upload(net, r.take(C), &g.beta));
    return IPDM_OK;
}

static int thin_cs_of(int c) { return c <= 8 ? 8 : (c <= 16 ? 16 : 32); }
static bool wants_tc(int cin_total, int cout) { return cout % 16 == 0 && cin_total >= 16 && (cout >= 64 || cin_total >= 64); }

// K layout of a tensor-core conv: channel c of the virtual concat -> column c (c < c0) or cs0 + (c - c0)
static inline uint16_t bf16_rn_host(float x) {
    uint32_t b; memcpy(&b, &x, 4);
    b += 0x7FFFu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}

// operand_tensor: the conv reads a tensor that exists only as its operand (GroupNorm-apply or upsample output), which the
// bf16 precision mode stores as bf16; raw_sources: the conv reads residual-stream tensors (fp32, virtual concat).
static int pack_conv(ipdm_unet* net, ConvW& c, int c0, int c1, bool raw_sources, int force = -1, bool operand_tensor = false) {
    const int kk = c.k * c.k;
    c.tc = force < 0 ? wants_tc(c.cin, c.cout) : force != 0;
    c.bf16 = c.tc && net->precision == IPDM_PREC_BF16 && (operand_tensor || !raw_sources);
    // thin tensor-core path (kind::tf32): only in the bf16 precision mode.  The full-resolution sinogram layers carry a large smooth
    // signal with ~2 % of fine detail, so a 10-bit operand mantissa there costs 6x on the whole-net error (measured: 4.3e-3 -> 2.8e-2
    // rel-L2 at 2000x912); the tf32 and fp32 modes keep these layers on the exact CUDA-core kernel.  Needs a single operand tensor
    // whose channel stride is 8 / 16 / 32 floats: a GroupNorm-apply or upsample output, or a raw tensor that already has one.
    c.thin = false;
    static const bool thin_off = getenv("IPDM_THIN") && atoi(getenv("IPDM_THIN")) == 0;   // experiment: exact CUDA-core thin layers in the bf16 mode too
    if (!c.tc && force < 0 && ((net->precision == IPDM_PREC_BF16 && !thin_off) || net->force_thin) && (c.cout == 8 || c.cout == 16) && c.cin <= 32 && c1 == 0) {
        if (operand_tensor || !raw_sources) { c.thin = true; c.thin_cs = thin_cs_of(c.cin); }
        else if (c.k == 1 && c0 >= 8 && (alloc_cs(c0) == 8 || alloc_cs(c0) == 16 || alloc_cs(c0) == 32)) { c.thin = true; c.thin_cs = alloc_cs(c0); }
    }
    if (c.thin) {
        std::vector<float> p((size_t)kk * 16 * c.thin_cs, 0.f);
        for (int co = 0; co < c.cout; ++co)
            for (int ci = 0; ci < c.cin; ++ci)
                for (int t = 0; t < kk; ++t) p[((size_t)t * 16 + co) * c.thin_cs + ci] = tf32_rn_host(c.w_host[((size_t)co * c.cin + ci) * kk + t]);
        IPDM_CHECK(upload(net, p, &c.w_dev));
        std::vector<float> pd((size_t)kk * c.cin * c.cout);
        for (int co = 0; co < c.cout; ++co)
            for (int ci = 0; ci < c.cin; ++ci)
                for (int t = 0; t < kk; ++t) pd[((size_t)t * c.cin + ci) * c.cout + co] = c.w_host[((size_t)co * c.cin + ci) * kk + t];
        IPDM_CHECK(upload(net, pd, &c.w_dev_direct));
        if (!c.b_host.empty()) IPDM_CHECK(upload(net, c.b_host, &c.b_dev));
        return IPDM_OK;
    }
    if (c.tc) {
        c.c0 = c0; c.c1 = c1;
        if (c.bf16) { c.c0 = c.cin; c.c1 = 0; c.cs0 = round_up(c.cin, 64); c.cs1 = 0; }
        else if (raw_sources) { c.cs0 = tc_src_cs(c0); c.cs1 = c1 ? tc_src_cs(c1) : 0; }
        else { c.c0 = c.cin; c.c1 = 0; c.cs0 = round_up(c.cin, 32); c.cs1 = 0; }
        IPDM_REQUIRE(c.cs0 % 32 == 0 && c.cs1 % 32 == 0, "pack_conv: source strides %d/%d are not multiples of 32", c.cs0, c.cs1);
        c.kpad = c.cs0 + c.cs1;
        const bool split = net->precision == IPDM_PREC_FP32;       // 3xTF32: w = w_hi + w_lo, both exactly representable in tf32
        std::vector<float> p((size_t)kk * c.cout * c.kpad, 0.f), plo(split ? p.size() : 0, 0.f);
        for (int co = 0; co < c.cout; ++co)
            for (int ci = 0; ci < c.cin; ++ci) {
                const int col = ci < c.c0 ? ci : c.cs0 + (ci - c.c0);
                for (int t = 0; t < kk; ++t) {
                    const float w = c.w_host[((size_t)co * c.cin + ci) * kk + t], hi = tf32_rn_host(w);
                    const size_t at = ((size_t)t * c.cout + co) * c.kpad + col;
                    p[at] = hi;
                    if (split) plo[at] = tf32_rn_host(w - hi);
                }
            }
        if (c.bf16) {                                   // same [tap][cout][K] layout, 2-byte elements
            std::vector<float> packed((p.size() + 1) / 2, 0.f);
            uint16_t* h = reinterpret_cast<uint16_t*>(packed.data());
            for (int co = 0; co < c.cout; ++co)
                for (int ci = 0; ci < c.cin; ++ci)
                    for (int t = 0; t < kk; ++t)
                        h[((size_t)t * c.cout + co) * c.kpad + ci] = bf16_rn_host(c.w_host[((size_t)co * c.cin + ci) * kk + t]);
            IPDM_CHECK(upload(net, packed, &c.w_dev));
        } else {
            IPDM_CHECK(upload(net, p, &c.w_dev));
            if (split) IPDM_CHECK(upload(net, plo, &c.w_dev_lo));
        }
    } else {
        std::vector<float> p((size_t)kk * c.cin * c.cout);
        for (int co = 0; co < c.cout; ++co)
            for (int ci = 0; ci < c.cin; ++ci)
                for (int t = 0; t < kk; ++t) p[((size_t)t * c.cin + ci) * c.cout + co] = c.w_host[((size_t)co * c.cin + ci) * kk + t];
        IPDM_CHECK(upload(net, p, &c.w_dev));
    }
    if (!c.b_host.empty()) IPDM_CHECK(upload(net, c.b_host, &c.b_dev));
    return IPDM_OK;
}

// ------------------------------------------------------------------------------------------------
// Width-folded thin layers.  The 2000x912 and 1000x456 levels of the projection UNet have 4 ... 16 channels (model.py channel_mult
// 1/16, 1/8, 1/4 of 64); as GEMMs they are K = 8 ... 24, N = 8 / 16, and every kernel written for that shape was bound by per-pixel
// instruction issue or by the extra GroupNorm passes (r01: half of the projection forward).  A dense NHWC tensor [H][W][C] IS the
// tensor [H][W/f][f*C]: choose f so that f*C_in and f*C_out are multiples of 32 and the layer becomes an ordinary 32/64-wide
// tensor-core layer on an image of W/f columns, with weights
//     Wf[dy][bdx][(bo, co)][(bi, ci)] = w[co][ci][dy][dx],  dx = f*(bdx - 1) + bi - bo in {-1, 0, 1}, else 0
// (bo / bi = pixel inside the output / input folded pixel).  It runs on the persistent halo kernels unchanged -- TMA zero fill of
// folded column -1 / W/f is the conv padding, the GroupNorm affine + SiLU is applied on the operand path, the epilogue adds the
// residual and emits the GroupNorm statistics -- except that the MMA issuer skips the k-steps (8 input columns) that are
// structurally zero: of the left / right neighbour only the last / first pixel contributes, so a 3x3 layer costs
// 3 * (f + 2) * C_in / 8 k-steps per 128 folded pixels instead of 9 * f * C_in / 8.
// Operands are tf32 (the accuracy-critical full-resolution layers never see bf16), accumulation fp32.
// ------------------------------------------------------------------------------------------------
static int fold_factor(const int* cs, int nsrc, int cout) {
    static const bool off = getenv("IPDM_FOLD") && atoi(getenv("IPDM_FOLD")) == 0;
    if (off || cout > 16) return 0;
    for (int f : {2, 4, 8}) {
        bool ok = f * cout == 32 || f * cout == 64;
        int chunks = 0;
        for (int s = 0; s < nsrc; ++s) { ok = ok && (f * cs[s]) % 32 == 0; chunks += f * cs[s] / 32; }
        if (ok && chunks <= 16) return f;
    }
    return 0;
}

// General form: the folded conv reads `nsrc` dense sources of cs[s] channels; source s is contracted with wsrc[s] (a 3x3 conv, or a
// 1x1 conv = the ResBlock shortcut folded into conv2: centre tap, same pixel only), whose input channels ci_off[s] ... belong to it.
// host part: folded weights p = [tap][f * C_out][f * sum(cs)] and the k-step masks; returns the fold factor (0: shape not eligible)
static int fold_pack_host(ConvW& c, int nsrc, const int* cs, const ConvW* const* wsrc, const int* ci_off, std::vector<float>& p) {
    const int f = fold_factor(cs, nsrc, c.cout);
    if (!f) return 0;
    int ctot = 0; for (int s = 0; s < nsrc; ++s) ctot += cs[s];
    const int kmain = wsrc[0]->k, N = f * c.cout, K = f * ctot, nt = kmain == 3 ? 9 : 1;
    p.assign((size_t)nt * N * K, 0.f);
    for (int t = 0; t < 9; ++t) c.fold_mask[t] = 0;
    for (int tap = 0; tap < nt; ++tap) {
        const int dy = kmain == 3 ? tap / 3 : 1, bdx = kmain == 3 ? tap % 3 : 1;
        int col0 = 0;
        for (int s = 0; s < nsrc; ++s) {
            const ConvW& w = *wsrc[s];
            const int kk = w.k * w.k;
            for (int lc = 0; lc < f * cs[s]; ++lc) {
                const int col = col0 + lc, bi = lc / cs[s], ci = ci_off[s] + lc % cs[s];
                bool any = false;
                for (int bo = 0; bo < f; ++bo) {
                    const int dx = f * (bdx - 1) + bi - bo;
                    if (dx < -1 || dx > 1) continue;
                    if (w.k == 1 && (dx != 0 || dy != 1)) continue;
                    any = true;
                    const int t = w.k == 3 ? dy * 3 + dx + 1 : 0;
                    for (int co = 0; co < c.cout; ++co)
                        p[((size_t)tap * N + bo * c.cout + co) * K + col] = tf32_rn_host(w.w_host[((size_t)co * w.cin + ci) * kk + t]);
                }
                if (any) c.fold_mask[tap] |= 1ull << (col / 8);              // bit = 4 * chunk + k-step
            }
            col0 += f * cs[s];
        }
    }
    c.fold_c0 = cs[0]; c.fold_c1 = nsrc > 1 ? cs[1] : 0; c.fold_c2 = nsrc > 2 ? cs[2] : 0; c.fold_k = K;
    return f;
}

static int pack_fold_multi(ipdm_unet* net, ConvW& c, int nsrc, const int* cs, const ConvW* const* wsrc, const int* ci_off, const float* bias_host) {
    c.fold = 0;
    if (!(net->precision == IPDM_PREC_BF16 || net->force_fold)) return IPDM_OK;
    std::vector<float> p;
    const int f = fold_pack_host(c, nsrc, cs, wsrc, ci_off, p);
    if (!f) return IPDM_OK;
    IPDM_CHECK(upload(net, p, &c.w_dev_fold));
    if (bias_host && !c.b_dev) IPDM_CHECK(upload(net, std::vector<float>(bias_host, bias_host + c.cout), &c.b_dev));
    c.fold = f;
    return IPDM_OK;
}

static int pack_fold(ipdm_unet* net, ConvW& c, int c0, int c1) {
    c.fold = 0;
    if ((c.k != 1 && c.k != 3) || c0 + c1 != c.cin) return IPDM_OK;
    const int cs[2] = {c0, c1}; const ConvW* ws[2] = {&c, &c}; const int off[2] = {0, c0};
    return pack_fold_multi(net, c, c1 ? 2 : 1, cs, ws, off, c.b_host.empty() ? nullptr : c.b_host.data());
}

// conv2 of a channel-changing ResBlock with its 1x1 shortcut folded in (model.py:110-130: out = conv2(act(norm2(h))) + shortcut(x)):
// K = [h | x_a | x_b]; the x columns carry the shortcut weights at the centre tap and zeros elsewhere; bias = conv2.bias + shortcut.bias.
// Dense form for conv_halo_fused_kernel (identity chunks, see ConvTcDesc::n_ident) and width-folded form for the thin levels.
static int pack_conv2_shortcut(ipdm_unet* net, ResW& w, int c0, int c1) {
    static const bool off = getenv("IPDM_SC_FOLD") && atoi(getenv("IPDM_SC_FOLD")) == 0;
    ConvW& c = w.conv2sc;
    c.cin = w.cout + c0 + c1; c.cout = w.cout; c.k = 3; c.tc = false; c.fold = 0;
    if (off || !w.has_shortcut || net->precision == IPDM_PREC_FP32) return IPDM_OK;
    std::vector<float> b(w.cout);
    for (int i = 0; i < w.cout; ++i) b[i] = w.conv2.b_host[i] + w.shortcut.b_host[i];
    IPDM_CHECK(upload(net, b, &c.b_dev));
    // Dense layers: built and measured (IPDM_SC_FOLD_DENSE=1), off by default -- every shortcut chunk costs a whole halo tile on the
    // operand path (TMA + transform slot) for ONE tap of MMAs, so on the HBM-bound 64-channel image layers the folded launch is no
    // faster than conv2 + the separate 1x1 launch (1315 vs 717 + 597 us at 16 x 512 x 512, 1722 vs 1332 us with a 192-channel input)
    // and x would be rounded to bf16 on the skip path.  The width-folded thin layers do gain (proj forward 56.7 -> 53.7 ms).
    static const bool dense_on = getenv("IPDM_SC_FOLD_DENSE") && atoi(getenv("IPDM_SC_FOLD_DENSE")) == 1;
    if (dense_on && w.conv2.tc && w.cout % 64 == 0 && c0 >= 16 && (c1 == 0 || c1 >= 16)) {
        const int csh = round_up(w.cout, 32), cs0 = round_up(c0, 32), cs1 = c1 ? round_up(c1, 32) : 0;
        c.kpad = csh + cs0 + cs1; c.c0 = w.cout; c.cs0 = csh; c.c1 = c0; c.cs1 = cs0;
        c.bf16 = net->precision == IPDM_PREC_BF16;
        std::vector<float> p((size_t)9 * w.cout * c.kpad, 0.f);
        for (int co = 0; co < w.cout; ++co) {
            for (int ci = 0; ci < w.cout; ++ci)
                for (int t = 0; t < 9; ++t) p[((size_t)t * w.cout + co) * c.kpad + ci] = w.conv2.w_host[((size_t)co * w.cout + ci) * 9 + t];
            for (int ci = 0; ci < c0 + c1; ++ci) {
                const int col = csh + (ci < c0 ? ci : cs0 + (ci - c0));
                p[((size_t)4 * w.cout + co) * c.kpad + col] = w.shortcut.w_host[(size_t)co * (c0 + c1) + ci];
            }
        }
        if (c.bf16) {
            std::vector<float> packed((p.size() + 1) / 2, 0.f);
            uint16_t* h = reinterpret_cast<uint16_t*>(packed.data());
            for (size_t i = 0; i < p.size(); ++i) h[i] = bf16_rn_host(p[i]);
            IPDM_CHECK(upload(net, packed, &c.w_dev));
        } else {
            for (float& v : p) v = tf32_rn_host(v);
            IPDM_CHECK(upload(net, p, &c.w_dev));
        }
        c.tc = true;
    }
    {
        const int cs[3] = {w.cout, c0, c1}; const ConvW* ws[3] = {&w.conv2, &w.shortcut, &w.shortcut}; const int offs[3] = {0, 0, c0};
        IPDM_CHECK(pack_fold_multi(net, c, c1 ? 3 : 2, cs, ws, offs, nullptr));
    }
    return IPDM_OK;
}

// Upsample(nearest 2x) + conv3x3 = four 2x2-tap convs on the low-resolution tensor, one per output parity (py, px)
// (ConvTcDesc::phase_up).  Output row 2i+py reads upsampled rows 2i+py-1 .. 2i+py+1 = source rows floor((2i+py+ky-1)/2): for py = 0
// {i-1: ky 0; i: ky 1, 2}, for py = 1 {i: ky 0, 1; i+1: ky 2}; columns alike.  In halo-tile coordinates (origin (i-1, j-1)) phase
// (py, px) uses the tap positions (py + a, px + b), a, b in {0, 1}, with the sums of the 3x3 weights that land there.
static void phase_pack_host(const ConvW& c, std::vector<float>& p) {          // p = [phase 4][tap 9][C_out][round_up(C_in, 32)], fp32 sums
    const int K = round_up(c.cin, 32);
    p.assign((size_t)4 * 9 * c.cout * K, 0.f);
    auto taps_of = [](int parity, int a, int* k) {           // 3x3 taps that land on source offset `a` of this parity; returns the count
        if (parity == 0) { if (a == 0) { k[0] = 0; return 1; } k[0] = 1; k[1] = 2; return 2; }
        if (a == 0) { k[0] = 0; k[1] = 1; return 2; } k[0] = 2; return 1;
    };
    for (int ph = 0; ph < 4; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
                int kys[2], kxs[2];
                const int ny = taps_of(py, a, kys), nx = taps_of(px, b, kxs);
                const int tap = (py + a) * 3 + px + b;
                for (int co = 0; co < c.cout; ++co)
                    for (int ci = 0; ci < c.cin; ++ci) {
                        double acc = 0;
                        for (int iy = 0; iy < ny; ++iy)
                            for (int ix = 0; ix < nx; ++ix) acc += c.w_host[((size_t)co * c.cin + ci) * 9 + kys[iy] * 3 + kxs[ix]];
                        p[(((size_t)ph * 9 + tap) * c.cout + co) * K + ci] = (float)acc;
                    }
            }
    }
}

static int pack_phase(ipdm_unet* net, ConvW& c) {
    static const bool off = getenv("IPDM_PHASE_UP") && atoi(getenv("IPDM_PHASE_UP")) == 0;
    c.w_dev_phase = nullptr;
    if (off || !c.tc || c.k != 3 || net->precision == IPDM_PREC_FP32 || c.cout % 64 != 0) return IPDM_OK;
    const int K = round_up(c.cin, 32);
    std::vector<float> p;
    phase_pack_host(c, p);
    if (net->precision == IPDM_PREC_BF16) {
        std::vector<float> packed((p.size() + 1) / 2, 0.f);
        uint16_t* h = reinterpret_cast<uint16_t*>(packed.data());
        for (size_t i = 0; i < p.size(); ++i) h[i] = bf16_rn_host(p[i]);
        IPDM_CHECK(upload(net, packed, &c.w_dev_phase));
    } else {
        for (float& v : p) v = tf32_rn_host(v);
        IPDM_CHECK(upload(net, p, &c.w_dev_phase));
    }
    c.phase_k = K;
    if (!c.b_host.empty() && !c.b_dev) IPDM_CHECK(upload(net, c.b_host, &c.b_dev));
    return IPDM_OK;
}

static inline float silu_h(float x) { return x / (1.0f + std::exp(-x)); }

// architecture walk shared by create() and param_count(): calls back for every parameter tensor in state_dict order
struct ArchVisitor {
    virtual void conv_in(int cin, int cout) = 0;
    virtual void res(int c0, int c1, int cout, int where) = 0;      // where: 0 down, 1 middle, 2 up ; input = cat(c0, c1)
    virtual void attn(int c, int where) = 0;
    virtual void down(int c) = 0;
    virtual void up(int c) = 0;
    virtual void begin_block(int where) = 0;
    virtual void out(int c, int cout) = 0;
    virtual void time_embed(int mc, int tdim) = 0;
    virtual ~ArchVisitor() {}
};

static void walk_arch(const ipdm_unet_config& cfg, ArchVisitor& v) {
    const int mc = cfg.model_channels, tdim = mc * 4;
    auto is_attn = [&](int ds) { for (int i = 0; i < cfg.n_attn; ++i) if (cfg.attention_resolutions[i] == ds) return true; return false; };
    v.time_embed(mc, tdim);
    int ch = (int)(cfg.channel_mult[0] * mc);
    v.begin_block(0); v.conv_in(cfg.in_channels, ch);
    std::vector<int> chans{ch};
    int ds = 1;
    const int nm = cfg.n_mult - 1;
    for (int level = 0; level < nm; ++level) {
        const int oc = (int)(cfg.channel_mult[level + 1] * mc);
        for (int i = 0; i < cfg.num_res_blocks; ++i) {
            v.begin_block(0); v.res(ch, 0, oc, 0); ch = oc;
            if (is_attn(ds)) v.attn(ch, 0);
            chans.push_back(ch);
        }
        if (level != nm - 1) { v.begin_block(0); v.down(ch); chans.push_back(ch); ds *= 2; }
    }
    v.begin_block(1); v.res(ch, 0, ch, 1); v.attn(ch, 1); v.res(ch, 0, ch, 1);
    for (int level = nm - 1; level >= 0; --level) {
        const int oc = (int)(mc * cfg.channel_mult[level + 1]);
        for (int i = 0; i <= cfg.num_res_blocks; ++i) {
            const int skip = chans.back(); chans.pop_back();
            v.begin_block(2); v.res(ch, skip, oc, 2); ch = oc;
            if (is_attn(ds)) v.attn(ch, 2);
            if (level && i == cfg.num_res_blocks) { v.up(ch); ds /= 2; }
        }
    }
    v.out(ch, cfg.out_channels);
}

struct CountVisitor : ArchVisitor {
    long long n = 0;
    void conv(int ci, int co, int k, bool b = true) { n += (long long)co * ci * k * k + (b ? co : 0); }
    void time_embed(int mc, int tdim) override { n += (long long)tdim * mc + tdim + (long long)tdim * tdim + tdim; tdim_ = tdim; }
    void conv_in(int ci, int co) override { conv(ci, co, 3); }
    void res(int c0, int c1, int co, int) override {
        const int ci = c0 + c1;
        n += 2 * ci; conv(ci, co, 3); n += (long long)co * tdim_ + co; n += 2 * co; conv(co, co, 3);
        if (ci != co) conv(ci, co, 1);
    }
    void attn(int c, int) override { n += 2 * c; conv(c, 3 * c, 1, false); conv(c, c, 1); }
    void down(int c) override { conv(c, c, 3); }
    void up(int c) override { conv(c, c, 3); }
    void begin_block(int) override {}
    void out(int c, int co) override { n += 2 * c; conv(c, co, 3); }
    int tdim_ = 0;
};

struct BuildVisitor : ArchVisitor {
    ipdm_unet* net; Reader& r; int rc = IPDM_OK; int tdim = 0, mc = 0;
    std::vector<float> te_w0, te_b0, te_w1, te_b1;
    Block* cur = nullptr;
    BuildVisitor(ipdm_unet* n, Reader& rd) : net(n), r(rd) {}
    void fail(int c) { if (rc == IPDM_OK) rc = c; }
    void time_embed(int mc_, int tdim_) override {
        mc = mc_; tdim = tdim_;
        te_w0 = r.take((size_t)tdim * mc); te_b0 = r.take(tdim); te_w1 = r.take((size_t)tdim * tdim); te_b1 = r.take(tdim);
    }
    void begin_block(int where) override {
        if (where == 0) { net->down_blocks.emplace_back(); cur = &net->down_blocks.back(); }
        else if (where == 1) cur = &net->middle;
        else { net->up_blocks.emplace_back(); cur = &net->up_blocks.back(); }
    }
    void conv_in(int ci, int co) override {
        net->convs.emplace_back(); read_conv(r, net->convs.back(), ci, co, 3, true);
        fail(pack_conv(net, net->convs.back(), ci, 0, true));
        cur->push_back({LayerRef::CONV_IN, (int)net->convs.size() - 1});
    }
    void res(int c0, int c1, int co, int) override {
        net->res.emplace_back(); ResW& w = net->res.back();
        const int ci = c0 + c1; w.cin = ci; w.cout = co;
        fail(read_gn(net, r, w.gn1, ci)); read_conv(r, w.conv1, ci, co, 3, true);
        w.temb_w = r.take((size_t)co * tdim); w.temb_b = r.take(co);
        fail(read_gn(net, r, w.gn2, co)); read_conv(r, w.conv2, co, co, 3, true);
        w.has_shortcut = ci != co;
        if (w.has_shortcut) read_conv(r, w.shortcut, ci, co, 1, true);
        fail(pack_conv(net, w.conv1, ci, 0, false));
        fail(pack_conv(net, w.conv2, co, 0, false));
        if (w.has_shortcut) fail(pack_conv(net, w.shortcut, c0, c1, true));
        fail(pack_fold(net, w.conv1, c0, c1)); fail(pack_fold(net, w.conv2, co, 0));
        if (w.has_shortcut) fail(pack_fold(net, w.shortcut, c0, c1));
        fail(pack_conv2_shortcut(net, w, c0, c1));
        cur->push_back({LayerRef::RES, (int)net->res.size() - 1});
    }
    void attn(int c, int) override {
        net->attn.emplace_back(); AttnW& a = net->attn.back(); a.C = c;
        fail(read_gn(net, r, a.norm, c)); read_conv(r, a.qkv, c, 3 * c, 1, false); read_conv(r, a.proj, c, c, 1, true);
        fail(pack_conv(net, a.qkv, c, 0, false));        // reads the GroupNorm-apply output (operand tensor)
        fail(pack_conv(net, a.proj, c, 0, true));         // reads the attention output (fp32 residual-stream layout)
        if (!a.qkv.tc || !a.proj.tc || c % (64 * net->heads) != 0 || c / net->heads != 64) {
            set_error("attention block with %d channels / %d heads is not supported (head_dim must be 64)", c, net->heads);
            fail(IPDM_ERR_UNSUPPORTED);
        }
        cur->push_back({LayerRef::ATTN, (int)net->attn.size() - 1});
    }
    void down(int c) override {
        net->convs.emplace_back(); read_conv(r, net->convs.back(), c, c, 3, true);
        fail(pack_conv(net, net->convs.back(), c, 0, true));
        cur->push_back({LayerRef::DOWN, (int)net->convs.size() - 1});
    }
    void up(int c) override {
        net->convs.emplace_back(); read_conv(r, net->convs.back(), c, c, 3, true);
        fail(pack_conv(net, net->convs.back(), c, 0, true, -1, true));
        fail(pack_fold(net, net->convs.back(), c, 0));
        fail(pack_phase(net, net->convs.back()));
        cur->push_back({LayerRef::UP, (int)net->convs.size() - 1});
    }
    void out(int c, int co) override {
        fail(read_gn(net, r, net->out_gn, c));
        net->convs.emplace_back(); read_conv(r, net->convs.back(), c, co, 3, true);
        fail(pack_conv(net, net->convs.back(), c, 0, false));
        net->out_conv = (int)net->convs.size() - 1;
    }
};

// bias1_t[t][co] = conv1.bias[co] + Linear(SiLU(time_embed(timestep_embedding(t))))[co]   (model.py:14-32, 218-222, 105-108, 128)
static int build_time_tables(ipdm_unet* net, const BuildVisitor& bv) {
    const int mc = bv.mc, tdim = bv.tdim, T = net->cfg.max_t, half = mc / 2;
    std::vector<float> emb((size_t)T * tdim);
    for (int t = 0; t < T; ++t) {
        std::vector<float> te(mc), h(tdim);
        for (int i = 0; i < half; ++i) {
            const float f = (float)std::exp(-std::log(10000.0) * i / half);
            te[i] = std::cos((float)t * f); te[half + i] = std::sin((float)t * f);
        }
        for (int o = 0; o < tdim; ++o) { double a = bv.te_b0[o]; for (int i = 0; i < mc; ++i) a += (double)bv.te_w0[(size_t)o * mc + i] * te[i]; h[o] = silu_h((float)a); }
        for (int o = 0; o < tdim; ++o) { double a = bv.te_b1[o]; for (int i = 0; i < tdim; ++i) a += (double)bv.te_w1[(size_t)o * tdim + i] * h[i]; emb[(size_t)t * tdim + o] = silu_h((float)a); }
    }   // emb now holds SiLU(time_embed(.)), the input of every block's time_emb Linear
    for (ResW& w : net->res) {
        std::vector<float> tab((size_t)T * w.cout);
        for (int t = 0; t < T; ++t)
            for (int o = 0; o < w.cout; ++o) {
                double a = w.temb_b[o];
                for (int i = 0; i < tdim; ++i) a += (double)w.temb_w[(size_t)o * tdim + i] * emb[(size_t)t * tdim + i];
                tab[(size_t)t * w.cout + o] = (float)((double)w.conv1.b_host[o] + a);
            }
        IPDM_CHECK(upload(net, tab, &w.bias1_t));
    }
    return IPDM_OK;
}

// ------------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------------
struct PlanBuilder {
    ipdm_unet* net; Plan* pl;
    int norm_slots = 0; int max_c = 0;
    int new_tensor(int n, int h, int w, int c, int cs) {
        VTensor t; t.n = n; t.h = h; t.w = w; t.c = c; t.cs = cs;
        pl->vt.push_back(t); return (int)pl->vt.size() - 1;
    }
    int act(int h, int w, int c) { return new_tensor(pl->B, h, w, c, alloc_cs(c)); }
    void touch(int id, int opi) { if (id < 0) return; VTensor& t = pl->vt[id]; if (t.def < 0) t.def = opi; t.last = std::max(t.last, opi); }
    int push(Op& o) {
        const int i = (int)pl->ops.size();
        for (int s = 0; s < o.nsrc; ++s) touch(o.src[s], i);
        touch(o.dst, i); touch(o.res, i); touch(o.aux, i); touch(o.aux2, i); touch(o.aux3, i);
        pl->ops.push_back(o); return i;
    }
    int cin_of(const int* src, int nsrc) { int c = 0; for (int i = 0; i < nsrc; ++i) c += pl->vt[src[i]].c; return c; }
    // width-folded path (pack_fold): every tensor the conv touches is dense fp32 with the channel split the weights were folded for,
    // and the row length is a multiple of the fold factor
    bool can_fold(const ConvW& cw, const int* src, int nsrc, int dst, int res) {
        if (!cw.fold) return false;
        const VTensor& s0 = pl->vt[src[0]];
        if (s0.w % cw.fold != 0 || s0.c != cw.fold_c0 || (nsrc > 1 ? pl->vt[src[1]].c : 0) != cw.fold_c1 || (nsrc > 2 ? pl->vt[src[2]].c : 0) != cw.fold_c2) return false;
        for (int i = 0; i < nsrc; ++i) { const VTensor& t = pl->vt[src[i]]; if (t.cs != t.c || t.bf16 || t.external) return false; }
        const VTensor& d = pl->vt[dst];
        if (d.cs != d.c || d.external || d.c != cw.cout) return false;
        if (res >= 0) { const VTensor& r = pl->vt[res]; if (r.cs != r.c || r.c != cw.cout || r.external) return false; }
        return true;
    }

    bool can_fold_shortcut(const ResW& w, int h1, const int* xs, int nxs, int dst) {
        const ConvW& cw = w.conv2sc;
        const int all[3] = {h1, xs[0], nxs > 1 ? xs[1] : -1};
        if (cw.fold && can_fold(cw, all, 1 + nxs, dst, -1)) return true;
        if (!cw.tc || net->precision == IPDM_PREC_FP32) return false;
        const VTensor& h = pl->vt[h1];
        if (!conv_tc_can_fuse_norm(h.h, h.w, pl->B, cw.cout, 9, 1)) return false;
        for (int i = 0; i < nxs; ++i) { const VTensor& t = pl->vt[xs[i]]; if (t.bf16 || t.external || t.c < 16) return false; }
        return true;
    }

    // GroupNorm(+SiLU) followed by a conv; returns nothing, writes `dst`
    // xs / nxs: the ResBlock input when its 1x1 shortcut is folded into this conv (cw then is ResW::conv2sc); the caller has checked can_fold_shortcut
    void norm_conv(const int* src, int nsrc, const GNW& gn, const ConvW& cw, int dst, int res, const float* bias, int bstride, bool use_t, int act_silu,
                   const int* xs = nullptr, int nxs = 0) {
        Op st; st.kind = Op::GN_STATS; st.nsrc = nsrc; st.src[0] = src[0]; st.src[1] = nsrc > 1 ? src[1] : -1; st.gn = &gn; st.norm_slot = norm_slots++;
        max_c = std::max(max_c, gn.C);
        push(st);
        const VTensor& s0 = pl->vt[src[0]];
        // IPDM_THIN_FUSE=1: the thin kernel normalises the TMA-landed tile itself (bit-identical results).  Off by default: with only
        // the two spare warps per CTA doing the transform it is SLOWER than the separate apply pass (measured at 16 slices: 8 -> 8 at
        // 2000x912 1025 us fused vs 600 + 311 us; 16 -> 16 at 1000x456 526 vs 283 + 155 us).
        static const bool thin_fuse = getenv("IPDM_THIN_FUSE") && atoi(getenv("IPDM_THIN_FUSE")) == 1;
        if (nxs) {
            // conv2 + folded shortcut: sources = [h | x ...], the last nxs are read raw and meet the centre tap only
            Op cv; cv.kind = Op::CONV_TC; cv.nsrc = 1 + nxs; cv.n_ident = nxs; cv.src[0] = src[0]; cv.src[1] = xs[0]; cv.src[2] = nxs > 1 ? xs[1] : -1;
            cv.cw = &cw; cv.dst = dst; cv.res = -1; cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t; cv.norm_slot = st.norm_slot; cv.gn = &gn; cv.act = act_silu;
            const int all[3] = {src[0], xs[0], nxs > 1 ? xs[1] : -1};
            cv.fold = can_fold(cw, all, 1 + nxs, dst, -1) ? cw.fold : 0;
            push(cv);
        } else if (cw.k == 3 && can_fold(cw, src, nsrc, dst, res)) {
            // width-folded: the persistent halo kernel normalises the raw tile(s) on its operand path and emits the output statistics
            Op cv; cv.kind = Op::CONV_TC; cv.nsrc = nsrc; cv.src[0] = src[0]; cv.src[1] = st.src[1]; cv.cw = &cw; cv.dst = dst; cv.res = res;
            cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t; cv.norm_slot = st.norm_slot; cv.gn = &gn; cv.act = act_silu; cv.fold = cw.fold;
            push(cv);
        } else if (cw.thin && thin_fuse && nsrc == 1 && s0.c == cw.thin_cs && s0.cs == cw.thin_cs && !s0.bf16) {
            // single dense source: the thin kernel normalises the TMA-landed tile itself, no operand tensor and no apply pass
            Op cv; cv.kind = Op::CONV_THIN; cv.nsrc = 1; cv.src[0] = src[0]; cv.cw = &cw; cv.dst = dst; cv.res = res; cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t;
            cv.norm_slot = st.norm_slot; cv.gn = &gn; cv.act = act_silu;
            push(cv);
        } else if (cw.thin) {
            const int a = new_tensor(pl->B, s0.h, s0.w, gn.C, cw.thin_cs);
            Op ap; ap.kind = Op::GN_APPLY; ap.nsrc = nsrc; ap.src[0] = src[0]; ap.src[1] = st.src[1]; ap.gn = &gn; ap.norm_slot = st.norm_slot; ap.dst = a; ap.act = act_silu;
            push(ap);
            Op cv; cv.kind = Op::CONV_THIN; cv.nsrc = 1; cv.src[0] = a; cv.cw = &cw; cv.dst = dst; cv.res = res; cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t;
            push(cv);
        } else if (cw.tc && net->precision != IPDM_PREC_FP32 && cw.k == 3 && (nsrc == 1 || s0.c % 32 == 0) &&
                   conv_tc_can_fuse_norm(s0.h, s0.w, pl->B, cw.cout, 9, 1)) {
            // the conv applies the GroupNorm affine + SiLU itself on its operand path (conv_halo_fused_kernel): no apply pass, no operand tensor
            Op cv; cv.kind = Op::CONV_TC; cv.nsrc = nsrc; cv.src[0] = src[0]; cv.src[1] = st.src[1]; cv.cw = &cw; cv.dst = dst; cv.res = res;
            cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t; cv.norm_slot = st.norm_slot; cv.gn = &gn; cv.act = act_silu;
            push(cv);
        } else if (cw.tc) {
            const int a = new_tensor(pl->B, s0.h, s0.w, gn.C, cw.bf16 ? round_up(gn.C, 64) : round_up(gn.C, 32));
            pl->vt[a].bf16 = cw.bf16;
            Op ap; ap.kind = Op::GN_APPLY; ap.nsrc = nsrc; ap.src[0] = src[0]; ap.src[1] = st.src[1]; ap.gn = &gn; ap.norm_slot = st.norm_slot; ap.dst = a; ap.act = act_silu;
            push(ap);
            Op cv; cv.kind = Op::CONV_TC; cv.nsrc = 1; cv.src[0] = a; cv.cw = &cw; cv.dst = dst; cv.res = res; cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t;
            push(cv);
        } else {
            Op cv; cv.kind = Op::CONV_DIRECT; cv.nsrc = nsrc; cv.src[0] = src[0]; cv.src[1] = st.src[1]; cv.cw = &cw; cv.dst = dst; cv.res = res;
            cv.bias = bias; cv.bias_t_stride = bstride; cv.use_t = use_t; cv.norm_slot = st.norm_slot; cv.gn = &gn;
            push(cv);
        }
    }
    void plain_conv(const int* src, int nsrc, const ConvW& cw, int dst, int res, int stride, int upsample) {
        if (stride == 1 && !upsample && can_fold(cw, src, nsrc, dst, res)) {
            Op cv; cv.kind = Op::CONV_TC; cv.nsrc = nsrc; cv.src[0] = src[0]; cv.src[1] = nsrc > 1 ? src[1] : -1;
            cv.cw = &cw; cv.dst = dst; cv.res = res; cv.bias = cw.b_dev; cv.fold = cw.fold;
            push(cv);
            return;
        }
        Op cv; cv.kind = cw.tc ? Op::CONV_TC : ((cw.thin && stride == 1 && !upsample && nsrc == 1 && pl->vt[src[0]].cs == cw.thin_cs) ? Op::CONV_THIN : Op::CONV_DIRECT); cv.nsrc = nsrc; cv.src[0] = src[0]; cv.src[1] = nsrc > 1 ? src[1] : -1;
        cv.cw = &cw; cv.dst = dst; cv.res = res; cv.bias = cw.b_dev; cv.stride = stride; cv.upsample = upsample;
        push(cv);
    }
    int res_block(const ResW& w, const int* src, int nsrc) {
        const VTensor s0 = pl->vt[src[0]];
        const int h1 = act(s0.h, s0.w, w.cout);
        norm_conv(src, nsrc, w.gn1, w.conv1, h1, -1, w.bias1_t, w.cout, true, 1);
        int resid = src[0];
        const int out = act(s0.h, s0.w, w.cout);
        if (w.has_shortcut && can_fold_shortcut(w, h1, src, nsrc, out)) {
            // out = conv2(act(norm2(h1))) + shortcut(x) as ONE launch: x rides along as extra K chunks of conv2 (no shortcut tensor, no
            // residual read in the epilogue)
            norm_conv(&h1, 1, w.gn2, w.conv2sc, out, -1, w.conv2sc.b_dev, 0, false, 1, src, nsrc);
            return out;
        }
        if (w.has_shortcut) { resid = act(s0.h, s0.w, w.cout); plain_conv(src, nsrc, w.shortcut, resid, -1, 1, 0); }
        norm_conv(&h1, 1, w.gn2, w.conv2, out, resid, w.conv2.b_dev, 0, false, 1);
        return out;
    }
    int attn_block(const AttnW& a, int x) {
        const VTensor s = pl->vt[x];
        const int T = s.h * s.w, tpad = round_up(T, 8);
        const bool split = net->precision == IPDM_PREC_FP32;
        const bool bf = net->precision == IPDM_PREC_BF16;          // q, k, v^T stored as bf16, attention runs kind::f16
        const int qk = new_tensor(pl->B, s.h, s.w, 3 * a.C, 3 * a.C);
        const int vt = new_tensor(pl->B, 1, 1, a.C * tpad, a.C * tpad);
        pl->vt[qk].bf16 = bf; pl->vt[vt].bf16 = bf;
        const int qk_lo = split ? new_tensor(pl->B, s.h, s.w, 3 * a.C, 3 * a.C) : -1;
        const int vt_lo = split ? new_tensor(pl->B, 1, 1, a.C * tpad, a.C * tpad) : -1;
        Op st; st.kind = Op::GN_STATS; st.nsrc = 1; st.src[0] = x; st.gn = &a.norm; st.norm_slot = norm_slots++; max_c = std::max(max_c, a.C); push(st);
        const int an = new_tensor(pl->B, s.h, s.w, a.C, a.C);
        pl->vt[an].bf16 = a.qkv.bf16;
        Op ap; ap.kind = Op::GN_APPLY; ap.nsrc = 1; ap.src[0] = x; ap.gn = &a.norm; ap.norm_slot = st.norm_slot; ap.dst = an; ap.act = 0; push(ap);
        Op q; q.kind = Op::CONV_TC; q.nsrc = 1; q.src[0] = an; q.cw = &a.qkv; q.dst = qk; q.aux = vt; q.qkv = 1; q.aux2 = qk_lo; q.aux3 = vt_lo; push(q);   // epilogue rounds q,k,v to tf32 (hi/lo pairs in fp32 mode)
        const int o = new_tensor(pl->B, s.h, s.w, a.C, a.C);
        Op at; at.kind = Op::ATTN; at.nsrc = 1; at.src[0] = qk; at.aux = vt; at.aux2 = qk_lo; at.aux3 = vt_lo; at.dst = o; push(at);
        const int out = act(s.h, s.w, a.C);
        Op pj; pj.kind = Op::CONV_TC; pj.nsrc = 1; pj.src[0] = o; pj.cw = &a.proj; pj.dst = out; pj.res = x; pj.bias = a.proj.b_dev; push(pj);
        return out;
    }
    int run_block(const Block& blk, const int* src_in, int nsrc_in, int up_h, int up_w) {
        int cur[2] = {src_in[0], nsrc_in > 1 ? src_in[1] : -1}; int ncur = nsrc_in;
        for (const LayerRef& l : blk) {
            int out = -1;
            const VTensor s0 = pl->vt[cur[0]];
            switch (l.kind) {
                case LayerRef::CONV_IN: out = act(s0.h, s0.w, net->convs[l.idx].cout); plain_conv(cur, 1, net->convs[l.idx], out, -1, 1, 0); break;
                case LayerRef::RES: out = res_block(net->res[l.idx], cur, ncur); break;
                case LayerRef::ATTN: out = attn_block(net->attn[l.idx], cur[0]); break;
                case LayerRef::DOWN: out = act((s0.h + 1) / 2, (s0.w + 1) / 2, s0.c); plain_conv(cur, 1, net->convs[l.idx], out, -1, 2, 0); break;
                case LayerRef::UP: {
                    const ConvW& cw = net->convs[l.idx];
                    out = act(up_h, up_w, s0.c);
                    if (cw.w_dev_phase && up_h == 2 * s0.h && up_w == 2 * s0.w && !s0.bf16 && !s0.external && s0.c >= 32 &&
                        conv_tc_can_fuse_norm(s0.h, s0.w, pl->B, cw.cout, 9, 1)) {
                        // exactly 2x: four 2x2-tap phases straight from the low-resolution tensor -- no upsampled operand tensor, 16/36 of the MMAs
                        Op cv; cv.kind = Op::CONV_TC; cv.nsrc = 1; cv.src[0] = cur[0]; cv.cw = &cw; cv.dst = out; cv.bias = cw.b_dev; cv.phase = 1;
                        push(cv);
                    } else if (cw.fold && up_w % cw.fold == 0 && pl->vt[out].cs == pl->vt[out].c) {
                        const int u = new_tensor(pl->B, up_h, up_w, s0.c, s0.c);          // dense tf32-rounded operand of the folded conv
                        Op up; up.kind = Op::UPSAMPLE; up.nsrc = 1; up.src[0] = cur[0]; up.dst = u; push(up);
                        plain_conv(&u, 1, cw, out, -1, 1, 0);
                    } else if (cw.tc || cw.thin) {
                        const int u = cw.thin ? new_tensor(pl->B, up_h, up_w, s0.c, cw.thin_cs)
                                              : (cw.bf16 ? new_tensor(pl->B, up_h, up_w, s0.c, round_up(s0.c, 64)) : act(up_h, up_w, s0.c));
                        pl->vt[u].bf16 = cw.bf16;
                        Op up; up.kind = Op::UPSAMPLE; up.nsrc = 1; up.src[0] = cur[0]; up.dst = u; push(up);
                        plain_conv(&u, 1, cw, out, -1, 1, 0);
                    } else plain_conv(cur, 1, cw, out, -1, 1, 1);
                } break;
            }
            cur[0] = out; cur[1] = -1; ncur = 1;
        }
        return cur[0];
    }
};

static TensorNHWC resolve(const Plan& pl, int id) {
    TensorNHWC t;
    if (id < 0) return t;
    const VTensor& v = pl.vt[id];
    t.n = v.n; t.h = v.h; t.w = v.w; t.c = v.c; t.cs = v.cs; t.bf16 = v.bf16;
    t.p = v.external ? nullptr : (float*)((char*)pl.arena + v.off);
    return t;
}

// turns the descriptor of a conv over dense thin tensors into its width-folded form (views, folded weights, k-step masks)
static void fold_view(TensorNHWC& t, int f) { if (t.p || t.c) { t.w /= f; t.c *= f; t.cs *= f; } }
static void fold_desc(ConvTcDesc& d, const ConvW& cw) {
    const int f = cw.fold;
    d.gn_mod[0] = d.src[0].c; d.gn_mod[1] = d.nsrc - d.n_ident > 1 ? d.src[1].c : 0; d.bias_mod = cw.cout;
    for (int s = 0; s < d.nsrc; ++s) fold_view(d.src[s], f);
    fold_view(d.out, f);
    if (d.res.p) fold_view(d.res, f);
    d.cout = f * cw.cout; d.n_tile = d.cout;
    d.w_packed = cw.w_dev_fold; d.w_packed_lo = nullptr; d.w_k = cw.fold_k; d.w_bf16 = 0;
    for (int t = 0; t < 9; ++t) d.kmask[t] = cw.k == 3 ? cw.fold_mask[t] : 0;
    d.fold = f;
    if (cw.k == 3 && !d.norm_scale) d.passthrough = 1;     // plain folded 3x3 (the Upsample conv): same kernel, identity operand path
}

static int build_plan(ipdm_unet* net, int B, int H, int W, Plan** out) {
    std::unique_ptr<Plan> pl(new Plan());
    pl->B = B; pl->H = H; pl->W = W;
    PlanBuilder pb{net, pl.get()};
    pl->x_id = pb.new_tensor(B, H, W, net->cfg.in_channels, net->cfg.in_channels); pl->vt[pl->x_id].external = true;
    pl->eps_id = pb.new_tensor(B, H, W, net->cfg.out_channels, net->cfg.out_channels); pl->vt[pl->eps_id].external = true;

    // ---- forward wiring (model.py:283-310) ----
    std::vector<int> hs;
    int h = pl->x_id;
    for (const Block& blk : net->down_blocks) { h = pb.run_block(blk, &h, 1, 0, 0); hs.push_back(h); }
    h = pb.run_block(net->middle, &h, 1, 0, 0);
    int h_ = hs.back(); hs.pop_back();
    for (const Block& blk : net->up_blocks) {
        const int cat[2] = {h, h_};
        if (!hs.empty()) { h_ = hs.back(); hs.pop_back(); }
        h = pb.run_block(blk, cat, 2, pl->vt[h_].h, pl->vt[h_].w);
    }
    pb.norm_conv(&h, 1, net->out_gn, net->convs[net->out_conv], pl->eps_id, -1, net->convs[net->out_conv].b_dev, 0, false, 1);
    pl->first_op = 0; pl->last_op = (int)pl->ops.size() - 1;
    // raw fp32 sources of tensor-core convs are read in 128-byte K chunks: widen the (16-channel) tensors they touch, and send
    // thin-path convs whose source no longer has the packed stride to the direct kernel
    for (const Op& o : pl->ops)
        if (o.kind == Op::CONV_TC && !o.fold)
            for (int s = 0; s < o.nsrc; ++s) {
                VTensor& t = pl->vt[o.src[s]];
                if (!t.bf16 && !t.external && t.cs % 32 != 0) t.cs = round_up(t.c, 32);
            }
    for (Op& o : pl->ops)
        if (o.kind == Op::CONV_THIN && pl->vt[o.src[0]].cs != o.cw->thin_cs) o.kind = Op::CONV_DIRECT;
    for (const Op& o : pl->ops)
        if (o.kind == Op::CONV_TC && o.fold)
            for (int id : {o.src[0], o.src[1], o.src[2], o.dst, o.res})
                IPDM_REQUIRE(id < 0 || pl->vt[id].cs == pl->vt[id].c, "unet: a width-folded conv shares a tensor with a layer that needs it padded (%d -> %d channels)",
                             pl->vt[id].c, pl->vt[id].cs);
    // GroupNorm statistics come from the epilogue of the conv that writes the tensor when that conv runs a persistent tensor-core
    // kernel (decided in conv_tc_prepare; the 3xTF32 and qkv kernels do not): a companion tensor holds the per-warp-row partials
    // from the producer to the last GroupNorm that reads the tensor, and that GroupNorm skips its own read of the tensor.
    if (net->precision != IPDM_PREC_FP32) {
        std::vector<int> producer(pl->vt.size(), -1);
        for (int i = 0; i < (int)pl->ops.size(); ++i) if (pl->ops[i].dst >= 0) producer[pl->ops[i].dst] = i;
        for (int gi = 0; gi < (int)pl->ops.size(); ++gi) {
            if (pl->ops[gi].kind != Op::GN_STATS) continue;
            for (int sidx = 0; sidx < pl->ops[gi].nsrc; ++sidx) {
                const int t = pl->ops[gi].src[sidx];
                const int pi = t < (int)producer.size() ? producer[t] : -1;
                // (the thin kernel's epilogue is its bottleneck: statistics there cost more than the separate read, measured)
                if (pi < 0 || pl->ops[pi].kind != Op::CONV_TC || pl->ops[pi].qkv || pl->vt[t].c % 4 != 0) continue;
                if (pl->vt[t].stats_t < 0) {
                    const int fold = pl->ops[pi].fold > 1 ? pl->ops[pi].fold : 1;
                    const int rows = conv_tc_stats_rows_bound(pl->vt[t].h, pl->vt[t].w / fold);
                    const int len = rows * 2 * pl->vt[t].c * fold;
                    pl->vt[t].stats_fold = fold;
                    const int id = pb.new_tensor(B, 1, 1, len, len);
                    pl->vt[id].def = pi; pl->vt[id].last = gi;
                    pl->vt[t].stats_t = id; pl->ops[pi].stats = id;
                } else {
                    pl->vt[pl->vt[t].stats_t].last = std::max(pl->vt[pl->vt[t].stats_t].last, gi);
                }
            }
        }
    }
    IPDM_REQUIRE(pl->ops[0].kind == Op::CONV_DIRECT && pl->ops.back().kind == Op::CONV_DIRECT,
                 "unet: first and last convolutions must be on the direct path (in/out channels too wide)");

    // ---- arena assignment: first-fit over tensors sorted by definition, freeing at last use ----
    struct Blk { size_t off, size; };
    std::vector<Blk> free_list; size_t top = 0;
    std::vector<std::vector<int>> expire(pl->ops.size() + 1);
    std::vector<int> order;
    for (int i = 0; i < (int)pl->vt.size(); ++i) if (!pl->vt[i].external && pl->vt[i].def >= 0) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return pl->vt[a].def < pl->vt[b].def; });
    size_t oi = 0;
    auto bytes_of = [&](const VTensor& t) { return ((size_t)t.n * t.h * t.w * t.cs * (t.bf16 ? 2 : 4) + 1023) & ~(size_t)1023; };
    for (int op = 0; op < (int)pl->ops.size(); ++op) {
        while (oi < order.size() && pl->vt[order[oi]].def == op) {
            VTensor& t = pl->vt[order[oi]];
            const size_t need = bytes_of(t);
            int best = -1;
            for (int f = 0; f < (int)free_list.size(); ++f)
                if (free_list[f].size >= need && (best < 0 || free_list[f].size < free_list[best].size)) best = f;
            if (best >= 0) {
                t.off = free_list[best].off;
                free_list[best].off += need; free_list[best].size -= need;
                if (free_list[best].size == 0) free_list.erase(free_list.begin() + best);
            } else { t.off = top; top += need; }
            expire[t.last].push_back(order[oi]);
            ++oi;
        }
        for (int id : expire[op]) {          // release after this op; merge neighbours
            const VTensor& t = pl->vt[id];
            free_list.push_back({t.off, bytes_of(t)});
            std::sort(free_list.begin(), free_list.end(), [](const Blk& a, const Blk& b) { return a.off < b.off; });
            for (size_t f = 0; f + 1 < free_list.size();)
                if (free_list[f].off + free_list[f].size == free_list[f + 1].off) { free_list[f].size += free_list[f + 1].size; free_list.erase(free_list.begin() + f + 1); }
                else ++f;
        }
    }
    pl->arena_bytes = top;
    IPDM_CHECK_CUDA(cudaMalloc(&pl->arena, std::max<size_t>(top, 1024)));
    IPDM_CHECK_CUDA(cudaMemset(pl->arena, 0, std::max<size_t>(top, 1024)));
    const size_t norm_floats = (size_t)pb.norm_slots * 2 * B * pb.max_c;
    IPDM_CHECK_CUDA(cudaMalloc(&pl->norm_buf, std::max<size_t>(norm_floats, 1) * sizeof(float)));
    IPDM_CHECK_CUDA(cudaMalloc(&pl->gn_partials, (size_t)B * GN_MAX_BLOCKS * pb.max_c * 2 * sizeof(double)));

    // ---- materialise descriptors ----
    for (Op& o : pl->ops) {
        float* nscale = o.norm_slot >= 0 ? pl->norm_buf + (size_t)o.norm_slot * 2 * B * pb.max_c : nullptr;
        float* nshift = nscale ? nscale + (size_t)B * pb.max_c : nullptr;
        switch (o.kind) {
            case Op::GN_STATS:
            case Op::GN_APPLY: {
                GroupNormDesc& g = o.gd;
                g.nsrc = o.nsrc; g.src[0] = resolve(*pl, o.src[0]); if (o.nsrc > 1) g.src[1] = resolve(*pl, o.src[1]);
                g.groups = o.gn->groups; g.gamma = o.gn->gamma; g.beta = o.gn->beta; g.scale = nscale; g.shift = nshift; g.partials = pl->gn_partials;
                for (int sidx = 0; sidx < o.nsrc; ++sidx) {
                    const VTensor& v = pl->vt[o.src[sidx]];
                    if (v.stats_t >= 0 && v.stats_rows > 0) { g.tile_stats[sidx] = resolve(*pl, v.stats_t).p; g.tile_rows[sidx] = v.stats_rows; g.tile_fold[sidx] = v.stats_fold; }
                }
                if (o.kind == Op::GN_APPLY) o.out_t = resolve(*pl, o.dst);
            } break;
            case Op::CONV_TC: {
                ConvTcDesc d;
                d.nsrc = o.nsrc; d.n_ident = o.n_ident;
                for (int sidx = 0; sidx < o.nsrc; ++sidx) d.src[sidx] = resolve(*pl, o.src[sidx]);
                d.ntaps = o.cw->k * o.cw->k; d.stride = o.stride; d.cout = o.cw->cout; d.w_packed = o.cw->w_dev; d.w_k = o.cw->kpad;
                d.w_packed_lo = o.cw->w_dev_lo;
                d.bias = o.bias; d.bias_t_stride = o.bias_t_stride; d.t_dev = o.use_t ? net->t_dev : nullptr;
                if (o.res >= 0) d.res = resolve(*pl, o.res);
                d.out = resolve(*pl, o.dst);
                if (o.qkv) {
                    const TensorNHWC v = resolve(*pl, o.aux);
                    d.qkv_mode = 1; d.vt = v.p; d.heads = net->heads; d.head_dim = o.cw->cin / net->heads; d.t_pad = v.c / o.cw->cin;
                    if (o.aux2 >= 0) { d.out_lo = resolve(*pl, o.aux2).p; d.vt_lo = resolve(*pl, o.aux3).p; }
                    d.qkv_bf16 = v.bf16;
                }
                if (o.stats >= 0) d.stats_out = resolve(*pl, o.stats).p;
                if (o.gn) { d.norm_scale = nscale; d.norm_shift = nshift; d.act_silu = o.act; d.w_bf16 = o.cw->bf16; }
                if (o.fold) fold_desc(d, *o.cw);
                if (o.phase) { d.phase_up = 1; d.w_packed = o.cw->w_dev_phase; d.w_packed_lo = nullptr; d.w_k = o.cw->phase_k; d.w_bf16 = o.cw->bf16; }
                IPDM_CHECK(conv_tc_prepare(o.tcp, d));
                pl->vt[o.dst].stats_rows = o.tcp.stats_out ? o.tcp.stats_rows : 0;
                o.flops = 2.0 * B * o.tcp.H * o.tcp.W * (o.fold ? o.fold : 1) * (double)o.cw->cin * o.cw->cout * d.ntaps;
                if (o.phase) o.flops *= 4.0;                         // (reference FLOPs: a 3x3 conv on the upsampled image)
                if (o.n_ident)                                       // conv2 (3x3 over C_out channels) + shortcut (1x1 over the block input)
                    o.flops = 2.0 * B * o.tcp.H * o.tcp.W * (o.fold ? o.fold : 1) * (double)o.cw->cout * (9.0 * o.cw->cout + (o.cw->cin - o.cw->cout));
            } break;
            case Op::CONV_THIN: {
                ConvThinDesc d;
                d.src = resolve(*pl, o.src[0]); d.ntaps = o.cw->k * o.cw->k; d.cout = o.cw->cout; d.w_packed = o.cw->w_dev;
                d.bias = o.bias; d.bias_t_stride = o.bias_t_stride; d.t_dev = o.use_t ? net->t_dev : nullptr;
                if (o.res >= 0) d.res = resolve(*pl, o.res);
                d.out = resolve(*pl, o.dst);
                if (o.gn) { d.norm_scale = nscale; d.norm_shift = nshift; d.act_silu = o.act; }
                IPDM_CHECK(conv_thin_prepare(o.thp, d));
                o.flops = 2.0 * B * d.out.h * d.out.w * (double)o.cw->cin * o.cw->cout * d.ntaps;
            } break;
            case Op::CONV_DIRECT: {
                ConvDirectDesc& d = o.cd;
                d.nsrc = o.nsrc; d.src[0] = resolve(*pl, o.src[0]); if (o.nsrc > 1) d.src[1] = resolve(*pl, o.src[1]);
                d.norm_scale = o.gn ? nscale : nullptr; d.norm_shift = o.gn ? nshift : nullptr;
                d.ksize = o.cw->k; d.stride = o.stride; d.upsample = o.upsample; d.cin = o.cw->cin; d.cout = o.cw->cout;
                d.w = o.cw->thin ? o.cw->w_dev_direct : o.cw->w_dev;
                d.bias = o.bias; d.bias_t_stride = o.bias_t_stride; d.t_dev = o.use_t ? net->t_dev : nullptr;
                if (o.res >= 0) d.res = resolve(*pl, o.res);
                d.out = resolve(*pl, o.dst);
                o.flops = conv_direct_flops(d);
            } break;
            case Op::UPSAMPLE: o.src_t = resolve(*pl, o.src[0]); o.out_t = resolve(*pl, o.dst); break;
            case Op::ATTN: {
                const TensorNHWC qk = resolve(*pl, o.src[0]), v = resolve(*pl, o.aux), ot = resolve(*pl, o.dst);
                AttentionDesc& a = o.ad;
                a.qk = qk.p; a.vt = v.p; a.out = ot.p; a.batch = B; a.T = qk.h * qk.w; a.C = ot.c; a.heads = net->heads; a.head_dim = a.C / a.heads;
                a.t_pad = v.c / a.C; a.bf16 = v.bf16;
                if (o.aux2 >= 0) { a.qk_lo = resolve(*pl, o.aux2).p; a.vt_lo = resolve(*pl, o.aux3).p; }
                IPDM_CHECK(attention_prepare(o.ap, a));
                o.flops = attention_flops(a);
            } break;
        }
        pl->flops += o.flops;
    }
    *out = pl.get();
    net->plans[std::make_tuple(B, H, W)] = std::move(pl);
    return IPDM_OK;
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

static int run_plan(ipdm_unet* net, Plan* pl, const float* x, int t, float* eps, cudaStream_t st) {
    set_int_kernel<<<1, 1, 0, st>>>(net->t_dev, t);
    count_launch();
    // IPDM_OP_TRACE=<file>: per-op CUDA-event timings of every forward, appended as text (tools/op_trace.py); not for timed runs
    static const char* trace_path = getenv("IPDM_OP_TRACE");
    std::vector<cudaEvent_t> ev;
    if (trace_path) {
        ev.resize(pl->ops.size() + 1);
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
    }
    for (size_t i = 0; i < pl->ops.size(); ++i) {
        Op& o = pl->ops[i];
        struct Rec { cudaEvent_t* e; cudaStream_t s; ~Rec() { if (e) cudaEventRecord(*e, s); } } rec{trace_path ? &ev[i + 1] : nullptr, st};
        switch (o.kind) {
            case Op::GN_STATS: IPDM_CHECK(groupnorm_stats_launch(o.gd, st)); break;
            case Op::GN_APPLY: IPDM_CHECK(groupnorm_apply_launch(o.gd, o.out_t, o.act, net->precision != IPDM_PREC_FP32, st)); break;
            case Op::CONV_TC: IPDM_CHECK(conv_tc_launch(o.tcp, st)); break;
            case Op::CONV_THIN: IPDM_CHECK(conv_thin_launch(o.thp, st)); break;
            case Op::CONV_DIRECT: {
                ConvDirectDesc d = o.cd;
                if ((int)i == pl->first_op) d.src[0].p = const_cast<float*>(x);
                if ((int)i == pl->last_op) d.out.p = eps;
                IPDM_CHECK(conv_direct_launch(d, st));
            } break;
            case Op::UPSAMPLE: IPDM_CHECK(upsample_nearest_launch(o.src_t, o.out_t, net->precision != IPDM_PREC_FP32, st)); break;
            case Op::ATTN: IPDM_CHECK(attention_launch(o.ap, st)); break;
        }
    }
    if (trace_path) {
        cudaStreamSynchronize(st);
        if (FILE* f = fopen(trace_path, "a")) {
            static const char* names[] = {"gn_stats", "gn_apply", "conv_tc", "conv_direct", "conv_thin", "upsample", "attention"};
            fprintf(f, "# forward B=%d\n", pl->B);
            for (size_t i = 0; i < pl->ops.size(); ++i) {
                const Op& o = pl->ops[i];
                float ms = 0.f; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
                const VTensor& s0 = pl->vt[o.src[0]];
                const VTensor* d = o.dst >= 0 ? &pl->vt[o.dst] : nullptr;
                fprintf(f, "%s src %dx%dx%d(+%d) dst %dx%dx%d k%d s%d %.1f us\n", names[(int)o.kind], s0.h, s0.w, s0.c, o.nsrc > 1 ? pl->vt[o.src[1]].c : 0,
                        d ? d->h : 0, d ? d->w : 0, d ? d->c : 0, o.cw ? o.cw->k : 0, o.stride, ms * 1e3f);
            }
            fclose(f);
        }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return IPDM_OK;
}

static int get_plan(ipdm_unet* net, int B, int H, int W, Plan** out) {
    auto it = net->plans.find(std::make_tuple(B, H, W));
    if (it != net->plans.end()) { *out = it->second.get(); return IPDM_OK; }
    return build_plan(net, B, H, W, out);
}

}  // namespace ipdm

extern "C" long long ipdm_unet_param_count(const ipdm_unet_config* cfg) {
    if (!cfg || cfg->n_mult < 2 || cfg->n_mult > 8) return -1;
    CountVisitor cv; walk_arch(*cfg, cv); return cv.n;
}

extern "C" int ipdm_unet_create(ipdm_unet** out, const ipdm_unet_config* cfg, const float* weights_host, size_t n_weights) {
    IPDM_REQUIRE(out && cfg && weights_host, "ipdm_unet_create: null argument");
    IPDM_REQUIRE(cfg->n_mult >= 2 && cfg->n_mult <= 8 && cfg->n_attn >= 0 && cfg->n_attn <= 8, "ipdm_unet_create: bad config");
    IPDM_REQUIRE(cfg->precision == IPDM_PREC_TF32 || cfg->precision == IPDM_PREC_FP32 || cfg->precision == IPDM_PREC_BF16,
                 "ipdm_unet_create: unknown precision %d", cfg->precision);
    const long long expect = ipdm_unet_param_count(cfg);
    IPDM_REQUIRE((long long)n_weights == expect, "ipdm_unet_create: got %zu weights, the config needs %lld", n_weights, expect);
    std::unique_ptr<ipdm_unet> net(new ipdm_unet());
    net->cfg = *cfg; net->heads = cfg->num_heads; net->precision = cfg->precision;
    if (net->cfg.max_t <= 0) net->cfg.max_t = 64;
    IPDM_CHECK_CUDA(cudaMalloc(&net->t_dev, sizeof(int)));
    Reader r{weights_host, n_weights};
    // conservative reserves so that references into the vectors stay valid while blocks are appended
    net->convs.reserve(256); net->res.reserve(256); net->attn.reserve(256); net->down_blocks.reserve(256); net->up_blocks.reserve(256);
    BuildVisitor bv(net.get(), r);
    walk_arch(net->cfg, bv);
    IPDM_CHECK(bv.rc);
    IPDM_REQUIRE(r.ok && r.pos == n_weights, "ipdm_unet_create: weight buffer walk ended at %zu of %zu", r.pos, n_weights);
    IPDM_CHECK(build_time_tables(net.get(), bv));
    *out = net.release();
    return IPDM_OK;
}

extern "C" int ipdm_unet_destroy(ipdm_unet* net) { delete net; return IPDM_OK; }

extern "C" int ipdm_unet_forward(ipdm_unet* net, const float* x, int t, float* eps, int batch, int h, int w, void* stream) {
    IPDM_REQUIRE(net && x && eps && batch > 0 && h > 0 && w > 0, "ipdm_unet_forward: bad arguments");
    IPDM_REQUIRE(t >= 0 && t < net->cfg.max_t, "ipdm_unet_forward: timestep %d outside the precomputed range [0, %d)", t, net->cfg.max_t);
    Plan* pl = nullptr;
    IPDM_CHECK(get_plan(net, batch, h, w, &pl));
    return run_plan(net, pl, x, t, eps, (cudaStream_t)stream);
}

extern "C" double ipdm_unet_flops(const ipdm_unet* cnet, int batch, int h, int w) {
    ipdm_unet* net = const_cast<ipdm_unet*>(cnet);
    Plan* pl = nullptr;
    if (!net || get_plan(net, batch, h, w, &pl) != IPDM_OK) return -1.0;
    return pl->flops;
}

// ------------------------------------------------------------------------------------------------
// single-op entry points for the per-kernel parity tests (tests/test_unet_kernels_gpu.py)
// ------------------------------------------------------------------------------------------------
static TensorNHWC mk(const float* p, int n, int h, int w, int c, int cs) {
    TensorNHWC t; t.p = const_cast<float*>(p); t.n = n; t.h = h; t.w = w; t.c = c; t.cs = cs; return t;
}

extern "C" int ipdm_debug_conv(const float* src0, int c0, int cs0, const float* src1, int c1, int cs1, int n, int h, int w,
                               const float* w_host, const float* bias_host, int cout, int k, int stride, int upsample_h,
                               int upsample_w, const float* norm_scale, const float* norm_shift, const float* res, int res_cs,
                               float* out, int out_cs, int use_tc_arg, void* stream) {
    IPDM_REQUIRE(src0 && w_host && out, "ipdm_debug_conv: null argument");
    const int use_tc = use_tc_arg & 0xff, variant = use_tc_arg >> 8;      // bits 8..: kernel variant of the tensor-core path (0 auto)
    ipdm_unet holder;
    holder.precision = use_tc == 2 ? IPDM_PREC_FP32 : (use_tc == 3 ? IPDM_PREC_BF16 : IPDM_PREC_TF32);
    ConvW cw; cw.cin = c0 + c1; cw.cout = cout; cw.k = k;
    cw.w_host.assign(w_host, w_host + (size_t)cout * cw.cin * k * k);
    if (bias_host) cw.b_host.assign(bias_host, bias_host + cout);
    if (use_tc == 7 || use_tc == 8) {                    // Upsample(2x nearest) + conv3x3 as four phases on the low-res source (7 bf16, 8 tf32)
        IPDM_REQUIRE(stride == 1 && upsample_h == 2 * h && upsample_w == 2 * w && c1 == 0 && cs0 % 32 == 0 && k == 3 && !res && !norm_scale,
                     "ipdm_debug_conv: the phase path is an exact 2x upsample + 3x3 conv of one padded source");
        holder.precision = use_tc == 7 ? IPDM_PREC_BF16 : IPDM_PREC_TF32;
        cw.tc = true; cw.bf16 = use_tc == 7;
        IPDM_CHECK(pack_phase(&holder, cw));
        IPDM_REQUIRE(cw.w_dev_phase, "ipdm_debug_conv: shape is not eligible for the phase path");
        ConvTcDesc d; d.nsrc = 1; d.src[0] = mk(src0, n, h, w, c0, cs0);
        d.ntaps = 9; d.stride = 1; d.cout = cout; d.bias = cw.b_dev;
        d.out = mk(out, n, 2 * h, 2 * w, cout, out_cs);
        d.phase_up = 1; d.w_packed = cw.w_dev_phase; d.w_k = cw.phase_k; d.w_bf16 = cw.bf16;
        float* stats = nullptr;
        IPDM_CHECK_CUDA(cudaMalloc(&stats, (size_t)n * conv_tc_stats_rows_bound(2 * h, 2 * w) * 2 * cout * sizeof(float)));
        d.stats_out = stats;
        ConvTcParams P;
        int rc2 = conv_tc_prepare(P, d);
        if (rc2 == IPDM_OK) rc2 = conv_tc_launch(P, (cudaStream_t)stream);
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaFree(stats);
        return rc2;
    }
    if (use_tc == 6) {                                   // width-folded tensor-core path: dense thin sources, optional fused GroupNorm + SiLU
        IPDM_REQUIRE(stride == 1 && upsample_h == 0 && cs0 == c0 && (c1 == 0 || cs1 == c1) && out_cs == cout && (!res || res_cs == cout),
                     "ipdm_debug_conv: the folded path takes dense tensors, stride 1");
        holder.force_fold = true;
        IPDM_CHECK(pack_fold(&holder, cw, c0, c1));
        IPDM_REQUIRE(cw.fold && w % cw.fold == 0, "ipdm_debug_conv: shape is not eligible for the folded path");
        ConvTcDesc d; d.nsrc = c1 ? 2 : 1; d.src[0] = mk(src0, n, h, w, c0, cs0); if (c1) d.src[1] = mk(src1, n, h, w, c1, cs1);
        if (norm_scale) { d.norm_scale = norm_scale; d.norm_shift = norm_shift; }
        d.ntaps = k * k; d.stride = 1; d.bias = cw.b_dev;
        if (res) d.res = mk(res, n, h, w, cout, res_cs);
        d.out = mk(out, n, h, w, cout, out_cs);
        fold_desc(d, cw);
        float* stats = nullptr;                          // exercise the statistics epilogue too
        IPDM_CHECK_CUDA(cudaMalloc(&stats, (size_t)n * conv_tc_stats_rows_bound(h, w / cw.fold) * 2 * d.cout * sizeof(float)));
        d.stats_out = stats;
        ConvTcParams P;
        int rc2 = conv_tc_prepare(P, d);
        if (rc2 == IPDM_OK) rc2 = conv_tc_launch(P, (cudaStream_t)stream);
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaFree(stats);
        return rc2;
    }
    if (use_tc == 4) {                                   // thin tensor-core path: src0 is the operand tensor (channel stride cs0 in {8,16,32})
        IPDM_REQUIRE(c1 == 0 && stride == 1 && upsample_h == 0 && !norm_scale, "ipdm_debug_conv: thin path takes one plain source");
        holder.force_thin = true;
        IPDM_CHECK(pack_conv(&holder, cw, c0, 0, false));
        IPDM_REQUIRE(cw.thin && cw.thin_cs == cs0, "ipdm_debug_conv: shape is not eligible for the thin path (stride %d expected)", cw.thin_cs);
        ConvThinDesc d; d.src = mk(src0, n, h, w, c0, cs0); d.ntaps = k * k; d.cout = cout; d.w_packed = cw.w_dev; d.bias = cw.b_dev;
        if (res) d.res = mk(res, n, h, w, cout, res_cs);
        d.out = mk(out, n, h, w, cout, out_cs);
        ConvThinParams TP;
        IPDM_CHECK(conv_thin_prepare(TP, d));
        int rc2 = conv_thin_launch(TP, (cudaStream_t)stream);
        cudaStreamSynchronize((cudaStream_t)stream);
        return rc2;
    }
    IPDM_CHECK(pack_conv(&holder, cw, c0, c1, true, use_tc != 0, use_tc == 3));
    cudaStream_t st = (cudaStream_t)stream;
    const int hin = upsample_h > 0 ? upsample_h : h, win = upsample_w > 0 ? upsample_w : w;
    const int ho = stride == 1 ? hin : (hin + 1) / 2, wo = stride == 1 ? win : (win + 1) / 2;
    int rc;
    if (use_tc) {
        if (norm_scale) IPDM_REQUIRE(cs0 % 32 == 0 && cs1 % 32 == 0 && cs0 + cs1 <= cw.kpad && (c1 == 0 || c0 == cs0), "ipdm_debug_conv: fused sources need channel strides that are multiples of 32");
        else IPDM_REQUIRE(cs0 == cw.cs0 && (c1 == 0 || cs1 == cw.cs1), "ipdm_debug_conv: tensor-core sources need channel strides %d / %d", cw.cs0, cw.cs1);
        IPDM_REQUIRE(upsample_h == 0, "ipdm_debug_conv: upsample is a direct-path feature");
        ConvTcDesc d; d.nsrc = c1 ? 2 : 1; d.src[0] = mk(src0, n, h, w, c0, cs0); if (c1) d.src[1] = mk(src1, n, h, w, c1, cs1);
        d.src[0].bf16 = cw.bf16 && !norm_scale;
        if (norm_scale) { d.norm_scale = norm_scale; d.norm_shift = norm_shift; d.w_bf16 = cw.bf16; }       // fused GroupNorm + SiLU: raw fp32 sources
        d.ntaps = k * k; d.stride = stride; d.cout = cout; d.w_packed = cw.w_dev; d.w_packed_lo = cw.w_dev_lo; d.w_k = cw.kpad; d.bias = cw.b_dev;
        if (res) d.res = mk(res, n, ho, wo, cout, res_cs);
        d.out = mk(out, n, ho, wo, cout, out_cs);
        d.variant = variant;
        ConvTcParams P;
        IPDM_CHECK(conv_tc_prepare(P, d));
        rc = conv_tc_launch(P, st);
    } else {
        ConvDirectDesc d; d.nsrc = c1 ? 2 : 1; d.src[0] = mk(src0, n, h, w, c0, cs0); if (c1) d.src[1] = mk(src1, n, h, w, c1, cs1);
        d.norm_scale = norm_scale; d.norm_shift = norm_shift; d.ksize = k; d.stride = stride; d.upsample = upsample_h > 0;
        d.cin = cw.cin; d.cout = cout; d.w = cw.w_dev; d.bias = cw.b_dev;
        if (res) d.res = mk(res, n, ho, wo, cout, res_cs);
        d.out = mk(out, n, ho, wo, cout, out_cs);
        rc = conv_direct_launch(d, st);
    }
    cudaStreamSynchronize(st);      // the packed weights are freed when `holder` goes out of scope
    return rc;
}

// times `iters` launches of one tensor-core conv (weights random, data whatever is in the buffers); ms_out = average per launch
// times `iters` launches of one tensor-core conv (weights random, data whatever is in the buffers); ms_out = average per launch.
// variant 5 = GroupNorm-fused persistent halo kernel (raw fp32 sources, scale = 1 / shift = 0 per channel)
extern "C" int ipdm_debug_conv_time(int c0, int c1, int n, int h, int w, int cout, int k, int stride, int mode, int variant,
                                    int with_res, int iters, float* ms_out, double* flops_out) {
    ipdm_unet holder;
    holder.precision = mode == 2 ? IPDM_PREC_FP32 : (mode == 3 ? IPDM_PREC_BF16 : IPDM_PREC_TF32);
    if (variant == 6 || variant == 7) {                  // width-folded thin layer, with (6) / without (7) the fused GroupNorm + SiLU
        holder.force_fold = true;
        ConvW cw; cw.cin = c0 + c1; cw.cout = cout; cw.k = k;
        cw.w_host.assign((size_t)cout * cw.cin * k * k, 0.01f);
        cw.b_host.assign(cout, 0.1f);
        IPDM_CHECK(pack_fold(&holder, cw, c0, c1));
        IPDM_REQUIRE(cw.fold && w % cw.fold == 0 && stride == 1, "ipdm_debug_conv_time: shape is not eligible for the folded path");
        float *s0 = nullptr, *s1 = nullptr, *out = nullptr, *res = nullptr, *nsc = nullptr, *nsh = nullptr, *stats = nullptr;
        const size_t px = (size_t)n * h * w;
        IPDM_CHECK_CUDA(cudaMalloc(&s0, px * c0 * 4)); IPDM_CHECK_CUDA(cudaMemset(s0, 0, px * c0 * 4));
        if (c1) { IPDM_CHECK_CUDA(cudaMalloc(&s1, px * c1 * 4)); IPDM_CHECK_CUDA(cudaMemset(s1, 0, px * c1 * 4)); }
        IPDM_CHECK_CUDA(cudaMalloc(&out, px * cout * 4));
        if (with_res) { IPDM_CHECK_CUDA(cudaMalloc(&res, px * cout * 4)); IPDM_CHECK_CUDA(cudaMemset(res, 0, px * cout * 4)); }
        ConvTcDesc d; d.nsrc = c1 ? 2 : 1; d.src[0] = mk(s0, n, h, w, c0, c0); if (c1) d.src[1] = mk(s1, n, h, w, c1, c1);
        if (variant == 6) {
            std::vector<float> one((size_t)n * cw.cin, 1.f);
            IPDM_CHECK_CUDA(cudaMalloc(&nsc, one.size() * 4)); IPDM_CHECK_CUDA(cudaMalloc(&nsh, one.size() * 4));
            IPDM_CHECK_CUDA(cudaMemcpy(nsc, one.data(), one.size() * 4, cudaMemcpyHostToDevice));
            IPDM_CHECK_CUDA(cudaMemset(nsh, 0, one.size() * 4));
            d.norm_scale = nsc; d.norm_shift = nsh;
        }
        d.ntaps = k * k; d.stride = 1; d.bias = cw.b_dev;
        if (res) d.res = mk(res, n, h, w, cout, cout);
        d.out = mk(out, n, h, w, cout, cout);
        fold_desc(d, cw);
        IPDM_CHECK_CUDA(cudaMalloc(&stats, (size_t)n * conv_tc_stats_rows_bound(h, w / cw.fold) * 2 * d.cout * sizeof(float)));
        d.stats_out = stats;
        ConvTcParams P;
        IPDM_CHECK(conv_tc_prepare(P, d));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int i = 0; i < 3; ++i) IPDM_CHECK(conv_tc_launch(P, nullptr));
        cudaEventRecord(a, nullptr);
        for (int i = 0; i < iters; ++i) IPDM_CHECK(conv_tc_launch(P, nullptr));
        cudaEventRecord(b, nullptr);
        IPDM_CHECK_CUDA(cudaDeviceSynchronize());
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        *ms_out = ms / iters;
        if (flops_out) *flops_out = 2.0 * px * (double)cw.cin * cout * k * k;
        cudaFree(s0); cudaFree(s1); cudaFree(out); cudaFree(res); cudaFree(nsc); cudaFree(nsh); cudaFree(stats); cudaEventDestroy(a); cudaEventDestroy(b);
        return IPDM_OK;
    }
    const bool fused = variant == 5;
    ConvW cw; cw.cin = c0 + c1; cw.cout = cout; cw.k = k;
    cw.w_host.assign((size_t)cout * cw.cin * k * k, 0.01f);
    cw.b_host.assign(cout, 0.1f);
    IPDM_CHECK(pack_conv(&holder, cw, c0, c1, true, 1, mode == 3));
    const int scs0 = fused ? round_up(c0, 32) : cw.cs0, scs1 = fused ? (c1 ? round_up(c1, 32) : 0) : cw.cs1;
    const int sc0 = fused ? c0 : cw.c0, sc1 = fused ? c1 : cw.c1;
    const int eb = (cw.bf16 && !fused) ? 2 : 4;
    const int ho = stride == 1 ? h : (h + 1) / 2, wo = stride == 1 ? w : (w + 1) / 2;
    float *s0 = nullptr, *s1 = nullptr, *out = nullptr, *res = nullptr, *nsc = nullptr, *nsh = nullptr;
    IPDM_CHECK_CUDA(cudaMalloc(&s0, (size_t)n * h * w * scs0 * eb));
    IPDM_CHECK_CUDA(cudaMemset(s0, 0, (size_t)n * h * w * scs0 * eb));
    if (scs1) { IPDM_CHECK_CUDA(cudaMalloc(&s1, (size_t)n * h * w * scs1 * 4)); IPDM_CHECK_CUDA(cudaMemset(s1, 0, (size_t)n * h * w * scs1 * 4)); }
    const int ocs = alloc_cs(cout);
    IPDM_CHECK_CUDA(cudaMalloc(&out, (size_t)n * ho * wo * ocs * 4));
    if (with_res) { IPDM_CHECK_CUDA(cudaMalloc(&res, (size_t)n * ho * wo * ocs * 4)); IPDM_CHECK_CUDA(cudaMemset(res, 0, (size_t)n * ho * wo * ocs * 4)); }
    ConvTcDesc d; d.nsrc = scs1 ? 2 : 1; d.src[0] = mk(s0, n, h, w, sc0, scs0); d.src[0].bf16 = cw.bf16 && !fused;
    if (scs1) d.src[1] = mk(s1, n, h, w, sc1, scs1);
    if (fused) {
        std::vector<float> one((size_t)n * cw.cin, 1.f);
        IPDM_CHECK_CUDA(cudaMalloc(&nsc, one.size() * 4)); IPDM_CHECK_CUDA(cudaMalloc(&nsh, one.size() * 4));
        IPDM_CHECK_CUDA(cudaMemcpy(nsc, one.data(), one.size() * 4, cudaMemcpyHostToDevice));
        IPDM_CHECK_CUDA(cudaMemset(nsh, 0, one.size() * 4));
        d.norm_scale = nsc; d.norm_shift = nsh; d.w_bf16 = cw.bf16;
    }
    d.ntaps = k * k; d.stride = stride; d.cout = cout; d.w_packed = cw.w_dev; d.w_packed_lo = cw.w_dev_lo; d.w_k = cw.kpad; d.bias = cw.b_dev;
    if (res) d.res = mk(res, n, ho, wo, cout, ocs);
    d.out = mk(out, n, ho, wo, cout, ocs);
    d.variant = fused ? 0 : variant;
    float* tstats = nullptr;                             // IPDM_TIME_STATS=1: with the GroupNorm statistics epilogue, as inside the network
    if (getenv("IPDM_TIME_STATS") && atoi(getenv("IPDM_TIME_STATS")) == 1) {
        IPDM_CHECK_CUDA(cudaMalloc(&tstats, (size_t)n * conv_tc_stats_rows_bound(ho, wo) * 2 * cout * sizeof(float)));
        d.stats_out = tstats;
    }
    ConvTcParams P;
    IPDM_CHECK(conv_tc_prepare(P, d));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) IPDM_CHECK(conv_tc_launch(P, nullptr));
    cudaEventRecord(a, nullptr);
    for (int i = 0; i < iters; ++i) IPDM_CHECK(conv_tc_launch(P, nullptr));
    cudaEventRecord(b, nullptr);
    IPDM_CHECK_CUDA(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / iters;
    if (flops_out) *flops_out = 2.0 * n * ho * wo * (double)cw.cin * cout * k * k;
    cudaFree(s0); cudaFree(s1); cudaFree(out); cudaFree(res); cudaFree(nsc); cudaFree(nsh); cudaFree(tstats); cudaEventDestroy(a); cudaEventDestroy(b);
    return IPDM_OK;
}

// Host-only views of the two weight transformations of round 2, for the CPU tests (tests/test_host.py): no device is touched.
// out == nullptr: only the dimensions are returned.
extern "C" int ipdm_debug_fold_pack(const float* w_host, int cout, int c0, int c1, int k, float* out, unsigned long long* masks_out,
                                    int* fold_out, int* n_out, int* k_out) {
    IPDM_REQUIRE(w_host && fold_out && n_out && k_out && (k == 1 || k == 3) && c0 > 0 && c1 >= 0 && cout > 0, "ipdm_debug_fold_pack: bad arguments");
    ConvW cw; cw.cin = c0 + c1; cw.cout = cout; cw.k = k;
    cw.w_host.assign(w_host, w_host + (size_t)cout * cw.cin * k * k);
    const int cs[2] = {c0, c1}; const ConvW* ws[2] = {&cw, &cw}; const int off[2] = {0, c0};
    std::vector<float> p;
    const int f = fold_pack_host(cw, c1 ? 2 : 1, cs, ws, off, p);
    *fold_out = f; *n_out = f * cout; *k_out = f * cw.cin;
    if (f && out) memcpy(out, p.data(), p.size() * sizeof(float));
    if (f && masks_out) for (int t = 0; t < 9; ++t) masks_out[t] = cw.fold_mask[t];
    return IPDM_OK;
}
extern "C" int ipdm_debug_phase_pack(const float* w_host, int cout, int cin, float* out, int* k_out) {
    IPDM_REQUIRE(w_host && k_out && cout > 0 && cin > 0, "ipdm_debug_phase_pack: bad arguments");
    ConvW cw; cw.cin = cin; cw.cout = cout; cw.k = 3;
    cw.w_host.assign(w_host, w_host + (size_t)cout * cin * 9);
    *k_out = round_up(cin, 32);
    if (out) { std::vector<float> p; phase_pack_host(cw, p); memcpy(out, p.data(), p.size() * sizeof(float)); }
    return IPDM_OK;
}

extern "C" int ipdm_debug_groupnorm(const float* src0, int c0, int cs0, const float* src1, int c1, int cs1, int n, int h, int w,
                                    const float* gamma_host, const float* beta_host, int act_silu, float* scale_out,
                                    float* shift_out, float* out, int out_cs, void* stream) {
    IPDM_REQUIRE(src0 && gamma_host && beta_host && scale_out && shift_out, "ipdm_debug_groupnorm: null argument");
    ipdm_unet holder;
    const int C = c0 + c1;
    GNW g; g.C = C; g.groups = gn_groups(C);
    IPDM_CHECK(upload(&holder, std::vector<float>(gamma_host, gamma_host + C), &g.gamma));
    IPDM_CHECK(upload(&holder, std::vector<float>(beta_host, beta_host + C), &g.beta));
    double* partials = nullptr;
    IPDM_CHECK_CUDA(cudaMalloc(&partials, (size_t)n * GN_MAX_BLOCKS * C * 2 * sizeof(double)));
    GroupNormDesc d; d.nsrc = c1 ? 2 : 1; d.src[0] = mk(src0, n, h, w, c0, cs0); if (c1) d.src[1] = mk(src1, n, h, w, c1, cs1);
    d.groups = g.groups; d.gamma = g.gamma; d.beta = g.beta; d.scale = scale_out; d.shift = shift_out; d.partials = partials;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = groupnorm_stats_launch(d, st);
    if (rc == IPDM_OK && out) rc = groupnorm_apply_launch(d, mk(out, n, h, w, C, out_cs), act_silu, 1, st);
    cudaStreamSynchronize(st);
    cudaFree(partials);
    return rc;
}

extern "C" int ipdm_debug_attention(const float* qk, const float* vt, const float* qk_lo, const float* vt_lo, float* out, int batch, int T,
                                    int t_pad, int heads, int C, void* stream) {
    AttentionDesc a; a.qk = qk; a.vt = vt; a.qk_lo = qk_lo; a.vt_lo = vt_lo; a.out = out; a.batch = batch; a.T = T; a.t_pad = t_pad; a.heads = heads; a.C = C; a.head_dim = C / heads;
    AttentionParams P;
    IPDM_CHECK(attention_prepare(P, a));
    return attention_launch(P, (cudaStream_t)stream);
}

extern "C" int ipdm_debug_attention_bf16(const void* qk, const void* vt, float* out, int batch, int T, int t_pad, int heads, int C, void* stream) {
    AttentionDesc a; a.qk = (const float*)qk; a.vt = (const float*)vt; a.out = out; a.batch = batch; a.T = T; a.t_pad = t_pad; a.heads = heads; a.C = C;
    a.head_dim = C / heads; a.bf16 = 1;
    AttentionParams P;
    IPDM_CHECK(attention_prepare(P, a));
    return attention_launch(P, (cudaStream_t)stream);
}

extern "C" int ipdm_debug_upsample(const float* src, int n, int hs, int ws, int cs, float* dst, int hd, int wd, void* stream) {
    return upsample_nearest_launch(mk(src, n, hs, ws, cs, cs), mk(dst, n, hd, wd, cs, cs), 0, (cudaStream_t)stream);
}
