// Thin convolutions (C_in <= 32, C_out <= 16, 3x3 or 1x1, stride 1) of the full-resolution levels with the GroupNorm affine + SiLU
// of the reference's `norm -> SiLU -> conv` (Model/model.py:98-101, 110-113) applied while the input tile is staged, the GroupNorm
// statistics of the OUTPUT produced in the epilogue, and fp32-class accuracy (3xTF32).
//
// Why a second thin kernel (conv_thin.cu is the tcgen05 one): the 2000x912 / 1000x456 levels of the projection UNet move
// 4 * (C_in + C_out) bytes per pixel with C = 4 ... 32 and are memory-bound by design, but round 1 spent three passes on every
// `GroupNorm -> SiLU -> conv`: a statistics read, an apply pass that writes an operand tensor, and a tcgen05 kernel whose epilogue
// (TMEM -> registers -> staging -> global, one pixel of 8-16 channels per thread) is bound by instruction issue.  With M = 128-row
// UMMA tiles and N = 8-16 the 5th-generation tensor core is 14 % busy and its operands must be whole TF32 words (the
// 10-bit mantissa on these large-signal, small-detail layers was the dominant error of the bf16 mode).  Here:
//   * one CTA stages a (TH+2) x 34 pixel halo tile of the raw fp32 source(s) (virtual concat), applies scale/shift + SiLU, zeroes the
//     pixels outside the image (the conv pads the ACTIVATION), splits every value into TF32 hi + lo and stores both to shared memory;
//   * eight warps run warp-level m16n8k8 TF32 MMAs (mma.sync): warp w owns output rows, an m-tile is 16 pixels of one row, A fragments
//     are two 64-bit shared loads per k-step (the K index of the MMA is permuted so that a lane's two channels are adjacent), the
//     weights sit in shared memory as TF32 hi + lo; D += A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation in registers;
//   * the accumulator fragment of a lane is 2 adjacent channels of 2 pixels: +bias[t] (+residual) and 64-bit stores that cover whole
//     32-byte sectors, no staging; per-lane sums / sums of squares of what was stored accumulate in fp64 over the CTA's tiles and
//     are written once per warp as the `stats` rows gn_tile_reduce folds (ConvTcDesc::stats_out format).
// The legacy tensor path is more than enough here: 9 * 3 MMAs per 16 pixels for C = 8 -> 8.
#include "common.cuh"
#include "unet_ops.cuh"

#include <algorithm>
#include <cstdlib>

namespace ipdm {

constexpr int CW_THREADS = 256, CW_TW = 32;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// pixel pitch of the staged tile / weight-row pitch in floats: rows 0..3 of a fragment must start in four different 8-word bank groups
__host__ __device__ constexpr int cw_pitch(int k) { return (k == 8 || k == 24) ? k : k + 8; }

template <int K, int NT, int TAPS, int TH>
struct CwSmem {
    static constexpr int HALO = TAPS == 9 ? 1 : 0;
    static constexpr int PITCH = cw_pitch(K);
    static constexpr int TPIX = (TH + 2 * HALO) * (CW_TW + 2 * HALO);
    static constexpr int TILE_FLOATS = TPIX * PITCH;
    static constexpr int W_FLOATS = TAPS * NT * 8 * PITCH;
    static constexpr int OFF_LO = TILE_FLOATS;                    // tile lo
    static constexpr int OFF_WHI = 2 * TILE_FLOATS;
    static constexpr int OFF_WLO = OFF_WHI + W_FLOATS;
    static constexpr int OFF_NORM = OFF_WLO + W_FLOATS;           // scale[K], shift[K] of the current slice
    static constexpr int TOTAL_BYTES = (OFF_NORM + 2 * K) * 4;
};

// grid (ctas_per_slice, batch): a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... of ONE slice, so its statistics rows belong to it
template <int K, int NT, int TAPS, int TH>
__global__ void __launch_bounds__(CW_THREADS)
conv_warp_kernel(const ConvWarpParams P) {
    using S = CwSmem<K, NT, TAPS, TH>;
    constexpr int HALO = S::HALO, PITCH = S::PITCH, TWH = CW_TW + 2 * HALO, RPW = TH / 8;   // output rows per warp
    constexpr int G = K / 4;                                                                 // 4-channel groups per pixel
    extern __shared__ __align__(16) float cw_smem[];
    float* t_hi = cw_smem; float* t_lo = cw_smem + S::OFF_LO;
    float* w_hi = cw_smem + S::OFF_WHI; float* w_lo = cw_smem + S::OFF_WLO;
    float* s_norm = cw_smem + S::OFF_NORM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y;
    const int Ctot = P.c0 + P.c1;

    // weights (already TF32 hi / lo, [tap][NT*8][PITCH], zero padded) and this slice's scale / shift -> shared memory, once per CTA
    for (int i = tid; i < S::W_FLOATS / 4; i += CW_THREADS) {
        reinterpret_cast<float4*>(w_hi)[i] = __ldg(reinterpret_cast<const float4*>(P.w_hi) + i);
        reinterpret_cast<float4*>(w_lo)[i] = __ldg(reinterpret_cast<const float4*>(P.w_lo) + i);
    }
    for (int i = tid; i < K; i += CW_THREADS) {
        const bool real = i < Ctot;
        s_norm[i] = (real && P.nscale) ? __ldg(P.nscale + (size_t)b * Ctot + i) : (real ? 1.f : 0.f);
        s_norm[K + i] = (real && P.nshift) ? __ldg(P.nshift + (size_t)b * Ctot + i) : 0.f;
    }
    // bias (+ time embedding row) of this lane's channels: n-tile j -> channels 8j + 2t, 8j + 2t + 1
    float bia[NT][2];
    {
        const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) { const int c = 8 * j + 2 * t + e; bia[j][e] = (bias && c < P.cout) ? __ldg(bias + c) : 0.f; }
    }
    double ssum[NT][2], ssq[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) { ssum[j][0] = ssum[j][1] = ssq[j][0] = ssq[j][1] = 0.0; }

    const int tiles = P.tiles_x * P.tiles_y;
    const size_t slice_pix = (size_t)b * P.H * P.W;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int tyi = tile / P.tiles_x, txi = tile - tyi * P.tiles_x;
        const int x0 = txi * CW_TW, y0 = tyi * TH;
        __syncthreads();                                     // the previous tile's MMAs are done with the staged tile (and the weights are in)
        // ---- stage: raw source(s) -> scale/shift (+SiLU) -> TF32 hi / lo, zero outside the image.  Batches of U independent 128-bit
        //      global loads per thread are issued before the first value is used. ----
        constexpr int U = 4, ITEMS = S::TPIX * G;
        for (int i0 = tid; i0 < ITEMS; i0 += CW_THREADS * U) {
            float4 v[U]; bool in[U]; int pixv[U], cv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * CW_THREADS;
                const int pix = i / G, c = 4 * (i - pix * G);
                const int ly = pix / TWH, lx = pix - ly * TWH;
                const int iy = y0 - HALO + ly, ix = x0 - HALO + lx;
                pixv[u] = pix; cv[u] = c;
                in[u] = i < ITEMS && (unsigned)iy < (unsigned)P.H && (unsigned)ix < (unsigned)P.W && c < Ctot;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in[u]) {
                    const size_t sp = slice_pix + (size_t)iy * P.W + ix;
                    v[u] = c < P.c0 ? ld_stream(reinterpret_cast<const float4*>(P.src0 + sp * P.cs0 + c))
                                    : ld_stream(reinterpret_cast<const float4*>(P.src1 + sp * P.cs1 + (c - P.c0)));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * CW_THREADS >= ITEMS) continue;
                float4 w = v[u];
                if (in[u]) {
                    const float4 sc = *reinterpret_cast<const float4*>(s_norm + cv[u]), sh = *reinterpret_cast<const float4*>(s_norm + K + cv[u]);
                    w.x = fmaf(w.x, sc.x, sh.x); w.y = fmaf(w.y, sc.y, sh.y); w.z = fmaf(w.z, sc.z, sh.z); w.w = fmaf(w.w, sc.w, sh.w);
                    if (P.act_silu && !(P.dbg & 4)) { w.x = silu(w.x); w.y = silu(w.y); w.z = silu(w.z); w.w = silu(w.w); }
                }
                float4 h, l;
                h.x = tf32_rn(w.x); h.y = tf32_rn(w.y); h.z = tf32_rn(w.z); h.w = tf32_rn(w.w);
                l.x = tf32_rn(w.x - h.x); l.y = tf32_rn(w.y - h.y); l.z = tf32_rn(w.z - h.z); l.w = tf32_rn(w.w - h.w);
                *reinterpret_cast<float4*>(t_hi + pixv[u] * PITCH + cv[u]) = h;
                *reinterpret_cast<float4*>(t_lo + pixv[u] * PITCH + cv[u]) = l;
            }
        }
        __syncthreads();
        // ---- MMA: warp `warp` owns output rows warp*RPW .. +RPW-1, two 16-pixel m-tiles per row ----
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int oy = warp * RPW + rr;                   // output row inside the tile
            float acc[2][NT][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int j = 0; j < NT; ++j) { acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f; }
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int dy = TAPS == 9 ? tap / 3 : 0, dx = TAPS == 9 ? tap - (tap / 3) * 3 : 0;
#pragma unroll
                for (int ks = 0; ks < K / 8; ++ks) {
                    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {            // B fragment: b0 = W[ch 2t][n = g], b1 = W[ch 2t+1][g] of this k-step
                        const int wo = ((tap * NT + j) * 8 + g) * PITCH + ks * 8 + 2 * t;
                        const float2 hh = *reinterpret_cast<const float2*>(w_hi + wo), ll = *reinterpret_cast<const float2*>(w_lo + wo);
                        bh[j][0] = __float_as_uint(hh.x); bh[j][1] = __float_as_uint(hh.y);
                        bl[j][0] = __float_as_uint(ll.x); bl[j][1] = __float_as_uint(ll.y);
                    }
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        // A fragment: rows g / g+8 = pixels, (a0, a2) = channels (2t, 2t+1) of pixel g, (a1, a3) of pixel g+8
                        const int p0 = (oy + dy) * TWH + m * 16 + g + dx;
                        const int ao = p0 * PITCH + ks * 8 + 2 * t;
                        const float2 h0 = *reinterpret_cast<const float2*>(t_hi + ao), h1 = *reinterpret_cast<const float2*>(t_hi + ao + 8 * PITCH);
                        const float2 l0 = *reinterpret_cast<const float2*>(t_lo + ao), l1 = *reinterpret_cast<const float2*>(t_lo + ao + 8 * PITCH);
                        const uint32_t ah[4] = {__float_as_uint(h0.x), __float_as_uint(h1.x), __float_as_uint(h0.y), __float_as_uint(h1.y)};
                        const uint32_t al[4] = {__float_as_uint(l0.x), __float_as_uint(l1.x), __float_as_uint(l0.y), __float_as_uint(l1.y)};
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            if (!(P.dbg & 1)) {
                                mma_tf32(acc[m][j], al, bh[j][0], bh[j][1]);      // small terms first
                                mma_tf32(acc[m][j], ah, bl[j][0], bl[j][1]);
                            }
                            if (!(P.dbg & 2)) mma_tf32(acc[m][j], ah, bh[j][0], bh[j][1]);
                        }
                    }
                }
            }
            // ---- epilogue: c0,c1 = pixel g, channels 2t, 2t+1; c2,c3 = pixel g+8 ----
            const int py = y0 + oy;
            if (py < P.H) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int hrow = 0; hrow < 2; ++hrow) {
                        const int px = x0 + m * 16 + g + 8 * hrow;
                        if (px < P.W) {
                            const size_t op = slice_pix + (size_t)py * P.W + px;
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const int c = 8 * j + 2 * t;
                                if (c < P.cout) {
                                    float2 v = make_float2(acc[m][j][2 * hrow] + bia[j][0], acc[m][j][2 * hrow + 1] + bia[j][1]);
                                    if (P.res) { const float2 r = __ldg(reinterpret_cast<const float2*>(P.res + op * P.res_cs + c)); v.x += r.x; v.y += r.y; }
                                    *reinterpret_cast<float2*>(P.out + op * P.out_cs + c) = v;
                                    ssum[j][0] += (double)v.x; ssum[j][1] += (double)v.y;
                                    ssq[j][0] += (double)v.x * v.x; ssq[j][1] += (double)v.y * v.y;
                                } else if (c < P.out_cs) {
                                    *reinterpret_cast<float2*>(P.out + op * P.out_cs + c) = make_float2(0.f, 0.f);     // channel padding stays zero
                                }
                            }
                            for (int c = NT * 8 + 2 * t; c < P.out_cs; c += 8) *reinterpret_cast<float2*>(P.out + op * P.out_cs + c) = make_float2(0.f, 0.f);
                        }
                    }
            }
        }
    }
    // ---- GroupNorm statistics of the output: one [2][cout] row per warp (fixed order: deterministic) ----
    if (P.stats_out) {
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double s = ssum[j][e], q = ssq[j][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
                const int c = 8 * j + 2 * t + e;
                if (g == 0 && c < P.cout) {
                    float* row = P.stats_out + ((size_t)b * P.stats_rows + blockIdx.x * 8 + warp) * 2 * P.cout;
                    row[c] = (float)s; row[P.cout + c] = (float)q;
                }
            }
    }
}

// ------------------------------------------------------------------------------------------------
int conv_warp_pack_weights(const float* w_host, int cout, int cin, int k, std::vector<float>& hi, std::vector<float>& lo, int* kpad_out, int* nt_out) {
    const int K = (cin + 7) / 8 * 8, NT = (cout + 7) / 8, taps = k * k, pitch = cw_pitch(K);
    hi.assign((size_t)taps * NT * 8 * pitch, 0.f); lo.assign(hi.size(), 0.f);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int tp = 0; tp < taps; ++tp) {
                const float w = w_host[((size_t)co * cin + ci) * taps + tp], h = tf32_rn_host(w);
                const size_t at = ((size_t)tp * NT * 8 + co) * pitch + ci;
                hi[at] = h; lo[at] = tf32_rn_host(w - h);
            }
    if (kpad_out) *kpad_out = K;
    if (nt_out) *nt_out = NT;
    return IPDM_OK;
}

bool conv_warp_supported(int cin, int cout, int k, int stride) {
    static const bool off = getenv("IPDM_CONV_WARP") && atoi(getenv("IPDM_CONV_WARP")) == 0;
    return !off && cin >= 4 && cin <= 32 && cin % 4 == 0 && (cout == 8 || cout == 16) && (k == 1 || k == 3) && stride == 1;
}

int conv_warp_stats_rows(int batch, int h, int w, int cin, int k) {          // rows per slice of the statistics tensor (plan builder)
    ConvWarpParams P; ConvWarpDesc d;
    d.src[0].n = batch; d.src[0].h = h; d.src[0].w = w; d.src[0].c = cin; d.src[0].cs = cin; d.ksize = k; d.cout = 8; d.out = d.src[0];
    d.dry = 1;
    if (conv_warp_prepare(P, d) != IPDM_OK) return 0;
    return P.grid_x * 8;
}

template <int K, int NT, int TAPS, int TH>
static int cw_launch(const ConvWarpParams& P, cudaStream_t st) {
    using S = CwSmem<K, NT, TAPS, TH>;
    static DeviceOnce once;
    if (once.need()) IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_warp_kernel<K, NT, TAPS, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL_BYTES));
    conv_warp_kernel<K, NT, TAPS, TH><<<dim3(P.grid_x, P.batch), CW_THREADS, S::TOTAL_BYTES, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

// tile height: 16 rows for the narrow inputs (halo overhead 1.2x), 8 rows when the staged tile would not leave room for two CTAs per SM
static int cw_tile_h(int K, int taps) { return (taps == 9 && K >= 16) ? 8 : 16; }

int conv_warp_prepare(ConvWarpParams& P, const ConvWarpDesc& d) {
    memset(&P, 0, sizeof(P));
    const TensorNHWC& s0 = d.src[0];
    const int c1 = d.nsrc > 1 ? d.src[1].c : 0, cin = s0.c + c1;
    IPDM_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv_warp: 1x1 or 3x3");
    IPDM_REQUIRE(cin >= 4 && cin <= 32 && (d.cout == 8 || d.cout == 16), "conv_warp: C_in %d / C_out %d outside the thin range", cin, d.cout);
    P.K = (cin + 7) / 8 * 8; P.NT = (d.cout + 7) / 8; P.taps = d.ksize * d.ksize; P.TH = cw_tile_h(P.K, P.taps);
    P.H = s0.h; P.W = s0.w; P.batch = s0.n;
    P.tiles_x = ceil_div(P.W, CW_TW); P.tiles_y = ceil_div(P.H, P.TH);
    // CTAs per slice: fill the machine ~3 CTAs per SM over the whole batch, never more than the slice has tiles
    const int tiles = P.tiles_x * P.tiles_y;
    P.grid_x = std::max(1, std::min(tiles, ceil_div(kNumSMs * 3, P.batch)));
    P.stats_rows = P.grid_x * 8;
    if (d.dry) return IPDM_OK;
    IPDM_REQUIRE(s0.c % 4 == 0, "conv_warp: source channel counts must be multiples of 4 (got %d)", s0.c);
    IPDM_REQUIRE(d.nsrc == 1 || (c1 % 4 == 0 && d.src[1].cs % 4 == 0 && d.src[1].h == s0.h && d.src[1].w == s0.w), "conv_warp: bad second source");
    IPDM_REQUIRE(s0.cs % 4 == 0 && d.out.cs % 2 == 0 && (!d.res.p || d.res.cs % 2 == 0), "conv_warp: channel strides must allow 128-bit loads / 64-bit stores");
    IPDM_REQUIRE(d.out.h == P.H && d.out.w == P.W && d.out.n == P.batch && d.out.cs >= d.cout, "conv_warp: output shape mismatch");
    IPDM_REQUIRE(d.w_hi && d.w_lo, "conv_warp: packed weights missing");
    P.src0 = s0.p; P.c0 = s0.c; P.cs0 = s0.cs;
    P.src1 = d.nsrc > 1 ? d.src[1].p : nullptr; P.c1 = c1; P.cs1 = d.nsrc > 1 ? d.src[1].cs : 0;
    P.nscale = d.norm_scale; P.nshift = d.norm_shift; P.act_silu = d.norm_scale ? d.act_silu : 0;
    P.w_hi = d.w_hi; P.w_lo = d.w_lo;
    P.bias = d.bias; P.bias_t_stride = d.bias_t_stride; P.t_dev = d.t_dev;
    P.res = d.res.p; P.res_cs = d.res.cs;
    P.out = d.out.p; P.out_cs = d.out.cs; P.cout = d.cout;
    P.stats_out = d.stats_out;
    static const int env_dbg = getenv("IPDM_WARP_DBG") ? atoi(getenv("IPDM_WARP_DBG")) : 0;     // experiments: 1 no lo terms, 2 no hi*hi, 4 no staging math
    P.dbg = env_dbg;
    return IPDM_OK;
}

int conv_warp_launch(const ConvWarpParams& P, cudaStream_t st) {
    ProfScope prof(PROF_CONV_DIRECT, st, 4.0 * P.batch * (double)P.H * P.W * (P.c0 + P.c1 + P.cout + (P.res ? P.cout : 0)));
#define CW_CASE(KK, NN, TT, HH) if (P.K == KK && P.NT == NN && P.taps == TT && P.TH == HH) return cw_launch<KK, NN, TT, HH>(P, st)
    CW_CASE(8, 1, 9, 16); CW_CASE(8, 2, 9, 16); CW_CASE(16, 1, 9, 8); CW_CASE(16, 2, 9, 8);
    CW_CASE(24, 1, 9, 8); CW_CASE(24, 2, 9, 8); CW_CASE(32, 1, 9, 8); CW_CASE(32, 2, 9, 8);
    CW_CASE(8, 1, 1, 16); CW_CASE(8, 2, 1, 16); CW_CASE(16, 1, 1, 16); CW_CASE(16, 2, 1, 16);
    CW_CASE(24, 1, 1, 16); CW_CASE(24, 2, 1, 16); CW_CASE(32, 1, 1, 16); CW_CASE(32, 2, 1, 16);
#undef CW_CASE
    set_error("conv_warp_launch: no kernel for K %d, N tiles %d, taps %d, rows %d", P.K, P.NT, P.taps, P.TH);
    return IPDM_ERR_UNSUPPORTED;
}

}  // namespace ipdm
