// CUDA-core kernels of the UNet noise predictor (NHWC fp32): direct convolution for the thin
// full-resolution layers, GroupNorm statistics / apply+SiLU, nearest-neighbour resize.
// Reference semantics: Model/model.py norm_layer :82-90, ResidualBlock :95-130, Upsample :160-171,
// Downsample :175-185, UNetModel.out :277-281.
#include "unet_ops.cuh"

#include <algorithm>
#include <cuda_bf16.h>

namespace ipdm {

// ------------------------------------------------------------------------------------------------
// direct convolution: one output pixel per thread, all (<= 16) output channels in registers.
// The CTA stages the input halo tile in shared memory as planes [c][y][x] (conflict-free reads for
// neighbouring pixels) and applies GroupNorm+SiLU / virtual concat / nearest upsample while loading.
// Memory-bound layers (C <= 24 at 2000x912 / 1000x456): 4*(C_in + C_out) bytes per pixel.
// ------------------------------------------------------------------------------------------------
struct ConvDirectParams {
    const float* src0; const float* src1;
    int c0, cs0, c1, cs1;                 // channels / channel stride of each source
    int hs, ws;                           // source spatial size
    int hin, win;                         // conv-input spatial size (== source, or the upsampled size)
    int hout, wout;
    float up_sy, up_sx; int upsample;
    const float* nscale; const float* nshift;
    int ksize, stride, cin, cout, co_tiles, cin_chunk, vec4;
    const float* w; const float* bias; int bias_t_stride; const int* t_dev;
    const float* res; int res_cs;
    float* out; int out_cs;
};

// CTA = 16 x 8 threads, each thread 4 horizontally adjacent output pixels x COUT_T channels in registers
// (64 x 8 output tile).  Input channels are processed in chunks (8, or 4 for stride 2) so the staged halo tile
// stays ~20-35 KB: per (row tap, channel) a thread reads (4-1)*S+K inputs and 3 weight vectors for 12*COUT_T FFMA.
constexpr int CD_BX = 16, CD_BY = 8, CD_PX = 4;
constexpr int CD_TW = CD_BX * CD_PX, CD_TH = CD_BY;

template <int COUT_T, int KS, int STRIDE>
__global__ void __launch_bounds__(CD_BX * CD_BY)
conv_direct_kernel(const ConvDirectParams P) {
    extern __shared__ __align__(16) float cd_smem[];
    constexpr int PAD = KS / 2;
    constexpr int TIN_W = (CD_TW - 1) * STRIDE + KS, TIN_H = (CD_TH - 1) * STRIDE + KS;
    constexpr int TIN_WP = (TIN_W + 3) & ~3;                    // row pitch: multiple of 4 floats -> 128-bit smem loads
    constexpr int PLANE = TIN_WP * TIN_H + 4;                   // +4: de-phase the planes across banks, keep 16-byte alignment
    constexpr int NI = (CD_PX - 1) * STRIDE + KS;               // inputs per row a thread needs
    const int chunk = P.cin_chunk;
    float* tile = cd_smem;                                      // [chunk][PLANE]
    float* wsm = cd_smem + (((size_t)chunk * PLANE + 3) & ~(size_t)3);   // [KS*KS][chunk][COUT_T], 16-byte aligned
    const int n = blockIdx.z / P.co_tiles;
    const int co_base = (blockIdx.z - n * P.co_tiles) * COUT_T; // C_out > 16: tiles of COUT_T output channels
    const int ox0 = blockIdx.x * CD_TW, oy0 = blockIdx.y * CD_TH;
    const int ix0 = ox0 * STRIDE - PAD, iy0 = oy0 * STRIDE - PAD;
    const int tid = threadIdx.y * CD_BX + threadIdx.x;
    constexpr int NT = CD_BX * CD_BY;

    float acc[CD_PX][COUT_T];
#pragma unroll
    for (int p = 0; p < CD_PX; ++p)
#pragma unroll
        for (int i = 0; i < COUT_T; ++i) acc[p][i] = 0.f;

    for (int c0 = 0; c0 < P.cin; c0 += chunk) {
        const int cc = min(chunk, P.cin - c0);
        __syncthreads();                                        // previous chunk fully consumed
        // weights of this chunk -> smem, zero-padded to COUT_T output channels
        for (int i = tid; i < KS * KS * cc * COUT_T; i += NT) {
            const int co = i % COUT_T, rest = i / COUT_T;       // rest = tap * cc + ci
            const int tap = rest / cc, ci = rest - tap * cc;
            wsm[(tap * chunk + ci) * COUT_T + co] = co_base + co < P.cout ? __ldg(P.w + ((size_t)tap * P.cin + c0 + ci) * P.cout + co_base + co) : 0.f;
        }
        // input halo tile of this chunk -> smem planes (GroupNorm+SiLU, concat, nearest upsample fused into the load)
        if (P.vec4) {
            // 128-bit path: one (pixel, 4-channel group) per item; cc is 4 or 8
            const int vpp = cc >> 2;                            // vectors per pixel (1 or 2)
            for (int i = tid; i < TIN_W * TIN_H * vpp; i += NT) {
                const int cv = i & (vpp - 1), pix = i >> (vpp >> 1);
                const int ly = pix / TIN_W, lx = pix - ly * TIN_W;
                const int iy = iy0 + ly, ix = ix0 + lx, c = c0 + 4 * cv;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (iy >= 0 && iy < P.hin && ix >= 0 && ix < P.win) {
                    int sy = iy, sx = ix;
                    if (P.upsample) {
                        sy = min((int)floorf((float)iy * P.up_sy), P.hs - 1);
                        sx = min((int)floorf((float)ix * P.up_sx), P.ws - 1);
                    }
                    const size_t sp = ((size_t)n * P.hs + sy) * P.ws + sx;
                    v = c < P.c0 ? __ldg(reinterpret_cast<const float4*>(P.src0 + sp * P.cs0 + c))
                                 : __ldg(reinterpret_cast<const float4*>(P.src1 + sp * P.cs1 + (c - P.c0)));
                    if (P.nscale) {
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(P.nscale + (size_t)n * P.cin + c));
                        const float4 sh = __ldg(reinterpret_cast<const float4*>(P.nshift + (size_t)n * P.cin + c));
                        v.x = silu(fmaf(v.x, sc.x, sh.x)); v.y = silu(fmaf(v.y, sc.y, sh.y));
                        v.z = silu(fmaf(v.z, sc.z, sh.z)); v.w = silu(fmaf(v.w, sc.w, sh.w));
                    }
                }
                float* tp = tile + (size_t)(4 * cv) * PLANE + ly * TIN_WP + lx;
                tp[0] = v.x; tp[PLANE] = v.y; tp[2 * PLANE] = v.z; tp[3 * PLANE] = v.w;
            }
        } else
        for (int i = tid; i < TIN_W * TIN_H * cc; i += NT) {
            const int ci = i % cc, pix = i / cc;
            const int ly = pix / TIN_W, lx = pix - ly * TIN_W;
            const int iy = iy0 + ly, ix = ix0 + lx, c = c0 + ci;
            float v = 0.f;
            if (iy >= 0 && iy < P.hin && ix >= 0 && ix < P.win) {
                int sy = iy, sx = ix;
                if (P.upsample) {
                    sy = min((int)floorf((float)iy * P.up_sy), P.hs - 1);
                    sx = min((int)floorf((float)ix * P.up_sx), P.ws - 1);
                }
                const size_t sp = ((size_t)n * P.hs + sy) * P.ws + sx;
                v = c < P.c0 ? __ldg(P.src0 + sp * P.cs0 + c) : __ldg(P.src1 + sp * P.cs1 + (c - P.c0));
                if (P.nscale) {
                    v = fmaf(v, __ldg(P.nscale + (size_t)n * P.cin + c), __ldg(P.nshift + (size_t)n * P.cin + c));
                    v = silu(v);
                }
            }
            tile[(size_t)ci * PLANE + ly * TIN_WP + lx] = v;
        }
        __syncthreads();
        const float* tbase = tile + threadIdx.y * STRIDE * TIN_WP + threadIdx.x * CD_PX * STRIDE;
        for (int ci = 0; ci < cc; ++ci) {
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
                float in[NI];
                const float* tp = tbase + (size_t)ci * PLANE + dy * TIN_WP;
#pragma unroll
                for (int k4 = 0; k4 < NI / 4; ++k4) {
                    const float4 t4 = reinterpret_cast<const float4*>(tp)[k4];
                    in[4 * k4] = t4.x; in[4 * k4 + 1] = t4.y; in[4 * k4 + 2] = t4.z; in[4 * k4 + 3] = t4.w;
                }
#pragma unroll
                for (int k = (NI / 4) * 4; k < NI; ++k) in[k] = tp[k];
#pragma unroll
                for (int dx = 0; dx < KS; ++dx) {
                    const float4* wp = reinterpret_cast<const float4*>(wsm + ((dy * KS + dx) * chunk + ci) * COUT_T);
#pragma unroll
                    for (int q = 0; q < COUT_T / 4; ++q) {
                        const float4 w4 = wp[q];
#pragma unroll
                        for (int p = 0; p < CD_PX; ++p) {
                            const float v = in[p * STRIDE + dx];
                            acc[p][4 * q] = fmaf(v, w4.x, acc[p][4 * q]);
                            acc[p][4 * q + 1] = fmaf(v, w4.y, acc[p][4 * q + 1]);
                            acc[p][4 * q + 2] = fmaf(v, w4.z, acc[p][4 * q + 2]);
                            acc[p][4 * q + 3] = fmaf(v, w4.w, acc[p][4 * q + 3]);
                        }
                    }
                }
            }
        }
    }
    const int oy = oy0 + threadIdx.y;
    if (oy >= P.hout) return;
    const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
#pragma unroll
    for (int p = 0; p < CD_PX; ++p) {
        const int ox = ox0 + threadIdx.x * CD_PX + p;
        if (ox >= P.wout) continue;
        const size_t op = ((size_t)n * P.hout + oy) * P.wout + ox;
        float o[COUT_T];
#pragma unroll
        for (int i = 0; i < COUT_T; ++i) {
            const int co = co_base + i;
            float v = 0.f;                                          // pad channels of the output stay zero
            if (co < P.cout) {
                v = acc[p][i] + (bias ? __ldg(bias + co) : 0.f);
                if (P.res) v += __ldg(P.res + op * P.res_cs + co);
            }
            o[i] = v;
        }
        if ((P.out_cs & 3) == 0) {
#pragma unroll
            for (int q = 0; q < COUT_T / 4; ++q)
                if (co_base + 4 * q < P.out_cs)
                    *reinterpret_cast<float4*>(P.out + op * P.out_cs + co_base + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < COUT_T; ++i)
                if (co_base + i < P.out_cs) P.out[op * P.out_cs + co_base + i] = o[i];
        }
        if (co_base + COUT_T >= P.cout)
            for (int co = co_base + COUT_T; co < P.out_cs; ++co) P.out[op * P.out_cs + co] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// Small direct convolution without a staged tile: the 1x1 shortcuts over a virtual concat, the stride-2
// Downsample convs of the thin levels and the 1 -> C stem convs.  These layers have no GroupNorm in
// front, so nothing is amortised by staging; they are pure HBM streams (4*(C_in + C_out) B per pixel) and
// the staged kernel above reaches only ~1.5 TB/s on them (two barriers per channel chunk).
// A thread owns NP horizontally adjacent output pixels x 4 output channels; the C_out/4 channel groups of a
// pixel sit in adjacent lanes, so a warp's 128-bit stores cover whole output pixels back to back (full
// sectors) and its input loads are warp-broadcast 128-bit __ldg (neighbouring taps hit in L1).  Weights:
// [tap][ci][C_out] in shared memory, one 128-bit load per (tap, ci).  A CTA walks image rows (grid-stride),
// so the index arithmetic per item is a shift and a mask.
// ------------------------------------------------------------------------------------------------
template <int KS, int STRIDE, int NP, int VEC>
__global__ void __launch_bounds__(256)
conv_small_kernel(const ConvDirectParams P, int cog_log2, int nrows) {
    extern __shared__ __align__(16) float cs_w[];                // [KS*KS][cin][ncog*4], zero beyond C_out
    constexpr int PAD = KS / 2, NI = (NP - 1) * STRIDE + KS;
    const int ncog = 1 << cog_log2, cpad = ncog * 4;
    for (int i = threadIdx.x; i < KS * KS * P.cin * cpad; i += 256) {
        const int co = i & (cpad - 1), rest = i >> (cog_log2 + 2);          // rest = tap * cin + ci
        cs_w[i] = co < P.cout ? __ldg(P.w + (size_t)rest * P.cout + co) : 0.f;
    }
    __syncthreads();
    const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
    const int wq = (P.wout + NP - 1) / NP;
    const int nv = P.cin / VEC;
    const int row_items = wq << cog_log2;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int n = row / P.hout, oy = row - n * P.hout;
        for (int item = threadIdx.x; item < row_items; item += 256) {
            const int cb = (item & (ncog - 1)) * 4;
            const int ox0 = (item >> cog_log2) * NP, ixb = ox0 * STRIDE - PAD;
            float acc[NP][4];
#pragma unroll
            for (int p = 0; p < NP; ++p) acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.f;
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
                const int iy = oy * STRIDE + dy - PAD;
                if (iy < 0 || iy >= P.hin) continue;
                const size_t rowp = ((size_t)n * P.hin + iy) * P.win;
#pragma unroll 2
                for (int v = 0; v < nv; ++v) {
                    const int c = v * VEC;
                    const float* sp; int scs;
                    if (c < P.c0) { sp = P.src0 + c; scs = P.cs0; } else { sp = P.src1 + (c - P.c0); scs = P.cs1; }
                    float in[NI][VEC];
#pragma unroll
                    for (int k = 0; k < NI; ++k) {
                        const int ix = ixb + k;
                        const bool ok = ix >= 0 && ix < P.win;
                        if constexpr (VEC == 4) {
                            const float4 t = ok ? __ldg(reinterpret_cast<const float4*>(sp + (rowp + ix) * scs)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            in[k][0] = t.x; in[k][1] = t.y; in[k][2] = t.z; in[k][3] = t.w;
                        } else {
                            in[k][0] = ok ? __ldg(sp + (rowp + ix) * scs) : 0.f;
                        }
                    }
#pragma unroll
                    for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const float4 w4 = *reinterpret_cast<const float4*>(cs_w + ((size_t)((dy * KS + dx) * P.cin + c + e)) * cpad + cb);
#pragma unroll
                            for (int p = 0; p < NP; ++p) {
                                const float x = in[p * STRIDE + dx][e];
                                acc[p][0] = fmaf(x, w4.x, acc[p][0]); acc[p][1] = fmaf(x, w4.y, acc[p][1]);
                                acc[p][2] = fmaf(x, w4.z, acc[p][2]); acc[p][3] = fmaf(x, w4.w, acc[p][3]);
                            }
                        }
                }
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int ox = ox0 + p;
                if (ox >= P.wout || cb >= P.out_cs) continue;
                const size_t op = ((size_t)n * P.hout + oy) * P.wout + ox;
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int co = cb + i;
                    float v = 0.f;                                    // pad channels of the output stay zero
                    if (co < P.cout) {
                        v = acc[p][i] + (bias ? __ldg(bias + co) : 0.f);
                        if (P.res) v += __ldg(P.res + op * P.res_cs + co);
                    }
                    o[i] = v;
                }
                *reinterpret_cast<float4*>(P.out + op * P.out_cs + cb) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

template <int KS, int STRIDE, int VEC>
static int launch_small(const ConvDirectParams& P, int batch, cudaStream_t st) {
    constexpr int NP = 2;
    int cog_log2 = 0;
    while ((4 << cog_log2) < std::max(P.cout, P.out_cs)) ++cog_log2;
    const size_t smem = (size_t)KS * KS * P.cin * (4 << cog_log2) * sizeof(float);
    IPDM_REQUIRE(smem <= 48 * 1024, "conv_small: weights (%zu B) do not fit in shared memory", smem);
    const int nrows = batch * P.hout;
    conv_small_kernel<KS, STRIDE, NP, VEC><<<std::min(nrows, 148 * 8), 256, smem, st>>>(P, cog_log2, nrows);
    return IPDM_OK;
}

// Fully unrolled form of the streaming kernel for the channel combinations the projection UNet actually has (ncu on the generic
// kernel above: 2.3-2.7 issued instructions per cycle, ~300 per item -- index arithmetic and loop control, not memory).  Channel
// counts are template parameters, a thread owns 2 adjacent output pixels x ALL output channels (inputs are loaded once per pixel),
// weights are broadcast 128-bit shared loads, every loop is compile-time.  Dense tensors only (out_cs == C_out).
template <int C0V, int C1V, int COUT, int KS, int STRIDE>
__global__ void __launch_bounds__(256)
conv_fixed_kernel(const ConvDirectParams P, int nrows) {
    constexpr int CIN = 4 * (C0V + C1V), NP = 2, PAD = KS / 2, NI = (NP - 1) * STRIDE + KS;
    extern __shared__ __align__(16) float cf_w[];                // [KS*KS][CIN][COUT], the packed layout as it is
    for (int i = threadIdx.x; i < KS * KS * CIN * COUT; i += blockDim.x) cf_w[i] = __ldg(P.w + i);
    __syncthreads();
    const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
    float bq[COUT];
#pragma unroll
    for (int i = 0; i < COUT; ++i) bq[i] = bias ? __ldg(bias + i) : 0.f;
    const int wq = (P.wout + NP - 1) / NP;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int n = row / P.hout, oy = row - n * P.hout;
        for (int oxp = threadIdx.x; oxp < wq; oxp += blockDim.x) {
            const int ox0 = oxp * NP, ixb = ox0 * STRIDE - PAD;
            float acc[NP][COUT];
#pragma unroll
            for (int p = 0; p < NP; ++p)
#pragma unroll
                for (int i = 0; i < COUT; ++i) acc[p][i] = bq[i];
#pragma unroll
            for (int dy = 0; dy < KS; ++dy) {
                const int iy = oy * STRIDE + dy - PAD;
                if (iy < 0 || iy >= P.hin) continue;
                const size_t rowp = ((size_t)n * P.hin + iy) * P.win;
#pragma unroll
                for (int v = 0; v < C0V + C1V; ++v) {
                    const float* sp = v < C0V ? P.src0 + rowp * P.cs0 + 4 * v : P.src1 + rowp * P.cs1 + 4 * (v - C0V);
                    const int scs = v < C0V ? P.cs0 : P.cs1;
                    float4 in[NI];
#pragma unroll
                    for (int k = 0; k < NI; ++k) {
                        const int ix = ixb + k;
                        in[k] = (ix >= 0 && ix < P.win) ? __ldg(reinterpret_cast<const float4*>(sp + (size_t)ix * scs)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int dx = 0; dx < KS; ++dx)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4* wp = reinterpret_cast<const float4*>(cf_w + ((dy * KS + dx) * CIN + 4 * v + e) * COUT);
#pragma unroll
                            for (int q = 0; q < COUT / 4; ++q) {
                                const float4 w4 = wp[q];
#pragma unroll
                                for (int p = 0; p < NP; ++p) {
                                    const float4 t = in[p * STRIDE + dx];
                                    const float x = e == 0 ? t.x : (e == 1 ? t.y : (e == 2 ? t.z : t.w));
                                    acc[p][4 * q] = fmaf(x, w4.x, acc[p][4 * q]); acc[p][4 * q + 1] = fmaf(x, w4.y, acc[p][4 * q + 1]);
                                    acc[p][4 * q + 2] = fmaf(x, w4.z, acc[p][4 * q + 2]); acc[p][4 * q + 3] = fmaf(x, w4.w, acc[p][4 * q + 3]);
                                }
                            }
                        }
                }
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int ox = ox0 + p;
                if (ox >= P.wout) continue;
                const size_t op = ((size_t)n * P.hout + oy) * P.wout + ox;
#pragma unroll
                for (int q = 0; q < COUT / 4; ++q) {
                    float4 o = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
                    if (P.res) {
                        const float4 r = __ldg(reinterpret_cast<const float4*>(P.res + op * P.res_cs + 4 * q));
                        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                    }
                    *reinterpret_cast<float4*>(P.out + op * COUT + 4 * q) = o;
                }
            }
        }
    }
}

template <int C0V, int C1V, int COUT, int KS, int STRIDE>
static int launch_fixed(const ConvDirectParams& P, int batch, cudaStream_t st) {
    const size_t smem = (size_t)KS * KS * 4 * (C0V + C1V) * COUT * sizeof(float);
    static_assert((size_t)KS * KS * 4 * (C0V + C1V) * COUT * sizeof(float) <= 48 * 1024, "conv_fixed: weights do not fit in shared memory");
    const int nrows = batch * P.hout;
    const int threads = std::min(256, ceil_div(ceil_div(P.wout, 2), 32) * 32);        // a CTA walks one image row at a time: no idle warps on narrow rows
    conv_fixed_kernel<C0V, C1V, COUT, KS, STRIDE><<<std::min(nrows, 148 * (2048 / threads)), threads, smem, st>>>(P, nrows);
    return IPDM_OK;
}

// the channel combinations of the shipped projection UNet; anything else takes the generic streaming kernel
static int try_fixed(const ConvDirectDesc& d, const ConvDirectParams& P, int batch, cudaStream_t st) {
    if (d.out.cs != d.cout || (d.res.p && d.res.cs % 4 != 0)) return 1;
    const int a = P.c0, b = P.c1, o = d.cout;
    if (d.ksize == 1) {
        if (a == 4 && b == 0 && o == 8) return launch_fixed<1, 0, 8, 1, 1>(P, batch, st);
        if (a == 8 && b == 4 && o == 8) return launch_fixed<2, 1, 8, 1, 1>(P, batch, st);
        if (a == 8 && b == 8 && o == 8) return launch_fixed<2, 2, 8, 1, 1>(P, batch, st);
        if (a == 16 && b == 8 && o == 8) return launch_fixed<4, 2, 8, 1, 1>(P, batch, st);
        if (a == 16 && b == 8 && o == 16) return launch_fixed<4, 2, 16, 1, 1>(P, batch, st);
        if (a == 16 && b == 16 && o == 16) return launch_fixed<4, 4, 16, 1, 1>(P, batch, st);
    } else if (d.stride == 2) {
        if (a == 8 && b == 0 && o == 8) return launch_fixed<2, 0, 8, 3, 2>(P, batch, st);
        if (a == 16 && b == 0 && o == 16) return launch_fixed<4, 0, 16, 3, 2>(P, batch, st);
    }
    return 1;
}

template <int COUT_T, int KS, int STRIDE>
static int launch_direct(const ConvDirectParams& P, int batch, cudaStream_t st) {
    constexpr int TIN_W = (CD_TW - 1) * STRIDE + KS, TIN_H = (CD_TH - 1) * STRIDE + KS, TIN_WP = (TIN_W + 3) & ~3;
    const size_t smem = ((((size_t)P.cin_chunk * (TIN_WP * TIN_H + 4) + 3) & ~(size_t)3) + (size_t)KS * KS * P.cin_chunk * COUT_T) * sizeof(float);
    static size_t configured[64] = {};                         // largest opt-in so far, per device
    int dev = 0;
    IPDM_CHECK_CUDA(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && (dev < 0 || dev >= 64 || smem > configured[dev])) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_direct_kernel<COUT_T, KS, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = smem;
    }
    dim3 grid(ceil_div(P.wout, CD_TW), ceil_div(P.hout, CD_TH), batch * P.co_tiles), block(CD_BX, CD_BY);
    conv_direct_kernel<COUT_T, KS, STRIDE><<<grid, block, smem, st>>>(P);
    return IPDM_OK;
}

int conv_direct_launch(const ConvDirectDesc& d, cudaStream_t st) {
    ConvDirectParams P{};
    const TensorNHWC& s0 = d.src[0];
    P.src0 = s0.p; P.c0 = s0.c; P.cs0 = s0.cs;
    if (d.nsrc == 2) { P.src1 = d.src[1].p; P.c1 = d.src[1].c; P.cs1 = d.src[1].cs; }
    IPDM_REQUIRE(d.cin == P.c0 + P.c1, "conv_direct: C_in %d != %d + %d", d.cin, P.c0, P.c1);
    IPDM_REQUIRE(d.cout >= 1, "conv_direct: C_out %d", d.cout);
    IPDM_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv_direct: 1x1 or 3x3");
    IPDM_REQUIRE(d.stride == 1 || (d.stride == 2 && d.ksize == 3), "conv_direct: stride 2 only for 3x3");
    P.hs = s0.h; P.ws = s0.w;
    P.upsample = d.upsample;
    P.hin = d.upsample ? d.out.h : s0.h; P.win = d.upsample ? d.out.w : s0.w;
    P.up_sy = (float)s0.h / (float)P.hin; P.up_sx = (float)s0.w / (float)P.win;
    P.hout = d.stride == 1 ? P.hin : (P.hin + 2 * (d.ksize / 2) - d.ksize) / 2 + 1;
    P.wout = d.stride == 1 ? P.win : (P.win + 2 * (d.ksize / 2) - d.ksize) / 2 + 1;
    IPDM_REQUIRE(d.out.h == P.hout && d.out.w == P.wout && d.out.n == s0.n, "conv_direct: output shape mismatch (%dx%d vs %dx%d)",
                 d.out.h, d.out.w, P.hout, P.wout);
    P.nscale = d.norm_scale; P.nshift = d.norm_shift;
    P.ksize = d.ksize; P.stride = d.stride; P.cin = d.cin; P.cout = d.cout;
    P.w = d.w; P.bias = d.bias; P.bias_t_stride = d.bias_t_stride; P.t_dev = d.t_dev;
    P.res = d.res.p; P.res_cs = d.res.cs;
    P.out = d.out.p; P.out_cs = d.out.cs;
    const int ct = d.cout <= 4 ? 4 : (d.cout <= 8 ? 8 : 16);
    P.co_tiles = (d.cout + ct - 1) / ct;
    P.cin_chunk = std::min(d.cin, d.stride == 2 ? 4 : 8);
    // 128-bit staging loads need every channel group of 4 to live in one source, 16-byte aligned
    P.vec4 = d.cin % 4 == 0 && P.c0 % 4 == 0 && P.cs0 % 4 == 0 && (d.nsrc == 1 || P.cs1 % 4 == 0) && P.cin_chunk % 4 == 0 &&
             ((uintptr_t)P.src0 % 16 == 0) && (d.nsrc == 1 || (uintptr_t)P.src1 % 16 == 0);
    ProfScope prof(PROF_CONV_DIRECT, st, 4.0 * s0.n * ((double)s0.h * s0.w * d.cin + (double)P.hout * P.wout * d.cout));   // bytes
    int rc = IPDM_ERR_UNSUPPORTED;
    // layers with nothing to amortise by staging (no fused GroupNorm / upsample): the streaming kernel
    const bool small_ok = !d.norm_scale && !d.upsample && d.cin <= 64 && d.out.cs % 4 == 0 &&
                          (size_t)d.ksize * d.ksize * d.cin * (size_t)(2 * std::max(d.cout, d.out.cs)) * 4 <= 48 * 1024;
    int small_rc = 1;                                             // 1 = not a streaming-kernel layer
    if (small_ok && P.vec4 && (d.ksize == 1 || d.stride == 2)) small_rc = try_fixed(d, P, s0.n, st);
    if (small_rc <= 0) { /* done */ }
    else if (small_ok && P.vec4 && d.ksize == 1) small_rc = launch_small<1, 1, 4>(P, s0.n, st);                          // 1x1 shortcut over a concat
    else if (small_ok && P.vec4 && d.ksize == 3 && d.stride == 2) small_rc = launch_small<3, 2, 4>(P, s0.n, st);    // Downsample
    else if (small_ok && d.cin == 1 && d.nsrc == 1 && d.ksize == 3 && d.stride == 1) small_rc = launch_small<3, 1, 1>(P, s0.n, st);   // stem 1 -> C
    if (small_rc <= 0) {
        IPDM_CHECK(small_rc);
        count_launch();
        IPDM_CHECK_LAUNCH();
        return IPDM_OK;
    }
#define IPDM_CD(CT) (d.ksize == 1 ? launch_direct<CT, 1, 1>(P, s0.n, st) : (d.stride == 1 ? launch_direct<CT, 3, 1>(P, s0.n, st) : launch_direct<CT, 3, 2>(P, s0.n, st)))
    rc = ct == 4 ? IPDM_CD(4) : (ct == 8 ? IPDM_CD(8) : IPDM_CD(16));
#undef IPDM_CD
    IPDM_CHECK(rc);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

double conv_direct_flops(const ConvDirectDesc& d) {
    return 2.0 * d.out.n * d.out.h * d.out.w * (double)d.cin * d.cout * d.ksize * d.ksize;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: HBM-bound single read of the tensor (4 B / element).
//   gn_partial_kernel   grid (nblk, slices): a CTA strides over pixel rows; thread t always sees the same
//                       4-channel vector (t % V), four independent 128-bit loads in flight per thread;
//                       fp32 per-thread sums, fixed-order fp64 combine per channel -> partials[n][blk][c][2]
//   gn_finalize_kernel  grid (groups, slices): fixed-order fp64 reduction over blocks x channels-in-group
//                       -> per-(slice, channel) scale = gamma*rstd, shift = beta - mean*gamma*rstd
// ------------------------------------------------------------------------------------------------
constexpr int GN_UNROLL = 4;

__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ src, int C, int cs, size_t npix, double* __restrict__ partials, int c_off, int Ctot) {
    __shared__ float red[256][8];
    const int n = blockIdx.y, V = C / 4;
    const int R = 256 / V > 0 ? 256 / V : 1;                    // pixel rows handled in parallel
    const int TA = R * V;
    const int t = threadIdx.x;
    float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
    if (t < TA) {
        const int cv = t % V, prow = t / V;
        const float* base = src + (size_t)n * npix * cs + 4 * cv;
        const size_t stride = (size_t)gridDim.x * R;
        size_t pix = (size_t)blockIdx.x * R + prow;
        for (; pix + (GN_UNROLL - 1) * stride < npix; pix += GN_UNROLL * stride) {
            float4 v[GN_UNROLL];
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) v[u] = ld_stream(reinterpret_cast<const float4*>(base + (pix + u * stride) * cs));
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
                q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y); q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
            }
        }
        for (; pix < npix; pix += stride) {
            const float4 v = ld_stream(reinterpret_cast<const float4*>(base + pix * cs));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
        }
    }
    red[t][0] = s.x; red[t][1] = s.y; red[t][2] = s.z; red[t][3] = s.w;
    red[t][4] = q.x; red[t][5] = q.y; red[t][6] = q.z; red[t][7] = q.w;
    __syncthreads();
    for (int c = t; c < C; c += 256) {
        const int cv = c / 4, comp = c % 4;
        double a = 0, b = 0;
        for (int r = 0; r < R; ++r) { a += red[r * V + cv][comp]; b += red[r * V + cv][4 + comp]; }
        double* o = partials + (((size_t)n * gridDim.x + blockIdx.x) * Ctot + c_off + c) * 2;
        o[0] = a; o[1] = b;
    }
}

// A source whose producer conv already wrote per-warp-row statistics (ConvTcDesc::stats_out: [slice][rows][2][C] fp32) is not read
// again: this kernel folds those rows (a few % of the tensor's bytes, coalesced) into the same fp64 partials gn_partial_kernel writes.
// grid (nblk, slices); thread t owns the 4-float vector t % V of a row (V = 2C/4), rows t / V, t / V + R, ...
// fold > 1: the producer was a width-folded conv, its rows are [2][fold][Creal] (pixel-major folded channels); C = fold * Creal here
// and the fold positions of a channel are summed in the final combine.
__global__ void __launch_bounds__(256)
gn_tile_reduce_kernel(const float* __restrict__ tile, int rows, int C, double* __restrict__ partials, int c_off, int Ctot, int fold) {
    __shared__ double red[256][4];
    const int n = blockIdx.y, V = 2 * C / 4;
    const int R = 256 / V > 0 ? 256 / V : 1;
    const int t = threadIdx.x;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (t < R * V) {
        const int cv = t % V;
        const float* base = tile + (size_t)n * rows * 2 * C + 4 * cv;
        const int stride = gridDim.x * R;
        int r = blockIdx.x * R + t / V;
        for (; r + 3 * stride < rows; r += 4 * stride) {            // four independent loads in flight (the loop is latency-bound otherwise)
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(r + u * stride) * 2 * C));
#pragma unroll
            for (int u = 0; u < 4; ++u) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
        }
        for (; r < rows; r += stride) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * 2 * C));
            a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
        }
    }
    red[t][0] = a0; red[t][1] = a1; red[t][2] = a2; red[t][3] = a3;
    __syncthreads();
    const int Cr = C / fold;                                      // real channels
    for (int i = t; i < 2 * Cr; i += 256) {                       // i indexes the [2][Cr] result: half = sum / sum of squares
        const int half = i / Cr, ch = i - half * Cr;
        double a = 0;
        for (int f = 0; f < fold; ++f) {
            const int j = half * C + f * Cr + ch, cv = j / 4, comp = j % 4;
            for (int r = 0; r < R; ++r) a += red[r * V + cv][comp];
        }
        partials[(((size_t)n * gridDim.x + blockIdx.x) * Ctot + c_off + ch) * 2 + half] = a;
    }
}

__global__ void __launch_bounds__(128)
gn_finalize_kernel(const double* __restrict__ partials, int nblk, int Ctot, int groups, double count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float* __restrict__ scale, float* __restrict__ shift) {
    __shared__ double sa[128], sb[128];
    const int g = blockIdx.x, n = blockIdx.y, cpg = Ctot / groups;
    double a = 0, b = 0;
    const int total = nblk * cpg;
    for (int i = threadIdx.x; i < total; i += 128) {
        const int blk = i / cpg, c = g * cpg + (i - blk * cpg);
        const double* o = partials + (((size_t)n * nblk + blk) * Ctot + c) * 2;
        a += o[0]; b += o[1];
    }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    const double mean = sa[0] / count;
    double var = sb[0] / count - mean * mean;
    var = var < 0 ? 0 : var;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += 128) {
        const double ga = gamma[c];
        scale[(size_t)n * Ctot + c] = (float)(ga * rstd);
        shift[(size_t)n * Ctot + c] = (float)((double)beta[c] - mean * ga * rstd);
    }
}

int groupnorm_stats_launch(const GroupNormDesc& d, cudaStream_t st) {
    int Ctot = 0;
    for (int s = 0; s < d.nsrc; ++s) Ctot += d.src[s].c;
    IPDM_REQUIRE(Ctot % d.groups == 0, "groupnorm: %d channels not divisible by %d groups", Ctot, d.groups);
    const TensorNHWC& s0 = d.src[0];
    const size_t npix = (size_t)s0.h * s0.w;
    double read_c = 0;
    for (int s = 0; s < d.nsrc; ++s) if (!d.tile_stats[s]) read_c += d.src[s].c;
    ProfScope prof(PROF_GROUPNORM, st, 4.0 * s0.n * (double)npix * read_c);
    const int R0 = std::max(1, 256 / (s0.c / 4));
    // one partial grid for all sources: enough CTAs to cover the machine a few times, at least 2 unrolled trips each
    int nblk = (int)std::min<size_t>(GN_MAX_BLOCKS, (npix + (size_t)R0 * 2 * GN_UNROLL - 1) / ((size_t)R0 * 2 * GN_UNROLL));
    nblk = std::max(1, std::min(nblk, std::max(1, kNumSMs * 6 / s0.n)));
    int c_off = 0;
    for (int s = 0; s < d.nsrc; ++s) {
        const TensorNHWC& t = d.src[s];
        IPDM_REQUIRE(t.c % 4 == 0 && t.cs % 4 == 0 && t.c <= 1024, "groupnorm: channel count %d must be a multiple of 4", t.c);
        if (d.tile_stats[s]) {                                    // statistics from the producer's epilogue: fold its partial rows
            const int fold = d.tile_fold[s] > 1 ? d.tile_fold[s] : 1;
            IPDM_REQUIRE(t.c * fold <= 512, "groupnorm: producer statistics support at most 512 channels per source");
            gn_tile_reduce_kernel<<<dim3(nblk, s0.n), 256, 0, st>>>(d.tile_stats[s], d.tile_rows[s], t.c * fold, d.partials, c_off, Ctot, fold);
        } else {                                                  // one read of the tensor
            gn_partial_kernel<<<dim3(nblk, s0.n), 256, 0, st>>>(t.p, t.c, t.cs, npix, d.partials, c_off, Ctot);
        }
        count_launch();
        c_off += t.c;
    }
    gn_finalize_kernel<<<dim3(d.groups, s0.n), 128, 0, st>>>(d.partials, nblk, Ctot, d.groups, (double)npix * (Ctot / d.groups), d.gamma, d.beta,
                                                            d.eps, d.scale, d.shift);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm apply + SiLU (+ virtual concat, + channel padding) -> operand tensor of a tensor-core conv.
// HBM-bound: 4 B read + 4 B written per element; two independent 128-bit loads in flight per thread.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_vec4(float* out, size_t elem_off, const float4& o, int bf16) {
    if (bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + elem_off) = pk;
    } else {
        *reinterpret_cast<float4*>(out + elem_off) = o;
    }
}

// grid (nblk, slices).  Thread t always handles the same 4-channel vector (t % V) of pixel rows t / V, t / V + R, ...: its
// scale / shift / source pointer live in registers, the loop body is U independent 128-bit loads, the affine + SiLU, and U
// stores -- no index arithmetic (the flat-index version spent two 64-bit divisions per vector and reached 3.7-4.6 TB/s).
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ s0, int c0, int cs0, const float* __restrict__ s1, int c1, int cs1,
                const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int ocs,
                size_t npix, int act, int rnd, int obf16) {
    const int Ctot = c0 + c1, V = ocs / 4;
    const int R = 256 / V > 0 ? 256 / V : 1;
    constexpr int U = 4;                                         // independent 128-bit loads in flight per thread
    const int n = blockIdx.y;
    for (int t = threadIdx.x; t < R * V; t += 256) {             // one pass unless V > 256
        const int cv = t % V, prow = t / V, c = 4 * cv;
        const bool real = c < Ctot;                              // pad channels of the operand tensor are written as zeros
        float4 sc = make_float4(0, 0, 0, 0), sh = sc;
        const float* src = nullptr; int scs = 0;
        if (real) {
            sc = __ldg(reinterpret_cast<const float4*>(scale + (size_t)n * Ctot + c));
            sh = __ldg(reinterpret_cast<const float4*>(shift + (size_t)n * Ctot + c));
            if (c < c0) { src = s0 + (size_t)n * npix * cs0 + c; scs = cs0; } else { src = s1 + (size_t)n * npix * cs1 + (c - c0); scs = cs1; }
        }
        const size_t obase = (size_t)n * npix * ocs + c;
        const size_t stride = (size_t)gridDim.x * R;
        auto finish = [&](float4 v, size_t pix) {
            float4 o = make_float4(0, 0, 0, 0);
            if (real) {
                o.x = fmaf(v.x, sc.x, sh.x); o.y = fmaf(v.y, sc.y, sh.y); o.z = fmaf(v.z, sc.z, sh.z); o.w = fmaf(v.w, sc.w, sh.w);
                if (act) { o.x = silu(o.x); o.y = silu(o.y); o.z = silu(o.z); o.w = silu(o.w); }
                if (rnd) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }   // tf32 mode: MMA operand
            }
            store_vec4(out, obase + pix * ocs, o, obf16);
        };
        size_t pix = (size_t)blockIdx.x * R + prow;
        for (; pix + (U - 1) * stride < npix; pix += U * stride) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = real ? ld_stream(reinterpret_cast<const float4*>(src + (pix + u * stride) * scs)) : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < U; ++u) finish(v[u], pix + u * stride);
        }
        for (; pix < npix; pix += stride)
            finish(real ? ld_stream(reinterpret_cast<const float4*>(src + pix * scs)) : make_float4(0, 0, 0, 0), pix);
    }
}

int groupnorm_apply_launch(const GroupNormDesc& d, const TensorNHWC& out, int act_silu, int round_tf32, cudaStream_t st) {
    const TensorNHWC& a = d.src[0];
    const int c1 = d.nsrc == 2 ? d.src[1].c : 0;
    IPDM_REQUIRE(out.cs % 4 == 0 && out.cs >= a.c + c1 && a.c % 4 == 0 && out.cs <= 1024, "groupnorm_apply: bad channel layout");
    const size_t npix = (size_t)a.h * a.w;
    const int V = out.cs / 4, R = std::max(256 / V, 1);
    // ~8 CTAs per SM over the whole batch, at least 4 row-iterations per thread
    const int nblk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ceil_div(kNumSMs * 8, a.n), (npix + (size_t)R * 4 - 1) / ((size_t)R * 4)));
    ProfScope prof(PROF_GROUPNORM, st, a.n * (double)npix * (4.0 * (a.c + c1) + (out.bf16 ? 2.0 : 4.0) * out.cs));
    gn_apply_kernel<<<dim3(nblk, a.n), 256, 0, st>>>(a.p, a.c, a.cs, d.nsrc == 2 ? d.src[1].p : nullptr, c1, d.nsrc == 2 ? d.src[1].cs : 0,
                                                     d.scale, d.shift, out.p, out.cs, npix, act_silu, round_tf32 && !out.bf16, out.bf16);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

// ------------------------------------------------------------------------------------------------
// nearest resize
// ------------------------------------------------------------------------------------------------
// grid (nblk, slices); thread t owns the 4-channel vector t % V (as gn_apply), 32-bit index arithmetic per pixel
__global__ void __launch_bounds__(256)
upsample_kernel(const float* __restrict__ src, int hs, int ws, int sc, int scs, float* __restrict__ dst, int hd, int wd, int dcs,
                float sy, float sx, int rnd, int obf16) {
    const int V = dcs / 4, R = 256 / V > 0 ? 256 / V : 1;
    const int n = blockIdx.y;
    const unsigned npix = (unsigned)hd * (unsigned)wd;
    for (int t = threadIdx.x; t < R * V; t += 256) {
        const int cv = t % V, c = 4 * cv;
        const bool real = c < sc;
        const float* sp = src + (size_t)n * hs * ws * scs + c;
        const size_t obase = (size_t)n * npix * dcs + c;
        for (unsigned pix = blockIdx.x * R + t / V; pix < npix; pix += gridDim.x * R) {
            const unsigned y = pix / (unsigned)wd, x = pix - y * (unsigned)wd;
            const int yy = min((int)floorf((float)y * sy), hs - 1), xx = min((int)floorf((float)x * sx), ws - 1);
            float4 v = make_float4(0, 0, 0, 0);
            if (real) v = __ldg(reinterpret_cast<const float4*>(sp + ((size_t)yy * ws + xx) * scs));
            if (rnd) { v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w); }   // tf32 mode: feeds a tensor-core conv only
            store_vec4(dst, obase + (size_t)pix * dcs, v, obf16);
        }
    }
}

int upsample_nearest_launch(const TensorNHWC& src, const TensorNHWC& dst, int round_tf32, cudaStream_t st) {
    IPDM_REQUIRE(src.cs % 4 == 0 && dst.cs % 4 == 0 && dst.cs >= src.c && src.c % 4 == 0 && src.n == dst.n && dst.cs <= 1024, "upsample: bad layout");
    const size_t npix = (size_t)dst.h * dst.w;
    IPDM_REQUIRE(npix < (1u << 31), "upsample: image too large");
    const int V = dst.cs / 4, R = std::max(256 / V, 1);
    const int nblk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ceil_div(kNumSMs * 8, dst.n), (npix + (size_t)R * 4 - 1) / ((size_t)R * 4)));
    ProfScope prof(PROF_UPSAMPLE, st, 4.0 * (double)src.elems() + (dst.bf16 ? 2.0 : 4.0) * dst.elems());
    upsample_kernel<<<dim3(nblk, dst.n), 256, 0, st>>>(src.p, src.h, src.w, src.c, src.cs, dst.p, dst.h, dst.w, dst.cs, (float)src.h / dst.h,
                                                        (float)src.w / dst.w, round_tf32 && !dst.bf16, dst.bf16);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

}  // namespace ipdm
