// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
// Covers every dense contraction of the reference UNet (Model/model.py: Conv2d 3x3 :101,113,165,
// stride-2 Conv2d :180, 1x1 shortcut / qkv / proj :117,142,143) whose channel counts reach the
// tensor-core path (C_in >= 16 after padding to 32, C_out in {16, 64, 128, ...}).
//
// Layout: activations NHWC fp32 with a channel stride that is a multiple of 32 (pad channels are 0),
// weights repacked as [tap][C_out][K] (K = input channels in virtual-concat order, K-major).
// GEMM view per CTA: D[128 pixels, BLOCK_N] += A_tap[128 pixels, 32 ch] * W_tap[BLOCK_N, 32 ch]^T over
// (tap, 32-channel chunk).  The A tile of a tap is ONE 4-D TMA box {32 ch, TW, TH, 1} of the NHWC
// tensor at pixel offset (dx-1, dy-1): out-of-image coordinates are zero-filled by TMA, which is exactly
// the convolution's zero padding and also covers ragged image sizes (228, 57, 29 ... columns).
// Stride-2 convolutions read four parity sub-lattices of the input through four tensor maps.
// Two virtual-concat sources (torch.cat([h, skip]) :306) are two tensor maps walked back to back.
//
// Operands are fp32 words rounded to TF32 (kind::tf32; tf32 and fp32 modes, K chunk = 32 channels) or bf16 (kind::f16;
// bf16 mode, K chunk = 64 channels); accumulation is fp32 in TMEM in every mode.
//
// Five kernels share this file (conv_tc_prepare picks one per layer, see there):
//   conv_halo_fused_kernel       the dominant kernel (round 2): GroupNorm + SiLU -> stride-1 3x3 as ONE launch over the raw fp32 tensors
//                                (operand transform warps between TMA and MMA), the width-folded thin layers of the 2000x912 / 1000x456
//                                levels (with the ResBlock shortcut as identity K chunks), the Upsample convs as four output-parity phases
//   conv_halo_persistent_kernel  stride-1 3x3 from an operand tensor: one halo tile per K chunk feeds all nine taps, persistent
//                                CTAs, double-buffered accumulator pairs, two epilogue warpgroups    (0.82-0.84 of the sustained bf16 peak)
//   conv_tc_persistent_kernel    per-tap variant for 1x1 / stride-2 / small images (L2 -> SM bound on 3x3 layers)
//   conv_halo_kernel             one halo tile per CTA, for the N = 16 layer
//   conv_tc_kernel               one tile per CTA: the 3xTF32 split of the fp32 mode and the qkv epilogue
// Roles in the one-tile kernel (128 threads): warp 0 lane* = TMA producer, warp 1 lane* = MMA issuer (tcgen05.mma,
// M=128, N=BLOCK_N, 32 bytes of K x4 per 128-byte stage), warp 2 = TMEM allocator; then all four warps run the
// epilogue: tcgen05.ld 32 lanes x 32 columns, + bias[t] (+ residual), 128-bit NHWC stores.  The persistent kernels
// add dedicated epilogue warps that transpose through shared memory for full-line stores and emit the GroupNorm
// statistics of their output (ConvTcDesc::stats_out).
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc.cuh"
#include "unet_ops.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace ipdm {

// ------------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int tmap_encode(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = get_encode();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return IPDM_ERR_CUDA; }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu, box %u %u %u %u)", (int)r, rank,
                  (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), (unsigned long long)(rank > 2 ? d[2] : 0),
                  (unsigned long long)(rank > 3 ? d[3] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
        return IPDM_ERR_CUDA;
    }
    return IPDM_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_THREADS = 128;
constexpr int TC_KC = 32;                 // fp32 elements per 128-byte operand row (bf16 operands: 64, P.kc)
constexpr int TC_A_BYTES = 128 * 128;

// bias (+ time-embedding row) of this CTA's output channels -> shared memory, once per CTA
template <int BLOCK_N>
__device__ __forceinline__ void stage_bias(const ConvTcParams& P, float* sbias, int n0) {
    const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) sbias[i] = (bias && n0 + i < P.cout) ? __ldg(bias + n0 + i) : 0.f;
}

// ---------------- epilogue of one 128-pixel sub-tile: TMEM -> registers -> (+bias, +residual) -> NHWC global ----------------
// Row m of the accumulator (== TMEM lane) is pixel (y0 + (m >> tw_log2), x0 + (m & (2^tw_log2 - 1))); columns >= tw_valid of a
// tile row are the padding columns of the halo kernel and are dropped.
template <int BLOCK_N, bool SPLIT>
__device__ __forceinline__ void tc_epilogue(const ConvTcParams& P, uint32_t tmem_acc, int quarter, int lane, int b, int x0, int y0, int n0,
                                            int tw_valid, const float* sbias) {
    const int m = quarter * 32 + lane;               // accumulator row == TMEM lane (a warp may only touch its own lane quarter)
    const int TW = 1 << P.tw_log2;
    constexpr int CHUNK = BLOCK_N < 32 ? 16 : 32;
    const int py = y0 + (m >> P.tw_log2), px = x0 + (m & (TW - 1));
    const bool valid = py < P.H && px < P.W && (m & (TW - 1)) < tw_valid;
    const size_t pix = ((size_t)b * P.H + py) * P.W + px;
    const bool has_res = valid && P.res != nullptr;
    // the residual of chunk c+1 is requested before chunk c is processed, so its latency hides behind the TMEM load + stores
    float4 rnext[CHUNK / 4];
    if (has_res) {
        const float4* rp = reinterpret_cast<const float4*>(P.res + pix * P.res_cs + n0);
#pragma unroll
        for (int i = 0; i < CHUNK / 4; ++i) rnext[i] = __ldg(rp + i);
    }
#pragma unroll 1
    for (int cc = 0; cc < BLOCK_N; cc += CHUNK) {
        __syncwarp();
        uint32_t r[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc;
        if constexpr (CHUNK == 32) tc::tmem_ld32(taddr, r);
        else { uint32_t r16[16]; tc::tmem_ld16(taddr, r16);
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = r16[i]; }
        float4 rcur[CHUNK / 4];
#pragma unroll
        for (int i = 0; i < CHUNK / 4; ++i) rcur[i] = rnext[i];
        if (has_res && cc + CHUNK < BLOCK_N && n0 + cc + CHUNK < P.cout) {
            const float4* rp = reinterpret_cast<const float4*>(P.res + pix * P.res_cs + n0 + cc + CHUNK);
#pragma unroll
            for (int i = 0; i < CHUNK / 4; ++i) rnext[i] = __ldg(rp + i);
        }
        tc::tmem_ld_wait();
        const int n = n0 + cc;
        if (!valid || n >= P.cout) continue;
        float v[CHUNK];
#pragma unroll
        for (int i = 0; i < CHUNK / 4; ++i) {
            const float4 bq = *reinterpret_cast<const float4*>(sbias + cc + 4 * i);
            v[4 * i] = __uint_as_float(r[4 * i]) + bq.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bq.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bq.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bq.w;
        }
        if (has_res) {
#pragma unroll
            for (int i = 0; i < CHUNK / 4; ++i) { v[4 * i] += rcur[i].x; v[4 * i + 1] += rcur[i].y; v[4 * i + 2] += rcur[i].z; v[4 * i + 3] += rcur[i].w; }
        }
        if (P.qkv_mode && !SPLIT && !P.qkv_bf16) {   // q, k, v are operands of the attention MMAs: round to nearest tf32
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) v[i] = tf32_rn(v[i]);
        }
        float lo[CHUNK];
        if (P.qkv_mode && SPLIT) {                   // fp32 mode: hand q, k, v to the attention kernel as tf32 hi / lo pairs
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) { const float h = tf32_rn(v[i]); lo[i] = tf32_rn(v[i] - h); v[i] = h; }
        }
        if (P.qkv_mode && P.qkv_bf16) {              // bf16 mode: q, k, v are bf16 operands of the attention MMAs
            const bool is_v = (n % (3 * P.head_dim)) >= 2 * P.head_dim;
            if (is_v) {
                const int head = n / (3 * P.head_dim), d0 = n % (3 * P.head_dim) - 2 * P.head_dim;
                __nv_bfloat16* vp = reinterpret_cast<__nv_bfloat16*>(P.vt) + (((size_t)b * P.heads + head) * P.head_dim + d0) * P.t_pad + ((size_t)py * P.W + px);
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) vp[(size_t)i * P.t_pad] = __float2bfloat16_rn(v[i]);
            } else {
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.out) + pix * P.out_cs + n);
#pragma unroll
                for (int i = 0; i < CHUNK / 8; ++i) {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __nv_bfloat162 pk = __floats2bfloat162_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
                        w[k] = *reinterpret_cast<const uint32_t*>(&pk);
                    }
                    op[i] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            continue;
        }
        if (P.qkv_mode && ((n % (3 * P.head_dim)) >= 2 * P.head_dim)) {
            // V of one head: write transposed, [b][head][d][t_pad] (token contiguous), so P.V^T is K-major for attention
            const int head = n / (3 * P.head_dim), d0 = n % (3 * P.head_dim) - 2 * P.head_dim;
            const size_t tok = (size_t)py * P.W + px;
            const size_t vo = (((size_t)b * P.heads + head) * P.head_dim + d0) * P.t_pad + tok;
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) P.vt[vo + (size_t)i * P.t_pad] = v[i];
            if (SPLIT) {
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) P.vt_lo[vo + (size_t)i * P.t_pad] = lo[i];
            }
        } else {
            if (P.qkv_mode && SPLIT) {
                float4* lp = reinterpret_cast<float4*>(P.out_lo + pix * P.out_cs + n);
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) lp[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
            float4* op = reinterpret_cast<float4*>(P.out + pix * P.out_cs + n);
#pragma unroll
            for (int i = 0; i < CHUNK / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    }
    if (valid && !P.qkv_mode && blockIdx.y == gridDim.y - 1)     // keep the channel padding of the output at zero
        for (int c = P.cout; c < P.out_cs; ++c) P.out[pix * P.out_cs + c] = 0.f;
}


// ---------------- coalesced epilogue (persistent kernel) ----------------
// tcgen05.ld hands every lane ONE pixel x 32 channels; writing that straight to NHWC makes each warp store touch 32
// different 128-byte lines with 16 bytes each (measured: ~8.7 us per 128x128 tile, the bottleneck of bf16 layers).
// Here each warp transposes its 32 x 32 block through a private shared-memory tile (pitch 36 floats, 128-bit accesses)
// so that 8 lanes cover one pixel's 128 contiguous bytes: every global load (residual) and store is a full line.
constexpr int EPI_PITCH = 36;
constexpr int EPI_WARP_FLOATS = 32 * EPI_PITCH;
constexpr int EPI_ROW_BYTES = 32 * 128;               // EpilogueRow: one [32 px][32 ch] SWIZZLE_128B tile per warp
template <int BLOCK_N, int CHUNK_W = 32>
__device__ __forceinline__ void tc_epilogue_coalesced(const ConvTcParams& P, uint32_t tmem_acc, int quarter, int lane, int b, int x0, int y0,
                                                      int n0, const float* sbias, float* stile, int tw_valid, float* stats_row) {
    const int TW = 1 << P.tw_log2;                    // tile row pitch; only the first tw_valid columns are outputs
    constexpr int CHUNK = BLOCK_N < 32 ? 16 : CHUNK_W; // columns per TMEM load: 16 halves the registers (r[], residual lines) of the 32-wide form
    constexpr int LPP = CHUNK / 4;                    // lanes per pixel row segment (8 for 32 channels, 4 for 16)
    constexpr int PPI = 32 / LPP;                     // pixels per warp instruction
    const int c4 = (lane % LPP) * 4, psub = lane / LPP;
    // per-row coordinates of the 32/PPI rows this lane touches in every chunk
    constexpr int NJ = 32 / PPI;
    size_t pixv[NJ]; bool okv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int m = quarter * 32 + j * PPI + psub;
        const int py = y0 + (m >> P.tw_log2), px = x0 + (m & (TW - 1));
        okv[j] = py < P.H && px < P.W && (m & (TW - 1)) < tw_valid;
        pixv[j] = ((size_t)b * P.H + py) * P.W + px;
    }
#pragma unroll 1
    for (int cc = 0; cc < BLOCK_N; cc += CHUNK) {
        const int n = n0 + cc + c4;
        const bool nok = n < P.cout;
        uint32_t r[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc;
        if constexpr (CHUNK == 32) tc::tmem_ld32(taddr, r);
        else { uint32_t r16[16]; tc::tmem_ld16(taddr, r16);
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = r16[i]; }
        float4 rres[NJ];                               // all residual lines of the chunk are requested before anything waits
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            rres[j] = (P.res && okv[j] && nok) ? __ldg(reinterpret_cast<const float4*>(P.res + pixv[j] * P.res_cs + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        tc::tmem_ld_wait();
        __syncwarp();                                  // previous chunk's readers are done with the tile
        float4* row = reinterpret_cast<float4*>(stile + lane * EPI_PITCH);
#pragma unroll
        for (int i = 0; i < CHUNK / 4; ++i)
            row[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        __syncwarp();
        const float4 bq = *reinterpret_cast<const float4*>(sbias + cc + c4);
        float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), sq = ss;          // GroupNorm partials of this lane's 4 channels
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (okv[j] && nok) {
                float4 v = *reinterpret_cast<const float4*>(stile + (j * PPI + psub) * EPI_PITCH + c4);
                v.x += bq.x + rres[j].x; v.y += bq.y + rres[j].y; v.z += bq.z + rres[j].z; v.w += bq.w + rres[j].w;
                *reinterpret_cast<float4*>(P.out + pixv[j] * P.out_cs + n) = v;
                ss.x += v.x; ss.y += v.y; ss.z += v.z; ss.w += v.w;
                sq.x = fmaf(v.x, v.x, sq.x); sq.y = fmaf(v.y, v.y, sq.y); sq.z = fmaf(v.z, v.z, sq.z); sq.w = fmaf(v.w, v.w, sq.w);
            }
        }
        if (stats_row) {
            // statistics of the values just written, for the GroupNorm that reads this tensor next: sum over the warp's 32 pixels
            // (lanes with equal lane % LPP hold the same channels), one [2][cout] row per warp, fixed order => deterministic
#pragma unroll
            for (int o = LPP; o < 32; o <<= 1) {
                ss.x += __shfl_xor_sync(0xffffffffu, ss.x, o); ss.y += __shfl_xor_sync(0xffffffffu, ss.y, o);
                ss.z += __shfl_xor_sync(0xffffffffu, ss.z, o); ss.w += __shfl_xor_sync(0xffffffffu, ss.w, o);
                sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
                sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o);
            }
            if (lane < LPP && nok) {
                *reinterpret_cast<float4*>(stats_row + n) = ss;
                *reinterpret_cast<float4*>(stats_row + P.cout + n) = sq;
            }
        }
    }
    // keep the channel padding of the output at zero (layers with C_out < channel stride: 16 -> 32)
    if (P.out_cs > P.cout && n0 + BLOCK_N >= P.cout) {
        const int m = quarter * 32 + lane;
        const int py = y0 + (m >> P.tw_log2), px = x0 + (m & (TW - 1));
        if (py < P.H && px < P.W && (m & (TW - 1)) < tw_valid) {
            const size_t pix = ((size_t)b * P.H + py) * P.W + px;
            for (int c = P.cout; c < P.out_cs; ++c) P.out[pix * P.out_cs + c] = 0.f;
        }
    }
}

// ---------------- coalesced epilogue of the halo kernels ----------------
// Same data flow as tc_epilogue_coalesced, specialised for the halo tiling (tile row pitch 32): the 32 accumulator rows of a warp are
// the 32 consecutive pixels x0 .. x0+31 of ONE image row, so every address is row base + pixel * channel stride -- one 64-bit base per
// tile and 32-bit offsets, where the generic form spent ~6 integer instructions per 128-bit access on (slice, y, x) arithmetic
// (ncu, width-folded layers: the epilogue warps issued 6.4 k instructions per 240-pixel x 32-channel tile and the SM was issue-bound).
template <int BLOCK_N, int CHUNK_W = 32>
struct EpilogueRow {
    static constexpr int CHUNK = BLOCK_N < 32 ? 16 : CHUNK_W;
    static constexpr int LPP = CHUNK / 4, PPI = 32 / LPP, NJ = 32 / PPI;
    int c4, psub, nvalid, n0, ostep, rstep, oo, ro, tx0, tpy, tb; uint32_t vmask; bool has_res;
    size_t pix0; float* obase; const float* rbase;
    float4 rres[NJ];                                   // residual lines of the chunk being processed

    // coordinates of this warp's row (pixels x0 .. x0+31 of image row py of slice b, output channels n0 ...)
    // ph: output-parity phase of an upsample conv (ph_log2 = 2): tile pixel (py, x) is output pixel (2 py + ph / 2, 2 x + ph % 2)
    __device__ __forceinline__ void setup(const ConvTcParams& P, int lane, int b, int x0, int py, int n0_, int tw_valid, int ph = 0) {
        c4 = (lane % LPP) * 4; psub = lane / LPP; n0 = n0_; tx0 = x0; tpy = py; tb = b;
        nvalid = py < P.H ? min(tw_valid, P.W - x0) : 0;
        vmask = 0;                                     // bit j: pixel j * PPI + psub is an output
#pragma unroll
        for (int j = 0; j < NJ; ++j) vmask |= (j * PPI + psub < nvalid ? 1u : 0u) << j;
        const int up = P.ph_log2 ? 2 : 1, opx = up * P.out_cs;           // elements between consecutive tile pixels in the output
        pix0 = ((size_t)b * P.H + py) * P.W + x0;
        const size_t opix0 = P.ph_log2 ? ((size_t)b * (2 * P.H) + 2 * py + (ph >> 1)) * (2 * P.W) + 2 * x0 + (ph & 1) : pix0;
        obase = P.out + opix0 * P.out_cs;
        has_res = P.res != nullptr;
        rbase = has_res ? P.res + pix0 * P.res_cs : P.out;               // (never read without has_res)
        ostep = PPI * opx; rstep = PPI * P.res_cs;
        oo = psub * opx + c4 + n0; ro = psub * P.res_cs + c4 + n0;
    }
    // request the residual lines of chunk cc.  Chunk 0 is requested BEFORE the warp waits for its accumulator: the addresses only depend
    // on the tile, so the HBM latency of the residual overlaps the MMAs instead of extending the epilogue (ncu, width-folded layers:
    // 73 % of the epilogue warps' stall samples sat on the first add that consumes a residual line)
    __device__ __forceinline__ void load_res(const ConvTcParams& P, int cc) {
        const bool nok = n0 + cc + c4 < P.cout;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            rres[j] = (has_res && nok && ((vmask >> j) & 1u)) ? ld_stream(reinterpret_cast<const float4*>(rbase + (ro + cc + j * rstep))) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (cc == 0 && has_res && BLOCK_N > CHUNK) {
            // the lines of the later chunks go to L2 now (no registers to hold them): their loads then cost an L2 hit, not an HBM round trip
#pragma unroll
            for (int c2 = CHUNK; c2 < BLOCK_N; c2 += CHUNK)
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (((vmask >> j) & 1u) && n0 + c2 + c4 < P.cout)
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(rbase + (ro + c2 + j * rstep)));
        }
    }
    // Staging tile: [32 px][128 B] with the 128-byte swizzle (16-byte chunk ^ (pixel & 7)) -- conflict-free both for the lane-per-pixel
    // writes of the accumulator rows and for the 8-lanes-per-pixel reads, and exactly the layout a SWIZZLE_128B tensor map stores from:
    // with P.tma_store the finished values go back into the tile and ONE cp.async.bulk.tensor per chunk writes the 30 x 32-channel box
    // (TMA clips at the image edge); otherwise the lanes store full lines themselves (upsample phases: strided pixels; N = 16 tiles).
    __device__ __forceinline__ void run(const ConvTcParams& P, uint32_t tmem_acc, int quarter, int lane, const float* sbias, float* stile, float* stats_row) {
        const uint32_t st = tc::smem_u32(stile);
        const bool tma = P.tma_store && CHUNK == 32;
        const uint32_t wr = st + (uint32_t)lane * 128u;                       // this lane's pixel row (staging writes)
        const int q = lane % LPP;                                             // 16-byte chunk this lane reads back
#pragma unroll 1
        for (int cc = 0; cc < BLOCK_N; cc += CHUNK) {
            const int n = n0 + cc + c4;
            const bool nok = n < P.cout;
            const uint32_t vm = nok ? vmask : 0u;
            uint32_t r[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc;
            if constexpr (CHUNK == 32) tc::tmem_ld32(taddr, r);
            else { uint32_t r16[16]; tc::tmem_ld16(taddr, r16);
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = r16[i]; }
            if (cc > 0) load_res(P, cc);
            tc::tmem_ld_wait();
            if (tma && lane == 0) tc::tma_store_wait_read();   // the previous chunk's bulk store has read the tile
            __syncwarp();                              // previous chunk's readers are done with the tile
#pragma unroll
            for (int i = 0; i < CHUNK / 4; ++i)
                tc::sts128(wr + (uint32_t)((i ^ (lane & 7)) << 4),
                           make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
            __syncwarp();
            const float4 bq = *reinterpret_cast<const float4*>(sbias + cc + c4);
            float4 ss = make_float4(0.f, 0.f, 0.f, 0.f), sq = ss;          // GroupNorm partials of this lane's 4 channels
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if ((vm >> j) & 1u) {
                    const int p = j * PPI + psub;
                    const uint32_t at = st + (uint32_t)p * 128u + (uint32_t)((q ^ (p & 7)) << 4);
                    float4 v = tc::lds128(at);
                    v.x += bq.x + rres[j].x; v.y += bq.y + rres[j].y; v.z += bq.z + rres[j].z; v.w += bq.w + rres[j].w;
                    if (tma) tc::sts128(at, v);
                    else st_stream(reinterpret_cast<float4*>(obase + (oo + cc + j * ostep)), v);
                    ss.x += v.x; ss.y += v.y; ss.z += v.z; ss.w += v.w;
                    sq.x = fmaf(v.x, v.x, sq.x); sq.y = fmaf(v.y, v.y, sq.y); sq.z = fmaf(v.z, v.z, sq.z); sq.w = fmaf(v.w, v.w, sq.w);
                }
            }
            if (tma) {
                tc::fence_proxy_async();               // generic-proxy writes of the tile -> visible to the TMA engine
                __syncwarp();
                if (lane == 0 && nvalid > 0) tc::tma_store_4d(&P.mapOut, st, n0 + cc, tx0, tpy, tb);
            }
            if (stats_row) {
#pragma unroll
                for (int o = LPP; o < 32; o <<= 1) {
                    ss.x += __shfl_xor_sync(0xffffffffu, ss.x, o); ss.y += __shfl_xor_sync(0xffffffffu, ss.y, o);
                    ss.z += __shfl_xor_sync(0xffffffffu, ss.z, o); ss.w += __shfl_xor_sync(0xffffffffu, ss.w, o);
                    sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
                    sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o);
                }
                if (lane < LPP && nok) {
                    *reinterpret_cast<float4*>(stats_row + n) = ss;
                    *reinterpret_cast<float4*>(stats_row + P.cout + n) = sq;
                }
            }
        }
        // keep the channel padding of the output at zero (layers with C_out < channel stride: 16 -> 32)
        if (P.out_cs > P.cout && n0 + BLOCK_N >= P.cout && lane < nvalid && !P.ph_log2) {
            float* o = P.out + (pix0 + lane) * P.out_cs;
            for (int c = P.cout; c < P.out_cs; ++c) o[c] = 0.f;
        }
    }
};

// SPLIT = fp32-accurate "3xTF32" mode: every fp32 operand x is used as x_hi + x_lo (x_hi = rn_tf32(x), x_lo =
// rn_tf32(x - x_hi)) and D += A_hi B_hi + A_hi B_lo + A_lo B_hi.  Weights are split on the host (two packed arrays, two
// TMA loads); the activation tile is split in shared memory by warps 2-3 right after the TMA lands (the split is
// elementwise, so the 128B swizzle is irrelevant to it), then handed to the MMA warp through a `ready` barrier.
template <int BLOCK_N, int STAGES, bool SPLIT>
struct TcSmem {
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int OFF_ALO = TC_A_BYTES;
    static constexpr int OFF_B = SPLIT ? 2 * TC_A_BYTES : TC_A_BYTES;
    static constexpr int OFF_BLO = OFF_B + B_BYTES;
    static constexpr int STAGE_BYTES = OFF_B + (SPLIT ? 2 : 1) * B_BYTES;
    static constexpr int TX_BYTES = TC_A_BYTES + (SPLIT ? 2 : 1) * B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = BAR_OFF + 256;
    static constexpr int TOTAL = BIAS_OFF + BLOCK_N * 4 + 16 + 1024;           // + alignment slack
    static_assert((3 * STAGES + 1) * 8 + 16 <= 256, "barrier block");
};
constexpr int TC_SPLIT_THREADS = 64;

template <int BLOCK_N, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ ConvTcParams P) {
    using S = TcSmem<BLOCK_N, STAGES, SPLIT>;
    constexpr int TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* ready = empty + STAGES;
    uint64_t* accum = ready + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(accum + 1);
    float* sbias = (float*)(smem + S::BIAS_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int b = blockIdx.x / tiles_per_img;
    const int tr = blockIdx.x - b * tiles_per_img;
    const int tyi = tr / P.tiles_x, txi = tr - tyi * P.tiles_x;
    const int TW = 1 << P.tw_log2, TH = 128 >> P.tw_log2;
    const int x0 = txi * TW, y0 = tyi * TH;
    const int n0 = blockIdx.y * BLOCK_N;
    stage_bias<BLOCK_N>(P, sbias, n0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); tc::mbar_init(&ready[i], TC_SPLIT_THREADS); }
        tc::mbar_init(accum, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&P.mapA[0]);
        tc::prefetch_tmap(&P.mapB);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nk = P.nk0 + P.nk1;
    const int total = P.ntaps * nk;

    if (warp == 0) {
        if (tc::elect_one()) {
            for (int it = 0; it < total; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                tc::mbar_wait(&empty[s], ph ^ 1u);
                tc::mbar_expect_tx(&full[s], S::TX_BYTES);
                const int tap = it / nk, kc = it - tap * nk;
                const int dy = P.ntaps == 9 ? tap / 3 : 1, dx = P.ntaps == 9 ? tap - (tap / 3) * 3 : 1;
                uint8_t* sA = smem + s * S::STAGE_BYTES;
                uint8_t* sB = sA + S::OFF_B;
                if (P.stride == 1) {
                    const bool first = kc < P.nk0;
                    tc::tma_load_4d(sA, first ? &P.mapA[0] : &P.mapA[1], &full[s], (first ? kc : kc - P.nk0) * P.kc,
                                    x0 + dx - 1, y0 + dy - 1, b);
                } else {
                    const int py = dy != 1, px = dx != 1;
                    tc::tma_load_4d(sA, &P.mapA[py * 2 + px], &full[s], kc * P.kc, x0 + (dx == 0 ? -1 : 0),
                                    y0 + (dy == 0 ? -1 : 0), b);
                }
                tc::tma_load_2d(sB, &P.mapB, &full[s], kc * P.kc, tap * P.cout_rows + n0);
                if constexpr (SPLIT) tc::tma_load_2d(sA + S::OFF_BLO, &P.mapBlo, &full[s], kc * P.kc, tap * P.cout_rows + n0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const bool bf16 = !SPLIT && P.bf16;      // bf16 operands: kind::f16, K = 16 per instruction (still 32 bytes)
            const uint32_t idesc = tc::make_idesc(bf16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, BLOCK_N);
            for (int it = 0; it < total; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                tc::mbar_wait(SPLIT ? &ready[s] : &full[s], ph);
                tc::tc_fence_after();
                const uint32_t sA = tc::smem_u32(smem + s * S::STAGE_BYTES);
                const uint64_t adesc = tc::smem_desc_k_sw128(sA), bdesc = tc::smem_desc_k_sw128(sA + S::OFF_B);
#pragma unroll
                for (int k = 0; k < 4; ++k) {        // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle atom
                    if (bf16) tc::umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((it | k) != 0));
                    else tc::umma_tf32(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((it | k) != 0));
                    if constexpr (SPLIT) {
                        const uint64_t alo = tc::smem_desc_k_sw128(sA + S::OFF_ALO), blo = tc::smem_desc_k_sw128(sA + S::OFF_BLO);
                        tc::umma_tf32(tmem_base, adesc + (uint64_t)(k * 2), blo + (uint64_t)(k * 2), idesc, 1u);
                        tc::umma_tf32(tmem_base, alo + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
                    }
                }
                tc::umma_commit(&empty[s]);          // frees the stage once these MMAs have read it
            }
            tc::umma_commit(accum);
        }
        __syncwarp();
    } else if constexpr (SPLIT) {
        // warps 2-3: split the activation tile of every stage into tf32 hi (in place) and lo (second buffer)
        const int st = threadIdx.x - 64;
        for (int it = 0; it < total; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            tc::mbar_wait(&full[s], ph);
            float4* a = reinterpret_cast<float4*>(smem + s * S::STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem + s * S::STAGE_BYTES + S::OFF_ALO);
#pragma unroll 4
            for (int i = st; i < TC_A_BYTES / 16; i += TC_SPLIT_THREADS) {
                const float4 v = a[i];
                float4 h, l;
                h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
                l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
                a[i] = h; lo[i] = l;
            }
            tc::fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
            tc::mbar_arrive(&ready[s]);
        }
    }

    // ---------------- epilogue: all four warps ----------------
    tc::mbar_wait(accum, 0);
    tc::tc_fence_after();
    tc_epilogue<BLOCK_N, SPLIT>(P, tmem_base, warp, lane, b, x0, y0, n0, TW, sbias);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}


// ================================================================================================
// Halo-reuse variant for stride-1 3x3 convolutions (the bulk of the FLOPs).
//
// The generic kernel above re-fetches a shifted 128-pixel activation tile for every tap: per 32-channel chunk it streams
// 9 x (16 KB + 16 KB) through shared memory and is bound by the L2 -> SM operand stream (profiles/r01_*: 1440 cycles per
// 256-cycle stage).  Here ONE halo tile [(8+2) rows x 32 pixels x 128 B] is loaded per chunk (one 4-D TMA box, OOB zero
// fill = padding) and all nine taps read it in place: a K-major SWIZZLE_128B UMMA operand may start at ANY 128-byte row of
// a TMA-written tile (the swizzle is a function of the absolute shared-memory address; verified on B200 by
// tools/experiments/shifted_desc.cu), so tap (dy,dx) of output sub-tile mt is simply the descriptor at row
// (4*mt + dy)*32 + dx.  Tile rows have a pitch of 32 pixels of which 30 are outputs (the two extra columns are computed and
// dropped), a CTA owns 8 x 30 outputs = two 128-row MMAs that share every weight tile.
// Operand bytes per chunk: 40 KB halo + 9 x 16 KB weights for 4608 MMA cycles = 40 B/cycle (generic kernel: 125 B/cycle).
// ================================================================================================
constexpr int HALO_RP = 32, HALO_TWV = 30, HALO_TH = 8;
constexpr int HALO_A_BYTES = (HALO_TH + 2) * HALO_RP * 128;          // 40 KB TMA box
constexpr int HALO_A_STRIDE = HALO_A_BYTES + 1024;                   // + the 2 pixels the last tap over-reads (garbage rows only)

template <int BLOCK_N, int NB>
struct HaloSmem {
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int OFF_B = 2 * HALO_A_STRIDE;
    static constexpr int BAR_OFF = OFF_B + NB * B_BYTES;
    static constexpr int BIAS_OFF = BAR_OFF + 256;
    static constexpr int TOTAL = BIAS_OFF + BLOCK_N * 4 + 16 + 1024;
    static_assert((2 * NB + 4 + 1) * 8 + 16 <= 256, "barrier block");
};
constexpr int HALO_THREADS = 256;                    // warps 0-3 / 4-7: epilogue of sub-tile 0 / 1 (warp % 4 = TMEM lane quarter)

template <int BLOCK_N, int NB>
__global__ void __launch_bounds__(HALO_THREADS)
conv_halo_kernel(const __grid_constant__ ConvTcParams P) {
    using S = HaloSmem<BLOCK_N, NB>;
    constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* a_full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_empty + 2;
    uint64_t* b_empty = b_full + NB;
    uint64_t* accum = b_empty + NB;
    uint32_t* tmem_slot = (uint32_t*)(accum + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int b = blockIdx.x / tiles_per_img;
    const int tr = blockIdx.x - b * tiles_per_img;
    const int tyi = tr / P.tiles_x, txi = tr - tyi * P.tiles_x;
    const int x0 = txi * HALO_TWV, y0 = tyi * HALO_TH;
    const int n0 = blockIdx.y * BLOCK_N;
    float* sbias = (float*)(smem + S::BIAS_OFF);
    stage_bias<BLOCK_N>(P, sbias, n0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NB; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
        tc::mbar_init(accum, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 0 && lane == 0) { tc::prefetch_tmap(&P.mapA[0]); tc::prefetch_tmap(&P.mapB); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nk = P.nk0 + P.nk1;

    if (warp == 0) {
        if (tc::elect_one()) {
            int j = 0;
            for (int kc = 0; kc < nk; ++kc) {
                const int sa = kc & 1;
                tc::mbar_wait(&a_empty[sa], (((uint32_t)kc >> 1) & 1u) ^ 1u);
                tc::mbar_expect_tx(&a_full[sa], HALO_A_BYTES);
                const bool first = kc < P.nk0;
                tc::tma_load_4d(smem + sa * HALO_A_STRIDE, first ? &P.mapA[0] : &P.mapA[1], &a_full[sa], (first ? kc : kc - P.nk0) * P.kc,
                                x0 - 1, y0 - 1, b);
                for (int tap = 0; tap < 9; ++tap, ++j) {
                    const int sb = j % NB;
                    tc::mbar_wait(&b_empty[sb], ((uint32_t)(j / NB) & 1u) ^ 1u);
                    tc::mbar_expect_tx(&b_full[sb], S::B_BYTES);
                    tc::tma_load_2d(smem + S::OFF_B + sb * S::B_BYTES, &P.mapB, &b_full[sb], kc * P.kc, tap * P.cout_rows + n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const bool bf16 = P.bf16;
            const uint32_t idesc = tc::make_idesc(bf16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, BLOCK_N);
            int j = 0;
            for (int kc = 0; kc < nk; ++kc) {
                const int sa = kc & 1;
                tc::mbar_wait(&a_full[sa], ((uint32_t)kc >> 1) & 1u);
                const uint32_t a_base = tc::smem_u32(smem + sa * HALO_A_STRIDE);
                for (int tap = 0; tap < 9; ++tap, ++j) {
                    const int sb = j % NB;
                    tc::mbar_wait(&b_full[sb], (uint32_t)(j / NB) & 1u);
                    tc::tc_fence_after();
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint64_t bdesc = tc::smem_desc_k_sw128(tc::smem_u32(smem + S::OFF_B + sb * S::B_BYTES));
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint64_t adesc = tc::smem_desc_k_sw128(a_base + (uint32_t)(((4 * mt + dy) * HALO_RP + dx) * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t acc = (uint32_t)((kc | tap | k) != 0);
                            if (bf16) tc::umma_f16(tmem_base + mt * BLOCK_N, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                            else tc::umma_tf32(tmem_base + mt * BLOCK_N, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                        }
                    }
                    tc::umma_commit(&b_empty[sb]);
                }
                tc::umma_commit(&a_empty[sa]);
            }
            tc::umma_commit(accum);
        }
        __syncwarp();
    }

    tc::mbar_wait(accum, 0);
    tc::tc_fence_after();
    {
        const int mt = warp >> 2;                    // two warpgroups drain the two accumulators concurrently
        tc_epilogue<BLOCK_N, false>(P, tmem_base + mt * BLOCK_N, warp & 3, lane, b, x0, y0 + 4 * mt, n0, HALO_TWV, sbias);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}


// ================================================================================================
// Persistent, warp-specialised variant of the per-tap kernel (the default for every non-split layer).
//
// Measured on B200 (tools/bench_conv.py): the one-tile-per-CTA kernel above sustains ~93 % tensor-pipe occupancy INSIDE its
// main loop but pays ~14 us of prologue / pipeline fill / epilogue per 128x128 tile, i.e. about as much as the 36-stage main
// loop of a 128-channel 3x3 layer.  Here one CTA per SM walks its tiles with
//   warp 0      TMA producer, 6-stage ring that keeps running across tile boundaries
//   warp 1      MMA issuer; accumulators alternate between two TMEM buffers
//   warps 4-7   epilogue of tile i (TMEM -> +bias/+residual -> NHWC) while the main loop of tile i+1 runs
// so the fixed costs are paid once per CTA and the epilogue is hidden.
// ================================================================================================
constexpr int PERS_THREADS = 256;
template <int BLOCK_N, int STAGES>
struct PersSmem {
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int BIAS_OFF = BAR_OFF + 256;
    static constexpr int MAX_COUT = 768;
    static constexpr int EPI_OFF = BIAS_OFF + MAX_COUT * 4;
    static constexpr int TOTAL = EPI_OFF + 4 * EPI_WARP_FLOATS * 4 + 16 + 1024;
    static_assert((2 * STAGES + 4) * 8 + 16 <= 256, "barrier block");
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(PERS_THREADS, 1)
conv_tc_persistent_kernel(const __grid_constant__ ConvTcParams P) {
    using S = PersSmem<BLOCK_N, STAGES>;
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;            // [2] accumulator ready for the epilogue
    uint64_t* tempty = tfull + 2;                // [2] accumulator drained (128 arrivals)
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* sbias = (float*)(smem + S::BIAS_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << P.tw_log2, TH = 128 >> P.tw_log2;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int n_ntiles = P.cout / BLOCK_N;
    const int total_tiles = tiles_per_img * P.batch * n_ntiles;
    const int nk = P.nk0 + P.nk1;
    const int iters_per_tile = P.ntaps * nk;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 128); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 0 && lane == 0) { tc::prefetch_tmap(&P.mapA[0]); tc::prefetch_tmap(&P.mapB); }
    {   // all bias (+ time-embedding) values of the layer -> smem, once per CTA
        const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
        const int bm = P.bias_mod ? P.bias_mod : P.cout;          // width-folded layers: output channel i is real channel i % bm
        for (int i = threadIdx.x; i < P.cout; i += PERS_THREADS) sbias[i] = bias ? __ldg(bias + i % bm) : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int tile, int& b, int& x0, int& y0, int& n0) {
        const int nt = tile % n_ntiles, mt = tile / n_ntiles;
        b = mt / tiles_per_img;
        const int tr = mt - b * tiles_per_img;
        const int tyi = tr / P.tiles_x, txi = tr - tyi * P.tiles_x;
        x0 = txi * TW; y0 = tyi * TH; n0 = nt * BLOCK_N;
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int it = 0;                                            // ring position, continuous across tiles
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
                for (int i = 0; i < iters_per_tile; ++i, ++it) {
                    const int s = it % STAGES;
                    tc::mbar_wait(&empty[s], ((uint32_t)(it / STAGES) & 1u) ^ 1u);
                    tc::mbar_expect_tx(&full[s], S::STAGE_BYTES);
                    const int tap = i / nk, kc = i - tap * nk;
                    const int dy = P.ntaps == 9 ? tap / 3 : 1, dx = P.ntaps == 9 ? tap - (tap / 3) * 3 : 1;
                    uint8_t* sA = smem + s * S::STAGE_BYTES;
                    if (P.stride == 1) {
                        const bool first = kc < P.nk0;
                        tc::tma_load_4d(sA, first ? &P.mapA[0] : &P.mapA[1], &full[s], (first ? kc : kc - P.nk0) * P.kc, x0 + dx - 1, y0 + dy - 1, b);
                    } else {
                        const int py = dy != 1, px = dx != 1;
                        tc::tma_load_4d(sA, &P.mapA[py * 2 + px], &full[s], kc * P.kc, x0 + (dx == 0 ? -1 : 0), y0 + (dy == 0 ? -1 : 0), b);
                    }
                    tc::tma_load_2d(sA + TC_A_BYTES, &P.mapB, &full[s], kc * P.kc, tap * P.cout_rows + n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const bool bf16 = P.bf16;
            const uint32_t idesc = tc::make_idesc(bf16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, BLOCK_N);
            int it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
                const int acc = tl & 1;
                tc::mbar_wait(&tempty[acc], (((uint32_t)tl >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int i = 0; i < iters_per_tile; ++i, ++it) {
                    const int s = it % STAGES;
                    tc::mbar_wait(&full[s], (uint32_t)(it / STAGES) & 1u);
                    tc::tc_fence_after();
                    const uint32_t sA = tc::smem_u32(smem + s * S::STAGE_BYTES);
                    const uint64_t adesc = tc::smem_desc_k_sw128(sA), bdesc = tc::smem_desc_k_sw128(sA + TC_A_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (bf16) tc::umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((i | k) != 0));
                        else tc::umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((i | k) != 0));
                    }
                    tc::umma_commit(&empty[s]);
                }
                tc::umma_commit(&tfull[acc]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        int tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            const int acc = tl & 1;
            int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
            tc::mbar_wait(&tfull[acc], ((uint32_t)tl >> 1) & 1u);
            tc::tc_fence_after();
            float* srow = nullptr;
            if (P.stats_out) {
                const int tr = tile / n_ntiles - b * tiles_per_img;                      // tile index inside the slice
                srow = P.stats_out + ((size_t)b * P.stats_rows + tr * 4 + (warp & 3)) * 2 * P.cout;
            }
            tc_epilogue_coalesced<BLOCK_N>(P, tmem_base + acc * ACC_COLS, warp & 3, lane, b, x0, y0, n0, sbias + n0,
                                           (float*)(smem + S::EPI_OFF) + (warp & 3) * EPI_WARP_FLOATS, 1 << P.tw_log2, srow);
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty[acc]);                           // 128 arrivals: the accumulator may be overwritten
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================
// Persistent halo-reuse kernel: the two ideas above combined, the default for stride-1 3x3 layers.
//
// The per-tap persistent kernel streams (16 KB A + B_BYTES B) per 4 MMAs: 75-80 B per MMA cycle and SM, while the L2 -> SM
// fabric delivers ~6300 B/cycle chip-wide = 42 B/cycle per SM (B300_MICROARCH "LTS throughput cap").  Measured: 57-66 % of
// the MMA rate on the large 128-channel layers, 42-54 % on the 64-channel ones, i.e. exactly L2-bound.  Here a CTA owns
// 8 x 30 output pixels (two 128-row accumulators): per K chunk ONE 40 KB halo tile and nine weight tiles feed 72 MMAs
// (24 B per MMA cycle at N = 128), weights are shared by both accumulators, and as in the persistent kernel
//   warp 0       TMA producer (A ring of 2 halo tiles, B ring of NB weight tiles, both continuous across tiles)
//   warp 1       MMA issuer; accumulator pairs alternate between two TMEM buffers (4 x BLOCK_N columns)
//   warps 4-11   two epilogue warpgroups, one per accumulator, draining tile i while tile i+1 is computed
// 30 of the 32 tile columns are outputs, so 6 % of the MMA work is discarded.
// ================================================================================================
constexpr int HP_THREADS = 384;
template <int BLOCK_N, int NB>
struct HaloPersSmem {
    static constexpr int B_BYTES = BLOCK_N * 128;
    static constexpr int OFF_B = 2 * HALO_A_STRIDE;
    static constexpr int BAR_OFF = OFF_B + NB * B_BYTES;
    static constexpr int BIAS_OFF = BAR_OFF + 512;
    static constexpr int MAX_COUT = 768;
    static constexpr int EPI_OFF = (BIAS_OFF + MAX_COUT * 4 + 1023) / 1024 * 1024;      // 8 x [32 px][128 B] SWIZZLE_128B tiles (EpilogueRow)
    static constexpr int TOTAL = EPI_OFF + 8 * EPI_ROW_BYTES + 1024;
    static_assert((2 * NB + 8) * 8 + 16 <= 512, "barrier block");
};

template <int BLOCK_N, int NB>
__global__ void __launch_bounds__(HP_THREADS, 1)
conv_halo_persistent_kernel(const __grid_constant__ ConvTcParams P) {
    using S = HaloPersSmem<BLOCK_N, NB>;
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int TMEM_COLS = 4 * ACC_COLS;                // 2 buffers x 2 accumulators
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* a_full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* a_empty = a_full + 2;
    uint64_t* b_full = a_empty + 2;
    uint64_t* b_empty = b_full + NB;
    uint64_t* tfull = b_empty + NB;              // [2] accumulator pair ready for the epilogue
    uint64_t* tempty = tfull + 2;                // [2] accumulator pair drained (256 arrivals)
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* sbias = (float*)(smem + S::BIAS_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int n_ntiles = P.cout / BLOCK_N;
    const int total_tiles = tiles_per_img * P.batch * n_ntiles;
    const int nk = P.nk0 + P.nk1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NB; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 256); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 0 && lane == 0) { tc::prefetch_tmap(&P.mapA[0]); tc::prefetch_tmap(&P.mapB); }
    {
        const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
        const int bm = P.bias_mod ? P.bias_mod : P.cout;
        for (int i = threadIdx.x; i < P.cout; i += HP_THREADS) sbias[i] = bias ? __ldg(bias + i % bm) : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int tile, int& b, int& x0, int& y0, int& n0) {
        const int nt = tile % n_ntiles, mt = tile / n_ntiles;
        b = mt / tiles_per_img;
        const int tr = mt - b * tiles_per_img;
        const int tyi = tr / P.tiles_x, txi = tr - tyi * P.tiles_x;
        x0 = txi * HALO_TWV; y0 = tyi * HALO_TH; n0 = nt * BLOCK_N;
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            int ia = 0, ib = 0;                                    // ring positions, continuous across tiles
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
                for (int kc = 0; kc < nk; ++kc, ++ia) {
                    const int sa = ia & 1;
                    tc::mbar_wait(&a_empty[sa], (((uint32_t)ia >> 1) & 1u) ^ 1u);
                    tc::mbar_expect_tx(&a_full[sa], HALO_A_BYTES);
                    const bool first = kc < P.nk0;
                    tc::tma_load_4d(smem + sa * HALO_A_STRIDE, first ? &P.mapA[0] : &P.mapA[1], &a_full[sa], (first ? kc : kc - P.nk0) * P.kc,
                                    x0 - 1, y0 - 1, b);
                    for (int tap = 0; tap < 9; ++tap) {
                        if (P.masked && !((P.kmask[tap] >> (4 * kc)) & 0xFull)) continue;     // structurally zero (tap, chunk): no tile, no MMAs
                        const int sb = ib % NB;
                        tc::mbar_wait(&b_empty[sb], ((uint32_t)(ib / NB) & 1u) ^ 1u);
                        tc::mbar_expect_tx(&b_full[sb], S::B_BYTES);
                        tc::tma_load_2d(smem + S::OFF_B + sb * S::B_BYTES, &P.mapB, &b_full[sb], kc * P.kc, tap * P.cout_rows + n0);
                        ++ib;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const bool bf16 = P.bf16;
            const uint32_t idesc = tc::make_idesc(bf16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, BLOCK_N);
            int ia = 0, ib = 0, tl = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
                const int buf = tl & 1;
                tc::mbar_wait(&tempty[buf], (((uint32_t)tl >> 1) & 1u) ^ 1u);     // both epilogue warpgroups have drained this pair
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 2 * ACC_COLS;
                uint32_t started = 0;                                                // the first MMA of a tile overwrites the accumulators
                for (int kc = 0; kc < nk; ++kc, ++ia) {
                    const int sa = ia & 1;
                    tc::mbar_wait(&a_full[sa], ((uint32_t)ia >> 1) & 1u);
                    const uint32_t a_base = tc::smem_u32(smem + sa * HALO_A_STRIDE);
                    if (!P.masked) {
                        // dense layers: straight-line issue, eight MMAs per tap (the masked form below costs a branch per k-step)
                        for (int tap = 0; tap < 9; ++tap, ++ib) {
                            const int sb = ib % NB;
                            tc::mbar_wait(&b_full[sb], (uint32_t)(ib / NB) & 1u);
                            tc::tc_fence_after();
                            const int dy = tap / 3, dx = tap - dy * 3;
                            const uint64_t bdesc = tc::smem_desc_k_sw128(tc::smem_u32(smem + S::OFF_B + sb * S::B_BYTES));
#pragma unroll
                            for (int mt = 0; mt < 2; ++mt) {
                                const uint64_t adesc = tc::smem_desc_k_sw128(a_base + (uint32_t)(((4 * mt + dy) * HALO_RP + dx) * 128));
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t acc = (uint32_t)((kc | tap | k) != 0);
                                    if (bf16) tc::umma_f16(d_tmem + mt * ACC_COLS, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                                    else tc::umma_tf32(d_tmem + mt * ACC_COLS, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                                }
                            }
                            tc::umma_commit(&b_empty[sb]);
                        }
                        tc::umma_commit(&a_empty[sa]);
                        continue;
                    }
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t km = (uint32_t)((P.kmask[tap] >> (4 * kc)) & 0xFull);
                        if (!km) continue;
                        const int sb = ib % NB;
                        tc::mbar_wait(&b_full[sb], (uint32_t)(ib / NB) & 1u);
                        tc::tc_fence_after();
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint64_t bdesc = tc::smem_desc_k_sw128(tc::smem_u32(smem + S::OFF_B + sb * S::B_BYTES));
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint64_t adesc = tc::smem_desc_k_sw128(a_base + (uint32_t)(((4 * mt + dy) * HALO_RP + dx) * 128));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (!((km >> k) & 1u)) continue;
                                const uint32_t acc = (started >> mt) & 1u;
                                if (bf16) tc::umma_f16(d_tmem + mt * ACC_COLS, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                                else tc::umma_tf32(d_tmem + mt * ACC_COLS, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                                started |= 1u << mt;
                            }
                        }
                        tc::umma_commit(&b_empty[sb]);
                        ++ib;
                    }
                    tc::umma_commit(&a_empty[sa]);
                }
                tc::umma_commit(&tfull[buf]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int mt = (warp - 4) >> 2;                              // warpgroup <-> accumulator (tile rows 4*mt .. 4*mt+3)
        float* stile = (float*)(smem + S::EPI_OFF + (warp - 4) * EPI_ROW_BYTES);
        int tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            const int buf = tl & 1;
            int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
            EpilogueRow<BLOCK_N> er;
            er.setup(P, lane, b, x0, y0 + 4 * mt + (warp & 3), n0, HALO_TWV);
            er.load_res(P, 0);
            tc::mbar_wait(&tfull[buf], ((uint32_t)tl >> 1) & 1u);
            tc::tc_fence_after();
            float* srow = nullptr;
            if (P.stats_out) {
                const int tr = tile / n_ntiles - b * tiles_per_img;
                srow = P.stats_out + ((size_t)b * P.stats_rows + (tr * 2 + mt) * 4 + (warp & 3)) * 2 * P.cout;
            }
            er.run(P, tmem_base + (buf * 2 + mt) * ACC_COLS, warp & 3, lane, sbias + n0, stile, srow);
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty[buf]);
        }
        if (lane == 0) tc::tma_store_wait_all();                     // the staging tile must outlive its last bulk store
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================
// GroupNorm-fused persistent halo kernel: the conv reads the RAW fp32 residual-stream tensor(s) and applies the GroupNorm
// affine + SiLU of the reference's `norm -> SiLU -> conv` (Model/model.py:98-101, 110-113) on the operand path, so the
// separate apply pass (4 B read + 2-4 B written per element, 18 % of the step in round 1) and the operand tensor disappear.
//
//   warp 2          TMA producer of the RAW halo tiles: ring of 2 [10 rows][32 px][32 ch fp32 = 128 B] (one 4-D box per 32-channel
//                   K chunk of the virtual concat, OOB zero fill)
//   warp 0          TMA producer of the weight tiles (ring of NB)
//   warps 12-19     transform: raw tile -> y = silu(x * scale[slice][c] + shift[slice][c]), zero outside the image (the conv's
//                   zero padding applies to the normalised activation), rounded exactly like the unfused apply pass:
//                     BF16: bf16, written as a K-major SWIZZLE_64B operand tile (64-byte pixel rows) into a second ring of 2
//                     TF32: tf32, written back IN PLACE (the TMA tile is already a SWIZZLE_128B operand tile)
//                   then fence.proxy.async + arrive on `a_ready`
//   warp 1          MMA issuer: as conv_halo_persistent_kernel, operands from the transformed tile
//   warps 4-11      two epilogue warpgroups (unchanged)
// The per-(slice, channel) scale / shift come from gn_finalize (statistics from the producer conv's epilogue).  The operand values equal
// those of the unfused apply pass up to the last bit of the SiLU (exponent argument as one fma; bf16 mode: tanh.approx form) -- within the
// operand rounding that follows, and checked against torch in tests/test_unet_kernels_gpu.py::test_tc_conv_fused_groupnorm.
// ================================================================================================
constexpr int HF_THREADS = 640;       // warps 0 weights TMA, 1 MMA, 2 TMEM alloc + raw-tile TMA, 4-11 epilogue, 12-19 transform
constexpr int HF_TWARPS = 8;
constexpr int HF_RAW_BYTES = HALO_A_BYTES;                            // 40 KB: [10][32 px][128 B]
constexpr int HF_OP_BYTES_BF16 = (HALO_TH + 2) * HALO_RP * 64;        // 20 KB: [10][32 px][64 B]
constexpr int HF_OP_STRIDE_BF16 = HF_OP_BYTES_BF16 + 1024;            // + the 2 pixels the last tap over-reads
template <int BLOCK_N, int NB, bool BF16, int NA>
struct HaloFusedSmem {
    static constexpr int B_BYTES = BLOCK_N * (BF16 ? 64 : 128);
    static constexpr int RAW_STRIDE = BF16 ? HF_RAW_BYTES : HALO_A_STRIDE;
    static constexpr int OFF_OP = NA * RAW_STRIDE;                     // BF16 only
    static constexpr int OFF_B = OFF_OP + (BF16 ? 2 * HF_OP_STRIDE_BF16 : 0);
    static constexpr int BAR_OFF = OFF_B + NB * B_BYTES;
    static constexpr int BIAS_OFF = BAR_OFF + 512;
    static constexpr int MAX_COUT = 512;
    static constexpr int EPI_OFF = (BIAS_OFF + MAX_COUT * 4 + 1023) / 1024 * 1024;      // 8 x [32 px][128 B] SWIZZLE_128B tiles (EpilogueRow)
    static constexpr int TOTAL = EPI_OFF + 8 * EPI_ROW_BYTES + 1024;
    static_assert((2 * NB + 4 * NA + 6) * 8 + 16 <= 512, "barrier block");

};

// NA = depth of the raw-tile ring (TF32: the tiles are transformed in place, so it is also the operand ring).  Layers with few K
// chunks per tile (the width-folded thin layers: one or three) are bound by HBM latency x bytes in flight and take NA = 4.
template <int BLOCK_N, int NB, bool BF16, int NA>
__global__ void __launch_bounds__(HF_THREADS, 1)
conv_halo_fused_kernel(const __grid_constant__ ConvTcParams P) {
    using S = HaloFusedSmem<BLOCK_N, NB, BF16, NA>;
    constexpr int ACC_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr int TMEM_COLS = 4 * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* raw_full = (uint64_t*)(smem + S::BAR_OFF);   // [NA] TMA landed a raw tile
    uint64_t* raw_empty = raw_full + NA;                   // [NA] BF16: the transform warps have read it (one arrival per warp)
    uint64_t* a_ready = raw_empty + NA;                    // [NA] operand tile written (one arrival per warp)
    uint64_t* a_empty = a_ready + NA;                      // [NA] MMAs have read the operand tile
    uint64_t* b_full = a_empty + NA;
    uint64_t* b_empty = b_full + NB;
    uint64_t* tfull = b_empty + NB;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    float* sbias = (float*)(smem + S::BIAS_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int n_ntiles = P.cout / BLOCK_N;
    const int ph_log2 = P.ph_log2;                                     // 2: upsample conv as four output-parity phases (see conv_tc_prepare)
    const int total_tiles = (tiles_per_img * P.batch * n_ntiles) << ph_log2;
    const int nk = P.nk0 + P.nk1 + P.nk2;

    if (threadIdx.x == 0) {
        const uint32_t issuers = (!BF16 && P.masked) ? 2u : 1u;        // masked (width-folded) layers: one MMA issuer per accumulator
        for (int i = 0; i < NA; ++i) { tc::mbar_init(&raw_full[i], 1); tc::mbar_init(&raw_empty[i], HF_TWARPS); tc::mbar_init(&a_ready[i], HF_TWARPS); tc::mbar_init(&a_empty[i], issuers); }
        for (int i = 0; i < NB; ++i) { tc::mbar_init(&b_full[i], 1); tc::mbar_init(&b_empty[i], issuers); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], issuers); tc::mbar_init(&tempty[i], 256); }
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 0 && lane == 0) { tc::prefetch_tmap(&P.mapA[0]); tc::prefetch_tmap(&P.mapB); }
    {
        const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
        const int bm = P.bias_mod ? P.bias_mod : P.cout;
        for (int i = threadIdx.x; i < P.cout; i += HF_THREADS) sbias[i] = bias ? __ldg(bias + i % bm) : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // 640 threads leave 96 registers per thread; the producer / issuer warpgroup needs far fewer and hands 32 per thread to each of the
    // two epilogue warpgroups, which then hold a whole 32-column accumulator chunk plus its residual lines (setmaxnreg, warpgroup-wide)

    // tile = blockIdx.x, blockIdx.x + gridDim.x, ...: (N tile, tile column, tile row, slice) advanced with adds and compares -- the four
    // integer divisions of a fresh decode cost ~90 instructions per tile in each of the 20 warps (ncu, width-folded layers)
    struct TileWalk {
        int nt, txi, tyi, b, s_nt, s_tx, s_ty, s_b, n_nt, n_tx, n_ty;
        __device__ __forceinline__ void split(int v, int& nt_, int& tx_, int& ty_, int& b_) const {
            nt_ = v % n_nt; v /= n_nt; tx_ = v % n_tx; v /= n_tx; ty_ = v % n_ty; b_ = v / n_ty;
        }
        __device__ __forceinline__ void init(int first, int step, int n_ntiles_, int tiles_x_, int tiles_y_) {
            n_nt = n_ntiles_; n_tx = tiles_x_; n_ty = tiles_y_;
            split(first, nt, txi, tyi, b); split(step, s_nt, s_tx, s_ty, s_b);
        }
        __device__ __forceinline__ void next() {
            nt += s_nt; if (nt >= n_nt) { nt -= n_nt; ++txi; }
            txi += s_tx; if (txi >= n_tx) { txi -= n_tx; ++tyi; }
            tyi += s_ty; if (tyi >= n_ty) { tyi -= n_ty; ++b; }
            b += s_b;
        }
    } walk;
    walk.init(blockIdx.x, gridDim.x, n_ntiles << ph_log2, P.tiles_x, P.tiles_y);      // fastest index = N tile x phase
    auto decode = [&](int, int& b, int& x0, int& y0, int& n0) {       // coordinates of the walker's current tile; callers advance it
        b = walk.b; x0 = walk.txi * HALO_TWV; y0 = walk.tyi * HALO_TH; n0 = (walk.nt >> ph_log2) * BLOCK_N;
    };
    auto phase = [&]() { return walk.nt & ((1 << ph_log2) - 1); };
    const bool is_transform = warp >= 12;

    if (warp < 4) {
    if (!BF16 && P.masked && (warp == 0 || warp == 1 || warp == 3)) {
        // Width-folded layers: 36 ... 100 small MMAs (N = 32 / 64, some k-steps structurally zero) per tile, so the kernel is bound by
        // how fast ONE thread can walk the issue loop (ncu: ~900 cycles per tap with the generic loop below).  Lean variant: taps
        // unrolled, masks in registers, ring positions kept incrementally, 32-bit descriptor arithmetic, predicated MMAs, and one
        // issuer warp per accumulator (warp 1: tile rows 0-3, warp 3: rows 4-7; every consumer barrier counts two arrivals).
        if (tc::elect_one()) {
            // (rolled tap loops on purpose: the unrolled form was 12 KB of straight-line code per issuer and spent 58 % of its cycles in
            // instruction-fetch stalls; the masks come from the constant bank, the tap's row offset is kept incrementally)
            int sb = 0; uint32_t phb = 0;
            if (warp == 0) {
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, walk.next()) {
                    const int n0 = (walk.nt >> ph_log2) * BLOCK_N;
                    for (int kc = 0; kc < nk; ++kc) {
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            if (!((uint32_t)(P.kmask[tap] >> (4 * kc)) & 0xFu)) continue;
                            tc::mbar_wait(&b_empty[sb], phb ^ 1u);
                            tc::mbar_expect_tx(&b_full[sb], S::B_BYTES);
                            tc::tma_load_2d(smem + S::OFF_B + sb * S::B_BYTES, &P.mapB, &b_full[sb], kc * 32, tap * P.cout_rows + n0);
                            if (++sb == NB) { sb = 0; phb ^= 1u; }
                        }
                    }
                }
            } else {
                const uint32_t mt = warp == 1 ? 0u : 1u;
                const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, BLOCK_N);
                const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(smem)) + mt * (4u * HALO_RP * 128u / 16u);
                const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(smem + S::OFF_B));
                int sa = 0; uint32_t pha = 0, tl = 0;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
                    const uint32_t buf = tl & 1u;
                    tc::mbar_wait(&tempty[buf], ((tl >> 1) & 1u) ^ 1u);
                    tc::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (buf * 2u + mt) * ACC_COLS;
                    uint32_t acc = 0;
                    for (int kc = 0; kc < nk; ++kc) {
                        tc::mbar_wait(&a_ready[sa], pha);
                        uint32_t a_tap = a_lo0 + (uint32_t)sa * (S::RAW_STRIDE / 16);      // descriptor of tap (0, 0); + 8 per pixel, + 256 per tile row
                        int dx = 0;
#pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t km = (uint32_t)(P.kmask[tap] >> (4 * kc)) & 0xFu;
                            if (km) {
                                tc::mbar_wait(&b_full[sb], phb);
                                tc::tc_fence_after();
                                const uint32_t b_tap = b_lo0 + (uint32_t)sb * (S::B_BYTES / 16);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t en = (km >> k) & 1u;
                                    tc::umma_tf32_lo(d_tmem, a_tap + 2u * k, b_tap + 2u * k, idesc, acc, en);
                                    acc |= en;
                                }
                                tc::umma_commit(&b_empty[sb]);
                                if (++sb == NB) { sb = 0; phb ^= 1u; }
                            }
                            a_tap += 8u;
                            if (++dx == 3) { dx = 0; a_tap += (HALO_RP - 3) * 8u; }
                        }
                        tc::umma_commit(&a_empty[sa]);
                        if (++sa == NA) { sa = 0; pha ^= 1u; }
                    }
                    tc::umma_commit(&tfull[buf]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        // weight tiles only.  The raw activation tiles have their own producer (warp 14): issued from this loop they would queue behind
        // the nine weight loads of the previous chunk, i.e. until the MMAs of that chunk start, and land one TMA latency + one
        // transform too late (measured: 0.82 ms instead of 0.53 for 128 -> 128 at 16 x 500 x 228)
        if (tc::elect_one()) {
            int ib = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, walk.next()) {
                int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
                for (int kc = 0; kc < nk; ++kc) {
                    const int ph = phase();
                    const uint32_t tm = P.tapmask[ph];
                    for (int tap = 0; tap < 9; ++tap) {
                        if (kc >= P.nk_gn && tap != 4) continue;              // shortcut chunk: centre tap only
                        if (!((tm >> tap) & 1u)) continue;                    // upsample phase: four of the nine tap positions
                        const int sb = ib % NB;
                        tc::mbar_wait(&b_empty[sb], ((uint32_t)(ib / NB) & 1u) ^ 1u);
                        tc::mbar_expect_tx(&b_full[sb], S::B_BYTES);
                        tc::tma_load_2d(smem + S::OFF_B + sb * S::B_BYTES, &P.mapB, &b_full[sb], kc * 32, (ph * 9 + tap) * P.cout_rows + n0);
                        ++ib;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        if (tc::elect_one()) {
            int ia = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, walk.next()) {
                int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
                for (int kc = 0; kc < nk; ++kc, ++ia) {
                    const int sa = ia % NA;
                    // the raw buffer is free once the transform warps have read it (BF16) / once the MMAs have read the in-place tile (TF32)
                    tc::mbar_wait(BF16 ? &raw_empty[sa] : &a_empty[sa], ((uint32_t)(ia / NA) & 1u) ^ 1u);
                    tc::mbar_expect_tx(&raw_full[sa], HF_RAW_BYTES);
                    const int si = kc < P.nk0 ? 0 : (kc < P.nk0 + P.nk1 ? 1 : 2);
                    const int kl = kc - (si == 0 ? 0 : (si == 1 ? P.nk0 : P.nk0 + P.nk1));
                    tc::tma_load_4d(smem + sa * S::RAW_STRIDE, &P.mapA[si], &raw_full[sa], kl * 32, x0 - 1, y0 - 1, b);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc(BF16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, BLOCK_N);
            constexpr int ROWB = BF16 ? 64 : 128;              // bytes per pixel row of the operand tile
            constexpr int KSTEPS = ROWB / 32;                  // 32 bytes of K per MMA
            int ia = 0, ib = 0, tl = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl, walk.next()) {
                const int buf = tl & 1;
                tc::mbar_wait(&tempty[buf], (((uint32_t)tl >> 1) & 1u) ^ 1u);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 2 * ACC_COLS;
                for (int kc = 0; kc < nk; ++kc, ++ia) {                              // (dense layers only: masked ones take the lean path above)
                    // operand slot: the bf16 form has two operand buffers whatever the depth NA of the raw ring; tf32 tiles are transformed in place
                    const int sa = BF16 ? (ia & 1) : ia % NA;
                    tc::mbar_wait(&a_ready[sa], (uint32_t)(BF16 ? ia >> 1 : ia / NA) & 1u);
                    const uint32_t a_base = tc::smem_u32(smem + (BF16 ? S::OFF_OP + sa * HF_OP_STRIDE_BF16 : sa * S::RAW_STRIDE));
                    const uint32_t tm = P.tapmask[phase()];
                    const int tap0 = __ffs(tm) - 1;                          // the first MMA of a tile overwrites the accumulators
                    for (int tap = 0; tap < 9; ++tap) {
                        if (kc >= P.nk_gn && tap != 4) continue;              // shortcut chunk: centre tap only
                        if (!((tm >> tap) & 1u)) continue;                    // upsample phase: four of the nine tap positions
                        const int sb = ib % NB;
                        tc::mbar_wait(&b_full[sb], (uint32_t)(ib / NB) & 1u);
                        tc::tc_fence_after();
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint32_t b_addr = tc::smem_u32(smem + S::OFF_B + sb * S::B_BYTES);
                        const uint64_t bdesc = BF16 ? tc::smem_desc_k(b_addr, 4, 512) : tc::smem_desc_k_sw128(b_addr);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint32_t a_addr = a_base + (uint32_t)(((4 * mt + dy) * HALO_RP + dx) * ROWB);
                            const uint64_t adesc = BF16 ? tc::smem_desc_k(a_addr, 4, 512) : tc::smem_desc_k_sw128(a_addr);
#pragma unroll
                            for (int k = 0; k < KSTEPS; ++k) {
                                const uint32_t acc = (uint32_t)((kc | (tap - tap0) | k) != 0);
                                if (BF16) tc::umma_f16(d_tmem + mt * ACC_COLS, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                                else tc::umma_tf32(d_tmem + mt * ACC_COLS, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc);
                            }
                        }
                        tc::umma_commit(&b_empty[sb]);
                        ++ib;
                    }
                    tc::umma_commit(&a_empty[sa]);
                }
                tc::umma_commit(&tfull[buf]);
            }
        }
        __syncwarp();
    }
    } else if (is_transform) {
        const int w4 = warp - 12;                                    // 0..7: 40 of the 320 tile pixels each
        const int c8 = lane & 7, psub = lane >> 3;                   // 16-byte chunk (4 channels) of the pixel row; pixel within a group of 4
        const int Ctot = P.gn_m0 + P.gn_m1;                          // real channels of the GroupNorm (== gn_c0 + gn_c1 unless width-folded)
        constexpr int PER = 320 / HF_TWARPS / 4;                     // pixels per lane and chunk
        const bool work = P.gn_act != 3 && P.gn_act != 4;            // 4: the tile already is the operand (tf32-rounded, no GroupNorm): hand it on
        // per-lane byte offsets of its PER pixels inside a raw tile / a bf16 operand tile: fixed for the whole kernel
        uint32_t roff[PER], ooff[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int r = w4 * (4 * PER) + j * 4 + psub;             // pixel of the halo tile: row r >> 5, column r & 31
            roff[j] = (uint32_t)r * 128u + (uint32_t)((c8 ^ (r & 7)) << 4);                                       // TMA SWIZZLE_128B: 16-byte chunk ^ (row & 7)
            ooff[j] = (uint32_t)r * 64u + (uint32_t)((((c8 >> 1) ^ ((r >> 1) & 3)) << 4) | ((c8 & 1) << 3));      // SWIZZLE_64B row: chunk (c8 >> 1) ^ ((r >> 1) & 3), half c8 & 1
        }
        constexpr float NL2E = -1.4426950408889634f;
        int ia = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, walk.next()) {
            int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
            // which of this lane's pixels lie inside the image (the conv pads the ACTIVATION with zeros): once per tile, not per chunk
            uint32_t inmask = 0;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int r = w4 * (4 * PER) + j * 4 + psub;
                const int yy = y0 - 1 + (r >> 5), xx = x0 - 1 + (r & 31);
                inmask |= ((unsigned)yy < (unsigned)P.H && (unsigned)xx < (unsigned)P.W ? 1u : 0u) << j;
            }
            for (int kc = 0; kc < nk; ++kc, ++ia) {
                const int sa = ia % NA;                                              // raw slot
                const int so = BF16 ? (ia & 1) : sa;                                 // operand slot (bf16: two buffers; tf32: in place)
                const uint32_t pho = (uint32_t)(BF16 ? ia >> 1 : ia / NA) & 1u;
                // per-lane affine of its 4 channels (pad channels of the source tensors: scale = shift = 0 -> silu(0) = 0)
                float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
                const bool ident = kc >= P.nk_gn || P.gn_act == 5;                   // shortcut chunk / plain conv: x itself, rounded to the operand type
                if (ident) sc = make_float4(1.f, 1.f, 1.f, 1.f);
                else {
                    const bool first = kc < P.nk0;
                    const int cl = (first ? kc : kc - P.nk0) * 32 + 4 * c8;          // channel inside its source
                    const int cg = first ? cl % P.gn_m0 : P.gn_m0 + cl % max(P.gn_m1, 1);   // channel of the GroupNorm (virtual concat; folded: pixel-major)
                    if (P.gn_act != 4 && cl < (first ? P.gn_c0 : P.gn_c1)) {
                        sc = __ldg(reinterpret_cast<const float4*>(P.gn_scale + (size_t)b * Ctot + cg));
                        sh = __ldg(reinterpret_cast<const float4*>(P.gn_shift + (size_t)b * Ctot + cg));
                    }
                }
                const bool do_silu = P.gn_act == 1 && !ident;
                // exponent argument of the sigmoid straight from x: -log2(e) * (x * sc + sh) as ONE fma with pre-scaled coefficients
                const float4 sce = make_float4(sc.x * NL2E, sc.y * NL2E, sc.z * NL2E, sc.w * NL2E), she = make_float4(sh.x * NL2E, sh.y * NL2E, sh.z * NL2E, sh.w * NL2E);
                // one warp watches the barrier, the other seven sleep on a named barrier: eight pollers cost a third of the SM's issue slots
                if (w4 == 0) tc::mbar_wait_idle(&raw_full[sa], (uint32_t)(ia / NA) & 1u);
                tc::named_bar_sync(1, HF_TWARPS * 32);
                const uint32_t raw_s = tc::smem_u32(smem + sa * S::RAW_STRIDE);
                const uint32_t op_s = tc::smem_u32(smem + S::OFF_OP + so * HF_OP_STRIDE_BF16);
                // A lane owns 10 pixels x 4 channels of the chunk: all ten loads are issued before the first SiLU so that the shared-memory and
                // MUFU latencies overlap (explicit ld/st.shared: generic accesses made the compiler serialise load -> store -> load).  BF16: the
                // raw slot goes back to the TMA producer as soon as the values sit in registers -- the 64-channel layers are bound by bytes in
                // flight (two 40 KB slots per SM) -- and only then the warp waits for its operand buffer.
                float4 v[PER];
                if (work) {
#pragma unroll
                    for (int j = 0; j < PER; ++j) v[j] = tc::lds128(raw_s + roff[j]);
                }
                if (BF16) {
                    // The slot goes back to the TMA producer only when the loaded values HAVE ARRIVED in registers: ld.shared is asynchronous
                    // and mbarrier.arrive does not wait for it, so every lane first consumes one word of each of its loads (an intermittent
                    // whole-tile corruption in the bf16 mode was traced to the producer overwriting a tile with loads still in flight).
                    if (work) {
                        uint32_t dep = 0;
#pragma unroll
                        for (int j = 0; j < PER; ++j) dep ^= __float_as_uint(v[j].x) ^ __float_as_uint(v[j].w);
                        asm volatile("" :: "r"(dep) : "memory");
                    }
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&raw_empty[sa]);                      // raw tile consumed (its values live in registers)
                    if (w4 == 0) tc::mbar_wait_idle(&a_empty[so], pho ^ 1u);                            // the operand buffer of this stage is free
                    tc::named_bar_sync(1, HF_TWARPS * 32);
                }
                if (work) {
                    // (the SiLU / no-SiLU decision is uniform for the chunk: two straight-line loops instead of a branch per pixel)
                    auto emit = [&](int j, float4 o) {
                        const bool inside = (inmask >> j) & 1u;
                        if (BF16) {
                            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                            tc::sts64_or_zero(op_s + ooff[j], *reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi), inside);
                        } else {
                            // (round to nearest tf32 = add half an ulp of the 10-bit mantissa; the tensor core drops the low 13 bits itself)
                            tc::sts128_or_zero(raw_s + roff[j], make_float4(tf32_rn_hw(o.x), tf32_rn_hw(o.y), tf32_rn_hw(o.z), tf32_rn_hw(o.w)), inside);
                        }
                    };
                    if (do_silu && BF16 && P.silu_tanh) {
                        // bf16 operands: silu(y) = h + h * tanh(h), h = y / 2 -- ONE MUFU (tanh.approx, |error| <= 2^-11) and two FMAs per element
                        // instead of ex2 + rcp and four; the approximation error stays below the bf16 rounding of the operand that follows
                        const float4 sch = make_float4(0.5f * sc.x, 0.5f * sc.y, 0.5f * sc.z, 0.5f * sc.w), shh = make_float4(0.5f * sh.x, 0.5f * sh.y, 0.5f * sh.z, 0.5f * sh.w);
#pragma unroll
                        for (int j = 0; j < PER; ++j) {
                            float4 h;
                            h.x = fmaf(v[j].x, sch.x, shh.x); h.y = fmaf(v[j].y, sch.y, shh.y); h.z = fmaf(v[j].z, sch.z, shh.z); h.w = fmaf(v[j].w, sch.w, shh.w);
                            emit(j, make_float4(fmaf(h.x, tanh_approx(h.x), h.x), fmaf(h.y, tanh_approx(h.y), h.y), fmaf(h.z, tanh_approx(h.z), h.z), fmaf(h.w, tanh_approx(h.w), h.w)));
                        }
                    } else if (do_silu) {
#pragma unroll
                        for (int j = 0; j < PER; ++j) {
                            float4 y, e;
                            y.x = fmaf(v[j].x, sc.x, sh.x); y.y = fmaf(v[j].y, sc.y, sh.y); y.z = fmaf(v[j].z, sc.z, sh.z); y.w = fmaf(v[j].w, sc.w, sh.w);
                            e.x = ex2_approx(fmaf(v[j].x, sce.x, she.x)); e.y = ex2_approx(fmaf(v[j].y, sce.y, she.y));
                            e.z = ex2_approx(fmaf(v[j].z, sce.z, she.z)); e.w = ex2_approx(fmaf(v[j].w, sce.w, she.w));
                            emit(j, make_float4(y.x * rcp_approx(1.0f + e.x), y.y * rcp_approx(1.0f + e.y), y.z * rcp_approx(1.0f + e.z), y.w * rcp_approx(1.0f + e.w)));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < PER; ++j)
                            emit(j, make_float4(fmaf(v[j].x, sc.x, sh.x), fmaf(v[j].y, sc.y, sh.y), fmaf(v[j].z, sc.z, sh.z), fmaf(v[j].w, sc.w, sh.w)));
                    }
                }
                tc::fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&a_ready[so]);
            }
        }
    } else if (warp >= 4 && warp < 12) {
        const int mt = (warp - 4) >> 2;
        float* stile = (float*)(smem + S::EPI_OFF + (warp - 4) * EPI_ROW_BYTES);
        int tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl, walk.next()) {
            const int buf = tl & 1;
            int b, x0, y0, n0; decode(tile, b, x0, y0, n0);
            EpilogueRow<BLOCK_N, 32> er;
            er.setup(P, lane, b, x0, y0 + 4 * mt + (warp & 3), n0, HALO_TWV, phase());
            er.load_res(P, 0);
            if ((warp & 3) == 0) tc::mbar_wait_idle(&tfull[buf], ((uint32_t)tl >> 1) & 1u);
            tc::named_bar_sync(2 + mt, 128);
            tc::tc_fence_after();
            float* srow = nullptr;
            if (P.stats_out) {
                const int tr = ((walk.tyi * P.tiles_x + walk.txi) << ph_log2) + phase();
                srow = P.stats_out + ((size_t)b * P.stats_rows + (tr * 2 + mt) * 4 + (warp & 3)) * 2 * P.cout;
            }
            er.run(P, tmem_base + (buf * 2 + mt) * ACC_COLS, warp & 3, lane, sbias + n0, stile, srow);
            tc::tc_fence_before();
            tc::mbar_arrive(&tempty[buf]);
        }
        if (lane == 0) tc::tma_store_wait_all();                     // the staging tile must outlive its last bulk store
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int NB, bool BF16, int NA = 2>
static int launch_halo_fused(const ConvTcParams& P, cudaStream_t st) {
    static DeviceOnce once;
    constexpr int smem = HaloFusedSmem<BN, NB, BF16, NA>::TOTAL;
    static_assert(smem <= 227 * 1024, "fused halo rings do not fit in shared memory");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_fused_kernel<BN, NB, BF16, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int total = (P.tiles_x * P.tiles_y * P.batch * (P.cout / BN)) << P.ph_log2;
    conv_halo_fused_kernel<BN, NB, BF16, NA><<<std::min(total, kNumSMs), HF_THREADS, smem, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

template <int BN, int NB>
static int launch_halo_pers(const ConvTcParams& P, cudaStream_t st) {
    static DeviceOnce once;
    constexpr int smem = HaloPersSmem<BN, NB>::TOTAL;
    static_assert(smem <= 227 * 1024, "persistent halo rings do not fit in shared memory");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_persistent_kernel<BN, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int total = P.tiles_x * P.tiles_y * P.batch * (P.cout / BN);
    conv_halo_persistent_kernel<BN, NB><<<std::min(total, kNumSMs), HP_THREADS, smem, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

template <int BN, int ST>
static int launch_pers(const ConvTcParams& P, cudaStream_t st) {
    static DeviceOnce once;
    constexpr int smem = PersSmem<BN, ST>::TOTAL;
    static_assert(smem <= 227 * 1024, "persistent ring does not fit in shared memory");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_persistent_kernel<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int total = P.tiles_x * P.tiles_y * P.batch * (P.cout / BN);
    conv_tc_persistent_kernel<BN, ST><<<std::min(total, kNumSMs), PERS_THREADS, smem, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int pick_tw_log2(int H, int W) {
    int best = 3; long long best_area = -1;
    for (int l = 3; l <= 7; ++l) {
        const int tw = 1 << l, th = 128 >> l;
        const long long area = (long long)ceil_div(W, tw) * tw * ceil_div(H, th) * th;
        if (best_area < 0 || area < best_area || (area == best_area && l > best)) { best = l; best_area = area; }
    }
    return best;
}

int conv_tc_stats_rows_bound(int h, int w) {
    const int halo = ceil_div(w, HALO_TWV) * ceil_div(h, HALO_TH) * 2, tap = ceil_div(w, 8) * ceil_div(h, 16);   // pick_tw_log2 never does worse than 8x16
    const int phases = 4 * ceil_div((w + 1) / 2, HALO_TWV) * ceil_div((h + 1) / 2, HALO_TH) * 2;                // upsample conv: four phases of the half-size tiling
    return 4 * std::max(std::max(halo, tap), phases);
}

// auto choice of the halo-reuse kernels (see conv_tc_prepare): stride-1 3x3, enough tiles to fill the machine, and an 8x30 tiling that wastes
// < 20 % of the MMA rows -- or a narrow (C_out < 64) layer
static bool halo_auto(int H, int W, int batch, int cout, int ntaps, int stride) {
    const long long htx = ceil_div(W, HALO_TWV), hty = ceil_div(H, HALO_TH);
    const bool halo_ok = stride == 1 && ntaps == 9 && htx * hty * batch >= kNumSMs / 2;
    const bool halo_fits = (double)W * H >= 0.8 * (double)(htx * HALO_TWV) * (double)(hty * HALO_TH);
    return halo_ok && (cout < 64 || halo_fits);
}
// true when a GroupNorm(+SiLU) -> conv pair of this shape runs as ONE conv_halo_fused_kernel launch (plan builder, unet.cu)
bool conv_tc_can_fuse_norm(int H, int W, int batch, int cout, int ntaps, int stride) {
    static const bool off = getenv("IPDM_GN_FUSE") && atoi(getenv("IPDM_GN_FUSE")) == 0;
    static const int env_variant = getenv("IPDM_CONV_VARIANT") ? atoi(getenv("IPDM_CONV_VARIANT")) : 0;
    return !off && env_variant == 0 && cout >= 64 && cout <= 512 && cout % 64 == 0 && halo_auto(H, W, batch, cout, ntaps, stride);
}

int conv_tc_prepare(ConvTcParams& P, const ConvTcDesc& d) {
    IPDM_REQUIRE(d.nsrc >= 1 && d.nsrc <= (d.n_ident ? 3 : 2) && d.n_ident >= 0 && d.n_ident < d.nsrc, "conv_tc: 1 or 2 sources (+ shortcut sources)");
    IPDM_REQUIRE(d.stride == 1 || (d.stride == 2 && d.nsrc == 1), "conv_tc: stride 2 takes one source");
    IPDM_REQUIRE(d.ntaps == 1 || d.ntaps == 9, "conv_tc: 1x1 or 3x3");
    memset(&P, 0, sizeof(P));
    const int Hin = d.src[0].h, Win = d.src[0].w;
    P.H = d.stride == 1 ? Hin : (Hin + 1) / 2;      // k=3, pad=1, stride 2 -> floor((H-1)/2)+1
    P.W = d.stride == 1 ? Win : (Win + 1) / 2;
    P.batch = d.src[0].n;
    // kernel choice (d.variant: 0 auto, 1 one-tile-per-CTA, 2 halo-reuse, 3 persistent per-tap, 4 persistent halo-reuse).  The 3xTF32
    // split and the qkv epilogue exist only in the one-tile-per-CTA kernel; everything else is persistent.
    static const int env_variant = getenv("IPDM_CONV_VARIANT") ? atoi(getenv("IPDM_CONV_VARIANT")) : 0;
    const int variant = d.variant ? d.variant : env_variant;
    const bool plain = !d.w_packed_lo && !d.qkv_mode && d.cout <= 768;
    // auto: stride-1 3x3 layers take a halo-reuse kernel when its 8x30 tiling wastes < 20 % of the MMA rows (measured with
    // tools/bench_conv.py at 16 slices, bf16: 788 -> 1081 TFLOP/s at 500x228x128, 394 -> 558 at 512x512x64; 64x64 and 32x32 images
    // lose 25-30 % to the tiling and stay per-tap).  N = 16 layers (144 -> 16 at 1000x456) have almost no epilogue: the
    // one-tile-per-CTA halo kernel with two CTAs per SM is the fastest there (1.59 ms against 1.85 persistent, 2.88 per-tap).
    const bool narrow = d.cout < 64 && d.n_tile != 32;      // (the N = 32 tile of the width-folded layers is persistent)
    const long long htx = ceil_div(P.W, HALO_TWV), hty = ceil_div(P.H, HALO_TH);
    const bool halo_ok = plain && d.stride == 1 && d.ntaps == 9 && (htx * hty * P.batch >= kNumSMs / 2 || d.fold);
    P.halo = halo_ok && (d.fold || variant == 2 || variant == 4 || (variant == 0 && halo_auto(P.H, P.W, P.batch, d.cout, d.ntaps, d.stride)));
    P.persistent = plain && (d.fold || (P.halo ? (variant == 4 || (variant == 0 && !narrow)) : (variant == 0 || variant == 3 || variant == 4)));
    P.tw_log2 = P.halo ? 5 : pick_tw_log2(P.H, P.W);
    const int TW = P.halo ? HALO_RP : 1 << P.tw_log2, TH = P.halo ? HALO_TH + 2 : 128 >> P.tw_log2;     // TMA box extent
    P.tiles_x = P.halo ? ceil_div(P.W, HALO_TWV) : ceil_div(P.W, TW);
    P.tiles_y = P.halo ? ceil_div(P.H, HALO_TH) : ceil_div(P.H, TH);
    P.ntaps = d.ntaps; P.stride = d.stride;
    P.cout = d.cout; P.block_n = d.n_tile ? d.n_tile : (d.cout >= 128 ? 128 : (d.cout >= 64 ? 64 : 16));
    IPDM_REQUIRE(P.block_n == 16 || P.block_n == 32 || P.block_n == 64 || P.block_n == 128, "conv_tc: N tile %d", P.block_n);
    IPDM_REQUIRE(P.block_n != 32 || P.persistent, "conv_tc: the N = 32 tile exists in the persistent kernels only");
    IPDM_REQUIRE(d.cout % P.block_n == 0, "conv_tc: C_out %d not a multiple of the N tile %d", d.cout, P.block_n);
    P.cout_rows = d.cout;
    int ktot = 0;
    // fused GroupNorm: the sources are RAW fp32 tensors whatever the operand type (the kernel converts on the operand path)
    P.fused = d.norm_scale != nullptr || d.passthrough || d.phase_up;
    for (int ph = 0; ph < 4; ++ph) P.tapmask[ph] = 0x1FFu;
    if (d.phase_up) {
        IPDM_REQUIRE(!d.norm_scale && !d.passthrough && d.nsrc == 1 && d.ntaps == 9 && d.stride == 1 && P.halo && P.persistent && P.block_n >= 64 && !d.res.p,
                     "conv_tc: the upsample phases run on a persistent halo layer (one source, 3x3, C_out >= 64, no residual)");
        P.gn_act = 5; P.gn_c0 = d.src[0].c; P.gn_c1 = 0; P.gn_m0 = P.gn_c0; P.gn_m1 = 0;
        P.ph_log2 = 2;
        for (int ph = 0; ph < 4; ++ph) {
            const int py = ph >> 1, px = ph & 1;
            P.tapmask[ph] = 0;
            for (int a = 0; a < 2; ++a)
                for (int b2 = 0; b2 < 2; ++b2) P.tapmask[ph] |= 1u << ((py + a) * 3 + px + b2);
        }
    } else if (d.passthrough) {
        // the conv_halo_fused_kernel pipeline (deep raw ring, lean masked MMA issue) with an identity operand path: tf32 tiles only
        IPDM_REQUIRE(!d.norm_scale && !d.w_bf16 && !d.src[0].bf16 && P.halo && P.persistent && P.block_n >= 32, "conv_tc: passthrough takes tf32-rounded fp32 sources on a persistent halo layer");
        P.gn_act = 4; P.gn_c0 = d.src[0].c; P.gn_c1 = d.nsrc > 1 ? d.src[1].c : 0; P.gn_m0 = P.gn_c0; P.gn_m1 = P.gn_c1;
    } else if (P.fused) {
        IPDM_REQUIRE(P.halo && P.persistent && P.block_n >= 32 && d.cout <= 512 && d.norm_shift, "conv_tc: GroupNorm fusion needs a persistent halo layer (3x3, stride 1, C_out 32..512)");
        const int ngn = d.nsrc - d.n_ident;                               // sources under the GroupNorm; the rest is the folded shortcut's input
        IPDM_REQUIRE(ngn == 1 || d.src[0].c == d.src[0].cs, "conv_tc: fused concat needs an unpadded first source (%d channels, stride %d)", d.src[0].c, d.src[0].cs);
        IPDM_REQUIRE(d.act_silu, "conv_tc: the fused operand path is GroupNorm + SiLU (the attention norm has no activation and feeds a 1x1 conv)");
        P.gn_scale = d.norm_scale; P.gn_shift = d.norm_shift; P.gn_act = d.act_silu;
        // experiments (tools only): IPDM_FUSE_DBG=2 affine without SiLU, 3 the transform warps only hand the barriers on (operands are garbage)
        static const int fuse_dbg = getenv("IPDM_FUSE_DBG") ? atoi(getenv("IPDM_FUSE_DBG")) : 0;
        if (fuse_dbg) P.gn_act = fuse_dbg;
        P.gn_c0 = d.src[0].c; P.gn_c1 = ngn > 1 ? d.src[1].c : 0;
        P.gn_m0 = d.gn_mod[0] ? d.gn_mod[0] : P.gn_c0; P.gn_m1 = ngn > 1 ? (d.gn_mod[1] ? d.gn_mod[1] : P.gn_c1) : 0;
        IPDM_REQUIRE(P.gn_c0 % 4 == 0 && P.gn_c1 % 4 == 0 && P.gn_m0 % 4 == 0 && (P.gn_c1 == 0 || P.gn_m1 % 4 == 0),
                     "conv_tc: fused GroupNorm needs channel counts that are multiples of 4");
    }
    P.bf16 = P.fused ? d.w_bf16 : d.src[0].bf16;
    P.kc = P.fused ? 32 : (P.bf16 ? 64 : 32);
    const int eb = (P.bf16 && !P.fused) ? 2 : 4;                      // element size of the activation tensors the TMA reads
    const CUtensorMapDataType dt = (P.bf16 && !P.fused) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    IPDM_REQUIRE(!(P.bf16 && d.w_packed_lo), "conv_tc: bf16 operands and the 3xTF32 split are exclusive");
    for (int s = 0; s < d.nsrc; ++s) {
        const TensorNHWC& t = d.src[s];
        IPDM_REQUIRE(t.bf16 == (P.bf16 && !P.fused), "conv_tc: concat sources differ in dtype");
        IPDM_REQUIRE(t.cs % P.kc == 0 && ((uintptr_t)t.p % 16) == 0, "conv_tc: source channel stride %d must be a multiple of %d", t.cs, P.kc);
        IPDM_REQUIRE(t.h == Hin && t.w == Win && t.n == P.batch, "conv_tc: concat sources differ in shape");
        (s == 0 ? P.nk0 : (s == 1 ? P.nk1 : P.nk2)) = t.cs / P.kc;
        if (s < d.nsrc - d.n_ident) P.nk_gn += t.cs / P.kc;
        ktot += t.cs;
        if (d.stride == 1) {
            const uint64_t dims[4] = {(uint64_t)t.cs, (uint64_t)t.w, (uint64_t)t.h, (uint64_t)t.n};
            const uint64_t str[3] = {(uint64_t)t.cs * eb, (uint64_t)t.w * t.cs * eb, (uint64_t)t.h * t.w * t.cs * eb};
            const uint32_t box[4] = {(uint32_t)P.kc, (uint32_t)TW, (uint32_t)TH, 1};
            IPDM_CHECK(tmap_encode(&P.mapA[s], dt, 4, t.p, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
        } else {
            for (int py = 0; py < 2; ++py)
                for (int px = 0; px < 2; ++px) {
                    const uint64_t wp = (uint64_t)(t.w - px + 1) / 2, hp = (uint64_t)(t.h - py + 1) / 2;
                    const uint64_t dims[4] = {(uint64_t)t.cs, wp ? wp : 1, hp ? hp : 1, (uint64_t)t.n};
                    const uint64_t str[3] = {(uint64_t)t.cs * 2 * eb, (uint64_t)t.w * t.cs * 2 * eb, (uint64_t)t.h * t.w * t.cs * eb};
                    const uint32_t box[4] = {(uint32_t)P.kc, (uint32_t)TW, (uint32_t)TH, 1};
                    const char* base = (const char*)t.p + ((size_t)py * t.w + px) * t.cs * eb;
                    IPDM_CHECK(tmap_encode(&P.mapA[py * 2 + px], dt, 4, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
                }
        }
    }
    // (fused bf16: the weights keep their operand-tensor K padding to 64; the trailing all-zero 32-column chunk is simply not walked)
    IPDM_REQUIRE(ktot == d.w_k || (P.fused && ktot <= d.w_k), "conv_tc: packed weight K %d != sum of source channel strides %d", d.w_k, ktot);
    {
        const int ebw = P.bf16 ? 2 : 4;
        const CUtensorMapDataType dtw = P.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        const uint64_t dims[2] = {(uint64_t)d.w_k, (uint64_t)d.ntaps * d.cout * (d.phase_up ? 4 : 1)};
        const uint64_t str[1] = {(uint64_t)d.w_k * ebw};
        const uint32_t box[2] = {(uint32_t)P.kc, (uint32_t)P.block_n};
        IPDM_CHECK(tmap_encode(&P.mapB, dtw, 2, d.w_packed, dims, str, box, (P.fused && P.bf16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B));
        P.split = d.w_packed_lo != nullptr;
        if (P.split) IPDM_CHECK(tmap_encode(&P.mapBlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d.w_packed_lo, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
    }
    P.out = d.out.p; P.out_cs = d.out.cs;
    const int up = d.phase_up ? 2 : 1;
    IPDM_REQUIRE(d.qkv_mode || (d.out.h == up * P.H && d.out.w == up * P.W && d.out.n == P.batch && d.out.cs % 4 == 0 && d.out.c >= d.cout),
                 "conv_tc: output tensor shape mismatch");
    P.bias = d.bias; P.bias_t_stride = d.bias_t_stride; P.t_dev = d.t_dev;
    P.res = d.res.p; P.res_cs = d.res.cs;
    P.qkv_mode = d.qkv_mode; P.vt = d.vt; P.t_pad = d.t_pad; P.heads = d.heads; P.head_dim = d.head_dim;
    P.out_lo = d.out_lo; P.vt_lo = d.vt_lo; P.qkv_bf16 = d.qkv_bf16;
    IPDM_REQUIRE(!d.n_ident || (P.fused && !d.passthrough && d.stride == 1), "conv_tc: a folded shortcut needs the GroupNorm-fused halo kernel");
    P.bias_mod = d.bias_mod; P.fold = d.fold;
    static const bool silu_tanh = !(getenv("IPDM_SILU_TANH") && atoi(getenv("IPDM_SILU_TANH")) == 0);
    P.silu_tanh = silu_tanh;
    P.masked = 0;
    for (int t = 0; t < 9; ++t) { P.kmask[t] = d.kmask[t]; if (d.kmask[t]) P.masked = 1; }
    IPDM_REQUIRE(!P.masked || (P.halo && P.persistent && !P.bf16 && d.ntaps == 9 && P.nk0 + P.nk1 + P.nk2 <= 16),
                 "conv_tc: k-step masks are a feature of the tf32 persistent halo kernels (<= 16 K chunks)");
    {   // MMA work actually issued (masked k-steps are skipped)
        double ksteps = 0;
        const int nkk = P.nk0 + P.nk1 + P.nk2, per = P.kc * (P.bf16 ? 2 : 4) / 32;      // k-steps (32 bytes of K) per chunk
        for (int t = 0; t < d.ntaps; ++t)
            for (int kc = 0; kc < nkk; ++kc) {
                if (!P.masked && kc >= P.nk_gn && t != 4) continue;               // shortcut chunks: centre tap only
                if (P.ph_log2 && !((0x1Bu >> t) & 1u)) continue;                  // upsample phases: 4 of the 9 taps each (x 4 phases below)
                const unsigned km = P.masked ? (unsigned)((P.kmask[t] >> (4 * kc)) & 0xF) : (1u << per) - 1u;
                ksteps += __builtin_popcount(km);
            }
        const int kel = P.bf16 ? 16 : 8;                                          // K elements per k-step
        // an upsample conv counts as the 3x3 conv on the upsampled image that it replaces (the algorithmic work of the layer):
        // 9 taps x 4 output pixels per source pixel; the tensor pipe issues 16/36 of that
        P.flops = 2.0 * P.batch * P.H * P.W * (double)P.cout * ksteps * kel * (P.ph_log2 ? 9.0 : 1.0);
    }
    // TMA-store epilogue of the persistent halo kernels: a [30 px][32 ch] box of one output row per warp and chunk (EpilogueRow)
    static const bool tma_store_off = getenv("IPDM_TMA_STORE") && atoi(getenv("IPDM_TMA_STORE")) == 0;
    P.tma_store = 0;
    if (!tma_store_off && P.halo && P.persistent && !d.phase_up && !d.qkv_mode && P.block_n >= 32 && d.cout % 32 == 0 && ((uintptr_t)d.out.p % 16) == 0) {
        const uint64_t od[4] = {(uint64_t)d.out.cs, (uint64_t)d.out.w, (uint64_t)d.out.h, (uint64_t)d.out.n};
        const uint64_t os[3] = {(uint64_t)d.out.cs * 4, (uint64_t)d.out.w * d.out.cs * 4, (uint64_t)d.out.h * d.out.w * d.out.cs * 4};
        const uint32_t ob[4] = {32, HALO_TWV, 1, 1};
        IPDM_CHECK(tmap_encode(&P.mapOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.out.p, od, os, ob, CU_TENSOR_MAP_SWIZZLE_128B));
        P.tma_store = 1;
    }
    P.stats_out = P.persistent ? d.stats_out : nullptr;              // only the persistent kernels' epilogue produces statistics
    P.stats_rows = P.stats_out ? (P.tiles_x * P.tiles_y * (P.halo ? 2 : 1) * 4) << P.ph_log2 : 0;
    IPDM_REQUIRE(!P.stats_out || (d.cout % 4 == 0 && P.stats_rows <= conv_tc_stats_rows_bound(up * P.H, up * P.W)), "conv_tc: statistics rows %d exceed the bound", P.stats_rows);
    IPDM_REQUIRE(!d.qkv_bf16 || (d.qkv_mode && !P.split && d.t_pad % 8 == 0 && d.out.cs % 8 == 0), "conv_tc: the bf16 qkv epilogue needs t_pad and the qk channel stride to be multiples of 8");
    IPDM_REQUIRE(!(d.qkv_mode && P.split) || (d.out_lo && d.vt_lo), "conv_tc: the fp32-mode qkv epilogue needs out_lo and vt_lo");
    return IPDM_OK;
}

template <int BN, int ST, bool SPLIT>
static int launch_tc(const ConvTcParams& P, cudaStream_t st) {
    static DeviceOnce once;
    constexpr int smem = TcSmem<BN, ST, SPLIT>::TOTAL;
    static_assert(smem <= 227 * 1024, "stage ring does not fit in shared memory");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, ST, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    dim3 grid(P.tiles_x * P.tiles_y * P.batch, P.cout / BN);
    conv_tc_kernel<BN, ST, SPLIT><<<grid, TC_THREADS, smem, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

template <int BN, int NB>
static int launch_halo(const ConvTcParams& P, cudaStream_t st) {
    static DeviceOnce once;
    constexpr int smem = HaloSmem<BN, NB>::TOTAL;
    static_assert(smem <= 227 * 1024, "halo ring does not fit in shared memory");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    dim3 grid(P.tiles_x * P.tiles_y * P.batch, P.cout / BN);
    conv_halo_kernel<BN, NB><<<grid, HALO_THREADS, smem, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

int conv_tc_launch(const ConvTcParams& P, cudaStream_t st) {
    // padded-K FLOPs actually issued to the tensor pipe; the persistent halo kernel (the dominant kernel of the step) is its own family
    // (width-folded thin layers are HBM-bound streaming layers: they are accounted in bytes with the other thin / direct convs)
    // algorithmic HBM bytes of the launch: every source once (raw fp32 when fused, else the operand type), the residual, the output
    const double px = (double)P.batch * P.H * P.W, opx = px * (P.ph_log2 ? 4.0 : 1.0);
    const double abytes = px * (P.nk0 + P.nk1 + P.nk2) * P.kc * ((P.bf16 && !P.fused) ? 2.0 : 4.0) + opx * P.cout * 4.0 * (P.res ? 2.0 : 1.0);
    ProfScope prof(P.fold ? PROF_CONV_DIRECT : (P.halo && P.persistent ? PROF_CONV_HALO_PERS : PROF_CONV_TC), st,
                   P.fold ? 4.0 * P.batch * (double)P.H * P.W * ((P.nk0 + P.nk1 + P.nk2) * P.kc + P.cout) : conv_tc_flops(P), abytes);
    if (P.fused) {
        if (P.bf16) {
            // N = 64 (image net, HBM-bound: 3.2 GB per launch): a third raw slot instead of half of the weight ring -- bytes in flight bound these layers
            // (IPDM_RAW3=1, measured at 16 x 512 x 512: 64 -> 64 0.735 -> 0.751 ms, 128 -> 64 1.811 -> 1.787 ms: these layers are not bound by
            // bytes in flight but by the transform stage, so the deeper weight ring stays the default)
            static const bool raw3 = getenv("IPDM_RAW3") && atoi(getenv("IPDM_RAW3")) == 1;
            if (P.block_n == 128) return launch_halo_fused<128, 8, true>(P, st);
            if (P.block_n == 64) return raw3 ? launch_halo_fused<64, 6, true, 3>(P, st) : launch_halo_fused<64, 12, true>(P, st);
        }
        else {
            if (P.block_n == 128) return launch_halo_fused<128, 6, false>(P, st);
            if (P.block_n == 64) return P.masked ? launch_halo_fused<64, 6, false, 3>(P, st) : launch_halo_fused<64, 10, false>(P, st);
            if (P.block_n == 32) return launch_halo_fused<32, 4, false, 4>(P, st);
        }
        set_error("conv_tc_launch: no fused kernel for N tile %d", P.block_n);
        return IPDM_ERR_UNSUPPORTED;
    }
    if (P.halo && P.persistent) {
        switch (P.block_n) {
            case 128: return launch_halo_pers<128, 6>(P, st);
            case 64: return launch_halo_pers<64, 10>(P, st);
            case 32: return launch_halo_pers<32, 12>(P, st);
            case 16: return launch_halo_pers<16, 12>(P, st);
        }
    }
    if (P.halo) {
        switch (P.block_n) {
            case 128: return launch_halo<128, 6>(P, st);
            case 64: return launch_halo<64, 8>(P, st);
            case 16: return launch_halo<16, 8>(P, st);
        }
    }
    if (P.persistent) {
        switch (P.block_n) {
            case 128: return launch_pers<128, 5>(P, st);
            case 64: return launch_pers<64, 7>(P, st);
            case 32: return launch_pers<32, 8>(P, st);
            case 16: return launch_pers<16, 9>(P, st);
        }
    }
    if (P.split) {
        switch (P.block_n) {
            case 128: return launch_tc<128, 3, true>(P, st);
            case 64: return launch_tc<64, 4, true>(P, st);
            case 16: return launch_tc<16, 4, true>(P, st);
        }
    } else {
        switch (P.block_n) {
            case 128: return launch_tc<128, 3, false>(P, st);
            case 64: return launch_tc<64, 4, false>(P, st);
            case 16: return launch_tc<16, 4, false>(P, st);
        }
    }
    set_error("conv_tc_launch: unsupported N tile %d", P.block_n);
    return IPDM_ERR_UNSUPPORTED;
}

double conv_tc_flops(const ConvTcParams& P) { return (P.split ? 3.0 : 1.0) * P.flops; }

}  // namespace ipdm
