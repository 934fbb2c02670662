// Internal op interfaces of the UNet noise predictor (host launchers; kernels live in *.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstring>

namespace ipdm {

// Activation tensor: NHWC, `cs` >= c elements per pixel (pad channels hold zeros); fp32, or bf16 for tensors that exist
// only as tensor-core operands in the bf16 precision mode (GroupNorm-apply and upsample outputs).
struct TensorNHWC {
    float* p = nullptr;
    int n = 0, h = 0, w = 0, c = 0, cs = 0;
    int bf16 = 0;
    size_t pixels() const { return (size_t)n * h * w; }
    size_t elems() const { return pixels() * cs; }
};

// ---- tensor-core implicit GEMM convolution (conv_tc.cu) ---------------------------------------------
struct ConvTcDesc {
    int nsrc = 1;
    TensorNHWC src[3];          // virtual concat along channels; all padded to multiples of 32 (three only with n_ident, see below)
    int ntaps = 9, stride = 1;
    int cout = 0;
    const float* w_packed = nullptr;   // [ntaps][cout][w_k] (tf32-rounded; the hi part in fp32 mode)
    const float* w_packed_lo = nullptr;// fp32 (3xTF32) mode: the lo part, same layout; selects the split kernel
    int w_k = 0;
    const float* bias = nullptr;       // [cout] or table [max_t][bias_t_stride]
    int bias_t_stride = 0;
    const int* t_dev = nullptr;        // device scalar: current timestep (selects the bias row)
    TensorNHWC res;                    // optional residual, same pixels as the output
    TensorNHWC out;
    int qkv_mode = 0; float* vt = nullptr; int t_pad = 0, heads = 0, head_dim = 0;
    float* out_lo = nullptr; float* vt_lo = nullptr;   // qkv epilogue in fp32 mode: q,k,v are written as tf32 hi / lo pairs
    int qkv_bf16 = 0;                                  // qkv epilogue writes q,k (out) and v^T (vt) as bf16, t_pad % 8 == 0
    int variant = 0;                                   // 0 auto, 1 one-tile-per-CTA, 2 halo-reuse, 3 persistent, 4 persistent halo (tests / benchmarks)
    // GroupNorm statistics of the OUTPUT, fused into the persistent kernels' epilogue: per (slice, 32-pixel warp row) partial sums
    // [batch][stats_rows][2][cout] (sum, sum of squares; fp32 over 32 pixels, reduced in fp64 by gn_finalize).  nullptr: off.
    float* stats_out = nullptr;
    // fused GroupNorm(+SiLU) of the input (conv_halo_fused_kernel): src[] are the RAW fp32 tensors, y = act(x*scale[n][c] + shift[n][c]) is
    // applied on the operand path; w_bf16 says whether w_packed (and hence the MMA) is bf16 or tf32
    const float* norm_scale = nullptr; const float* norm_shift = nullptr; int act_silu = 1; int w_bf16 = 0;
    // width-folded thin layers (unet.cu pack_conv_folded): a dense [H][W][C] tensor with C in {4, 8, 16, ...} is read as
    // [H][W/f][f*C]; src / res / out are those VIEWS.  kmask[tap] holds 4 bits per 32-channel K chunk: the 8-column k-steps of that
    // (tap, chunk) whose packed weights are structurally non-zero (0 = no mask); gn_mod / bias_mod = real channel counts behind the
    // folded ones (per-channel GroupNorm affine and bias are indexed modulo them); n_tile forces the N tile (0 auto)
    uint64_t kmask[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int gn_mod[2] = {0, 0}; int bias_mod = 0; int n_tile = 0; int fold = 0;
    // 1x1 shortcut folded into the conv (model.py:130 `h + self.shortcut(x)`): the LAST n_ident sources are the ResBlock input x, read
    // raw (no GroupNorm) on the operand path and contracted with the centre tap only -- w_packed carries the shortcut weights in the
    // K columns of those sources at tap 4 and zeros at the other taps, bias = conv bias + shortcut bias (conv_halo_fused_kernel only)
    int n_ident = 0;
    // Upsample (nearest, exactly 2x) + 3x3 conv as four output-parity phases on the LOW-resolution source (model.py:163-170): output
    // pixel (2i+py, 2j+px) only sees source rows {i-1+py, i+py} and columns {j-1+px, j+px}, so each phase is a 2x2-tap conv whose
    // weights are sums of the 3x3 taps that land on the same source pixel: 16/36 of the MMAs, no upsampled tensor.  src = the raw fp32
    // low-res tensor (converted on the operand path), out = the high-res tensor, w_packed = [phase 4][tap 9][cout][K]
    int phase_up = 0;
    int passthrough = 0;               // run a plain (no GroupNorm) tf32 layer through conv_halo_fused_kernel with an identity operand path
};
int conv_tc_stats_rows_bound(int h, int w);
bool conv_tc_can_fuse_norm(int H, int W, int batch, int cout, int ntaps, int stride);            // upper bound of stats_rows for an h x w output, any kernel variant

struct ConvTcParams {
    CUtensorMap mapA[4];
    CUtensorMap mapB, mapBlo;
    CUtensorMap mapOut; int tma_store;                 // halo kernels: the epilogue stores [30 px][32 ch] boxes of the output by TMA
    int split, bf16, kc, halo, persistent;
    int H, W, tiles_x, tiles_y, tw_log2, batch;
    int ntaps, stride, nk0, nk1, nk2, nk_gn;           // K chunks per source; chunks >= nk_gn are identity (shortcut) chunks
    int cout, cout_rows, block_n;
    float* out; int out_cs;
    const float* bias; int bias_t_stride; const int* t_dev;
    const float* res; int res_cs;
    int qkv_mode; float* vt; int t_pad, heads, head_dim;
    float* out_lo; float* vt_lo; int qkv_bf16;
    float* stats_out; int stats_rows;                  // set by prepare only for the persistent kernels (else nullptr / 0)
    int fused; const float* gn_scale; const float* gn_shift; int gn_c0, gn_c1, gn_act;    // conv_halo_fused_kernel
    uint64_t kmask[9]; int gn_m0, gn_m1, bias_mod, masked, fold;                          // width-folded thin layers
    int silu_tanh;                                                                        // bf16 operand path: SiLU through tanh.approx (one MUFU)
    int ph_log2; uint32_t tapmask[4];                                                     // upsample conv phases: tap positions used by each
    double flops;                                                                         // MMA work issued per launch (prepare)
};

int conv_tc_prepare(ConvTcParams& P, const ConvTcDesc& d);
int conv_tc_launch(const ConvTcParams& P, cudaStream_t st);
double conv_tc_flops(const ConvTcParams& P);

// ---- tensor-core convolution for thin layers: C_in <= 32, C_out in {8, 16}, 3x3 / 1x1 stride 1 (conv_thin.cu) ------
struct ConvThinDesc {
    TensorNHWC src;                    // fp32 operand tensor, channel stride 8, 16 or 32 (pad channels zero)
    int ntaps = 9, cout = 0;
    const float* w_packed = nullptr;   // [ntaps][16][src.cs], tf32-rounded, rows >= cout and columns >= C_in zero
    const float* bias = nullptr; int bias_t_stride = 0; const int* t_dev = nullptr;
    TensorNHWC res, out;
    // optional fused GroupNorm(+SiLU) of the source: y = act(x*scale[n][c] + shift[n][c]), rounded to TF32, applied to the TMA-landed
    // tile in shared memory by two helper warps (src is then the RAW activation tensor, dense: cs == C_in)
    const float* norm_scale = nullptr; const float* norm_shift = nullptr; int act_silu = 1;
};
struct ConvThinParams {
    CUtensorMap mapA, mapW;
    int H, W, tiles_x, tiles_y, batch, ntaps, cs, cout, rows;
    float* out; int out_cs;
    const float* bias; int bias_t_stride; const int* t_dev;
    const float* res; int res_cs;
    const float* norm_scale; const float* norm_shift; int act_silu;
    int dbg;                           // IPDM_THIN_DBG experiments (tools only): 1 no stores, 2 no residual loads, 4 no MMAs, 8 no operand loads
};
int conv_thin_prepare(ConvThinParams& P, const ConvThinDesc& d);
int conv_thin_launch(const ConvThinParams& P, cudaStream_t st);

// ---- direct (CUDA-core) convolution for thin layers (unet_kernels.cu) -------------------------------
struct ConvDirectDesc {
    int nsrc = 1;
    TensorNHWC src[2];                 // virtual concat
    const float* norm_scale = nullptr; // optional fused GroupNorm+SiLU on load: y = silu(x*scale[n][c] + shift[n][c])
    const float* norm_shift = nullptr; //   both [batch][cin_total]
    int ksize = 3, stride = 1;
    int upsample = 0;                  // 1: sources are nearest-upsampled to (out.h, out.w) before the conv
    int cin = 0, cout = 0;
    const float* w = nullptr;          // [ksize*ksize][cin][cout]; C_out > 16 is processed in tiles of 16
    const float* bias = nullptr; int bias_t_stride = 0; const int* t_dev = nullptr;
    TensorNHWC res;
    TensorNHWC out;                    // pad channels (c..cs) are written as zeros
};
int conv_direct_launch(const ConvDirectDesc& d, cudaStream_t st);
double conv_direct_flops(const ConvDirectDesc& d);

// ---- GroupNorm (unet_kernels.cu) ---------------------------------------------------------------------
// statistics over (H, W, C/G) per slice and group -> per-(slice, channel) affine y = x*scale + shift
struct GroupNormDesc {
    int nsrc = 1;
    TensorNHWC src[2];                 // virtual concat
    int groups = 0;
    const float* gamma = nullptr;      // [C]
    const float* beta = nullptr;
    float eps = 1e-5f;
    float* scale = nullptr;            // [batch][C] out
    float* shift = nullptr;
    double* partials = nullptr;        // workspace [batch][GN_MAX_BLOCKS][C][2]
    // per source: statistics already produced by the conv that wrote the tensor (ConvTcDesc::stats_out), [batch][rows][2][c];
    // such a source is not read again
    const float* tile_stats[2] = {nullptr, nullptr};
    int tile_rows[2] = {0, 0};
    int tile_fold[2] = {0, 0};         // > 1: the producer was width-folded, its rows hold [2][fold][c] (see gn_tile_reduce_kernel)
};
constexpr int GN_MAX_BLOCKS = 888;        // 6 CTAs per SM
int groupnorm_stats_launch(const GroupNormDesc& d, cudaStream_t st);
// out[n][y][x][c] = act(src*scale + shift) (c < C; act = SiLU or identity), 0 for pad channels; out.cs may exceed C
int groupnorm_apply_launch(const GroupNormDesc& d, const TensorNHWC& out, int act_silu, int round_tf32, cudaStream_t st);

// nearest-neighbour resize (F.interpolate(mode="nearest") index rule), NHWC
int upsample_nearest_launch(const TensorNHWC& src, const TensorNHWC& dst, int round_tf32, cudaStream_t st);

// ---- attention (attention.cu) --------------------------------------------------------------------------
struct AttentionDesc {
    const float* qk = nullptr;  // NHWC [B][T][3*C]: per head h, q at channel 3*d*h, k at 3*d*h + d
    const float* vt = nullptr;  // [B][heads][d][t_pad]
    float* out = nullptr;       // NHWC [B][T][C], channel = h*d + dd
    const float* qk_lo = nullptr; const float* vt_lo = nullptr;   // fp32 (3xTF32) mode: tf32 lo parts, same layouts
    int batch = 0, T = 0, t_pad = 0, heads = 0, head_dim = 0, C = 0;
    int bf16 = 0;               // qk and vt hold bf16 elements (same layouts, t_pad a multiple of 8)
};
struct AttentionParams {
    CUtensorMap mapQ, mapK, mapV, mapQlo, mapKlo, mapVlo;
    int split, bf16, idle;
    float* out; int batch, T, heads, C; float scale_log2;
};
int attention_prepare(AttentionParams& P, const AttentionDesc& d);
int attention_launch(const AttentionParams& P, cudaStream_t st);
double attention_flops(const AttentionDesc& d);

}  // namespace ipdm
