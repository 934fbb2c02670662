// Flash-style self-attention on tcgen05/TMEM for the UNet's AttentionBlock (Model/model.py:135-155):
//   attn = softmax((q*s)(k*s)^T), s = d^-1/4 ;  h = attn . v         heads = 4, d = 64, T up to 7125
// The [T,T] score matrix is never materialised (the reference allocates 812 MB per slice at T = 7125).
//
// Inputs come straight from the qkv 1x1 GEMM epilogue (conv_tc.cu): q and k live in the NHWC qkv tensor
// [B][T][3C] (head h: q at channel 3dh, k at 3dh+d, the reference's reshape(B*heads, 3d, T).chunk(3)),
// v is stored transposed [B][heads][d][t_pad] so that P.V^T is a K-major UMMA like everything else.
//
// CTA = 128 queries of one (slice, head); 64-key blocks stream through a TMA ring.
//   warp 4 lane*: TMA producer            warp 5 lane*: MMA issuer (+ TMEM alloc)
//   warps 0-3  : one query row per thread (row == TMEM lane): online softmax in registers,
//                P written to shared memory in the 128B-swizzled K-major layout UMMA expects,
//                O accumulated in registers (O_blk = P.V is a fresh TMEM tile per key block).
// S = QK^T and O_blk = P V: M=128 N=64, 32 bytes of K per instruction (8 tf32 / 16 bf16).  Out-of-range keys are
// zero-filled by TMA and masked to -inf; out-of-range query rows are computed and dropped.
// Two kernels: attention_kernel (serial schedule; kept for the 3xTF32 split of the fp32 mode) and
// attention_pipe_kernel (tf32 / bf16 modes: S, P and O double-buffered, see its comment).
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"
#include "unet_ops.cuh"

namespace ipdm {

constexpr int AT_BQ = 128, AT_BKV = 64, AT_D = 64;
constexpr int AT_THREADS = 192;
constexpr int AT_TF32 = 0, AT_SPLIT = 1, AT_BF16 = 2;
// MODE tf32 : q, k, v, P are fp32 words rounded to tf32; every operand tile is 2 chunks of 128-byte rows (32 elements each).
// MODE split: fp32-accurate 3xTF32: q, k, v arrive as tf32 hi/lo pairs (written by the qkv GEMM epilogue), the softmax threads
//             write P as hi/lo, and every product is hi*hi + hi*lo + lo*hi.  One K/V stage (192 KB of smem).
// MODE bf16 : q, k, v, P are bf16 (kind::f16, fp32 accumulate); a 64-element row is ONE 128-byte swizzled row, so every tile is
//             half the size, the MMA count halves, and three CTAs fit on an SM: one CTA's softmax hides behind the others' MMAs.
template <int MODE>
struct AtL {
    static constexpr bool SPLIT = MODE == AT_SPLIT, BF = MODE == AT_BF16;
    static constexpr int M = SPLIT ? 2 : 1;                   // hi (+ lo) copies of every operand tile
    static constexpr int NCH = BF ? 1 : 2;                    // 128-byte chunks per 64-element operand row
    static constexpr int ELEMS = BF ? 64 : 32;                // elements per chunk row
    static constexpr int Q_BYTES = NCH * AT_BQ * 128;
    static constexpr int K_BYTES = NCH * AT_BKV * 128;
    static constexpr int V_BYTES = NCH * AT_D * 128;
    static constexpr int P_BYTES = NCH * AT_BQ * 128;
    static constexpr int NST = SPLIT ? 1 : 2;
    static constexpr int STAGE = M * (K_BYTES + V_BYTES);
    static constexpr int OFF_QLO = Q_BYTES;
    static constexpr int OFF_KV = M * Q_BYTES;
    static constexpr int OFF_KLO = K_BYTES;                   // inside a stage: K_hi [K_lo] V_hi [V_lo]
    static constexpr int OFF_V = M * K_BYTES;
    static constexpr int OFF_VLO = OFF_V + V_BYTES;
    static constexpr int OFF_P = OFF_KV + NST * STAGE;
    static constexpr int OFF_PLO = OFF_P + P_BYTES;
    static constexpr int OFF_BAR = OFF_P + M * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 16 * 8 + 1024;
    static constexpr int MAX_REGS = BF ? 112 : 255;             // bf16: 3 CTAs x 192 threads x 112 registers = 63 K of the 64 K file
};

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS) __maxnreg__(AtL<MODE>::MAX_REGS)
attention_kernel(const __grid_constant__ AttentionParams P) {
    using L = AtL<MODE>;
    constexpr bool SPLIT = L::SPLIT, BF = L::BF;
    constexpr int NST = L::NST, NCH = L::NCH;
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)at_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + L::OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // 2
    uint64_t* kv_empty = bars + 3;      // 2
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;        // 128 arrivals
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_slot = (uint32_t*)(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, head = blockIdx.y, b = blockIdx.z;
    const int nkv = (P.T + AT_BKV - 1) / AT_BKV;

    if (threadIdx.x == 0) {
        tc::mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&kv_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
        tc::mbar_init(s_full, 1); tc::mbar_init(p_full, 128); tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc(tmem_slot, 128);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;

    if (warp == 4) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(q_full, L::M * L::Q_BYTES);
            for (int c = 0; c < NCH; ++c) {
                tc::tma_load_3d(smem + c * (AT_BQ * 128), &P.mapQ, q_full, head * 3 * AT_D + c * 32, q0, b);
                if constexpr (SPLIT) tc::tma_load_3d(smem + L::OFF_QLO + c * (AT_BQ * 128), &P.mapQlo, q_full, head * 3 * AT_D + c * 32, q0, b);
            }
            for (int j = 0; j < nkv; ++j) {
                const int s = j % NST;
                tc::mbar_wait(&kv_empty[s], ((uint32_t)(j / NST) & 1u) ^ 1u);
                tc::mbar_expect_tx(&kv_full[s], L::STAGE);
                uint8_t* sk = smem + L::OFF_KV + s * L::STAGE;
                uint8_t* sv = sk + L::OFF_V;
                for (int c = 0; c < NCH; ++c) {
                    tc::tma_load_3d(sk + c * (AT_BKV * 128), &P.mapK, &kv_full[s], head * 3 * AT_D + AT_D + c * 32, j * AT_BKV, b);
                    tc::tma_load_3d(sv + c * (AT_D * 128), &P.mapV, &kv_full[s], j * AT_BKV + c * 32, head * AT_D, b);
                    if constexpr (SPLIT) {
                        tc::tma_load_3d(sk + L::OFF_KLO + c * (AT_BKV * 128), &P.mapKlo, &kv_full[s], head * 3 * AT_D + AT_D + c * 32, j * AT_BKV, b);
                        tc::tma_load_3d(sk + L::OFF_VLO + c * (AT_D * 128), &P.mapVlo, &kv_full[s], j * AT_BKV + c * 32, head * AT_D, b);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc(BF ? tc::FMT_BF16 : tc::FMT_TF32, 128, 64);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t bb, uint32_t acc) {
                if constexpr (BF) tc::umma_f16(d, a, bb, idesc, acc); else tc::umma_tf32(d, a, bb, idesc, acc);
            };
            const uint32_t sQ = tc::smem_u32(smem), sP = tc::smem_u32(smem + L::OFF_P);
            auto issue_S = [&](int j) {
                const int s = j % NST;
                tc::mbar_wait(&kv_full[s], (uint32_t)(j / NST) & 1u);
                tc::tc_fence_after();
                const uint32_t sK = tc::smem_u32(smem + L::OFF_KV + s * L::STAGE);
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                      // 32 bytes of K per instruction (8 tf32 / 16 bf16)
                        const uint64_t qd = tc::smem_desc_k_sw128(sQ + c * (AT_BQ * 128)) + (uint64_t)(k * 2);
                        const uint64_t kd = tc::smem_desc_k_sw128(sK + c * (AT_BKV * 128)) + (uint64_t)(k * 2);
                        mma(tmem_S, qd, kd, (uint32_t)((c | k) != 0));
                        if constexpr (SPLIT) {
                            mma(tmem_S, qd, tc::smem_desc_k_sw128(sK + L::OFF_KLO + c * (AT_BKV * 128)) + (uint64_t)(k * 2), 1u);
                            mma(tmem_S, tc::smem_desc_k_sw128(sQ + L::OFF_QLO + c * (AT_BQ * 128)) + (uint64_t)(k * 2), kd, 1u);
                        }
                    }
                tc::umma_commit(s_full);
            };
            tc::mbar_wait(q_full, 0);
            issue_S(0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % NST;
                tc::mbar_wait(p_full, (uint32_t)j & 1u);
                tc::tc_fence_after();
                const uint32_t sV = tc::smem_u32(smem + L::OFF_KV + s * L::STAGE + L::OFF_V);
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t pd = tc::smem_desc_k_sw128(sP + c * (AT_BQ * 128)) + (uint64_t)(k * 2);
                        const uint64_t vd = tc::smem_desc_k_sw128(sV + c * (AT_D * 128)) + (uint64_t)(k * 2);
                        mma(tmem_O, pd, vd, (uint32_t)((c | k) != 0));
                        if constexpr (SPLIT) {
                            mma(tmem_O, pd, tc::smem_desc_k_sw128(sV + L::V_BYTES + c * (AT_D * 128)) + (uint64_t)(k * 2), 1u);
                            mma(tmem_O, tc::smem_desc_k_sw128(sP + L::P_BYTES + c * (AT_BQ * 128)) + (uint64_t)(k * 2), vd, 1u);
                        }
                    }
                tc::umma_commit(o_full);
                tc::umma_commit(&kv_empty[s]);
                if (j + 1 < nkv) issue_S(j + 1);
            }
        }
        __syncwarp();
    } else {
        // ---------------- softmax / accumulate: one query row per thread ----------------
        const int r = warp * 32 + lane;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        float o_acc[AT_D];
#pragma unroll
        for (int i = 0; i < AT_D; ++i) o_acc[i] = 0.f;
        uint8_t* prow = smem + L::OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
        const int sw = r & 7;
        for (int j = 0; j < nkv; ++j) {
            tc::mbar_wait(s_full, (uint32_t)j & 1u);
            tc::tc_fence_after();
            const int nvalid = P.T - j * AT_BKV;           // keys beyond T are masked
            float rs = 0.f, alpha, m_new;
            if constexpr (BF) {
                // 32 score registers live at a time (three CTAs share the SM's register file): the first half is read once for the
                // row maximum and again, after the second half has been exponentiated, for its own exponentials.
                uint32_t sh[32];
                auto load_half = [&](int h) {
                    tc::tmem_ld32(tmem_S + lane_off + h * 32, sh);
                    tc::tmem_ld_wait();
                    if (nvalid < AT_BKV) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (h * 32 + i >= nvalid) sh[i] = __float_as_uint(-INFINITY);
                    }
                };
                float mx = -INFINITY;
                load_half(0);
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sh[i]));
                load_half(1);
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(sh[i]));
                m_new = fmaxf(m_run, mx);
                alpha = exp2f((m_run - m_new) * P.scale_log2);
                const float mb = m_new * P.scale_log2;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int h = 1 - hh;
                    if (hh == 1) load_half(0);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {            // 8 keys = one 16-byte unit of the 128-byte row
                        uint32_t w[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float a = exp2f(fmaf(__uint_as_float(sh[8 * u + 2 * i]), P.scale_log2, -mb));
                            const float c = exp2f(fmaf(__uint_as_float(sh[8 * u + 2 * i + 1]), P.scale_log2, -mb));
                            const __nv_bfloat162 pk = __floats2bfloat162_rn(a, c);
                            const float2 back = __bfloat1622float2(pk);      // the denominator sums what the MMA multiplies
                            rs += back.x + back.y;
                            w[i] = *reinterpret_cast<const uint32_t*>(&pk);
                        }
                        *reinterpret_cast<uint4*>(prow + (((h * 4 + u) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            } else {
                uint32_t sr[2][32];
                tc::tmem_ld32(tmem_S + lane_off, sr[0]);
                tc::tmem_ld32(tmem_S + lane_off + 32, sr[1]);
                tc::tmem_ld_wait();
                float mx = -INFINITY;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float v = __uint_as_float(sr[h][i]);
                        v = (h * 32 + i < nvalid) ? v : -INFINITY;
                        sr[h][i] = __float_as_uint(v);
                        mx = fmaxf(mx, v);
                    }
                m_new = fmaxf(m_run, mx);
                alpha = exp2f((m_run - m_new) * P.scale_log2);
                const float mb = m_new * P.scale_log2;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float4 pv;
                        pv.x = exp2f(fmaf(__uint_as_float(sr[h][4 * u]), P.scale_log2, -mb));
                        pv.y = exp2f(fmaf(__uint_as_float(sr[h][4 * u + 1]), P.scale_log2, -mb));
                        pv.z = exp2f(fmaf(__uint_as_float(sr[h][4 * u + 2]), P.scale_log2, -mb));
                        pv.w = exp2f(fmaf(__uint_as_float(sr[h][4 * u + 3]), P.scale_log2, -mb));
                        float4 ph;
                        ph.x = tf32_rn(pv.x); ph.y = tf32_rn(pv.y); ph.z = tf32_rn(pv.z); ph.w = tf32_rn(pv.w);   // P is an MMA operand
                        if constexpr (SPLIT) {
                            rs += (pv.x + pv.y) + (pv.z + pv.w);
                            float4 pl;
                            pl.x = tf32_rn(pv.x - ph.x); pl.y = tf32_rn(pv.y - ph.y); pl.z = tf32_rn(pv.z - ph.z); pl.w = tf32_rn(pv.w - ph.w);
                            *reinterpret_cast<float4*>(prow + L::P_BYTES + h * (AT_BQ * 128) + ((u ^ sw) << 4)) = pl;
                        } else {
                            rs += (ph.x + ph.y) + (ph.z + ph.w);
                        }
                        *reinterpret_cast<float4*>(prow + h * (AT_BQ * 128) + ((u ^ sw) << 4)) = ph;
                    }
                }
            }
            l_run = fmaf(l_run, alpha, rs);
            m_run = m_new;
            tc::fence_proxy_async();                         // generic-proxy smem writes -> visible to the MMA (async proxy)
            tc::tc_fence_before();
            tc::mbar_arrive(p_full);
            tc::mbar_wait(o_full, (uint32_t)j & 1u);
            tc::tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t orr[32];
                tc::tmem_ld32(tmem_O + lane_off + h * 32, orr);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[h * 32 + i] = fmaf(o_acc[h * 32 + i], alpha, __uint_as_float(orr[i]));
            }
        }
        const int t = q0 + r;
        if (t < P.T) {
            const float inv = 1.0f / l_run;
            float4* op = reinterpret_cast<float4*>(P.out + ((size_t)b * P.T + t) * P.C + head * AT_D);
#pragma unroll
            for (int i = 0; i < AT_D / 4; ++i)
                op[i] = SPLIT ? make_float4(o_acc[4 * i] * inv, o_acc[4 * i + 1] * inv, o_acc[4 * i + 2] * inv, o_acc[4 * i + 3] * inv)
                              : make_float4(tf32_rn(o_acc[4 * i] * inv), tf32_rn(o_acc[4 * i + 1] * inv), tf32_rn(o_acc[4 * i + 2] * inv),
                                            tf32_rn(o_acc[4 * i + 3] * inv));      // tf32 / bf16 mode: operand of the proj 1x1 GEMM
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------------------------------
// Pipelined schedule (tf32 and bf16 modes).  In the kernel above a key block is one serial chain
//   S MMA -> softmax -> P -> PV MMA -> O read, and only co-resident CTAs overlap MMAs with softmax (measured 380 TFLOP/s in
// bf16, ~1500 cycles per block and SM against 600 cycles of MMA).  Here S, P and O are double-buffered (TMEM: S0 S1 O0 O1 =
// 256 columns; two P tiles in shared memory; three K/V stages), the MMA warp runs S two blocks ahead, and a softmax thread folds
// O_blk(j-1) into its accumulator AFTER it has handed P(j) to the tensor core:
//   MMA warp      : S(0) S(1) | wait P(j) -> PV(j) -> S(j+2)
//   softmax thread: wait S(j) -> exp -> P(j) -> arrive | wait O(j-1) -> o = o*alpha(j-1) + O_blk(j-1)
// so neither side waits for the other's latency.  Buffer reuse is ordered by the same barriers: P(j) is written after
// O(j-2) was consumed (PV(j-2) done reading P[j&1]); PV(j) is issued after P(j) arrived, i.e. after O(j-2) was read from
// O[j&1]; S(j+2) overwrites S[j&1] after P(j) arrived, i.e. after S(j) was read.
// ------------------------------------------------------------------------------------------------------------------------
template <int MODE>
struct AtP {
    static constexpr bool BF = MODE == AT_BF16;
    static constexpr int NCH = BF ? 1 : 2;
    static constexpr int Q_BYTES = NCH * AT_BQ * 128;
    static constexpr int K_BYTES = NCH * AT_BKV * 128;
    static constexpr int V_BYTES = NCH * AT_D * 128;
    static constexpr int P_BYTES = NCH * AT_BQ * 128;
    static constexpr int NST = 3;
    static constexpr int STAGE = K_BYTES + V_BYTES;
    static constexpr int OFF_KV = Q_BYTES;
    static constexpr int OFF_P = OFF_KV + NST * STAGE;
    static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 16 * 8 + 1024;
    static constexpr int MAX_REGS = BF ? 168 : 255;             // bf16: 2 CTAs x 192 threads x 168 registers
};

template <int MODE>
__global__ void __launch_bounds__(AT_THREADS) __maxnreg__(AtP<MODE>::MAX_REGS)
attention_pipe_kernel(const __grid_constant__ AttentionParams P) {
    using L = AtP<MODE>;
    constexpr bool BF = L::BF;
    constexpr int NST = L::NST, NCH = L::NCH;
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)at_smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + L::OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // 3
    uint64_t* kv_empty = bars + 4;      // 3
    uint64_t* s_full = bars + 7;        // 2
    uint64_t* p_full = bars + 9;        // 2, 128 arrivals each
    uint64_t* o_full = bars + 11;       // 2
    uint32_t* tmem_slot = (uint32_t*)(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BQ, head = blockIdx.y, b = blockIdx.z;
    const int nkv = (P.T + AT_BKV - 1) / AT_BKV;

    if (threadIdx.x == 0) {
        tc::mbar_init(q_full, 1);
        for (int i = 0; i < NST; ++i) { tc::mbar_init(&kv_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&s_full[i], 1); tc::mbar_init(&p_full[i], 128); tc::mbar_init(&o_full[i], 1); }
        tc::fence_barrier_init();
    }
    if (warp == 5) tc::tmem_alloc(tmem_slot, 256);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;                     // S0 | S1 | O0 | O1, 64 columns each

    if (warp == 4) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(q_full, L::Q_BYTES);
            for (int c = 0; c < NCH; ++c) tc::tma_load_3d(smem + c * (AT_BQ * 128), &P.mapQ, q_full, head * 3 * AT_D + c * 32, q0, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % NST;
                tc::mbar_wait(&kv_empty[s], ((uint32_t)(j / NST) & 1u) ^ 1u);
                tc::mbar_expect_tx(&kv_full[s], L::STAGE);
                uint8_t* sk = smem + L::OFF_KV + s * L::STAGE;
                uint8_t* sv = sk + L::K_BYTES;
                for (int c = 0; c < NCH; ++c) {
                    tc::tma_load_3d(sk + c * (AT_BKV * 128), &P.mapK, &kv_full[s], head * 3 * AT_D + AT_D + c * 32, j * AT_BKV, b);
                    tc::tma_load_3d(sv + c * (AT_D * 128), &P.mapV, &kv_full[s], j * AT_BKV + c * 32, head * AT_D, b);
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::make_idesc(BF ? tc::FMT_BF16 : tc::FMT_TF32, 128, 64);
            auto mma = [&](uint32_t d, uint64_t a, uint64_t bb, uint32_t acc) {
                if constexpr (BF) tc::umma_f16(d, a, bb, idesc, acc); else tc::umma_tf32(d, a, bb, idesc, acc);
            };
            const uint32_t sQ = tc::smem_u32(smem);
            auto issue_S = [&](int j) {
                const int s = j % NST;
                tc::mbar_wait(&kv_full[s], (uint32_t)(j / NST) & 1u);
                tc::tc_fence_after();
                const uint32_t sK = tc::smem_u32(smem + L::OFF_KV + s * L::STAGE);
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma(tmem_base + (j & 1) * 64, tc::smem_desc_k_sw128(sQ + c * (AT_BQ * 128)) + (uint64_t)(k * 2),
                            tc::smem_desc_k_sw128(sK + c * (AT_BKV * 128)) + (uint64_t)(k * 2), (uint32_t)((c | k) != 0));
                tc::umma_commit(&s_full[j & 1]);
            };
            tc::mbar_wait(q_full, 0);
            issue_S(0);
            if (nkv > 1) issue_S(1);
            for (int j = 0; j < nkv; ++j) {
                const int s = j % NST, bsel = j & 1;
                tc::mbar_wait(&p_full[bsel], ((uint32_t)j >> 1) & 1u);
                tc::tc_fence_after();
                const uint32_t sV = tc::smem_u32(smem + L::OFF_KV + s * L::STAGE + L::K_BYTES);
                const uint32_t sP = tc::smem_u32(smem + L::OFF_P + bsel * L::P_BYTES);
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma(tmem_base + 128 + bsel * 64, tc::smem_desc_k_sw128(sP + c * (AT_BQ * 128)) + (uint64_t)(k * 2),
                            tc::smem_desc_k_sw128(sV + c * (AT_D * 128)) + (uint64_t)(k * 2), (uint32_t)((c | k) != 0));
                tc::umma_commit(&o_full[bsel]);
                tc::umma_commit(&kv_empty[s]);
                if (j + 2 < nkv) issue_S(j + 2);
            }
        }
        __syncwarp();
    } else {
        // ---------------- softmax / accumulate: one query row per thread ----------------
        const int r = warp * 32 + lane;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f, alpha_pend = 0.f;
        float o_acc[AT_D];
#pragma unroll
        for (int i = 0; i < AT_D; ++i) o_acc[i] = 0.f;
        const int sw = r & 7;
        auto fold_O = [&](int j) {                               // o = o * alpha(j) + O_blk(j)
            if (P.idle) tc::mbar_wait_idle(&o_full[j & 1], ((uint32_t)j >> 1) & 1u); else tc::mbar_wait(&o_full[j & 1], ((uint32_t)j >> 1) & 1u);
            tc::tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t orr[32];
                tc::tmem_ld32(tmem_base + 128 + (j & 1) * 64 + lane_off + h * 32, orr);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[h * 32 + i] = fmaf(o_acc[h * 32 + i], alpha_pend, __uint_as_float(orr[i]));
            }
        };
        for (int j = 0; j < nkv; ++j) {
            const int bsel = j & 1;
            const uint32_t tS = tmem_base + bsel * 64 + lane_off;
            uint8_t* prow = smem + L::OFF_P + bsel * L::P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
            if (P.idle) tc::mbar_wait_idle(&s_full[bsel], ((uint32_t)j >> 1) & 1u); else tc::mbar_wait(&s_full[bsel], ((uint32_t)j >> 1) & 1u);
            tc::tc_fence_after();
            const int nvalid = P.T - j * AT_BKV;           // keys beyond T are masked
            uint32_t sr[2][32];
            tc::tmem_ld32(tS, sr[0]);
            tc::tmem_ld32(tS + 32, sr[1]);
            tc::tmem_ld_wait();
            if (nvalid < AT_BKV) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (h * 32 + i >= nvalid) sr[h][i] = __float_as_uint(-INFINITY);
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;      // four chains instead of one of 64
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(sr[h][i])); mx1 = fmaxf(mx1, __uint_as_float(sr[h][i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(sr[h][i + 2])); mx3 = fmaxf(mx3, __uint_as_float(sr[h][i + 3]));
                }
            const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
            const float alpha = ex2_approx((m_run - m_new) * P.scale_log2);
            const float mb = m_new * P.scale_log2;
            float rs0 = 0.f, rs1 = 0.f;
            if constexpr (BF) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {                // 8 keys = one 16-byte unit of the 128-byte row
                    const uint32_t* sp = &sr[u >> 2][(u & 3) * 8];
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float a = ex2_approx(fmaf(__uint_as_float(sp[2 * i]), P.scale_log2, -mb));
                        const float c = ex2_approx(fmaf(__uint_as_float(sp[2 * i + 1]), P.scale_log2, -mb));
                        const __nv_bfloat162 pk = __floats2bfloat162_rn(a, c);
                        rs0 += a; rs1 += c;                  // unrounded: the bf16 rounding of P is zero-mean over 64 keys
                        w[i] = *reinterpret_cast<const uint32_t*>(&pk);
                    }
                    *reinterpret_cast<uint4*>(prow + ((u ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float4 ph;
                        ph.x = tf32_rn(ex2_approx(fmaf(__uint_as_float(sr[h][4 * u]), P.scale_log2, -mb)));
                        ph.y = tf32_rn(ex2_approx(fmaf(__uint_as_float(sr[h][4 * u + 1]), P.scale_log2, -mb)));
                        ph.z = tf32_rn(ex2_approx(fmaf(__uint_as_float(sr[h][4 * u + 2]), P.scale_log2, -mb)));
                        ph.w = tf32_rn(ex2_approx(fmaf(__uint_as_float(sr[h][4 * u + 3]), P.scale_log2, -mb)));      // P is an MMA operand
                        rs0 += ph.x + ph.z; rs1 += ph.y + ph.w;
                        *reinterpret_cast<float4*>(prow + h * (AT_BQ * 128) + ((u ^ sw) << 4)) = ph;
                    }
            }
            l_run = fmaf(l_run, alpha, rs0 + rs1);
            m_run = m_new;
            tc::fence_proxy_async();                         // generic-proxy smem writes -> visible to the MMA (async proxy)
            tc::tc_fence_before();
            tc::mbar_arrive(&p_full[bsel]);
            if (j >= 1) fold_O(j - 1);                       // the previous block's P.V, with the previous block's rescale factor
            alpha_pend = alpha;
        }
        fold_O(nkv - 1);
        const int t = q0 + r;
        if (t < P.T) {
            const float inv = 1.0f / l_run;
            float4* op = reinterpret_cast<float4*>(P.out + ((size_t)b * P.T + t) * P.C + head * AT_D);
#pragma unroll
            for (int i = 0; i < AT_D / 4; ++i)
                op[i] = make_float4(tf32_rn(o_acc[4 * i] * inv), tf32_rn(o_acc[4 * i + 1] * inv), tf32_rn(o_acc[4 * i + 2] * inv),
                                    tf32_rn(o_acc[4 * i + 3] * inv));              // operand of the proj 1x1 GEMM
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem_base, 256);
}

int attention_prepare(AttentionParams& P, const AttentionDesc& d) {
    IPDM_REQUIRE(d.head_dim == AT_D && d.C == d.heads * AT_D, "attention: head_dim must be 64 (got %d)", d.head_dim);
    IPDM_REQUIRE(d.t_pad % (d.bf16 ? 8 : 4) == 0 && d.t_pad >= d.T, "attention: t_pad must be a multiple of %d", d.bf16 ? 8 : 4);
    IPDM_REQUIRE(!(d.bf16 && d.qk_lo), "attention: bf16 operands and the 3xTF32 split are exclusive");
    memset(&P, 0, sizeof(P));
    const uint64_t eb = d.bf16 ? 2 : 4;                          // bytes per q/k/v element
    const uint32_t row = d.bf16 ? 64 : 32;                        // elements per 128-byte operand row
    const CUtensorMapDataType dt = d.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const uint64_t dq[3] = {(uint64_t)3 * d.C, (uint64_t)d.T, (uint64_t)d.batch};
    const uint64_t sq[2] = {(uint64_t)3 * d.C * eb, (uint64_t)d.T * 3 * d.C * eb};
    const uint32_t bq[3] = {row, AT_BQ, 1}, bk[3] = {row, AT_BKV, 1};
    IPDM_CHECK(tmap_encode(&P.mapQ, dt, 3, d.qk, dq, sq, bq, CU_TENSOR_MAP_SWIZZLE_128B));
    IPDM_CHECK(tmap_encode(&P.mapK, dt, 3, d.qk, dq, sq, bk, CU_TENSOR_MAP_SWIZZLE_128B));
    const uint64_t dv[3] = {(uint64_t)d.T, (uint64_t)d.heads * AT_D, (uint64_t)d.batch};
    const uint64_t sv[2] = {(uint64_t)d.t_pad * eb, (uint64_t)d.heads * AT_D * d.t_pad * eb};
    const uint32_t bv[3] = {row, AT_D, 1};
    IPDM_CHECK(tmap_encode(&P.mapV, dt, 3, d.vt, dv, sv, bv, CU_TENSOR_MAP_SWIZZLE_128B));
    P.split = d.qk_lo != nullptr;
    P.bf16 = d.bf16;
    // softmax warps sleep (try_wait with a suspend hint) instead of polling while they wait for S / O_blk: 1465 -> 1402 us at T = 7125, 16 slices
    static const int env_idle = getenv("IPDM_ATTN_IDLE") ? atoi(getenv("IPDM_ATTN_IDLE")) : 1;
    P.idle = env_idle;
    if (P.split) {
        IPDM_REQUIRE(d.vt_lo != nullptr, "attention: fp32 mode needs both qk_lo and vt_lo");
        IPDM_CHECK(tmap_encode(&P.mapQlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d.qk_lo, dq, sq, bq, CU_TENSOR_MAP_SWIZZLE_128B));
        IPDM_CHECK(tmap_encode(&P.mapKlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d.qk_lo, dq, sq, bk, CU_TENSOR_MAP_SWIZZLE_128B));
        IPDM_CHECK(tmap_encode(&P.mapVlo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d.vt_lo, dv, sv, bv, CU_TENSOR_MAP_SWIZZLE_128B));
    }
    P.out = d.out; P.batch = d.batch; P.T = d.T; P.heads = d.heads; P.C = d.C;
    P.scale_log2 = 1.4426950408889634f / sqrtf((float)AT_D);      // (d^-1/4)^2 * log2(e)
    return IPDM_OK;
}

int attention_launch(const AttentionParams& P, cudaStream_t st) {
    static DeviceOnce once;
    static_assert(AtL<AT_SPLIT>::SMEM <= 227 * 1024, "fp32-mode attention tiles do not fit in shared memory");
    static_assert(3 * (AtL<AT_BF16>::SMEM + 1024) <= 227 * 1024, "bf16 attention is sized for three CTAs per SM");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<AT_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtL<AT_TF32>::SMEM));
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<AT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtL<AT_SPLIT>::SMEM));
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<AT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtL<AT_BF16>::SMEM));
        static_assert(AtP<AT_TF32>::SMEM <= 227 * 1024 && 2 * (AtP<AT_BF16>::SMEM + 1024) <= 227 * 1024, "pipelined attention tiles do not fit");
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<AT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtP<AT_BF16>::SMEM));
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(attention_pipe_kernel<AT_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtP<AT_TF32>::SMEM));
    }
    dim3 grid((P.T + AT_BQ - 1) / AT_BQ, P.heads, P.batch);
    ProfScope prof(PROF_ATTENTION, st, (P.split ? 3.0 : 1.0) * 4.0 * P.batch * P.heads * (double)P.T * P.T * AT_D);
    static const bool pipe = !(getenv("IPDM_ATTN_PIPE") && atoi(getenv("IPDM_ATTN_PIPE")) == 0);      // 0: the serial schedule (experiments)
    if (P.split) attention_kernel<AT_SPLIT><<<grid, AT_THREADS, AtL<AT_SPLIT>::SMEM, st>>>(P);
    else if (P.bf16 && pipe) attention_pipe_kernel<AT_BF16><<<grid, AT_THREADS, AtP<AT_BF16>::SMEM, st>>>(P);
    else if (pipe) attention_pipe_kernel<AT_TF32><<<grid, AT_THREADS, AtP<AT_TF32>::SMEM, st>>>(P);
    else if (P.bf16) attention_kernel<AT_BF16><<<grid, AT_THREADS, AtL<AT_BF16>::SMEM, st>>>(P);
    else attention_kernel<AT_TF32><<<grid, AT_THREADS, AtL<AT_TF32>::SMEM, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

double attention_flops(const AttentionDesc& d) { return 4.0 * d.batch * d.heads * (double)d.T * d.T * AT_D; }

}  // namespace ipdm
