// Thin convolutions (C_in <= 32, C_out in {8, 16}) of the full-resolution levels on the tensor cores.
//
// The projection UNet spends its first and last two levels (2000x912 and 1000x456) in layers with 4...24 channels
// (Model/model.py channel_mult = [1/16, 1/8, 1/4, ...]).  On CUDA cores they were a third of the whole step
// (profiles/r01_*: conv_direct), although each of them only moves 4*(C_in + C_out) bytes per pixel.  This kernel maps them
// to tcgen05:  D[128 pixels, 16] += A_tap[128 pixels, K = C_in] * W_tap[16, K]^T  with
//   * a pixel row of C_in in {8, 16, 32} fp32 channels = 32 / 64 / 128 bytes = one K-major operand row in the matching
//     SWIZZLE_32B / 64B / 128B mode (TMA writes it, UMMA reads it; verified by tools/experiments/shifted_desc_narrow.cu),
//   * ONE halo tile [(4+2) rows x 32 pixels] per output tile, read in place by all nine taps through descriptors that
//     start at row dy*32 + dx (30 of the 32 columns are outputs),
//   * all tap weights (<= 18 KB) resident in shared memory for the lifetime of the persistent CTA,
//   * eight TMEM accumulators (16 columns each) so that epilogues overlap the next tiles' MMAs.
// Per tile the tensor pipe issues 9 * C_in/8 instructions of ~45 cycles (14 % busy); measured at 16 slices, 2000x912, 8 -> 8:
// 583 us for 2.3 GB (3.9 TB/s) against ~1.0 ms for the CUDA-core kernel.  The kernel is bound by the per-tile latency chain of its
// epilogue warps (barrier wait -> tcgen05.ld -> staging -> stores, ~1350 warp instructions per tile; ablations with no loads,
// MMAs or stores still take 430 us), not by TMA, the tensor pipe or HBM: see profiles/r01_thin_gnapply_attention_ncu_full.md.
#include "common.cuh"
#include "tc.cuh"
#include "unet_ops.cuh"

#include <algorithm>
#include <cstdlib>

namespace ipdm {

constexpr int TH_RP = 32, TH_TWV = 30;                  // row pitch (pixels), valid columns; output rows per tile: template ROWS (4 or 8)
constexpr int TH_NACC = 8;                              // accumulators in flight (16 TMEM columns each)
constexpr int TH_ACC_COLS = 16, TH_TMEM_COLS = TH_NACC * TH_ACC_COLS;
constexpr int TH_THREADS = 256;                         // warp 0 producer, 1 MMA, 2 TMEM alloc, 4-7 epilogue
// The epilogue (TMEM -> registers -> +bias +residual -> global) is a latency chain of ~1.7 k cycles per tile on one warpgroup, four
// times the MMA issue time of a C = 8 tile; two or three co-resident CTAs per SM (short operand rings, 128 TMEM columns each) keep
// up to twelve epilogue warps in flight instead of four.
constexpr int TH_STG_PITCH = 20;                         // floats per pixel row of the epilogue staging tile (80 B: conflict-free)
constexpr int TH_STG_BYTES = 4 * 32 * TH_STG_PITCH * 4;  // one 32-pixel staging tile per epilogue warp

template <int RB, int ROWS> struct ThinCfg {            // RB = bytes per pixel row of the operand tensor; ROWS = output rows per tile
    static constexpr int SUB = ROWS / 4;                // 128-row MMA groups (accumulators) per tile
    static constexpr int NT = TH_NACC / SUB;            // tiles whose accumulators are in flight
    static constexpr int CTAS = RB == 128 ? 2 : 3;      // co-resident CTAs per SM (4 was tried for RB = 32: 64 registers spill, 733 vs 583 us)
    static constexpr int W_BYTES = (9 * 16 * RB + 1023) / 1024 * 1024;
    static constexpr int BOX_BYTES = (ROWS + 2) * TH_RP * RB;
    static constexpr int SLOT = (BOX_BYTES + 2 * RB + 1023) / 1024 * 1024;   // + the 2 pixels the last tap over-reads
    static constexpr int NSA = ROWS == 4 ? (RB == 32 ? 8 : (RB == 64 ? 4 : 3)) : (RB == 32 ? 5 : 2);
    static constexpr int LAYOUT = RB == 32 ? 6 : (RB == 64 ? 4 : 2);         // UMMA layout_type: SWIZZLE_32B / 64B / 128B
    static constexpr int SBO = 8 * RB;
    static constexpr int OFF_A = W_BYTES;
    static constexpr int BAR_OFF = OFF_A + NSA * SLOT;
    static constexpr int STG_OFF = BAR_OFF + 512;
    static constexpr int TOTAL = STG_OFF + TH_STG_BYTES + 1024;
};

__device__ __forceinline__ uint64_t thin_desc(uint32_t saddr, int layout, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

template <int RB, int ROWS>
__global__ void __launch_bounds__(TH_THREADS, ThinCfg<RB, ROWS>::CTAS)
conv_thin_kernel(const __grid_constant__ ConvThinParams P) {
    using C = ThinCfg<RB, ROWS>;
    constexpr int SUB = C::SUB, NT = C::NT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* a_full = (uint64_t*)(smem + C::BAR_OFF);
    uint64_t* a_empty = a_full + C::NSA;
    uint64_t* t_full = a_empty + C::NSA;
    uint64_t* t_empty = t_full + TH_NACC;
    uint64_t* w_full = t_empty + TH_NACC;
    uint64_t* a_ready = w_full + 1;                          // fused GroupNorm: tile normalised in place (64 arrivals)
    uint32_t* tmem_slot = (uint32_t*)(a_ready + C::NSA);
    float* sbias = (float*)(smem + C::BAR_OFF + 384);     // 16-byte aligned: the epilogue reads it as float4

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int total_tiles = tiles_per_img * P.batch;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NSA; ++i) { tc::mbar_init(&a_full[i], 1); tc::mbar_init(&a_empty[i], 1); tc::mbar_init(&a_ready[i], 64); }
        for (int i = 0; i < TH_NACC; ++i) { tc::mbar_init(&t_full[i], 1); tc::mbar_init(&t_empty[i], 4); }      // one arrival per epilogue warp
        tc::mbar_init(w_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, TH_TMEM_COLS);
    if (threadIdx.x < 16) {
        const float* bias = P.bias ? P.bias + (P.t_dev ? (size_t)(*P.t_dev) * P.bias_t_stride : 0) : nullptr;
        sbias[threadIdx.x] = (bias && (int)threadIdx.x < P.cout) ? __ldg(bias + threadIdx.x) : 0.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile = blockIdx.x, blockIdx.x + gridDim.x, ...: (slice, tile row, tile column) kept with adds and compares instead of two
    // integer divisions per tile and warp
    struct TileWalk {
        int b, tyi, txi, sb, sy, sx, tx, ty;
        __device__ void init(int tile, int g, int tiles_x, int tiles_y) {
            tx = tiles_x; ty = tiles_y;
            const int per = tx * ty;
            b = tile / per; const int tr = tile - b * per; tyi = tr / tx; txi = tr - tyi * tx;
            sb = g / per; const int gr = g - sb * per; sy = gr / tx; sx = gr - sy * tx;
        }
        __device__ void next() {
            txi += sx; if (txi >= tx) { txi -= tx; ++tyi; }
            tyi += sy; if (tyi >= ty) { tyi -= ty; ++b; }
            b += sb;
        }
    };
    auto decode = [&](int tile, int& b, int& x0, int& y0) {
        b = tile / tiles_per_img;
        const int tr = tile - b * tiles_per_img;
        const int tyi = tr / P.tiles_x, txi = tr - tyi * P.tiles_x;
        x0 = txi * TH_TWV; y0 = tyi * ROWS;
    };

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::mbar_expect_tx(w_full, P.ntaps * 16 * RB);
            tc::tma_load_2d(smem, &P.mapW, w_full, 0, 0);
            const int off = P.ntaps == 9 ? 1 : 0;                 // 1x1: the tile is its own halo
            int it = 0;
            TileWalk tw; tw.init(blockIdx.x, gridDim.x, P.tiles_x, P.tiles_y);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it, tw.next()) {
                const int b = tw.b, x0 = tw.txi * TH_TWV, y0 = tw.tyi * ROWS;
                const int s = it % C::NSA;
                tc::mbar_wait(&a_empty[s], ((uint32_t)(it / C::NSA) & 1u) ^ 1u);
                if (P.dbg & 8) { tc::mbar_arrive(&a_full[s]); continue; }          // experiment: no operand loads at all
                tc::mbar_expect_tx(&a_full[s], C::BOX_BYTES);
                tc::tma_load_4d(smem + C::OFF_A + s * C::SLOT, &P.mapA, &a_full[s], 0, x0 - off, y0 - off, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (tc::elect_one()) {
            const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, 16);
            const uint32_t w_base = tc::smem_u32(smem);
            tc::mbar_wait(w_full, 0);
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int s = it % C::NSA, slot = it % NT;
                tc::mbar_wait(&t_empty[slot], ((uint32_t)(it / NT) & 1u) ^ 1u);
                tc::mbar_wait(P.norm_scale ? &a_ready[s] : &a_full[s], (uint32_t)(it / C::NSA) & 1u);
                tc::tc_fence_after();
                const uint32_t a_base = tc::smem_u32(smem + C::OFF_A + s * C::SLOT);
                for (int tap = 0; tap < ((P.dbg & 4) ? 1 : P.ntaps); ++tap) {
                    const int dy = P.ntaps == 9 ? tap / 3 : 0, dx = P.ntaps == 9 ? tap - (tap / 3) * 3 : 0;
                    const uint64_t bd = thin_desc(w_base + (uint32_t)(tap * 16 * RB), C::LAYOUT, C::SBO);
#pragma unroll
                    for (int sub = 0; sub < SUB; ++sub) {          // tile rows 4*sub .. 4*sub+3: their own 128-row accumulator
                        const uint64_t ad = thin_desc(a_base + (uint32_t)(((4 * sub + dy) * TH_RP + dx) * RB), C::LAYOUT, C::SBO);
#pragma unroll
                        for (int k = 0; k < RB / 32; ++k)
                            tc::umma_tf32(tmem_base + (slot * SUB + sub) * TH_ACC_COLS, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc,
                                          (uint32_t)((tap | k) != 0));
                    }
                }
                tc::umma_commit(&a_empty[s]);
                tc::umma_commit(&t_full[slot]);
            }
        }
        __syncwarp();
    } else if (warp < 4) {
        // warps 2-3, fused GroupNorm(+SiLU): normalise the landed halo tile in place.  The transform is elementwise, so the swizzle
        // only matters for finding the channel of a 16-byte unit: physical unit = logical ^ f(row) with f = (row>>2)&1 / (row>>1)&3 /
        // row&7 for 32 / 64 / 128-byte rows.  Pixels outside the image stay 0 (the conv pads the NORMALISED tensor with zeros).
        if (P.norm_scale) {
            constexpr int UPR = RB / 16, NU = (ROWS + 2) * TH_RP * UPR;
            const int t2 = (warp - 2) * 32 + lane;
            const int off = P.ntaps == 9 ? 1 : 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                int b, x0, y0; decode(tile, b, x0, y0);
                const int s = it % C::NSA;
                tc::mbar_wait(&a_full[s], (uint32_t)(it / C::NSA) & 1u);
                uint8_t* base = smem + C::OFF_A + s * C::SLOT;
                // this slice's scale / shift for all C_in channels in registers (one load per tile, not per unit)
                float4 scv[UPR], shv[UPR];
#pragma unroll
                for (int i = 0; i < UPR; ++i) {
                    scv[i] = __ldg(reinterpret_cast<const float4*>(P.norm_scale + (size_t)b * P.cs) + i);
                    shv[i] = __ldg(reinterpret_cast<const float4*>(P.norm_shift + (size_t)b * P.cs) + i);
                }
                constexpr int BATCH = 6;                       // NU / 64 = 6, 12, 24 units per thread: batches of 6 loads in flight
#pragma unroll 1
                for (int u0 = t2; u0 < NU; u0 += 64 * BATCH) {
                    float4 v[BATCH]; bool in[BATCH];
#pragma unroll
                    for (int k = 0; k < BATCH; ++k) {
                        const int u = u0 + 64 * k, row = u / UPR;
                        const int gy = y0 - off + (row >> 5), gx = x0 - off + (row & 31);
                        in[k] = u < NU && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W;
                        v[k] = in[k] ? *reinterpret_cast<const float4*>(base + u * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int k = 0; k < BATCH; ++k) {
                        if (!in[k]) continue;
                        const int u = u0 + 64 * k, row = u / UPR, up = u - row * UPR;
                        const int swz = RB == 32 ? (row >> 2) & 1 : (RB == 64 ? (row >> 1) & 3 : row & 7);
                        const int lu = (up ^ swz) & (UPR - 1);
                        float4 a = scv[0], d = shv[0];
#pragma unroll
                        for (int i = 1; i < UPR; ++i) if (lu == i) { a = scv[i]; d = shv[i]; }
                        float4 w = v[k];
                        w.x = fmaf(w.x, a.x, d.x); w.y = fmaf(w.y, a.y, d.y); w.z = fmaf(w.z, a.z, d.z); w.w = fmaf(w.w, a.w, d.w);
                        if (P.act_silu) { w.x = silu(w.x); w.y = silu(w.y); w.z = silu(w.z); w.w = silu(w.w); }
                        w.x = tf32_rn(w.x); w.y = tf32_rn(w.y); w.z = tf32_rn(w.z); w.w = tf32_rn(w.w);
                        *reinterpret_cast<float4*>(base + u * 16) = w;
                    }
                }
                tc::fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async-proxy reads
                tc::mbar_arrive(&a_ready[s]);
            }
        }
    } else {
        const int q = warp & 3;                                  // TMEM lane quarter == output row of the tile
        // Coalesced path (the output and residual rows are exactly C_out floats wide, the normal case): a warp's 30 pixels are one
        // contiguous run of 30*C_out floats.  Each lane parks its pixel in a padded shared-memory tile and the warp then moves
        // 128-bit vectors k = lane, lane+32, ... of the run: loads (residual) and stores cover whole lines, where the lane-per-pixel
        // form touched 16 B of every 32 / 64 B (measured 2.1 TB/s for C_out = 16 against 4 TB/s for C_out = 8).
        // The residual of tile i+1 is requested before tile i is finished: its HBM latency hides behind a whole epilogue.
        const bool packed = P.out_cs == P.cout && (!P.res || P.res_cs == P.cout);
        const int cg_log2 = P.cout == 8 ? 1 : 2, cg = 1 << cg_log2;          // 128-bit vectors per pixel
        float* stg = reinterpret_cast<float*>(smem + C::STG_OFF) + q * (32 * TH_STG_PITCH);
        bool nvalid = false; size_t npix = 0; int nrun = 0; float4 nrr[4];
        TileWalk tw; tw.init(blockIdx.x, gridDim.x, P.tiles_x, P.tiles_y);
        int rt = blockIdx.x, rs = 0;                                           // tile and 4-row group of the NEXT request
        auto request = [&]() {                                                 // (tile, group) in processing order
            nvalid = false; nrun = 0;
            if (rt >= total_tiles) return;
            const int b = tw.b, x0 = tw.txi * TH_TWV, y0 = tw.tyi * ROWS + 4 * rs;
            const int py = y0 + q, px = x0 + lane;
            nvalid = lane < TH_TWV && py < P.H && px < P.W;
            if (packed) {
                npix = ((size_t)b * P.H + py) * P.W + x0;                      // first pixel of the warp's run
                nrun = py < P.H ? min(TH_TWV, P.W - x0) << cg_log2 : 0;        // vectors in the run
                if (P.res) {
                    const float4* rp = reinterpret_cast<const float4*>(P.res + npix * P.cout);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (lane + 32 * j < nrun) nrr[j] = __ldg(rp + lane + 32 * j);
                }
            } else {
                npix = ((size_t)b * P.H + py) * P.W + px;
                if (nvalid && P.res) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (4 * i < P.cout) nrr[i] = __ldg(reinterpret_cast<const float4*>(P.res + npix * P.res_cs) + i);
                }
            }
            if (++rs == SUB) { rs = 0; rt += gridDim.x; tw.next(); }
        };
        request();
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
          const int slot = it % NT;
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            const bool valid = nvalid; const size_t pix = npix; const int run = nrun;
            float4 rr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) rr[i] = nrr[i];
            request();
            if (sub == 0) {                                        // one barrier round trip per tile (SUB accumulators)
                tc::mbar_wait(&t_full[slot], (uint32_t)(it / NT) & 1u);
                tc::tc_fence_after();
            }
            uint32_t r[16];
            tc::tmem_ld16(tmem_base + (slot * SUB + sub) * TH_ACC_COLS + ((uint32_t)(q * 32) << 16), r);
            tc::tmem_ld_wait();
            if (sub == SUB - 1) {                                  // values are in registers: release the accumulators early
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&t_empty[slot]);    // one arrival per warp, not 32 serialized ones on the same barrier
            }
            if (packed) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < cg)
                        *reinterpret_cast<float4*>(stg + lane * TH_STG_PITCH + 4 * i) =
                            make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                __syncwarp();
                float4* op = reinterpret_cast<float4*>(P.out + pix * P.cout);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = lane + 32 * j;
                    if (k < run) {
                        const int part = k & (cg - 1);
                        float4 v = *reinterpret_cast<const float4*>(stg + (k >> cg_log2) * TH_STG_PITCH + 4 * part);
                        const float4 bq = *reinterpret_cast<const float4*>(sbias + 4 * part);
                        v.x += bq.x; v.y += bq.y; v.z += bq.z; v.w += bq.w;
                        if (P.res) { v.x += rr[j].x; v.y += rr[j].y; v.z += rr[j].z; v.w += rr[j].w; }
                        if (!(P.dbg & 1)) op[k] = v;
                    }
                }
                __syncwarp();                                          // the staging tile is rewritten by the next tile
            } else if (valid) {
                float4* op = reinterpret_cast<float4*>(P.out + pix * P.out_cs);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (4 * i < P.cout) {
                        float4 v = make_float4(__uint_as_float(r[4 * i]) + sbias[4 * i], __uint_as_float(r[4 * i + 1]) + sbias[4 * i + 1],
                                               __uint_as_float(r[4 * i + 2]) + sbias[4 * i + 2], __uint_as_float(r[4 * i + 3]) + sbias[4 * i + 3]);
                        if (P.res) { v.x += rr[i].x; v.y += rr[i].y; v.z += rr[i].z; v.w += rr[i].w; }
                        op[i] = v;
                    } else if (4 * i < P.out_cs) {
                        op[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // channel padding of the output stays zero
                    }
                }
                for (int c = 16; c < P.out_cs; c += 4) op[c / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, TH_TMEM_COLS);
}

int conv_thin_prepare(ConvThinParams& P, const ConvThinDesc& d) {
    memset(&P, 0, sizeof(P));
    const TensorNHWC& t = d.src;
    IPDM_REQUIRE(t.cs == 8 || t.cs == 16 || t.cs == 32, "conv_thin: operand channel stride %d must be 8, 16 or 32", t.cs);
    IPDM_REQUIRE(d.cout == 8 || d.cout == 16, "conv_thin: C_out %d must be 8 or 16", d.cout);
    IPDM_REQUIRE(d.ntaps == 1 || d.ntaps == 9, "conv_thin: 1x1 or 3x3");
    IPDM_REQUIRE(!t.bf16 && ((uintptr_t)t.p % 16) == 0 && d.out.cs % 4 == 0 && d.out.cs >= d.cout, "conv_thin: bad tensor layout");
    IPDM_REQUIRE(d.out.h == t.h && d.out.w == t.w && d.out.n == t.n, "conv_thin: output shape mismatch");
    P.H = t.h; P.W = t.w; P.batch = t.n; P.ntaps = d.ntaps; P.cs = t.cs; P.cout = d.cout;
    // output rows per tile: 4 (one accumulator) or 8 (two accumulators per barrier round trip; IPDM_THIN_ROWS=8)
    static const int env_rows = getenv("IPDM_THIN_ROWS") ? atoi(getenv("IPDM_THIN_ROWS")) : 4;
    P.rows = env_rows == 8 ? 8 : 4;
    P.tiles_x = ceil_div(P.W, TH_TWV); P.tiles_y = ceil_div(P.H, P.rows);
    const CUtensorMapSwizzle sw = t.cs == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : (t.cs == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    const uint64_t dims[4] = {(uint64_t)t.cs, (uint64_t)t.w, (uint64_t)t.h, (uint64_t)t.n};
    const uint64_t str[3] = {(uint64_t)t.cs * 4, (uint64_t)t.w * t.cs * 4, (uint64_t)t.h * t.w * t.cs * 4};
    const uint32_t box[4] = {(uint32_t)t.cs, TH_RP, (uint32_t)P.rows + 2, 1};
    IPDM_CHECK(tmap_encode(&P.mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, t.p, dims, str, box, sw));
    const uint64_t wd[2] = {(uint64_t)t.cs, (uint64_t)d.ntaps * 16};
    const uint64_t ws[1] = {(uint64_t)t.cs * 4};
    const uint32_t wb[2] = {(uint32_t)t.cs, (uint32_t)d.ntaps * 16};
    IPDM_CHECK(tmap_encode(&P.mapW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d.w_packed, wd, ws, wb, sw));
    P.out = d.out.p; P.out_cs = d.out.cs;
    P.bias = d.bias; P.bias_t_stride = d.bias_t_stride; P.t_dev = d.t_dev;
    P.res = d.res.p; P.res_cs = d.res.cs;
    P.norm_scale = d.norm_scale; P.norm_shift = d.norm_shift; P.act_silu = d.act_silu;
    IPDM_REQUIRE(!d.norm_scale || (d.norm_shift && t.c == t.cs), "conv_thin: the fused GroupNorm needs a dense source (C_in == channel stride)");
    static const int env_dbg = getenv("IPDM_THIN_DBG") ? atoi(getenv("IPDM_THIN_DBG")) : 0;
    P.dbg = env_dbg;
    if (P.dbg & 2) P.res = nullptr;
    IPDM_REQUIRE(!P.res || P.res_cs % 4 == 0, "conv_thin: residual channel stride must be a multiple of 4");
    return IPDM_OK;
}

template <int RB, int ROWS>
static int launch_thin(const ConvThinParams& P, cudaStream_t st) {
    static DeviceOnce once;
    using C = ThinCfg<RB, ROWS>;
    constexpr int smem = C::TOTAL;
    static_assert(C::CTAS * (smem + 1024) <= 227 * 1024, "thin conv: operand rings of the co-resident CTAs do not fit in shared memory");
    static_assert((3 * C::NSA + 2 * TH_NACC + 1) * 8 + 16 <= 384, "barrier block");
    if (once.need()) {
        IPDM_CHECK_CUDA(cudaFuncSetAttribute(conv_thin_kernel<RB, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    const int total = P.tiles_x * P.tiles_y * P.batch;
    conv_thin_kernel<RB, ROWS><<<std::min(total, kNumSMs * C::CTAS), TH_THREADS, smem, st>>>(P);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

int conv_thin_launch(const ConvThinParams& P, cudaStream_t st) {
    ProfScope prof(PROF_CONV_DIRECT, st, 4.0 * P.batch * (double)P.H * P.W * (P.cs + P.cout));    // bytes (same family as the direct convs)
    switch (P.cs * 16 + P.rows) {
        case 8 * 16 + 4: return launch_thin<32, 4>(P, st);
        case 16 * 16 + 4: return launch_thin<64, 4>(P, st);
        case 32 * 16 + 4: return launch_thin<128, 4>(P, st);
        case 8 * 16 + 8: return launch_thin<32, 8>(P, st);
        case 16 * 16 + 8: return launch_thin<64, 8>(P, st);
        case 32 * 16 + 8: return launch_thin<128, 8>(P, st);
    }
    set_error("conv_thin_launch: unsupported channel stride %d / rows %d", P.cs, P.rows);
    return IPDM_ERR_UNSUPPORTED;
}

}  // namespace ipdm
