// Experiment: can a K-major SWIZZLE_128B UMMA A-operand start at an arbitrary 128-byte row of a larger TMA-written tile?
// D[128 x 32] = A[rows r0 .. r0+127, 32 fp32] * B^T with B = I(32) (as tf32), so D must equal the selected rows of A.
// Tries base_offset = 0 and base_offset = (addr >> 7) & 7 for several r0.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ipdm-pytorch_b200/csrc/tc.cuh"
using namespace ipdm;

struct Params { CUtensorMap mapA, mapB; int r0; int use_base_offset; float* out; };

__global__ void __launch_bounds__(128) k(const __grid_constant__ Params P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 192 rows x 128 B = 24 KB
    uint8_t* sB = smem + 32768;               // 32 rows x 128 B
    uint64_t* bar = (uint64_t*)(smem + 40960);
    uint64_t* done = bar + 1;
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(done, 1); tc::fence_barrier_init(); }
    if (warp == 2) tc::tmem_alloc(slot, 32);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tm = *slot;
    if (warp == 0 && tc::elect_one()) {
        tc::mbar_expect_tx(bar, 192 * 128 + 32 * 128);
        tc::tma_load_2d(sA, &P.mapA, bar, 0, 0);
        tc::tma_load_2d(sB, &P.mapB, bar, 0, 0);
        tc::mbar_wait(bar, 0);
        tc::tc_fence_after();
        const uint32_t a_addr = tc::smem_u32(sA) + P.r0 * 128;
        uint64_t ad = tc::smem_desc_k_sw128(a_addr);
        if (P.use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
        const uint64_t bd = tc::smem_desc_k_sw128(tc::smem_u32(sB));
        const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, 32);
        for (int kk = 0; kk < 4; ++kk) tc::umma_tf32(tm, ad + kk * 2, bd + kk * 2, idesc, kk != 0);
        tc::umma_commit(done);
    }
    __syncwarp();
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
    uint32_t r[32];
    tc::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), r);
    tc::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) P.out[(warp * 32 + lane) * 32 + i] = __uint_as_float(r[i]);
    tc::tc_fence_before(); __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tm, 32);
}

namespace ipdm { void set_error(const char*, ...) {} }
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    const int R = 192;
    std::vector<float> hA(R * 32), hB(32 * 32, 0.f);
    for (int r = 0; r < R; ++r) for (int c = 0; c < 32; ++c) hA[r * 32 + c] = (float)(r * 32 + c);   // exact in tf32 up to 2048.. use small ints
    for (auto& v : hA) v = (float)((int)v % 1021);
    for (int i = 0; i < 32; ++i) hB[i * 32 + i] = 1.f;
    float *dA, *dB, *dO;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dO, 128 * 32 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    Params P;
    cuuint64_t dims[2] = {32, (cuuint64_t)R}, str[1] = {128}; cuuint32_t box[2] = {32, (cuuint32_t)R}, es[2] = {1, 1};
    enc(&P.mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dimsb[2] = {32, 32}; cuuint32_t boxb[2] = {32, 32};
    enc(&P.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    P.out = dO;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 + 1024);
    std::vector<float> hO(128 * 32);
    for (int ubo = 0; ubo < 2; ++ubo)
        for (int r0 : {0, 1, 2, 3, 5, 8, 9, 33, 34, 63}) {
            P.r0 = r0; P.use_base_offset = ubo;
            k<<<1, 128, 48 * 1024 + 1024>>>(P);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("r0=%d base_offset=%d: CUDA error %s\n", r0, ubo, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m) for (int c = 0; c < 32; ++c) if (hO[m * 32 + c] != hA[(r0 + m) * 32 + c]) ++bad;
            printf("r0=%2d base_offset=%d: %s (%d mismatches; D[0][0..3] = %g %g %g %g, expect %g %g)\n", r0, ubo, bad ? "MISMATCH" : "ok", bad,
                   hO[0], hO[1], hO[2], hO[3], hA[r0 * 32], hA[r0 * 32 + 1]);
        }
    return 0;
}
