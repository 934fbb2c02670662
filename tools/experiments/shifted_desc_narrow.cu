// Experiment: K-major UMMA operands with 32-byte (SWIZZLE_32B) and 64-byte (SWIZZLE_64B) rows, A starting at an arbitrary row.
// D[128 x 16] = A[rows r0.., KW fp32] * B^T, B[16 x KW] = first 16 columns... B = selection matrix (B[n][k] = (k == n % KW)).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../ipdm-pytorch_b200/csrc/tc.cuh"
using namespace ipdm;
namespace ipdm { void set_error(const char*, ...) {} }

struct Params { CUtensorMap mapA, mapB; int r0, kw, layout, sbo; float* out; };

__device__ __forceinline__ uint64_t desc(uint32_t addr, int layout, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}

__global__ void __launch_bounds__(128) k(const __grid_constant__ Params P) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem; uint8_t* sB = smem + 16384;
    uint64_t* bar = (uint64_t*)(smem + 20480); uint64_t* done = bar + 1; uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(done, 1); tc::fence_barrier_init(); }
    if (warp == 2) tc::tmem_alloc(slot, 32);
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tm = *slot;
    if (warp == 0 && tc::elect_one()) {
        tc::mbar_expect_tx(bar, 192 * P.kw * 4 + 16 * P.kw * 4);
        tc::tma_load_2d(sA, &P.mapA, bar, 0, 0);
        tc::tma_load_2d(sB, &P.mapB, bar, 0, 0);
        tc::mbar_wait(bar, 0);
        tc::tc_fence_after();
        const uint32_t a_addr = tc::smem_u32(sA) + P.r0 * P.kw * 4;
        const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, 16);
        for (int kk = 0; kk < P.kw / 8; ++kk)
            tc::umma_tf32(tm, desc(a_addr, P.layout, P.sbo) + kk * 2, desc(tc::smem_u32(sB), P.layout, P.sbo) + kk * 2, idesc, kk != 0);
        tc::umma_commit(done);
    }
    __syncwarp();
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
    uint32_t r[16];
    tc::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), r);
    tc::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) P.out[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    tc::tc_fence_before(); __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tm, 32);
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    for (int kw : {8, 16}) {
        const int R = 192;
        std::vector<float> hA(R * kw), hB(16 * kw, 0.f);
        for (int i = 0; i < R * kw; ++i) hA[i] = (float)(i % 1021);
        for (int n = 0; n < 16; ++n) hB[n * kw + (n % kw)] = 1.f;          // D[m][n] = A[m][n % kw]
        float *dA, *dB, *dO;
        cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dO, 128 * 16 * 4);
        cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
        Params P; P.kw = kw; P.layout = kw == 8 ? 6 : 4; P.sbo = 8 * kw * 4; P.out = dO;
        const CUtensorMapSwizzle sw = kw == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
        cuuint64_t dims[2] = {(cuuint64_t)kw, (cuuint64_t)R}, str[1] = {(cuuint64_t)kw * 4}; cuuint32_t box[2] = {(cuuint32_t)kw, (cuuint32_t)R}, es[2] = {1, 1};
        CUresult r1 = enc(&P.mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t dimsb[2] = {(cuuint64_t)kw, 16}; cuuint32_t boxb[2] = {(cuuint32_t)kw, 16};
        CUresult r2 = enc(&P.mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("kw=%d encode rc %d %d\n", kw, (int)r1, (int)r2);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
        std::vector<float> hO(128 * 16);
        for (int r0 : {0, 1, 2, 3, 4, 5, 7, 8, 9, 33, 34, 35}) {
            P.r0 = r0;
            k<<<1, 128, 32 * 1024>>>(P);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kw=%d r0=%d: CUDA error %s\n", kw, r0, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) if (hO[m * 16 + n] != hA[(r0 + m) * kw + (n % kw)]) ++bad;
            printf("rows of %2d B (swizzle %s) r0=%2d: %s (%d mismatches)\n", kw * 4, kw == 8 ? "32B" : "64B", r0, bad ? "MISMATCH" : "ok", bad);
        }
    }
    return 0;
}
