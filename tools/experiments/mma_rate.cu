// Experiment: tcgen05.mma kind::tf32 issue rate (cycles per M128 x N x K8 instruction) as a function of the A-operand start row
// inside a SWIZZLE_128B tile (aligned to the 8-row swizzle atom or shifted by 1-2 rows) and of N.  Data is whatever is in smem.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../ipdm-pytorch_b200/csrc/tc.cuh"
using namespace ipdm;
namespace ipdm { void set_error(const char*, ...) {} }

template <int N>
__global__ void __launch_bounds__(128) k(int r0, int iters, long long* out, int bf16) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* done = (uint64_t*)(smem + 96 * 1024);
    uint32_t* slot = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 256; i += 128) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { tc::mbar_init(done, 1); tc::fence_barrier_init(); }
    if (warp == 2) tc::tmem_alloc(slot, 256);
    tc::fence_proxy_async();
    tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
    const uint32_t tm = *slot;
    if (warp == 0 && tc::elect_one()) {
        const uint32_t a = tc::smem_u32(smem) + r0 * 128, b = tc::smem_u32(smem + 48 * 1024);
        const uint64_t ad = tc::smem_desc_k_sw128(a), bd = tc::smem_desc_k_sw128(b);
        const uint32_t idesc = tc::make_idesc(bf16 ? tc::FMT_BF16 : tc::FMT_TF32, 128, N);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i)
            for (int kk = 0; kk < 4; ++kk) {
                if (bf16) tc::umma_f16(tm, ad + kk * 2, bd + kk * 2, idesc, 1u);
                else tc::umma_tf32(tm, ad + kk * 2, bd + kk * 2, idesc, 1u);
            }
        tc::umma_commit(done);
        tc::mbar_wait(done, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
    tc::tc_fence_before(); __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tm, 256);
}

template <int N> void run(int bf16) {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int r0 : {0, 1, 2, 4, 8, 33}) {
        const int iters = 2000;
        k<N><<<148, 128, 100 * 1024>>>(r0, iters, d, bf16);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%s N=%3d r0=%2d: %.1f cycles per MMA (%s)\n", bf16 ? "bf16" : "tf32", N, r0, (double)c / (iters * 4.0), cudaGetErrorString(e));
    }
}
int main() { run<128>(0); run<256>(0); run<64>(0); run<128>(1); run<256>(1); return 0; }
