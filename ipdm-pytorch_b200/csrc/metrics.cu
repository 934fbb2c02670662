// On-device image-quality metrics (SURVEY "next" row N1): PSNR and SSIM of reconstructed slices against the full-dose
// image, with the definitions the reference uses in metric_calculate (Utils/train_test_utils.py:789-799):
//   both images in "pixel" units = miu2pixel(mu) (Dataset/npz_data_loader.py:20-36: mu -> HU -> [-1024,3072] HU window
//   mapped to [0,1] and clipped), NaNs of the test image replaced by 0.5 (:792);
//   PSNR = skimage.peak_signal_noise_ratio(data_range=1) = 10 log10(1 / mean((ref - test)^2));
//   SSIM = skimage.structural_similarity(win_size=11, data_range=1), skimage 0.19 defaults: uniform 11x11 window,
//          K1 = 0.01, K2 = 0.03, sample covariance (N/(N-1)), mean of the SSIM map cropped by (win-1)/2 pixels.
// Only the cropped interior enters the mean, so the filter's boundary mode never matters: every window lies inside.
// Window sums are accumulated in fp64 (variance = E[x^2] - E[x]^2 cancels badly in fp32 on smooth CT images).
#include "common.cuh"

namespace ipdm {

constexpr int MT_TILE = 32, MT_WIN_MAX = 15, MT_HALO = MT_TILE + MT_WIN_MAX - 1;

__global__ void __launch_bounds__(256)
miu2pixel_kernel(const float* __restrict__ mu, float* __restrict__ pix, size_t n, float lo, float hi) {
    const float miu_water = 0.183f;                             // npz_data_loader.py:5
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float hu = (mu[i] - miu_water) * 1e3f / miu_water - 24.f;
        float p = (hu - lo) / (hi - lo);
        p = hu < lo ? 0.f : (hu > hi ? 1.f : p);
        pix[i] = isnan(p) ? 0.5f : p;
    }
}

// grid (tiles_x, tiles_y, batch), 256 threads = 32 x 8, four rows per thread.  partials[b][blk] = {sum of squared errors over
// the tile, sum of SSIM over the tile's interior pixels}; fixed-order reductions => deterministic.
__global__ void __launch_bounds__(256)
metrics_tile_kernel(const float* __restrict__ test, const float* __restrict__ ref, int H, int W, int win, double* __restrict__ partials) {
    __shared__ float st[MT_HALO][MT_HALO + 1], sr[MT_HALO][MT_HALO + 1];
    __shared__ double red[2][256];
    const int b = blockIdx.z, r = win / 2, ext = MT_TILE + 2 * r;
    const int x0 = blockIdx.x * MT_TILE, y0 = blockIdx.y * MT_TILE;
    const float* tb = test + (size_t)b * H * W;
    const float* rb = ref + (size_t)b * H * W;
    for (int i = threadIdx.x; i < ext * ext; i += 256) {
        const int ly = i / ext, lx = i - ly * ext;
        const int gy = y0 - r + ly, gx = x0 - r + lx;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        float t = in ? tb[(size_t)gy * W + gx] : 0.f;
        st[ly][lx] = isnan(t) ? 0.5f : t;                           // metric_calculate :792
        sr[ly][lx] = in ? rb[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double np = (double)win * win, cov = np / (np - 1.0), c1 = 0.01 * 0.01, c2 = 0.03 * 0.03;
    double se = 0, ss = 0;
    for (int k = 0; k < 4; ++k) {
        const int ly = ty * 4 + k, gy = y0 + ly, gx = x0 + tx;
        if (gy >= H || gx >= W) continue;
        const double d = (double)sr[ly + r][tx + r] - (double)st[ly + r][tx + r];
        se += d * d;
        if (gy < r || gy >= H - r || gx < r || gx >= W - r) continue;      // outside the cropped SSIM map
        double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
        for (int wy = 0; wy < win; ++wy)
            for (int wx = 0; wx < win; ++wx) {
                const double x = st[ly + wy][tx + wx], y = sr[ly + wy][tx + wx];
                sx += x; sy += y; sxx += x * x; syy += y * y; sxy += x * y;
            }
        const double ux = sx / np, uy = sy / np;
        const double vx = cov * (sxx / np - ux * ux), vy = cov * (syy / np - uy * uy), vxy = cov * (sxy / np - ux * uy);
        ss += ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2));
    }
    red[0][threadIdx.x] = se; red[1][threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x, nblk = (size_t)gridDim.x * gridDim.y;
        partials[((size_t)b * nblk + blk) * 2] = red[0][0];
        partials[((size_t)b * nblk + blk) * 2 + 1] = red[1][0];
    }
}

__global__ void metrics_finalize_kernel(const double* __restrict__ partials, int nblk, double npix, double nint, double* __restrict__ out) {
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    double se = 0, ss = 0;
    for (int i = 0; i < nblk; ++i) { se += partials[((size_t)b * nblk + i) * 2]; ss += partials[((size_t)b * nblk + i) * 2 + 1]; }
    const double mse = se / npix;
    out[2 * b] = mse > 0 ? 10.0 * log10(1.0 / mse) : INFINITY;
    out[2 * b + 1] = nint > 0 ? ss / nint : NAN;
}

}  // namespace ipdm

using namespace ipdm;

extern "C" size_t ipdm_metrics_workspace_bytes(int batch, int h, int w) {
    if (batch <= 0 || h <= 0 || w <= 0) return 0;
    return (size_t)batch * ceil_div(w, MT_TILE) * ceil_div(h, MT_TILE) * 2 * sizeof(double);
}

extern "C" int ipdm_psnr_ssim(const float* test, const float* ref, int batch, int h, int w, int win_size, double* out, void* workspace, void* stream) {
    IPDM_REQUIRE(test && ref && out && workspace && batch > 0, "ipdm_psnr_ssim: null argument");
    IPDM_REQUIRE(win_size >= 3 && win_size <= MT_WIN_MAX && (win_size & 1) && h >= win_size && w >= win_size,
                 "ipdm_psnr_ssim: win_size %d must be odd, in [3, %d] and not larger than the image", win_size, MT_WIN_MAX);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(ceil_div(w, MT_TILE), ceil_div(h, MT_TILE), batch);
    metrics_tile_kernel<<<grid, 256, 0, st>>>(test, ref, h, w, win_size, (double*)workspace);
    count_launch();
    const int r = win_size / 2;
    metrics_finalize_kernel<<<batch, 32, 0, st>>>((const double*)workspace, (int)(grid.x * grid.y), (double)h * w, (double)(h - 2 * r) * (w - 2 * r), out);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}

extern "C" int ipdm_miu2pixel(const float* mu, float* pix, size_t n, float hu_lo, float hu_hi, void* stream) {
    IPDM_REQUIRE(mu && pix && hu_hi > hu_lo, "ipdm_miu2pixel: bad arguments");
    if (n == 0) return IPDM_OK;
    const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 8);
    miu2pixel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mu, pix, n, hu_lo, hu_hi);
    count_launch();
    IPDM_CHECK_LAUNCH();
    return IPDM_OK;
}
