"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the image-quality metrics the reference computes in
`progressive_domain_denoiser.metric_calculate` (Utils/train_test_utils.py:789-799) through scikit-image:

    compare_psnr(fdct, ld, data_range=1)                     skimage.metrics.peak_signal_noise_ratio
    compare_ssim(fdct, ld, win_size=11, data_range=1)        skimage.metrics.structural_similarity

scikit-image is a third-party dependency that is NOT vendored in /root/reference and not installed in this image (the reference's
requirements pin scikit-image 0.19.x); its published algorithm is restated here:
  PSNR = 10 log10(data_range^2 / mean((a - b)^2))
  SSIM (Wang et al. 2004 with skimage defaults): uniform win x win filter for the local means and second moments, sample covariance
  (cov_norm = N / (N - 1), N = win^2), C1 = (0.01 R)^2, C2 = (0.03 R)^2, S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),
  result = mean of S cropped by (win - 1) // 2 pixels on every side (so the filter's boundary mode never enters).
Pinning (tests/test_oracle.py): (1) the published definition and its analytic properties (brute-force windows); (2) `ssim_skimage_path`
below, which replays skimage 0.19's `structural_similarity` statement by statement ON THE SAME THIRD-PARTY FILTER it calls
(`scipy.ndimage.uniform_filter`, mode "reflect", float64; scipy IS installed here) followed by its crop-and-mean -- so the window
alignment, the boundary handling and the crop are checked against the library arithmetic skimage itself runs.  skimage's own wrapper
code (argument checks, dtype promotion) is the only part not executed: PARITY UNPINNED against the skimage package itself.
`miu2pixel` follows Dataset/npz_data_loader.py:20-36; NaN -> 0.5 follows metric_calculate :792."""
import numpy as np

MIU_WATER = np.float32(0.183)


def miu2pixel(mu, hu_range=(-1024.0, 3072.0)):
    mu = np.asarray(mu, dtype=np.float32)
    hu = (mu - MIU_WATER) * np.float32(1e3) / MIU_WATER - np.float32(24)
    lo, hi = np.float32(hu_range[0]), np.float32(hu_range[1])
    img = (hu - lo) / (hi - lo)
    img[hu < lo] = 0
    img[hu > hi] = 1
    return img


def psnr(ref, test, data_range=1.0):
    test = np.where(np.isnan(test), 0.5, test).astype(np.float64)
    return 10.0 * np.log10(data_range ** 2 / np.mean((np.asarray(ref, np.float64) - test) ** 2))


def _box(a, win):
    c = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), axis=0), axis=1)
    return (c[win:, win:] - c[:-win, win:] - c[win:, :-win] + c[:-win, :-win]) / (win * win)


def ssim(ref, test, win_size=11, data_range=1.0):
    x = np.where(np.isnan(test), 0.5, test).astype(np.float64)
    y = np.asarray(ref, np.float64)
    n = win_size * win_size
    cov = n / (n - 1.0)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    ux, uy = _box(x, win_size), _box(y, win_size)                       # valid windows only == the cropped interior
    vx, vy, vxy = cov * (_box(x * x, win_size) - ux * ux), cov * (_box(y * y, win_size) - uy * uy), cov * (_box(x * y, win_size) - ux * uy)
    return float(np.mean(((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))))


def ssim_skimage_path(ref, test, win_size=11, data_range=1.0):
    """skimage.metrics.structural_similarity (0.19.x: skimage/metrics/_structural_similarity.py) replayed on scipy.ndimage.uniform_filter,
    the function skimage calls for gaussian_weights=False: same filter, same sample covariance, same crop, float64 throughout."""
    from scipy.ndimage import uniform_filter
    im1 = np.asarray(ref, dtype=np.float64)
    im2 = np.where(np.isnan(test), 0.5, test).astype(np.float64)
    K1, K2 = 0.01, 0.03
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)                                   # use_sample_covariance=True
    ux, uy = uniform_filter(im1, size=win_size), uniform_filter(im2, size=win_size)
    uxx, uyy, uxy = uniform_filter(im1 * im1, size=win_size), uniform_filter(im2 * im2, size=win_size), uniform_filter(im1 * im2, size=win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    A1, A2, B1, B2 = 2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
