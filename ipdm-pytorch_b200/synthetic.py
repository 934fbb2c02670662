"""Seeded synthetic inputs for the IPDM progressive path (host side, numpy only).

There is no Mayo data offline, so tests, the oracle's golden generator and
bench.py all draw their inputs from here: an analytic ellipse phantom, its
equiangular fan-beam line integrals in the geometry of the reference FBP
(Recon/FBP_kernel.py:32-44: D = 59.5 cm, 912 detectors at pitch 0.0010125 rad
offset by 3.75 bins, 2000 views of 0.18 deg, 512^2 image over 42 cm), and the
low-dose noise model of Utils/Low_dose_CT_simulate.py:38-44.

Frames.  The reference ``FBP.convert(flip=True)`` reverses the detector axis on
the way in and the column axis on the way out (FBP_kernel.py:100,118).  The
sinogram returned here is in the *input* frame of ``convert`` and the phantom
raster in its *output* frame, so ``convert(sinogram)`` approximates ``raster``.
"""
import numpy as np

N_VIEWS = 2000
N_DET = 912
N_PIX = 512
SRC_ISO = 59.5            # cm, FBP_kernel.py:32
DET_PITCH = 0.0010125     # rad, FBP_kernel.py:35
HALF_FOV = 21.0           # cm, FBP_kernel.py:44
MU_WATER = 0.183          # 1/cm, Dataset/npz_data_loader.py:15


def view_angles():
    """theta_k = k * 0.18 deg in radians, fp64 (FBP_kernel.py:38)."""
    return np.arange(0, 359.82 + 0.18, 0.18) / 180 * np.pi


def detector_angles():
    """nda[k] = (-451.75 + k) * da as f32 (FBP_kernel.py:39-40)."""
    m, da = N_DET, DET_PITCH
    return np.arange((-m / 2 + 0.5 + 3.75) * da, (m / 2 - 0.5 + 3.75 + 1) * da, da).astype("float32")


def phantom_ellipses(slice_id=0):
    """Shepp-Logan-like body section: rows of (x0, y0, a, b, angle_rad, mu) in cm / cm^-1.

    Densities are additive.  ``slice_id`` jitters the inner structures so that
    different slices of a synthetic volume differ.
    """
    rng = np.random.default_rng(1000 + int(slice_id))
    j = lambda s: float(rng.uniform(-s, s))
    w = MU_WATER
    e = [
        (0.0, 0.0, 16.5, 12.5, 0.0, 1.9 * w),                      # dense outer shell
        (0.0, -0.2, 15.6, 11.7, 0.0, -0.88 * w),                   # soft tissue interior (~1.02 w)
        (5.2 + j(.4), 0.3 + j(.4), 2.6, 4.8, np.deg2rad(-18 + j(6)), -0.22 * w),   # "lung" right
        (-5.2 + j(.4), 0.3 + j(.4), 3.2, 5.4, np.deg2rad(18 + j(6)), -0.22 * w),   # "lung" left
        (0.0 + j(.3), 5.6 + j(.3), 3.4, 2.2, 0.0, 0.06 * w),
        (0.0 + j(.3), 1.6 + j(.3), 0.9, 0.9, 0.0, 0.10 * w),
        (0.0 + j(.3), -1.6 + j(.3), 0.7, 0.7, 0.0, 0.10 * w),
        (-1.9 + j(.3), -7.2 + j(.3), 0.9, 0.45, 0.0, 0.08 * w),
        (0.0 + j(.3), -7.2 + j(.3), 0.45, 0.45, 0.0, 0.08 * w),
        (1.6 + j(.3), -7.2 + j(.3), 0.45, 0.9, 0.0, 0.08 * w),
        (0.0, -10.2, 1.6, 1.1, 0.0, 0.75 * w),                     # "spine"
    ]
    return np.asarray(e, dtype=np.float64)


def fan_sinogram(ellipses, views=None):
    """Exact line integrals of the ellipse phantom, [n_views, 912] f64, convert()-input frame.

    Filtered sample j (after convert's detector flip) is read by the
    backprojector at fan angle nda[j] + da/2 (FBP_kernel.py:158-163: weight
    lam = u - k on q[k], 1-lam on q[k-1], u = (alpha - nda[0])/da + 0.5), so
    rays are traced at those angles and the detector axis is reversed at the end.
    """
    th = view_angles() if views is None else np.asarray(views, dtype=np.float64)
    al = detector_angles().astype(np.float64) + 0.5 * DET_PITCH
    ct, st = np.cos(th)[:, None], np.sin(th)[:, None]
    ca, sa = np.cos(al)[None, :], np.sin(al)[None, :]
    # source and unit ray direction in the backprojector's native frame:
    #   s = x sin(th) + y cos(th), c = D + x cos(th) - y sin(th), alpha = atan(s / c)
    sx, sy = -SRC_ISO * ct, SRC_ISO * st
    dx, dy = ca * ct + sa * st, -ca * st + sa * ct
    out = np.zeros((th.size, al.size), dtype=np.float64)
    for x0, y0, a, b, ang, mu in np.asarray(ellipses, dtype=np.float64):
        c, s = np.cos(ang), np.sin(ang)
        px, py = sx - x0, sy - y0
        q0x, q0y = (c * px + s * py) / a, (-s * px + c * py) / b
        qdx, qdy = (c * dx + s * dy) / a, (-s * dx + c * dy) / b
        A = qdx * qdx + qdy * qdy
        Bq = q0x * qdx + q0y * qdy
        C = q0x * q0x + q0y * q0y - 1.0
        disc = Bq * Bq - A * C
        out += mu * 2.0 * np.sqrt(np.maximum(disc, 0.0)) / A
    return out[:, ::-1].copy()


def rasterize(ellipses, n=N_PIX, oversample=2):
    """mu image [n, n] f32 of the phantom in convert()'s output frame (column-flipped)."""
    m = n * oversample
    px = 2 * HALF_FOV / m
    idx = np.arange(m, dtype=np.float64)
    x = (idx - (m - 1) / 2) * px
    y = ((m - 1) / 2 - idx) * px
    X, Y = np.meshgrid(x, y)
    img = np.zeros((m, m), dtype=np.float64)
    for x0, y0, a, b, ang, mu in np.asarray(ellipses, dtype=np.float64):
        c, s = np.cos(ang), np.sin(ang)
        u = (c * (X - x0) + s * (Y - y0)) / a
        v = (-s * (X - x0) + c * (Y - y0)) / b
        img += mu * ((u * u + v * v) <= 1.0)
    img = img.reshape(n, oversample, n, oversample).mean(axis=(1, 3))
    return img[:, ::-1].astype(np.float32)


def add_noise(sino, factor=0.25, rng=None):
    """Low-dose noise of Low_dose_CT_simulate.py:38-44 (Ne = 5.8, N0 = 1.4e5)."""
    rng = np.random.default_rng(0) if rng is None else rng
    ne, n0 = 5.8, 1.4e5
    n = rng.standard_normal(sino.shape)
    ex = np.exp(sino)
    return sino + np.sqrt((1 - factor) * ex * (1 + ((1 + factor) * ne * ex) / (factor * n0)) / (factor * n0)) * n


def make_slice(slice_id=0, dose=0.25):
    """Returns (ld_sinogram [2000,912] f32 >= 0, nd_sinogram f32, nd_image [512,512] f32)."""
    ell = phantom_ellipses(slice_id)
    clean = fan_sinogram(ell)
    noisy = add_noise(clean, dose, np.random.default_rng(5000 + int(slice_id)))
    return (np.clip(noisy, 0, None).astype(np.float32), clean.astype(np.float32), rasterize(ell))


def cheap_sinogram(batch, seed=0):
    """Gaussian-profile sinograms [batch, 2000, 912] f32 for kernel-only benchmarks."""
    rng = np.random.default_rng(seed)
    d = np.arange(N_DET, dtype=np.float32)[None, None, :]
    v = np.arange(N_VIEWS, dtype=np.float32)[None, :, None]
    base = 4 * np.exp(-((d - 456) / 260) ** 2) * (1 + 0.1 * np.sin(2 * np.pi * v / N_VIEWS))
    out = base + rng.normal(0, 0.05, size=(batch, N_VIEWS, N_DET)).astype(np.float32)
    return np.clip(out, 0, None).astype(np.float32)


def noise_tape(shape, count, seed):
    """``count`` standard-normal tensors of ``shape`` from a seeded torch CPU generator.

    The reference draws with torch.randn_like; parity runs replace those draws
    with this tape, consumed in call order.
    """
    import torch
    g = torch.Generator().manual_seed(int(seed))
    return [torch.randn(shape, generator=g, dtype=torch.float32) for _ in range(count)]
