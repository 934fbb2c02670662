"""Unit conversions and the per-slice .npy/.npz dataset of the reference's Dataset/npz_data_loader.py.

Conversions (:9-52) are elementwise host/torch helpers used by metrics and by the 1-HU parity
tolerance; they accept numpy arrays and torch tensors alike.  The dataset class keeps the reference's
constructor and item layout (ld_img, fd_proj, fd_img, ld_proj) but splits paths with os.sep instead
of the reference's hard-coded backslash (:122-126).
"""
import glob
import os

import numpy as np
import torch
from torch.utils.data.dataset import Dataset

MIU_WATER = 0.183
DEFAULT_WINDOW = (-1024, 3072)


def pixel2HU(img, window=None):
    lo, hi = DEFAULT_WINDOW if window is None else window
    return img * (hi - lo) + lo


def HU2miu(HU):
    return MIU_WATER + ((HU + 24) * MIU_WATER / 1e3)


def miu2HU(miu):
    return (miu - MIU_WATER) * 1e3 / MIU_WATER - 24


def HU2pixel(HU, new_window=None):
    lo, hi = DEFAULT_WINDOW if new_window is None else new_window
    img = (HU - lo) / (hi - lo)
    img[HU < lo] = 0
    img[HU > hi] = 1
    return img


def miu2pixel(miu, HU_range=None):
    return HU2pixel(miu2HU(miu), HU_range)


def pixel2miu(pix):
    return HU2miu(pixel2HU(pix))


def reset_window_centre(img, new_window=None, origin_window=None):
    origin_window = list(DEFAULT_WINDOW) if origin_window is None else origin_window
    new_window = origin_window if new_window is None else new_window
    hu = img * (origin_window[1] - origin_window[0]) + origin_window[0]
    out = (hu - new_window[0]) / (new_window[1] - new_window[0])
    out[hu < new_window[0]] = 0
    out[hu > new_window[1]] = 1
    return out


def _as_chw(a):
    """torchvision ToTensor() on a 2-D float array only prepends a channel axis (no scaling)."""
    return torch.from_numpy(np.ascontiguousarray(a))[None]


class Siemens_dataset_npz(Dataset):
    """One file per slice under <root>/<patient>/<file>; float32 mu image [512,512] or sinogram [2000,912]."""

    _ORDER = ("ldimg", "fdproj", "fdimg", "ldproj")

    def __init__(self, ldproj_path=None, ldimg_path=None, fdproj_path=None, fdimg_path=None, proj_clip=False,
                 img_clip=True, data_type='siemens', patch=None, patch_per_image=None, assign=None):
        self.patch, self.patch_per_image = patch, patch_per_image
        self.proj_clip, self.img_clip, self.data_type = proj_clip, img_clip, data_type
        self.patient_name = self.slice_name = None
        self.roots = dict(ldimg=ldimg_path, fdproj=fdproj_path, fdimg=fdimg_path, ldproj=ldproj_path)
        self.ldproj_path, self.ldimg_path, self.fdproj_path, self.fdimg_path = ldproj_path, ldimg_path, fdproj_path, fdimg_path
        self.files = {}
        for key in ("fdimg", "fdproj", "ldimg", "ldproj"):          # the reference's precedence for names / length
            root = self.roots[key]
            if root is None:
                continue
            names = sorted(glob.glob(os.path.join(root, "*", "*")))
            if assign is not None and key.startswith("fd"):
                names = [n for n in names if os.path.basename(os.path.dirname(n)) in assign]
            self.files[key] = names
            if self.patient_name is None:
                self.patient_name = [os.path.basename(os.path.dirname(n)) for n in names]
                stem = (lambda n: os.path.basename(n).split(".")[-4]) if data_type == "mayo" else \
                       (lambda n: os.path.basename(n).split(".")[0])
                self.slice_name = [stem(n) for n in names]
        setattr(self, "fdimg_file_name", self.files.get("fdimg"))
        setattr(self, "fdproj_file_name", self.files.get("fdproj"))
        setattr(self, "ldimg_file_name", self.files.get("ldimg"))
        setattr(self, "ldproj_file_name", self.files.get("ldproj"))

    @staticmethod
    def get_data(file_path):
        return np.load(file_path)["arr_0"] if file_path.endswith("npz") else np.load(file_path)

    def _load(self, key, path):
        a = self.get_data(path)
        if key.endswith("proj") and self.proj_clip:
            a = a / 10
        t = _as_chw(a)
        return self.get_patch(t) if self.patch is not None else t

    def __getitem__(self, idx):
        return [self._load(k, self.files[k][idx]) if k in self.files else None for k in self._ORDER]

    def __len__(self):
        for key in ("fdimg", "fdproj", "ldimg", "ldproj"):
            if key in self.files:
                return len(self.files[key])
        return 0

    def get_data_from_name(self, patient_name, slice_name):
        out = []
        for k in self._ORDER:
            if k not in self.files:
                out.append(None)
                continue
            path = [n for n in self.files[k] if patient_name in n and slice_name in n][0]
            a = self.get_data(path)
            out.append(_as_chw(a / 10 if (k.endswith("proj") and self.proj_clip) else a))
        return out

    def get_patch(self, data):
        ph, pw = self.patch
        out = torch.zeros((self.patch_per_image, ph, pw))
        for i in range(self.patch_per_image):
            y = int(torch.randint(0, data.shape[-2] - ph + 1, (1,)))
            x = int(torch.randint(0, data.shape[-1] - pw + 1, (1,)))
            out[i] = data[0, y:y + ph, x:x + pw]
        return out

    @staticmethod
    def collate(batch_data):
        cols = []
        for j in range(4):
            items = [b[j] for b in batch_data]
            cols.append(torch.stack(items, dim=0) if items[0] is not None else None)
        return tuple(cols)
