"""Sinogram data sets: ``proj_geom.toml`` + ``sinogram.npy [angles, H, W]`` (``ctdr/dataset/dataset.py:54-130``)."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset

from ..utils import util_sino


class SinoDataset(Dataset):
    def __init__(self, root_dir, nmaterials=2, eta=0, transform=None, dtype=torch.float32):
        self.nmaterials, self.eta, self.p_original = nmaterials, eta, None
        toml_path, pkl_path = os.path.join(root_dir, "proj_geom.toml"), os.path.join(root_dir, "proj_geom.pkl")
        if os.path.exists(toml_path):
            self.proj_geom = util_sino.read_from_toml(toml_path)
        elif os.path.exists(pkl_path):
            with open(pkl_path, "rb") as fh:
                self.proj_geom = pickle.load(fh)
        else:
            raise Exception("There is no file of proj_geom.toml or proj_geom.pkl")
        try:
            p = np.load(os.path.join(root_dir, "sinogram.npy"))
        except OSError:
            raise Exception("Projection data not found")
        self._finish(p, dtype)

    @classmethod
    def from_arrays(cls, proj_geom, sinogram, nmaterials=2, eta=0, dtype=torch.float32):
        """Build a data set from memory (synthetic targets: the reference ships no sinograms)."""
        self = cls.__new__(cls)
        self.nmaterials, self.eta, self.proj_geom = nmaterials, eta, proj_geom
        self._finish(np.asarray(sinogram), dtype)
        return self

    def _finish(self, p, dtype):
        self.p_original = p.copy()
        self.p = p
        if self.eta > 0.0:
            self.impose_noise()
        self.p = torch.tensor(self.p, dtype=dtype)
        self.nangles = self.p.shape[0] if self.p.dim() == 3 else self.p.shape[1]

    def __len__(self):
        return self.p.shape[0]

    def __getitem__(self, idx):
        return idx, self.p[idx, :]

    def impose_noise(self):
        """Gaussian noise of relative level eta, seed 1, clipped at 0 (``dataset.py:121-130``)."""
        np.random.seed(1)
        e_hat = np.random.normal(size=self.p_original.shape)
        relative = np.sqrt(np.sum(self.p_original ** 2)) / np.sqrt(np.sum(e_hat ** 2))
        self.p = self.p_original + self.eta * e_hat * relative
        self.p[self.p < 0.0] = 0.0
        print("succesfully impose noise")
