"""Relevancy store — what `generate_relevancy.py dataset` writes per RGB frame and what
`SceneUnderstandDataset.load_patches` reads back (reference generate_relevancy.py:63-146, dataset.py:817-872):

  write: maps [P,H,W] -> nearest-exact resize to the storage grid (128 x 128) -> append the mean map over the labels as
         row "mean" -> rows appended to the scene's `saliencies` dataset; labels (+ "mean") and L2-normalised text
         features (+ their mean, re-normalised) stored under `data/saliencies/<rgb>|<config>|...`;
  read : selected rows minus the stored mean map -> bilinear (align_corners=False) back to the image size (x50 at
         dataset.py:1049-1054 = `gain`).

The arithmetic runs on the GPU (csrc/store.cu); the container is self-describing and mirrors the reference's HDF5 key
layout.  h5py is not installed in this image, so the default backend is a NumPy `.npz` archive (deflate, the same codec
family as the reference's gzip-9 datasets) in which an HDF5 *region reference* to row i of `saliencies` is stored as
the row index i; when h5py is importable `RelevancyStore(path, backend="h5")` writes the reference's exact layout."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ._lib import check, f32, i32, lib, ptr, stream_ptr

STORAGE_SHAPE = (128, 128)


def pack_maps(maps: torch.Tensor, storage_shape=STORAGE_SHAPE) -> torch.Tensor:
    """[P,H,W] fp32 (device) -> [P+1,SH,SW] fp32 (device): nearest-exact resize + mean row (generate_relevancy.py:95-111)."""
    assert maps.is_cuda and maps.dtype == torch.float32 and maps.dim() == 3
    maps = maps.contiguous()
    P, H, W = maps.shape
    SH, SW = storage_shape
    out = torch.empty(P + 1, SH, SW, device=maps.device)
    check(lib().semabs_relevancy_store_pack(ptr(maps), i32(P), i32(H), i32(W), i32(SH), i32(SW), ptr(out), stream_ptr()))
    return out


def pack_text_features(feats: torch.Tensor) -> torch.Tensor:
    """[P,E] -> [P+1,E]: append the mean feature, L2-normalise every row (generate_relevancy.py:112-121); 17 x 768 numbers:
    plain torch ops, not a kernel."""
    f = torch.cat([feats, feats.mean(dim=0, keepdim=True)], dim=0)
    return f / f.norm(dim=-1, keepdim=True)


def unpack_maps(stored: torch.Tensor, rows: Sequence[int], image_shape, mean_row: Optional[int] = None, gain: float = 1.0):
    """stored [N,SH,SW] fp32 (device) -> [K,H,W] fp32 (device): (stored[rows] - stored[mean_row]) bilinearly up-sampled to
    `image_shape`, times `gain` (dataset.py:817-834,866-871; gain = 50 at dataset.py:1049-1054)."""
    assert stored.is_cuda and stored.dtype == torch.float32 and stored.dim() == 3
    stored = stored.contiguous()
    H, W = int(image_shape[0]), int(image_shape[1])
    idx = torch.as_tensor(list(rows), dtype=torch.int32, device=stored.device)
    out = torch.empty(len(idx), H, W, device=stored.device)
    check(lib().semabs_relevancy_store_unpack(ptr(stored), ptr(idx), i32(len(idx)), i32(-1 if mean_row is None else mean_row),
                                              i32(stored.shape[1]), i32(stored.shape[2]), i32(H), i32(W), f32(gain), ptr(out),
                                              stream_ptr()))
    return out


class RelevancyStore:
    """One scene file.  Keys mirror the reference's HDF5 layout:
        saliencies                                   float32 [N, 128, 128]
        data/saliencies/<rgb>|<config>               row indices into `saliencies` (HDF5: region references)
        data/saliencies/<rgb>|<config>|saliency_text_labels          bytes [P+1]   (last = b"mean")
        data/saliencies/<rgb>|<config>|saliency_text_label_features  float32 [P+1, E]"""

    def __init__(self, path: str, storage_shape=STORAGE_SHAPE, backend: str = "npz"):
        self.path, self.storage_shape, self.backend = path, tuple(storage_shape), backend
        self.arrays: Dict[str, np.ndarray] = {"saliencies": np.zeros((0,) + self.storage_shape, np.float32)}
        if backend == "h5":
            import h5py  # noqa: F401  (absent in this image; the layout below is the reference's)
        if os.path.exists(path) and backend == "npz":
            with np.load(path, allow_pickle=False) as z:
                self.arrays = {k: z[k] for k in z.files}

    # ---- writer (generate_saliency_helper) ----
    def add(self, rgb_name: str, config_name: str, maps: torch.Tensor, text_labels: List[str], text_feats: torch.Tensor,
            replace: bool = False) -> np.ndarray:
        prefix = f"data/saliencies/{rgb_name}|{config_name}"
        if prefix in self.arrays and not replace:
            raise Exception(f"{prefix} already present")  # write_to_hdf5's behaviour (utils.py:300-304)
        packed = pack_maps(maps.to("cuda", torch.float32), self.storage_shape).cpu().numpy()
        base = self.arrays["saliencies"].shape[0]
        self.arrays["saliencies"] = np.concatenate([self.arrays["saliencies"], packed], axis=0)  # resize_and_add_data
        refs = np.arange(base, base + packed.shape[0], dtype=np.int64)
        self.arrays[prefix] = refs
        self.arrays[prefix + "|saliency_text_labels"] = np.array(list(text_labels) + ["mean"]).astype("S")
        self.arrays[prefix + "|saliency_text_label_features"] = pack_text_features(text_feats.float().cpu()).numpy()
        return refs

    def flush(self):
        if self.backend == "npz":
            tmp = self.path + ".tmp.npz"
            np.savez_compressed(tmp, **self.arrays)
            os.replace(tmp, self.path)
        else:
            import h5py

            with h5py.File(self.path, "a") as f:
                if "saliencies" in f:
                    del f["saliencies"]
                ds = f.create_dataset("saliencies", data=self.arrays["saliencies"], chunks=(1,) + self.storage_shape,
                                      compression="gzip", compression_opts=9, maxshape=(None,) + self.storage_shape)
                grp = f.require_group("data").require_group("saliencies")
                for k, v in self.arrays.items():
                    if not k.startswith("data/saliencies/"):
                        continue
                    name = k[len("data/saliencies/"):]
                    if name in grp:
                        del grp[name]
                    if "|saliency_text" in name:
                        grp.create_dataset(name, data=v, compression="gzip", compression_opts=9)
                    else:
                        grp.create_dataset(name, data=[ds.regionref[int(i), ...] for i in v], dtype=h5py.regionref_dtype)

    # ---- reader (SceneUnderstandDataset.load_patches) ----
    def load_patches(self, rgb_name: str, config_name: str, image_shape, labels: Optional[Sequence[str]] = None,
                     subtract_mean_relevancy: bool = True, gain: float = 1.0, device="cuda"):
        """-> dict(patch_labels [K] str, patch_saliencies [K,H,W] fp32 device, patch_label_features [K,E])."""
        prefix = f"data/saliencies/{rgb_name}|{config_name}"
        all_labels = self.arrays[prefix + "|saliency_text_labels"].astype(str).tolist()
        refs = self.arrays[prefix]
        sel = [i for i, l in enumerate(all_labels) if l != "mean"] if labels is None else sorted(all_labels.index(l) for l in labels)
        mean_row = int(refs[all_labels.index("mean")]) if subtract_mean_relevancy else None
        stored = torch.from_numpy(self.arrays["saliencies"]).to(device)
        maps = unpack_maps(stored, [int(refs[i]) for i in sel], image_shape, mean_row, gain)
        feats = torch.from_numpy(self.arrays[prefix + "|saliency_text_label_features"]).float()[sel]
        return {"patch_labels": np.array(all_labels)[sel], "patch_saliencies": maps, "patch_label_features": feats,
                "num_patches": len(sel)}
