"""Inter-stage store of the hot path on disk: the per-view products of `render_reverse` in the
reference's own file layout, so a run can skip stage A exactly like the reference does when the folders exist.

Layout and dtypes follow
  * gaussctrl/gc_dataparser_ns.py:408-420  file names, 1-based: `depth_npy/frame_%05d.npy`, `z_0/frame_%05d.npy`,
    `mask_npy/frame_%05d.npy`, `unedited/frame_%05d.jpg` (a folder that is absent is simply not loaded);
  * gaussctrl/gc_render.py:217-221, 834-838  writer of depth: `np.save(path, outputs["depth"].cpu().numpy())` = [H,W,1] f32;
  * gaussctrl/gc_dataset.py:35-66           readers: depth `np.load(p)[:, :, 0][None]` -> [1,H,W]; z_0 and mask as stored
    ([1,4,h,w] f32 and [H,W]);
  * gaussctrl/gc_dataset.py:110-127,156-159 un-edited image: uint8 JPEG -> float32/255 torch tensor [H,W,3];
  * gaussctrl/gc_pipeline.py:268-274        the `train_data[i]` entries these files round-trip into.
Pure host I/O (numpy + PIL): nothing here touches the GPU."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

FOLDERS = {"depth_image": ("depth_npy", "npy"), "z_0_image": ("z_0", "npy"), "mask_image": ("mask_npy", "npy"),
           "unedited_image": ("unedited", "jpg")}


def frame_path(root: str, key: str, idx: int) -> str:
    """Path of view `idx` (0-based) for train_data key `key`; file names count from 1 (gc_dataparser_ns.py:410)."""
    folder, ext = FOLDERS[key]
    return os.path.join(root, folder, f"frame_{idx + 1:05d}.{ext}")


def available(root: str, load_mask: bool = True) -> List[str]:
    """Keys whose folder exists under `root` (gc_dataparser_ns.py:408-420; masks only with `load_mask`)."""
    keys = [k for k, (folder, _) in FOLDERS.items() if os.path.isdir(os.path.join(root, folder))]
    return [k for k in keys if k != "mask_image" or load_mask]


def save_view(root: str, idx: int, entry: Dict, jpeg_quality: int = 95) -> None:
    """Write one `train_data[idx]` entry (the keys it has) in the reference's layout."""
    if "depth_image" in entry:   # train_data holds [1,H,W]; the file holds the model output [H,W,1]
        d = np.asarray(entry["depth_image"], dtype=np.float32)
        _save_npy(frame_path(root, "depth_image", idx), np.ascontiguousarray(np.transpose(d, (1, 2, 0))))
    if "z_0_image" in entry:
        _save_npy(frame_path(root, "z_0_image", idx), np.asarray(entry["z_0_image"], dtype=np.float32))
    if entry.get("mask_image") is not None:
        _save_npy(frame_path(root, "mask_image", idx), np.asarray(entry["mask_image"]))
    if "unedited_image" in entry:
        from PIL import Image
        img = torch.as_tensor(entry["unedited_image"]).float().clamp(0, 1)
        u8 = (img * 255.0 + 0.5).to(torch.uint8).numpy()
        path = frame_path(root, "unedited_image", idx)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        Image.fromarray(u8).save(path, quality=jpeg_quality)


def _save_npy(path: str, arr: np.ndarray) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.save(path, arr)


def load_view(root: str, idx: int, keys: Optional[Sequence[str]] = None, load_mask: bool = True) -> Dict:
    """Read the stored products of view `idx` into `train_data` form (GCDataset.get_metadata, gc_dataset.py:129-162)."""
    out: Dict = {}
    for key in (available(root, load_mask) if keys is None else keys):
        path = frame_path(root, key, idx)
        if key == "depth_image":
            out[key] = np.load(path)[:, :, 0][None]                         # [1,H,W]
        elif key in ("z_0_image", "mask_image"):
            out[key] = np.load(path)
        else:
            from PIL import Image
            img = np.array(Image.open(path), dtype="uint8")
            if img.ndim == 2:
                img = img[:, :, None].repeat(3, axis=2)
            out[key] = torch.from_numpy(img.astype("float32") / 255.0)      # [H,W,3] 0..1
    return out


def save_train_data(root: str, train_data: Sequence[Dict]) -> None:
    for i, entry in enumerate(train_data):
        save_view(root, int(entry.get("image_idx", i)), entry)


def load_train_data(root: str, n_views: int, train_data: Optional[List[Dict]] = None, load_mask: bool = True) -> List[Dict]:
    """Fill (or create) a `train_data` list from the folders that exist under `root`."""
    td = train_data if train_data is not None else [{"image_idx": i} for i in range(n_views)]
    keys = available(root, load_mask)
    for i in range(n_views):
        td[i].update(load_view(root, int(td[i].get("image_idx", i)), keys))
    return td


def has_stage_a(root: str) -> bool:
    """True when `edit_images` can run without `render_reverse`: depth and z_0 folders are present."""
    have = available(root)
    return "depth_image" in have and "z_0_image" in have
