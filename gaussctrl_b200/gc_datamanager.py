"""GaussCtrlDataManagerConfig / GaussCtrlDataManager: host-side mirror of gaussctrl/gc_datamanager.py:54-111, 213-235.

Only what the editing hot path and its caller (the fine-tune loop) touch is implemented here: the config fields with the
reference's defaults, the view sub-sampling that decides WHICH views are edited (`cameras`, `train_data`,
`train_unseen_cameras`), and `next_train`.  Image caching / undistortion stays nerfstudio's FullImageDatamanager (the
reference copies that method verbatim from nerfstudio, gc_datamanager.py:113-188); it is I/O outside the path."""
from __future__ import annotations

import random
from copy import deepcopy
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple, Type

from ._compat import FullImageDatamanager, FullImageDatamanagerConfig


@dataclass
class GaussCtrlDataManagerConfig(FullImageDatamanagerConfig):
    """gaussctrl/gc_datamanager.py:54-66 (same field names and defaults)."""
    _target: Type = field(default_factory=lambda: GaussCtrlDataManager)
    patch_size: int = 32
    """Size of patch to sample from. If >1, patch-based sampling will be used."""
    subset_num: int = 4
    """The scene's views are split into this many contiguous subsets before sampling."""
    sampled_views_every_subset: int = 10
    """Views sampled (without replacement) from every subset: 4 x 10 = 40 edited views by default."""
    load_all: bool = False
    """Edit every image of the dataset instead of the sampled subset."""


def sample_view_subset(view_num: int, subset_num: int, per_subset: int, rng=random) -> List[int]:
    """gc_datamanager.py:95-103: anchors every `view_num // subset_num` views, only the first FOUR anchors are used
    regardless of `subset_num` (SURVEY §8a gotcha 5), `per_subset` sorted samples from every bucket.  The reference
    draws with the unseeded global `random`; pass a seeded `random.Random` for reproducible runs."""
    anchors = list(range(0, view_num, view_num // subset_num))[:4] + [view_num]
    picked: List[int] = []
    for cur, nxt in zip(anchors[:-1], anchors[1:]):
        picked += sorted(rng.sample(list(range(cur, nxt)), per_subset))
    return picked


class GaussCtrlDataManager(FullImageDatamanager):
    """gc_datamanager.py:69-111: after nerfstudio's FullImageDatamanager has cached the training images, choose the
    views to edit and expose them as `cameras` / `train_data` (a list of dicts the pipeline fills with
    `unedited_image`, `depth_image`, `z_0_image`, `mask_image`, `image`)."""

    config: GaussCtrlDataManagerConfig

    def __init__(self, config: GaussCtrlDataManagerConfig, device="cpu", test_mode="val", world_size: int = 1,
                 local_rank: int = 0, **kwargs):
        super().__init__(config, device, test_mode, world_size, local_rank)
        self.sample_idx: List[int] = []
        self.step_every = 1
        self.edited_image_dict: Dict = {}
        self._select_views()

    def _num_source_views(self) -> int:
        ds = self.train_dataset
        dpo = getattr(ds, "_dataparser_outputs", None)
        return len(dpo.image_filenames) if dpo is not None else len(self.cached_train)

    def _uses_all_views(self) -> bool:
        c = self.config
        return self._num_source_views() <= c.subset_num * c.sampled_views_every_subset or c.load_all

    def _select_views(self) -> None:
        c = self.config
        if self._uses_all_views():
            self.cameras = self.train_dataset.cameras if self.train_dataset is not None else []
            self.train_data = self.cached_train
            self.train_unseen_cameras = list(range(len(self.train_data)))
            return
        sampled = sample_view_subset(self._num_source_views(), c.subset_num, c.sampled_views_every_subset)
        self.sample_idx = sampled
        self.cameras = [self.train_dataset.cameras[i:i + 1] for i in sampled]
        self.train_data = []
        for i, src in enumerate(sampled):
            data = self.cached_train[src]
            data["image_idx"] = i
            self.train_data.append(data)
        self.train_unseen_cameras = list(range(c.subset_num * c.sampled_views_every_subset))

    def next_train(self, step: int) -> Tuple[object, Dict]:
        """gc_datamanager.py:213-235: a random not-yet-seen view (re-filled when exhausted) and a copy of its entry."""
        image_idx = self.train_unseen_cameras.pop(random.randint(0, len(self.train_unseen_cameras) - 1))
        if len(self.train_unseen_cameras) == 0:
            self.train_unseen_cameras = list(range(len(self.train_data)))
        data = deepcopy(self.train_data[image_idx])
        data["image"] = data["image"].to(self.device)
        if self._uses_all_views():
            camera = self.cameras[image_idx:image_idx + 1].to(self.device)
        else:
            camera = self.cameras[image_idx:image_idx + 1][0].to(self.device)
        if getattr(camera, "metadata", None) is None:
            camera.metadata = {}
        camera.metadata["cam_idx"] = image_idx
        return camera, data
