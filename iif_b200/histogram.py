"""Per-class frequencies and IIF weight vectors, computed on the GPU.

Mirrors the init-time numpy of the reference: `LT_Dataset.__init__` / `get_cls_num_list`
(classification/imbalanced_dataset.py:100-144) and `IIFLoss.__init__` (classification/custom.py:14-26);
for detection, the `img_freq` / `instance_freq` columns and the 14 weight columns of
`lvis_files/idf_1204.csv` (their generator is not in the reference repo).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

VARIANTS = ("raw", "smooth", "rel", "normit", "gombit", "base2", "base10")
CSV_NAME = {"rel": "prob"}  # the CSV tables call the `rel` variant `prob`


def class_counts(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """counts[c] = #{labels == c}  (int64 [C], bit-exact; shared-memory integer atomics)."""
    return ops.hist_labels(labels, num_classes)


def lt_class_map(counts: torch.Tensor) -> np.ndarray:
    """Descending-frequency class re-index (imbalanced_dataset.py:115-120).  Kept on the host with
    the reference's own `np.argsort(-counts)` call so tie order is identical."""
    c = counts.detach().cpu().numpy()
    order = np.argsort(-c)
    cmap = np.zeros(len(c), dtype=np.int64)
    cmap[order] = np.arange(len(c))
    return cmap


def image_instance_freq(image_ids: torch.Tensor, categories: torch.Tensor, num_images: int, num_classes: int):
    """(img_freq, instance_freq) int64 [C] from per-annotation (image, category) pairs."""
    return ops.hist_images_dedup(image_ids, categories, num_images, num_classes)


def iif_weights(counts: torch.Tensor, variant: str, *, total: int = 0, iif_norm: float = 0.0) -> torch.Tensor:
    """One variant as an fp32 [1,C] row (float64 arithmetic, one rounding), optional p-norm."""
    return ops.weights_from_counts(counts, variant, total=total, norm_p=iif_norm).unsqueeze(0)


def iif_weight_dict(counts: torch.Tensor, *, total: int = 0, iif_norm: float = 0.0) -> dict:
    """All seven variants: the `.iif` dict of the reference criterion (custom.py:15-26)."""
    return {v: iif_weights(counts, v, total=total, iif_norm=iif_norm) for v in VARIANTS}


def detection_weight_table(img_freq: torch.Tensor, instance_freq: torch.Tensor, num_images: int) -> dict:
    """The 14 weight columns of idf_1204.csv from the two frequency columns (N_img = #images,
    N_obj = sum(instance_freq)); keys use the CSV column names."""
    out = {}
    for v in VARIANTS:
        name = CSV_NAME.get(v, v)
        out[name] = ops.weights_from_counts(img_freq, v, total=int(num_images))
        out[name + "_obj"] = ops.weights_from_counts(instance_freq, v, total=0)
    return out
