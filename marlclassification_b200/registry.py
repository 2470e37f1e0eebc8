"""Dataset name -> (dataset constructor, feature-extractor constructor) (reference:
registry.py:27-57).  The ``cnn_constructor`` half is on the hot path; the dataset half serves
the train / test CLI (image-folder datasets only, see data.py)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable

import torch as th

from .data import (  # noqa: F401  (default_image_pipeline is re-exported as in the reference)
    default_image_pipeline, folder_dataset_constructor, u8_image_pipeline, unsupported_dataset_constructor,
)
from .networks.vision import (
    AidCnn, KneeMriCnn, MnistCnn, Resisc45Cnn, SkinCancerCnn, VisionCnnModule, WorldStratCnn,
)


@dataclass(frozen=True)
class DatasetSpec:
    dataset_constructor: Callable[[str, Callable[[Any], th.Tensor]], Any]
    cnn_constructor: Callable[[int], VisionCnnModule]


DATASET_REGISTRY: dict[str, DatasetSpec] = {
    "mnist": DatasetSpec(folder_dataset_constructor("mnist"), MnistCnn),
    "resisc45": DatasetSpec(folder_dataset_constructor("resisc45"), Resisc45Cnn),
    "kneemri": DatasetSpec(unsupported_dataset_constructor("kneemri", "pickled 3-D volumes"), KneeMriCnn),
    "aid": DatasetSpec(folder_dataset_constructor("aid"), AidCnn),
    "worldstrat": DatasetSpec(unsupported_dataset_constructor("worldstrat", "CSV metadata + land-cover masks"),
                              WorldStratCnn),
    "skin_cancer": DatasetSpec(folder_dataset_constructor("skin_cancer"), SkinCancerCnn),
}


def get_dataset_spec(name: str) -> DatasetSpec:
    assert name in DATASET_REGISTRY, f'Unknown dataset "{name}", expected one of {sorted(DATASET_REGISTRY)}'
    return DATASET_REGISTRY[name]
