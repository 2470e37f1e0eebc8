"""Dataset name -> feature-extractor constructor (reference: registry.py; only the
``cnn_constructor`` half touches the hot path -- datasets are out of scope here)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

from .networks.vision import (
    AidCnn, KneeMriCnn, MnistCnn, Resisc45Cnn, SkinCancerCnn, VisionCnnModule, WorldStratCnn,
)


@dataclass(frozen=True)
class DatasetSpec:
    cnn_constructor: Callable[[int], VisionCnnModule]


DATASET_REGISTRY: dict[str, DatasetSpec] = {
    "mnist": DatasetSpec(MnistCnn),
    "resisc45": DatasetSpec(Resisc45Cnn),
    "kneemri": DatasetSpec(KneeMriCnn),
    "aid": DatasetSpec(AidCnn),
    "worldstrat": DatasetSpec(WorldStratCnn),
    "skin_cancer": DatasetSpec(SkinCancerCnn),
}


def get_dataset_spec(name: str) -> DatasetSpec:
    assert name in DATASET_REGISTRY, f'Unknown dataset "{name}", expected one of {sorted(DATASET_REGISTRY)}'
    return DATASET_REGISTRY[name]
