"""Dataset name -> (dataset constructor, feature-extractor constructor) (reference:
registry.py:27-57).  The ``cnn_constructor`` half is on the hot path; the dataset half serves
the train / test CLI (image-folder datasets only, see data.py)."""
from __future__ import annotations

from typing import Any, Callable, NamedTuple

from . import data
from .data import default_image_pipeline, u8_image_pipeline  # noqa: F401  (re-exported, registry.py:56-57)
from .networks import vision


class DatasetSpec(NamedTuple):
    dataset_constructor: Callable[[str, Callable[[Any], Any]], Any]   # (resources dir, image transform) -> dataset
    cnn_constructor: Callable[[int], vision.VisionCnnModule]          # window size f -> feature extractor


def _folder(name: str, cnn) -> DatasetSpec:
    return DatasetSpec(data.folder_dataset_constructor(name), cnn)


def _not_a_folder(name: str, why: str, cnn) -> DatasetSpec:
    return DatasetSpec(data.unsupported_dataset_constructor(name, why), cnn)


DATASET_REGISTRY: dict[str, DatasetSpec] = dict(
    mnist=_folder("mnist", vision.MnistCnn),
    resisc45=_folder("resisc45", vision.Resisc45Cnn),
    aid=_folder("aid", vision.AidCnn),
    skin_cancer=_folder("skin_cancer", vision.SkinCancerCnn),
    worldstrat=_not_a_folder("worldstrat", "CSV metadata + land-cover masks", vision.WorldStratCnn),
    kneemri=_not_a_folder("kneemri", "pickled 3-D volumes", vision.KneeMriCnn),
)


def get_dataset_spec(name: str) -> DatasetSpec:
    assert name in DATASET_REGISTRY, f'Unknown dataset "{name}", expected one of {sorted(DATASET_REGISTRY)}'
    return DATASET_REGISTRY[name]
