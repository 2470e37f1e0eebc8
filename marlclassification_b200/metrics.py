"""Sliding-window meters used by the Trainer (reference: metrics.py).  Host-side
bookkeeping only; the confusion-matrix picture is drawn with PIL (no matplotlib needed)."""
from __future__ import annotations

from typing import Optional

import torch as th


def format_metric(metric: th.Tensor, class_map: dict) -> str:
    idx_to_class = {v: k for k, v in class_map.items()}
    return ", ".join(f'"{idx_to_class[i]}" : {metric[i].item() * 100.0:.1f}%' for i in range(metric.size(0)))


class ConfusionMeter:
    """Keeps (argmax prediction, target) pairs of the last ``window_size`` batches
    and derives per-class precision / recall (metrics.py:25-108)."""

    def __init__(self, nb_class: int, window_size: Optional[int] = None) -> None:
        self.__nb_class = nb_class
        self.__window_size = window_size
        self.__results: list[tuple[th.Tensor, th.Tensor]] = []

    def add(self, y_proba: th.Tensor, y_true: th.Tensor) -> None:
        self.__results.append((y_proba.argmax(dim=1).detach(), y_true.detach()))
        if self.__window_size is not None and len(self.__results) > self.__window_size:
            self.__results.pop(0)

    def conf_mat(self) -> th.Tensor:
        pred = th.cat([p for p, _ in self.__results])
        true = th.cat([t for _, t in self.__results])
        flat = true * self.__nb_class + pred
        return th.bincount(flat, minlength=self.__nb_class**2).view(self.__nb_class, self.__nb_class)

    def precision(self) -> th.Tensor:
        cm = self.conf_mat().to(th.float)
        tot = cm.sum(dim=0)
        return th.where(tot != 0, cm.diagonal() / tot.clamp(min=1.0), th.zeros_like(tot))

    def recall(self) -> th.Tensor:
        cm = self.conf_mat().to(th.float)
        tot = cm.sum(dim=1)
        return th.where(tot != 0, cm.diagonal() / tot.clamp(min=1.0), th.zeros_like(tot))

    def save_conf_matrix(self, epoch: int, output_dir: str, stage: str) -> None:
        """Row-normalised confusion matrix as ``confusion_matrix_epoch_{e}_{stage}.png``
        (metrics.py:110-129).  Drawn with PIL: matplotlib is not a dependency of this build."""
        import os

        from .visualization import heatmap_image

        cm = self.conf_mat().to(th.float).cpu()
        norm = cm / cm.sum(dim=1, keepdim=True).clamp(min=1.0)
        img = heatmap_image(norm, title=f"confusion matrix epoch {epoch} - {stage}",
                            xlabel="Predicated Label", ylabel="True Label")
        img.save(os.path.join(output_dir, f"confusion_matrix_epoch_{epoch}_{stage}.png"))


class LossMeter:
    def __init__(self, window_size: Optional[int] = None) -> None:
        self.__window_size = window_size
        self.__values: list[float] = []

    def add(self, value: float) -> None:
        self.__values.append(value)
        if self.__window_size is not None and len(self.__values) > self.__window_size:
            self.__values.pop(0)

    def loss(self) -> float:
        return sum(self.__values) / max(1, len(self.__values))
