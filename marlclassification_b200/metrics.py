"""Sliding-window meters used by the Trainer (reference: metrics.py).  Host-side
bookkeeping only; the confusion-matrix picture is drawn with PIL (no matplotlib needed)."""
from __future__ import annotations

from collections import deque
from typing import Optional

import torch as th


def format_metric(metric: th.Tensor, class_map: dict) -> str:
    idx_to_class = {v: k for k, v in class_map.items()}
    return ", ".join(f'"{idx_to_class[i]}" : {metric[i].item() * 100.0:.1f}%' for i in range(metric.size(0)))


class ConfusionMeter:
    """Confusion matrix over the last ``window_size`` batches (all batches when ``None``) and the
    per-class precision / recall derived from it (metrics.py:25-108).

    The reference keeps the (prediction, target) pairs and rebuilds the matrix from all of them on
    every query (``cat`` over the window + ``bincount``, twice per training iteration).  Here the
    matrix is kept as a running count on the tensors' device: a batch adds its ``bincount``, the
    batch that leaves the window subtracts its own, so a query costs nothing that grows with the
    window.  Same numbers, same device placement, no host synchronisation."""

    def __init__(self, nb_class: int, window_size: Optional[int] = None) -> None:
        self.__nb_class = nb_class
        self.__window_size = window_size
        self.__cells: deque[th.Tensor] = deque()   # per batch: flat cell index true * Nc + pred
        self.__counts: Optional[th.Tensor] = None  # int64[Nc * Nc], lives where the batches live

    def __bincount(self, cells: th.Tensor) -> th.Tensor:
        out = th.bincount(cells, minlength=self.__nb_class**2)
        if out.numel() != self.__nb_class**2:
            raise RuntimeError(f"ConfusionMeter: a label is outside [0, {self.__nb_class}) "
                               "(nb_class smaller than the dataset's number of classes?)")
        return out

    def add(self, y_proba: th.Tensor, y_true: th.Tensor) -> None:
        cells = (y_true.detach() * self.__nb_class + y_proba.detach().argmax(dim=1)).to(th.int64)
        if self.__counts is None or self.__counts.device != cells.device:
            self.__counts = th.zeros(self.__nb_class**2, dtype=th.int64, device=cells.device)
            for old in self.__cells:
                self.__counts += self.__bincount(old.to(cells.device))
        self.__counts += self.__bincount(cells)
        self.__cells.append(cells)
        if self.__window_size is not None and len(self.__cells) > self.__window_size:
            self.__counts -= self.__bincount(self.__cells.popleft().to(cells.device))

    def conf_mat(self) -> th.Tensor:
        """int64[Nc, Nc], rows = true class, columns = predicted class."""
        if self.__counts is None:
            return th.zeros(self.__nb_class, self.__nb_class, dtype=th.int64)
        return self.__counts.view(self.__nb_class, self.__nb_class).clone()

    @staticmethod
    def __ratio(hits: th.Tensor, totals: th.Tensor) -> th.Tensor:
        """hits / totals where totals != 0, else 0 (metrics.py:71-108)."""
        return th.where(totals != 0, hits / totals.clamp(min=1.0), th.zeros_like(totals))

    def precision(self) -> th.Tensor:
        cm = self.conf_mat().to(th.float)
        return self.__ratio(cm.diagonal(), cm.sum(dim=0))

    def recall(self) -> th.Tensor:
        cm = self.conf_mat().to(th.float)
        return self.__ratio(cm.diagonal(), cm.sum(dim=1))

    def mean_precision_recall(self) -> th.Tensor:
        """f32[2] = (mean precision, mean recall) on the matrix's device: what the training loop logs,
        computed from ONE matrix and ready to be packed into a single device->host read."""
        cm = self.conf_mat().to(th.float)
        diag = cm.diagonal()
        return th.stack((self.__ratio(diag, cm.sum(dim=0)).mean(), self.__ratio(diag, cm.sum(dim=1)).mean()))

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
