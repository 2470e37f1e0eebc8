"""Feature extractors (reference: networks/vision.py).

The modules own the parameters (reference-compatible ``state_dict`` keys, e.g.
``_Generic2dCnnModule__layers.0.weight``) and describe their structure to the
CUDA engine through ``cnn_spec``; the arithmetic runs in csrc/cnn.cu.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch as th
from torch import nn


class VisionCnnModule(nn.Module, ABC):
    """Injection seam of the reference (vision.py:8-15)."""

    @property
    @abstractmethod
    def out_size(self) -> int: ...

    @abstractmethod
    def forward(self, o_t: th.Tensor) -> th.Tensor: ...


class _Generic2dCnnModule(VisionCnnModule):
    """k x [Conv2d(3x3, s2, p1) -> GroupNorm -> SiLU] -> Flatten (vision.py:23-53)."""

    first_channel_only = False

    def __init__(self, f: int, layers: list[tuple[int, int]], group_norm_nums: list[int]) -> None:
        super().__init__()
        blocks: list[nn.Module] = []
        side = f
        for (c_in, c_out), groups in zip(layers, group_norm_nums):
            blocks += [nn.Conv2d(c_in, c_out, 3, 2, 1), nn.GroupNorm(groups, c_out), nn.SiLU()]
            side = (side - 1) // 2 + 1
        blocks.append(nn.Flatten(1, -1))
        self.__layers = nn.Sequential(*blocks)
        self.__out_size = layers[-1][1] * side * side
        self.__spec = (f, [tuple(l) for l in layers], list(group_norm_nums))

    @property
    def out_size(self) -> int:
        return self.__out_size

    @property
    def cnn_spec(self) -> tuple[int, list[tuple[int, int]], list[int], bool]:
        """(window, [(c_in, c_out)], [groups], first_channel_only) for the engine."""
        f, layers, groups = self.__spec
        return f, layers, groups, self.first_channel_only

    def forward(self, o_t: th.Tensor) -> th.Tensor:
        """[N, C, f, f] -> [N, out_size] through the fused CUDA CNN kernel."""
        from ..engine import cnn_forward

        return cnn_forward(self, o_t)


class MnistCnn(_Generic2dCnnModule):
    first_channel_only = True  # grey scale: reads channel 0 (vision.py:64)

    def __init__(self, f: int) -> None:
        super().__init__(f, [(1, 8), (8, 16)], [2, 4])


class Resisc45Cnn(_Generic2dCnnModule):
    def __init__(self, f: int) -> None:
        super().__init__(f, [(3, 16), (16, 32), (32, 64)], [2, 4, 8])


class AidCnn(_Generic2dCnnModule):
    def __init__(self, f: int) -> None:
        super().__init__(f, [(3, 16), (16, 32), (32, 64), (64, 128)], [2, 4, 8, 16])


class WorldStratCnn(_Generic2dCnnModule):
    def __init__(self, f: int) -> None:
        super().__init__(f, [(3, 16), (16, 32), (32, 64), (64, 128), (128, 256)], [2, 4, 8, 16, 32])


class SkinCancerCnn(_Generic2dCnnModule):
    def __init__(self, f: int) -> None:
        super().__init__(f, [(3, 16), (16, 32), (32, 64)], [2, 4, 8])


class KneeMriCnn(VisionCnnModule):
    """3-D volumes (Conv3d + BatchNorm3d) are outside the accelerated hot path
    (SURVEY section 2 row 5, section 8b): refuse loudly instead of mis-computing."""

    def __init__(self, f: int = 16) -> None:  # pragma: no cover - not supported
        super().__init__()
        raise NotImplementedError("KneeMriCnn (3-D, BatchNorm) is out of scope of the B200 hot path")

    @property
    def out_size(self) -> int:  # pragma: no cover
        raise NotImplementedError

    def forward(self, o_t: th.Tensor) -> th.Tensor:  # pragma: no cover
        raise NotImplementedError
