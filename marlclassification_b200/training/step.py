"""One optimisation step for a fixed episode geometry, replayed as CUDA graphs.

The step is   rollout -> loss phase A -> [all-reduce 3 doubles] -> loss phase B
-> BPTT -> [all-reduce flat grads] -> fused Adam.   Everything between the
collectives is one captured graph (the engine never allocates or synchronises),
so a 16-step episode costs a handful of host calls instead of ~10^4 launches.
"""
from __future__ import annotations

from typing import Optional

import torch as th

from ..engine import EpisodeEngine
from ..input_pipeline import StagedBatch
from ..parallel import DataParallelContext
from .optim import FlatAdam


class TrainStep:
    def __init__(self, engine: EpisodeEngine, optim: FlatAdam, dp: Optional[DataParallelContext] = None,
                 use_graph: bool = True) -> None:
        self.engine, self.optim = engine, optim
        self.dp = dp if dp is not None else DataParallelContext()
        self.use_graph = use_graph
        dev = engine.device
        self.static_img = th.zeros(engine.nb, engine.C, engine.H, engine.W, dtype=th.float32, device=dev)
        self.static_y = th.zeros(engine.nb, dtype=th.int64, device=dev)
        self._graphs: Optional[list] = None
        self._warm = 0

    # ---- the segments between collectives -----------------------------------------
    def _seg_forward(self) -> None:
        self.engine.forward(self.static_img)
        self.engine.loss_phase_a(self.static_y)

    def _seg_backward(self) -> None:
        self.engine.loss_phase_b()
        self.engine.backward(self.static_img)

    def _seg_update(self) -> None:
        self.optim.step(1.0 / self.dp.world_size)

    def _segments(self):
        if self.dp.enabled:
            return [self._seg_forward, self._seg_backward, self._seg_update]
        return [lambda: (self._seg_forward(), self._seg_backward(), self._seg_update())]

    def _capture(self) -> None:
        graphs = []
        for seg in self._segments():
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                seg()
            graphs.append(g)
        self._graphs = graphs

    @property
    def launches_per_step(self) -> int:
        l = self.engine.launches
        return l["forward"] + l["loss"] + l["backward"] + 2  # + Adam (2 kernels)

    def run_static(self) -> th.Tensor:
        """Run one step on whatever is in static_img / static_y."""
        eng, dp = self.engine, self.dp
        if self.use_graph and self._graphs is None and self._warm >= 2:
            self._capture()
        if self._graphs is not None:
            segs = [g.replay for g in self._graphs]
        else:
            segs = self._segments()
            self._warm += 1
        if dp.enabled:
            segs[0]()
            dp.all_reduce_stats(eng.loss_stats)
            segs[1]()
            dp.all_reduce_grads_sum(eng.model.flat_grads)
            segs[2]()
        else:
            segs[0]()
        return eng.loss_out

    def __call__(self, img: th.Tensor, y: th.Tensor, **inject) -> th.Tensor:
        """img / y may live on the host (pinned -> async H2D) or on the device, or be a
        StagedBatch from the DevicePrefetcher (copy already in flight on the copy stream)."""
        if isinstance(img, StagedBatch):
            img.deliver(self.static_img, self.static_y)
            return self.run_static()
        if inject.get("pos0") is not None or inject.get("hidden0") is not None or inject.get("actions") is not None:
            return self._run_injected(img, y, inject)
        self.static_img.copy_(img, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        return self.run_static()

    def _run_injected(self, img, y, inject) -> th.Tensor:
        """Parity path: explicit draws, eager (pointers differ per call)."""
        eng, dp = self.engine, self.dp
        self.static_img.copy_(img, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        eng.forward(self.static_img, inject.get("pos0"), inject.get("hidden0"), inject.get("actions"))
        eng.loss_phase_a(self.static_y)
        dp.all_reduce_stats(eng.loss_stats)
        eng.loss_phase_b()
        eng.backward(self.static_img)
        dp.all_reduce_grads_sum(eng.model.flat_grads)
        self._seg_update()
        return eng.loss_out


class EvalStep:
    """Forward-only rollout for a fixed geometry (trainer.py:164-196: ``no_grad`` episode, vote =
    mean over agents of the last step's predictions), replayed as one CUDA graph."""

    def __init__(self, engine: EpisodeEngine, use_graph: bool = True) -> None:
        self.engine, self.use_graph = engine, use_graph
        dev = engine.device
        self.static_img = th.zeros(engine.nb, engine.C, engine.H, engine.W, dtype=th.float32, device=dev)
        self.static_y = th.zeros(engine.nb, dtype=th.int64, device=dev)
        self.vote = th.zeros(engine.nb, engine.step_preds.shape[-1], dtype=th.float32, device=dev)
        self._graph: Optional[th.cuda.CUDAGraph] = None
        self._warm = 0

    def _body(self) -> None:
        self.engine.forward(self.static_img)
        th.mean(self.engine.step_preds[-1], dim=0, out=self.vote)  # prediction.mean(0), trainer.py:180

    def run_static(self) -> th.Tensor:
        if self.use_graph and self._graph is None and self._warm >= 2:
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g):
                self._body()
            self._graph = g
        if self._graph is not None:
            self._graph.replay()
        else:
            self._body()
            self._warm += 1
        return self.vote

    def __call__(self, img) -> th.Tensor:
        """Returns the [Nb, Nc] vote (a static buffer: clone it to keep it across calls)."""
        if isinstance(img, StagedBatch):
            img.deliver(self.static_img, self.static_y)
        else:
            self.static_img.copy_(img, non_blocking=True)
        return self.run_static()
