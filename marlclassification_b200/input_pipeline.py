"""Host -> device input pipeline.

The reference feeds the episode from a ``DataLoader`` whose workers run PIL decode and
torchvision's ``ToTensor`` on the host (registry.py:56-57, train.py:91-107) and then moves the
fp32 batch with ``x.to(device)`` inside the loop (trainer.py:66-68): the copy of batch i+1 only
starts when step i has finished, and fp32 pixels are 4x the bytes of the decoded uint8 image.

Here a :class:`DevicePrefetcher` wraps the same iterable:

* batches are moved on a dedicated copy stream from pinned memory, one batch ahead of the
  compute stream (two-slot ring, stream-ordered with events: no host synchronisation);
* ``uint8`` batches (``[B,H,W,C]`` as PIL yields them, or ``[B,C,H,W]``) travel as bytes and are
  converted on the device by ``marlc_images_u8_to_f32`` (bit-identical to ``ToTensor``);
  ``float32`` ``[B,C,H,W]`` batches - what the reference's DataLoader yields - are accepted as they
  are, so the class is a drop-in around an unmodified loader.

``Trainer.train_epoch`` uses it for every loader; ``Trainer.train_step`` accepts the
:class:`StagedBatch` objects it yields.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Iterator, Optional, Tuple

import torch as th

from . import _lib


def images_u8_to_f32(src: th.Tensor, out: Optional[th.Tensor] = None, *, hwc: bool = True) -> th.Tensor:
    """``ToTensor`` for a whole uint8 batch on the device: u8[B,H,W,C] (``hwc``) or u8[B,C,H,W]
    -> f32[B,C,H,W] in [0,1].  Runs on the current stream; ``out`` may be given (e.g. the static
    input buffer of a captured step)."""
    src = _lib.require_cuda(src, "images", th.uint8)
    if src.dim() != 4:
        raise RuntimeError(f"images: expected a 4-D uint8 batch, got shape {tuple(src.shape)}")
    if hwc:
        b, h, w, c = src.shape
    else:
        b, c, h, w = src.shape
    if out is None:
        out = th.empty(b, c, h, w, dtype=th.float32, device=src.device)
    else:
        _lib.require_cuda(out, "out", th.float32)
        if tuple(out.shape) != (b, c, h, w) or not out.is_contiguous():
            raise RuntimeError(f"out: expected contiguous f32[{b},{c},{h},{w}], got {tuple(out.shape)}")
    _lib.check(_lib.lib().marlc_images_u8_to_f32(src.data_ptr(), out.data_ptr(), b, c, h, w, 1 if hwc else 0,
                                                 _lib.stream_ptr(src.device)))
    return out


@dataclass
class StagedBatch:
    """A batch whose host->device copy has been issued on the copy stream."""
    raw: th.Tensor            # device: u8[B,H,W,C] / u8[B,C,H,W] / f32[B,C,H,W]
    y: Optional[th.Tensor]    # device int64[B] (None for unlabeled batches)
    hwc: bool
    ready: th.cuda.Event      # recorded on the copy stream after both copies
    consumed: th.cuda.Event   # recorded by deliver(): the slot may be overwritten after it
    h2d_bytes: int

    @property
    def image_shape(self) -> Tuple[int, int, int, int]:
        """(B, C, H, W) of the fp32 batch this will become."""
        if self.raw.dtype == th.uint8 and self.hwc:
            b, h, w, c = self.raw.shape
            return b, c, h, w
        return tuple(self.raw.shape)  # type: ignore[return-value]

    @property
    def shape(self) -> Tuple[int, int, int, int]:
        return self.image_shape

    def deliver(self, img_out: Optional[th.Tensor] = None, y_out: Optional[th.Tensor] = None):
        """Make the batch available to the CURRENT stream as f32[B,C,H,W] (+ labels), writing into
        ``img_out`` / ``y_out`` when given.  Stream-ordered; never blocks the host."""
        s = th.cuda.current_stream(self.raw.device)
        s.wait_event(self.ready)
        if self.raw.dtype == th.uint8:
            img = images_u8_to_f32(self.raw, img_out, hwc=self.hwc)
        elif img_out is not None:
            img = img_out.copy_(self.raw, non_blocking=True)
        else:
            img = self.raw.clone()
        y = self.y
        if y is not None:
            y = y_out.copy_(y, non_blocking=True) if y_out is not None else y.clone()
        self.consumed.record(s)
        return img, y


class DevicePrefetcher:
    """Iterate over ``(images, labels)`` host batches, yielding :class:`StagedBatch` objects whose
    copy runs one batch ahead of the consumer.  ``hwc`` says how uint8 batches are laid out."""

    def __init__(self, batches: Iterable, device: th.device, *, hwc: bool = True, depth: int = 2) -> None:
        device = th.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"DevicePrefetcher: expected a CUDA device (no CPU path exists), got {device}")
        if depth < 2:
            raise ValueError("DevicePrefetcher needs at least two slots")
        self._src, self._dev, self._hwc, self._depth = batches, device, hwc, depth
        self._stream = th.cuda.Stream(device)
        self._slots: list = [None] * depth   # per slot: dict(pin_x, pin_y, dev_x, dev_y, copied, consumed)
        self._n = 0

    def __len__(self) -> int:
        return len(self._src)  # type: ignore[arg-type]

    # ---- one slot = pinned staging (for pageable inputs) + device buffers + two events
    def _slot_for(self, x: th.Tensor, y: Optional[th.Tensor]) -> dict:
        i = self._n % self._depth
        self._n += 1
        sl = self._slots[i]
        if sl is None or sl["dev_x"].shape != x.shape or sl["dev_x"].dtype != x.dtype:
            sl = {"dev_x": th.empty(x.shape, dtype=x.dtype, device=self._dev),
                  "dev_y": None, "pin_x": None, "pin_y": None,
                  "copied": th.cuda.Event(), "consumed": None}
            self._slots[i] = sl
        if y is not None and (sl["dev_y"] is None or sl["dev_y"].shape != y.shape):
            sl["dev_y"] = th.empty(y.shape, dtype=th.int64, device=self._dev)
        return sl

    @staticmethod
    def _pinned(sl: dict, key: str, t: th.Tensor) -> th.Tensor:
        """Pageable memory cannot be copied asynchronously: bounce through a pinned buffer that is
        reused once the previous copy out of it has completed."""
        if t.is_pinned():
            return t
        buf = sl[key]
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = th.empty(t.shape, dtype=t.dtype).pin_memory()
            sl[key] = buf
        else:
            sl["copied"].synchronize()
        buf.copy_(t)
        return buf

    def stage(self, x: th.Tensor, y: Optional[th.Tensor] = None) -> StagedBatch:
        if x.dtype not in (th.uint8, th.float32) or x.dim() != 4:
            raise RuntimeError(f"DevicePrefetcher: expected a uint8 or float32 4-D batch, got {x.dtype} {tuple(x.shape)}")
        if y is not None and y.dtype != th.int64:
            y = y.to(th.int64)
        x = x.contiguous()
        if x.is_cuda:  # already resident: nothing to move
            ev = th.cuda.Event()
            ev.record(th.cuda.current_stream(self._dev))
            return StagedBatch(x, None if y is None else y.to(self._dev), self._hwc, ev, th.cuda.Event(), 0)
        sl = self._slot_for(x, y)
        hx = self._pinned(sl, "pin_x", x)
        hy = None if y is None else self._pinned(sl, "pin_y", y.contiguous())
        with th.cuda.stream(self._stream):
            if sl["consumed"] is not None:
                self._stream.wait_event(sl["consumed"])  # the previous occupant has been delivered
            sl["dev_x"].copy_(hx, non_blocking=True)
            if hy is not None:
                sl["dev_y"].copy_(hy, non_blocking=True)
            sl["copied"].record(self._stream)
        sl["consumed"] = th.cuda.Event()
        nbytes = x.numel() * x.element_size() + (0 if y is None else y.numel() * 8)
        return StagedBatch(sl["dev_x"], None if y is None else sl["dev_y"], self._hwc, sl["copied"], sl["consumed"], nbytes)

    @staticmethod
    def _split(item):
        if isinstance(item, (tuple, list)):
            return item[0], (item[1] if len(item) > 1 else None)
        return item, None

    def __iter__(self) -> Iterator[StagedBatch]:
        it = iter(self._src)
        try:
            nxt = self.stage(*self._split(next(it)))
        except StopIteration:
            return
        for item in it:
            cur, nxt = nxt, self.stage(*self._split(item))  # batch i+1 is in flight while i is consumed
            yield cur
        yield nxt
