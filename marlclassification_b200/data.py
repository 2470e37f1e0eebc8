"""Image-folder datasets feeding the episode (reference: data/datasets.py:16-77, 98-116 and
registry.py:56-57).

The reference's 2-D datasets are all "one directory per class under
``<res>/downloaded/<name>``" read with PIL and converted by torchvision ``ToTensor`` in the
DataLoader workers, so fp32 pixels cross PCIe.  Here one :class:`FolderDataset` covers them and
offers two item formats:

* ``default_image_pipeline()`` - the reference's: f32[C,H,W] in [0,1] on the host;
* ``u8_image_pipeline()`` - the decoded bytes u8[H,W,C]; ``Trainer.prefetch`` copies them as
  bytes (4x less H2D traffic) and ``marlc_images_u8_to_f32`` runs ``ToTensor`` on the device,
  bit-identical to the host conversion (tests/test_gpu_input.py).

Datasets that are not image folders (WorldStrat CSV + masks, KneeMRI pickled 3-D volumes,
Kinetics video) are outside the 2-D hot path and raise.
"""
from __future__ import annotations

import os
from os.path import isdir, join
from typing import Any, Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch as th
from PIL import Image
from torch.utils.data import Dataset

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def pil_rgb_loader(path: str) -> Image.Image:
    """Every image becomes 3-channel RGB, as in datasets.py:16-21 (MNIST PNGs included; MnistCnn
    then reads channel 0 only, vision.py:64)."""
    with open(path, "rb") as fh:
        return Image.open(fh).convert("RGB")


def to_u8_hwc(img: Image.Image) -> th.Tensor:
    """PIL image -> u8[H,W,C]: the decoded bytes, no arithmetic."""
    arr = np.array(img, dtype=np.uint8)  # a writable copy (torch refuses read-only buffers)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    return th.from_numpy(arr)


def to_f32_chw(img: Image.Image) -> th.Tensor:
    """torchvision ``ToTensor`` restated (registry.py:56-57): u8 HWC -> f32 CHW, x / 255."""
    return to_u8_hwc(img).permute(2, 0, 1).contiguous().to(th.float32).div(255)


def default_image_pipeline() -> Callable[[Image.Image], th.Tensor]:
    return to_f32_chw


def u8_image_pipeline() -> Callable[[Image.Image], th.Tensor]:
    return to_u8_hwc


class FolderDataset(Dataset):
    """``root/<class name>/<image>``; classes sorted by name -> ``class_to_idx`` (what
    torchvision ``ImageFolder`` yields for the reference, so ``class_to_idx.json`` files agree)."""

    def __init__(self, root: str, transform: Callable[[Any], th.Tensor],
                 loader: Callable[[str], Image.Image] = pil_rgb_loader) -> None:
        assert os.path.exists(root) and isdir(root), f"{root} does not exist or is not a directory"
        classes = sorted(e.name for e in os.scandir(root) if e.is_dir())
        if not classes:
            raise FileNotFoundError(f"Couldn't find any class folder in {root}.")
        self.root = root
        self.classes: List[str] = classes
        self.class_to_idx: Dict[str, int] = {c: i for i, c in enumerate(classes)}
        self.samples: List[Tuple[str, int]] = []
        for c in classes:
            for dirpath, _, files in sorted(os.walk(join(root, c), followlinks=True)):
                for fn in sorted(files):
                    if fn.lower().endswith(IMG_EXTENSIONS):
                        self.samples.append((join(dirpath, fn), self.class_to_idx[c]))
        if not self.samples:
            raise FileNotFoundError(f"Found no image file under {root} (extensions {IMG_EXTENSIONS}).")
        self.targets = [t for _, t in self.samples]
        self.transform = transform
        self.loader = loader

    def __len__(self) -> int:
        return len(self.samples)

    def __getitem__(self, index: int) -> Tuple[th.Tensor, int]:
        path, target = self.samples[index]
        return self.transform(self.loader(path)), target


# sub-directory of ``<res>/downloaded`` per dataset (datasets.py:28, 48, 68, 104)
FOLDER_DATASETS: Dict[str, Sequence[str]] = {
    "mnist": ("mnist_png", "all_png"),
    "resisc45": ("NWPU-RESISC45",),
    "aid": ("AID",),
    "skin_cancer": ("skin_cancer",),
}


def folder_dataset_constructor(name: str) -> Callable[[str, Callable[[Any], th.Tensor]], FolderDataset]:
    def make(res_path: str, img_transform: Callable[[Any], th.Tensor]) -> FolderDataset:
        return FolderDataset(join(res_path, "downloaded", *FOLDER_DATASETS[name]), img_transform)

    make.__name__ = f"{name}_dataset"
    return make


def unsupported_dataset_constructor(name: str, why: str):
    def make(res_path: str, img_transform: Callable[[Any], th.Tensor]):
        raise NotImplementedError(f'dataset "{name}" is outside the 2-D image-folder path of this build: {why}')

    return make


class ShardedBatchSampler:
    """Batch sampler for data-parallel runs: every rank draws the SAME global
    permutation (seed + epoch) and cuts it into global batches of ``batch_size``; rank r decodes
    only images ``[r*B/G, (r+1)*B/G)`` of each.  Shards are always equal (the gradient average
    over ranks equals the global-batch mean, trainer.py:111): a ragged last batch is trimmed to a
    multiple of the world size.  With world_size 1 it is a plain shuffled batch sampler with
    ``drop_last=False``, like the reference's loaders (train.py:91-107)."""

    def __init__(self, length: int, batch_size: int, rank: int = 0, world_size: int = 1, shuffle: bool = True,
                 seed: int = 0) -> None:
        if batch_size % world_size != 0:
            raise ValueError(f"batch size {batch_size} is not divisible by world size {world_size}")
        if not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside [0, {world_size})")
        self.length, self.batch_size, self.rank, self.world_size = length, batch_size, rank, world_size
        self.shuffle, self.seed, self.epoch = shuffle, seed, 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def _global_batches(self) -> List[List[int]]:
        if self.shuffle:
            g = th.Generator().manual_seed(self.seed * 1_000_003 + self.epoch)
            order = th.randperm(self.length, generator=g).tolist()
        else:
            order = list(range(self.length))
        out = []
        for i in range(0, self.length, self.batch_size):
            chunk = order[i:i + self.batch_size]
            chunk = chunk[:len(chunk) - len(chunk) % self.world_size]
            if chunk:
                out.append(chunk)
        return out

    def __iter__(self):
        for chunk in self._global_batches():
            per = len(chunk) // self.world_size
            yield chunk[self.rank * per:(self.rank + 1) * per]

    def __len__(self) -> int:
        return len(self._global_batches())


def collate_images(items: Sequence[Tuple[th.Tensor, int]]) -> Tuple[th.Tensor, th.Tensor]:
    """Stack equally-sized images (u8 HWC or f32 CHW) and their labels (int64)."""
    return th.stack([x for x, _ in items]), th.tensor([int(y) for _, y in items], dtype=th.int64)


class ResidentImages:
    """A whole image-folder split decoded ONCE and kept as ``u8[N,H,W,C]`` on the device.

    The reference re-decodes every file with PIL in DataLoader workers each epoch and ships fp32
    pixels over PCIe each step (train.py:91-107).  At B200 rates (thousands of image-episodes per
    second per GPU) the decode workers become the limit long before the GPU does, while the
    datasets themselves are small next to 180 GB of HBM (RESISC45: 31 500 x 256 x 256 x 3 bytes =
    6.2 GB; AID: 10 000 x 600 x 600 x 3 = 10.8 GB; MNIST: 0.16 GB).  So: decode once (thread pool -
    PIL releases the GIL while decoding), one H2D copy of the bytes, and every batch afterwards is
    an ``index_select`` in HBM followed by the device ``ToTensor`` kernel - no host work and no
    PCIe traffic per step.  Iterating yields ``(u8[B,H,W,C], int64[B])`` device tensors, which
    ``Trainer.prefetch`` passes through untouched.

    All images must share one size (true for the reference's folder datasets).  Under data
    parallelism every rank keeps the whole split (any image can land in any rank's shard after a
    reshuffle); the sampler still hands each rank only its slice of every global batch."""

    def __init__(self, dataset, indices: Sequence[int], batch_sampler: ShardedBatchSampler, device,
                 decode_threads: int = 8, hbm_fraction: float = 0.5) -> None:
        from concurrent.futures import ThreadPoolExecutor

        self.batch_sampler = batch_sampler
        self.device = th.device(device)
        if batch_sampler.length != len(indices):
            raise ValueError(f"batch sampler covers {batch_sampler.length} samples, the split holds {len(indices)}")
        samples = [dataset.samples[i] for i in indices]
        if not samples:
            self.images = th.empty(0, 0, 0, 0, dtype=th.uint8, device=self.device)
            self.labels = th.empty(0, dtype=th.int64, device=self.device)
            return
        first = to_u8_hwc(dataset.loader(samples[0][0]))
        need = len(samples) * first.numel()
        if self.device.type == "cuda":
            free, _ = th.cuda.mem_get_info(self.device)
            if need > hbm_fraction * free:
                raise MemoryError(f"resident split needs {need / 2**30:.1f} GiB of HBM but only "
                                  f"{hbm_fraction:.0%} of the free {free / 2**30:.1f} GiB may be used: "
                                  "drop --resident (streaming loader) or shard over more GPUs")
        host = th.empty((len(samples), *first.shape), dtype=th.uint8, pin_memory=self.device.type == "cuda")

        def decode(k: int) -> None:
            img = first if k == 0 else to_u8_hwc(dataset.loader(samples[k][0]))
            if img.shape != first.shape:
                raise ValueError(f"{samples[k][0]}: size {tuple(img.shape)} differs from {tuple(first.shape)}; "
                                 "a resident split needs equally sized images")
            host[k] = img

        with ThreadPoolExecutor(max_workers=max(1, decode_threads)) as pool:
            list(pool.map(decode, range(len(samples))))
        self.images = host.to(self.device, non_blocking=True)
        self.labels = th.tensor([t for _, t in samples], dtype=th.int64).to(self.device)
        if self.device.type == "cuda":
            th.cuda.current_stream(self.device).synchronize()  # the pinned staging buffer is freed on return

    @property
    def nbytes(self) -> int:
        return self.images.numel()

    def __len__(self) -> int:
        return len(self.batch_sampler)

    def __iter__(self):
        for batch in self.batch_sampler:
            idx = th.as_tensor(batch, dtype=th.int64, device=self.device)
            yield self.images.index_select(0, idx), self.labels.index_select(0, idx)
