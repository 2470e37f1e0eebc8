"""``train`` mode of the CLI (reference: train.py:18-168): build the networks from the options,
split the dataset 85/15, alternate ``Trainer.train_epoch`` / ``eval_epoch``, save ``marl.json``,
``class_to_idx.json``, a confusion matrix and a ``state_dict`` per epoch, and finish with one
visualised episode.

Differences that come from the B200 path, not from the interface: batches travel as decoded
bytes and become fp32 on the device (``u8_image_pipeline`` + ``Trainer.prefetch``); under
``torch.distributed.run`` every rank trains on its shard of each global batch with the flat
gradient all-reduce; ``--resident`` decodes each split once and serves every batch from HBM
(``data.ResidentImages``); MLflow is optional (``tracking.RunTracker``)."""
from __future__ import annotations

import json
import random
from os import makedirs
from os.path import isdir, join
from typing import Dict, List, Optional, Tuple

import torch as th
from torch.utils.data import DataLoader, Subset

from .config import MainConfig, ModelConfig, TrainConfig
from .core import EpisodeSampler
from .data import ResidentImages, ShardedBatchSampler, collate_images, to_f32_chw, u8_image_pipeline
from .parallel import DataParallelContext
from .registry import get_dataset_spec
from .runtime import cuda_device, data_parallel, shutdown
from .tracking import RunTracker
from .training import Trainer
from .visualization import visualize_steps

SPLIT_SEED = 0x5eed
TRAIN_FRACTION = 0.85  # train.py:82


def split_indices(n: int, fraction: float = TRAIN_FRACTION, seed: int = SPLIT_SEED) -> Tuple[List[int], List[int]]:
    """Seeded random split (the reference draws an unseeded ``randperm``, train.py:83-87; every
    data-parallel rank must cut the same way, so the permutation is fixed here)."""
    order = th.randperm(n, generator=th.Generator().manual_seed(seed)).tolist()
    cut = int(fraction * n)
    return order[:cut], order[cut:]


def _batches(dataset, indices: List[int], batch_size: int, dp: DataParallelContext, seed: int, workers: int,
             resident_on=None):
    """Batches of one split: streamed by DataLoader workers (decoded bytes, pinned), or - with
    ``resident_on`` a device - decoded once and served from HBM (``data.ResidentImages``)."""
    sampler = ShardedBatchSampler(len(indices), batch_size, dp.rank, dp.world_size, shuffle=True, seed=seed)
    if resident_on is not None:
        return ResidentImages(dataset, indices, sampler, resident_on, decode_threads=max(1, workers))
    return DataLoader(Subset(dataset, indices), batch_sampler=sampler, num_workers=workers, pin_memory=True,
                      collate_fn=collate_images, persistent_workers=workers > 0)


def _open_run(main_config, model_config, train_config, dataset, device, dp) -> RunTracker:
    """Rank 0 only: run directory, tracker, ``marl.json`` and ``class_to_idx.json`` (train.py:30-75)."""
    out = train_config.output_dir
    makedirs(join(out, "models"), exist_ok=True)
    if not isdir(join(out, "models")):
        raise NotADirectoryError(f'"{join(out, "models")}" is not a directory.')
    tracker = RunTracker("MARLClassification", f"train_{main_config.run_id}", out)
    tracker.log_params({"output_dir": out, "model_dir": join(out, "models"), **dict(main_config), **dict(model_config),
                        **dict(train_config), "device": device.type, "world_size": dp.world_size})
    model_config.save_marl_config(join(out, "marl.json"))
    with open(join(out, "class_to_idx.json"), "w", encoding="utf-8") as fh:
        json.dump(dataset.class_to_idx, fh)
    return tracker


def train_main(main_config: MainConfig, model_config: ModelConfig, train_config: TrainConfig,
               num_workers: int = 6, resident: bool = False) -> None:
    assert model_config.state_dim == 2, (
        "Only 2D is supported by the CUDA episode (the reference also admits 3-D volumes, train.py:23-28)"
    )
    device = cuda_device(main_config.cuda)
    dp = data_parallel(device)
    out = train_config.output_dir

    dataset = get_dataset_spec(model_config.ft_extr_str).dataset_constructor(train_config.resources_dir,
                                                                             u8_image_pipeline())
    networks, agents, env = model_config.build_marl(main_config.nb_agent)
    tracker: Optional[RunTracker] = None
    if dp.rank == 0:
        tracker = _open_run(main_config, model_config, train_config, dataset, device, dp)

    networks.to(device)
    networks.ensure_flat()
    dp.broadcast_params(networks.flat_params)  # replicas start from rank 0's initialisation

    train_idx, held_out_idx = split_indices(len(dataset))
    where = device if resident else None
    train_batches = _batches(dataset, train_idx, train_config.batch_size, dp, 1, num_workers, where)
    held_out_batches = _batches(dataset, held_out_idx, train_config.batch_size, dp, 2, num_workers, where)

    def log(step: int, metrics: Dict[str, float]) -> None:
        if tracker is not None:
            tracker.log_metrics(step=step, metrics=metrics)

    sampler = EpisodeSampler(agents, env, main_config.step)
    trainer = Trainer(networks, agents.nb_class, train_config.learning_rate, train_config.gamma,
                      metric_logger=log, dp=dp)

    for epoch in range(train_config.nb_epoch):
        for loader in (train_batches, held_out_batches):
            loader.batch_sampler.set_epoch(epoch)
        trainer.train_epoch(train_batches, epoch, sampler)
        meter = trainer.eval_epoch(held_out_batches, epoch, sampler)
        if tracker is None:
            continue  # replicas hold identical weights; rank 0 reports its shard of the held-out split
        meter.save_conf_matrix(epoch, out, "eval")
        log(trainer.curr_step, {"eval_prec": meter.precision().mean().item(),
                                "eval_recs": meter.recall().mean().item()})
        th.save(networks.state_dict(), join(out, "models", f"nn_models_epoch_{epoch}.pt"))

    if tracker is not None:
        if held_out_idx:  # one visualised episode on a held-out image (train.py:149-166)
            path, _ = dataset.samples[random.choice(held_out_idx)]
            pixels = to_f32_chw(dataset.loader(path))
            visualize_steps(sampler, pixels.to(device), pixels, model_config.window_size, out, dataset.class_to_idx)
        tracker.end()
    shutdown()
