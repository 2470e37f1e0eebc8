"""``train`` mode of the CLI (reference: train.py:18-168): build the networks from the options,
split the dataset 85/15, alternate ``Trainer.train_epoch`` / ``eval_epoch``, save ``marl.json``,
``class_to_idx.json``, a confusion matrix and a ``state_dict`` per epoch, and finish with one
visualised episode.

Differences that come from the B200 path, not from the interface: batches travel as decoded
bytes and become fp32 on the device (``u8_image_pipeline`` + ``Trainer.prefetch``); under
``torch.distributed.run`` every rank trains on its shard of each global batch with the flat
gradient all-reduce; MLflow is optional (``tracking.RunTracker``)."""
from __future__ import annotations

import json
from os import makedirs
from os.path import isdir, join
from random import randrange
from typing import Dict

import torch as th
from torch.utils.data import DataLoader, Subset

from .config import MainConfig, ModelConfig, TrainConfig
from .core import EpisodeSampler
from .data import ShardedBatchSampler, collate_images, default_image_pipeline, u8_image_pipeline
from .registry import get_dataset_spec
from .runtime import cuda_device, data_parallel, shutdown
from .tracking import RunTracker
from .training import Trainer
from .visualization import visualize_steps

SPLIT_SEED = 0x5eed


def _loader(dataset, batch_size: int, dp, seed: int, num_workers: int) -> DataLoader:
    sampler = ShardedBatchSampler(len(dataset), batch_size, dp.rank, dp.world_size, shuffle=True, seed=seed)
    return DataLoader(dataset, batch_sampler=sampler, num_workers=num_workers, pin_memory=True,
                      collate_fn=collate_images, persistent_workers=num_workers > 0)


def train_main(main_config: MainConfig, model_config: ModelConfig, train_config: TrainConfig,
               num_workers: int = 6) -> None:
    assert model_config.state_dim == 2, (
        "Only 2D is supported by the CUDA episode (the reference also admits 3-D volumes, train.py:23-28)"
    )
    device = cuda_device(main_config.cuda)
    dp = data_parallel(device)
    lead = dp.rank == 0

    output_dir = train_config.output_dir
    model_dir = join(output_dir, "models")
    if lead:
        makedirs(model_dir, exist_ok=True)
        if not isdir(model_dir):
            raise NotADirectoryError(f'"{model_dir}" is not a directory.')

    tracker = RunTracker("MARLClassification", f"train_{main_config.run_id}", output_dir) if lead else None

    dataset_spec = get_dataset_spec(model_config.ft_extr_str)
    nn_models, marl_m, env = model_config.build_marl(main_config.nb_agent)
    dataset = dataset_spec.dataset_constructor(train_config.resources_dir, u8_image_pipeline())

    if lead:
        tracker.log_params({"output_dir": output_dir, "model_dir": model_dir, **dict(main_config),
                            **dict(model_config), **dict(train_config), "device": device.type,
                            "world_size": dp.world_size})
        model_config.save_marl_config(join(output_dir, "marl.json"))
        with open(join(output_dir, "class_to_idx.json"), "w", encoding="utf-8") as json_f:
            json.dump(dataset.class_to_idx, json_f)

    nn_models.to(device)
    nn_models.ensure_flat()
    dp.broadcast_params(nn_models.flat_params)  # replicas start from rank 0's initialisation

    # 85 % train / 15 % eval (train.py:82-91); the permutation is seeded so every rank agrees
    ratio_eval = 0.85
    idx = th.randperm(len(dataset), generator=th.Generator().manual_seed(SPLIT_SEED))
    cut = int(ratio_eval * idx.size(0))
    idx_train, idx_test = idx[:cut].tolist(), idx[cut:].tolist()
    train_dataset, test_dataset = Subset(dataset, idx_train), Subset(dataset, idx_test)
    train_dataloader = _loader(train_dataset, train_config.batch_size, dp, 1, num_workers)
    test_dataloader = _loader(test_dataset, train_config.batch_size, dp, 2, num_workers)

    def metric_logger(step: int, metrics: Dict[str, float]) -> None:
        if tracker is not None:
            tracker.log_metrics(step=step, metrics=metrics)

    episode_sampler = EpisodeSampler(marl_m, env, main_config.step)
    trainer = Trainer(nn_models, marl_m.nb_class, train_config.learning_rate, train_config.gamma,
                      metric_logger=metric_logger, dp=dp)

    for e in range(train_config.nb_epoch):
        train_dataloader.batch_sampler.set_epoch(e)
        test_dataloader.batch_sampler.set_epoch(e)
        trainer.train_epoch(train_dataloader, e, episode_sampler)
        conf_meter_eval = trainer.eval_epoch(test_dataloader, e, episode_sampler)
        if not lead:
            continue  # replicas hold identical weights; rank 0 reports its shard of the eval split
        precs, recs = conf_meter_eval.precision(), conf_meter_eval.recall()
        conf_meter_eval.save_conf_matrix(e, output_dir, "eval")
        tracker.log_metrics(step=trainer.curr_step,
                            metrics={"eval_prec": precs.mean().item(), "eval_recs": recs.mean().item()})
        th.save(nn_models.state_dict(), join(model_dir, f"nn_models_epoch_{e}.pt"))

    if lead and len(idx_test) > 0:
        to_f32 = default_image_pipeline()
        path, _ = dataset.samples[idx_test[randrange(len(idx_test))]]
        x = to_f32(dataset.loader(path)).to(device)
        visualize_steps(episode_sampler, x, x, model_config.window_size, output_dir, dataset.class_to_idx)

    if tracker is not None:
        tracker.end()
    shutdown()
