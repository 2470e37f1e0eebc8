"""``test`` mode of the CLI (reference: eval.py:14-87): load ``marl.json`` + a ``state_dict``,
run forward-only episodes over an image folder and print the confusion matrix, per-class
precision and recall."""
from __future__ import annotations

from os import makedirs
from os.path import exists, isdir, isfile

import torch as th
from torch.utils.data import DataLoader

from .config import EvalConfig, MainConfig, ModelConfig
from .core import EpisodeSampler
from .data import FolderDataset, ShardedBatchSampler, collate_images, u8_image_pipeline
from .metrics import ConfusionMeter, format_metric
from .runtime import cuda_device
from .training import Trainer


def eval_main(main_config: MainConfig, eval_config: EvalConfig, num_workers: int = 8) -> ConfusionMeter:
    assert exists(eval_config.json_path), f'JSON path "{eval_config.json_path}" does not exist'
    assert isfile(eval_config.json_path), f'"{eval_config.json_path}" is not a file'
    assert exists(eval_config.state_dict_path), f"State dict path {eval_config.state_dict_path} does not exist"
    assert isfile(eval_config.state_dict_path), f"{eval_config.state_dict_path} is not a file"
    if exists(eval_config.output_dir) and not isdir(eval_config.output_dir):
        raise NotADirectoryError(f'"{eval_config.output_dir}" is not a directory')
    if exists(eval_config.output_dir):
        print(f"File in {eval_config.output_dir} will be overwritten")
    else:
        print(f'Create "{eval_config.output_dir}"')
        makedirs(eval_config.output_dir)

    device = cuda_device(main_config.cuda)
    test_dataset = FolderDataset(eval_config.dataset_path, u8_image_pipeline())
    marl_config = ModelConfig.load_marl_config(eval_config.json_path)
    nn_models, marl_m, env = marl_config.build_marl(main_config.nb_agent)
    nn_models.load_state_dict(th.load(eval_config.state_dict_path, map_location="cpu"))
    nn_models.eval()
    nn_models.to(device)

    data_loader = DataLoader(
        test_dataset, batch_sampler=ShardedBatchSampler(len(test_dataset), eval_config.batch_size, shuffle=True),
        num_workers=num_workers, pin_memory=True, collate_fn=collate_images,
    )
    episode_sampler = EpisodeSampler(marl_m, env, main_config.step)
    # forward-only graph replay over byte-wise prefetched batches; lr / gamma are unused in eval
    trainer = Trainer(nn_models, nn_models.nb_class, 0.0, 0.99)
    conf_meter = trainer.eval_epoch(data_loader, 0, episode_sampler)  # vote = mean over agents, eval.py:73

    print(conf_meter.conf_mat())
    precs, recs = conf_meter.precision(), conf_meter.recall()
    print(f"Precision : {format_metric(precs, test_dataset.class_to_idx)}")
    print(f"Precision mean = {precs.mean()}")
    print(f"Recall : {format_metric(recs, test_dataset.class_to_idx)}")
    print(f"Recall mean : {recs.mean()}")
    conf_meter.save_conf_matrix(0, eval_config.output_dir, "test")
    return conf_meter
