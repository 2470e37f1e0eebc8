"""``test`` mode of the CLI (reference: eval.py:14-87): load ``marl.json`` + a ``state_dict``,
run forward-only episodes over an image folder and print the confusion matrix, per-class
precision and recall."""
from __future__ import annotations

from os import makedirs
from os.path import exists, isdir

from torch.utils.data import DataLoader

from .config import EvalConfig, MainConfig, ModelConfig
from .core import EpisodeSampler
from .data import FolderDataset, ShardedBatchSampler, collate_images, u8_image_pipeline
from .metrics import ConfusionMeter, format_metric
from .runtime import cuda_device
from .training import Trainer


def eval_main(main_config: MainConfig, eval_config: EvalConfig, num_workers: int = 8) -> ConfusionMeter:
    out_dir = eval_config.output_dir
    if exists(out_dir) and not isdir(out_dir):
        raise NotADirectoryError(f'"{out_dir}" is not a directory')
    print(f"File in {out_dir} will be overwritten" if exists(out_dir) else f'Create "{out_dir}"')
    makedirs(out_dir, exist_ok=True)

    device = cuda_device(main_config.cuda)
    _, networks, agents, env = ModelConfig.load_trained(eval_config.json_path, eval_config.state_dict_path,
                                                        main_config.nb_agent, device)
    images = FolderDataset(eval_config.dataset_path, u8_image_pipeline())  # decoded bytes; ToTensor on the device
    loader = DataLoader(images, batch_sampler=ShardedBatchSampler(len(images), eval_config.batch_size, shuffle=True),
                        num_workers=num_workers, pin_memory=True, collate_fn=collate_images)
    # forward-only graph replay over byte-wise prefetched batches (lr / gamma are unused in eval);
    # the vote is the mean over agents of the last step's prediction, eval.py:66-73
    evaluator = Trainer(networks, networks.nb_class, 0.0, 0.99)
    meter = evaluator.eval_epoch(loader, 0, EpisodeSampler(agents, env, main_config.step))

    print(meter.conf_mat())
    for title, per_class in (("Precision", meter.precision()), ("Recall", meter.recall())):
        print(f"{title} : {format_metric(per_class, images.class_to_idx)}")
        print(f"{title} mean = {per_class.mean()}")
    meter.save_conf_matrix(0, out_dir, "test")
    return meter
