"""``infer`` mode of the CLI (reference: infer.py:17-95): one visualised episode per image."""
from __future__ import annotations

import glob
import json
from datetime import datetime
from os import makedirs
from os.path import exists, getmtime, isfile, join, split

import torch as th
from tqdm import tqdm

from .config import InferConfig, MainConfig, ModelConfig
from .core import EpisodeSampler
from .data import default_image_pipeline, pil_rgb_loader
from .runtime import cuda_device
from .visualization import visualize_steps


def infer_main(main_config: MainConfig, infer_config: InferConfig) -> None:
    assert exists(infer_config.json_path), f'JSON path "{infer_config.json_path}" does not exist'
    assert isfile(infer_config.json_path), f'"{infer_config.json_path}" is not a file'
    assert exists(infer_config.state_dict_path), f"State dict path {infer_config.state_dict_path} does not exist"
    assert isfile(infer_config.state_dict_path), f"{infer_config.state_dict_path} is not a file"
    print("Will use :\n"
          f"- JSON of : {datetime.fromtimestamp(getmtime(infer_config.json_path))}\n"
          f"- state_dict of : {datetime.fromtimestamp(getmtime(infer_config.state_dict_path))}\n"
          f"- class_to_idx of : {datetime.fromtimestamp(getmtime(infer_config.class_to_idx))}")
    with open(infer_config.class_to_idx, "r", encoding="utf-8") as json_f:
        class_to_idx = json.load(json_f)

    device = cuda_device(main_config.cuda)
    marl_config = ModelConfig.load_marl_config(infer_config.json_path)
    nn_models, marl_m, env = marl_config.build_marl(main_config.nb_agent)
    nn_models.load_state_dict(th.load(infer_config.state_dict_path, map_location="cpu"))
    nn_models.eval()
    nn_models.to(device)
    episode_sampler = EpisodeSampler(marl_m, env, main_config.step)
    img_pipeline = default_image_pipeline()

    paths = [p for pattern in infer_config.images_path for p in sorted(glob.glob(pattern, recursive=True))]
    for img_path in tqdm(paths):
        x = img_pipeline(pil_rgb_loader(img_path))
        curr_img_path = join(infer_config.output_dir, split(img_path)[-1])
        makedirs(curr_img_path, exist_ok=True)
        with open(join(curr_img_path, "info.txt"), "w", encoding="utf-8") as info_f:
            info_f.writelines([f"{img_path}\n", f"{infer_config.json_path}\n", f"{infer_config.state_dict_path}\n"])
        visualize_steps(episode_sampler, x.to(device), x, marl_config.window_size, curr_img_path, class_to_idx)
