"""``infer`` mode of the CLI (reference: infer.py:17-95): one visualised episode per image, each
in its own sub-directory of the output directory together with an ``info.txt`` naming the inputs."""
from __future__ import annotations

import glob
import json
from datetime import datetime
from os import makedirs
from os.path import basename, getmtime, join

from tqdm import tqdm

from .config import InferConfig, MainConfig, ModelConfig
from .core import EpisodeSampler
from .data import pil_rgb_loader, to_f32_chw
from .runtime import cuda_device
from .visualization import visualize_steps


def infer_main(main_config: MainConfig, infer_config: InferConfig) -> None:
    device = cuda_device(main_config.cuda)
    config, _, agents, env = ModelConfig.load_trained(infer_config.json_path, infer_config.state_dict_path,
                                                      main_config.nb_agent, device)
    used = {"JSON": infer_config.json_path, "state_dict": infer_config.state_dict_path,
            "class_to_idx": infer_config.class_to_idx}
    print("Will use :\n" + "\n".join(f"- {k} of : {datetime.fromtimestamp(getmtime(v))}" for k, v in used.items()))
    with open(infer_config.class_to_idx, "r", encoding="utf-8") as fh:
        class_to_idx = json.load(fh)

    sampler = EpisodeSampler(agents, env, main_config.step)
    matches = [path for pattern in infer_config.images_path for path in sorted(glob.glob(pattern, recursive=True))]
    for path in tqdm(matches):
        pixels = to_f32_chw(pil_rgb_loader(path))
        target = join(infer_config.output_dir, basename(path))
        makedirs(target, exist_ok=True)
        with open(join(target, "info.txt"), "w", encoding="utf-8") as fh:
            fh.write("\n".join((path, infer_config.json_path, infer_config.state_dict_path)) + "\n")
        visualize_steps(sampler, pixels.to(device), pixels, config.window_size, target, class_to_idx)
