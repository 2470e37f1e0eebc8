"""Synthetic image-folder datasets for the CLI tests (PNG files written with PIL)."""
import os

import numpy as np
from PIL import Image


def make_image_folder(root: str, classes: int, per_class: int, size: int, grey_every: int = 0, nested: bool = False,
                      seed: int = 0) -> str:
    """``root/class_{k}/img_{i}.png``; class k images are noise around a class-specific level so a
    classifier has something to learn.  ``grey_every`` > 0 writes every n-th image as 8-bit grey
    (MNIST PNGs are); ``nested`` puts some files one directory deeper (ImageFolder walks them)."""
    rng = np.random.default_rng(seed)
    n = 0
    for k in range(classes):
        for i in range(per_class):
            d = os.path.join(root, f"class_{k}", "part") if nested and i % 2 else os.path.join(root, f"class_{k}")
            os.makedirs(d, exist_ok=True)
            level = int(255 * (k + 0.5) / classes)
            if grey_every and n % grey_every == 0:
                arr = np.clip(rng.normal(level, 20, (size, size)), 0, 255).astype(np.uint8)
                Image.fromarray(arr, "L").save(os.path.join(d, f"img_{i}.png"))
            else:
                arr = np.clip(rng.normal(level, 20, (size, size, 3)), 0, 255).astype(np.uint8)
                Image.fromarray(arr, "RGB").save(os.path.join(d, f"img_{i}.png"))
            n += 1
    return root
