"""Episode visualisation (reference: visualization.py:14-99): what the agents have uncovered after
each step, the vote of that step in the title, and an animated GIF.  Same file names as the
reference (``pred_original.png``, ``pred_step_{t}.png``, ``animated_gif.gif``); frames are drawn
with PIL instead of matplotlib, so the CLI runs without a plotting stack."""
from __future__ import annotations

from os.path import join
from typing import Any, List, Mapping, Optional

import numpy as np
import torch as th
from PIL import Image, ImageDraw

from .core import EpisodeSampler

_TITLE_H = 18
_MIN_SIDE = 256


def _frame(rgba: np.ndarray, title: str) -> Image.Image:
    """u8[H,W,4] canvas + title strip, scaled up (nearest) so small images stay readable."""
    h, w = rgba.shape[:2]
    scale = max(1, _MIN_SIDE // max(h, w))
    body = Image.fromarray(rgba, "RGBA").resize((w * scale, h * scale), Image.NEAREST)
    backdrop = Image.new("RGBA", body.size, (255, 255, 255, 255))
    backdrop.alpha_composite(body)
    out = Image.new("RGB", (max(body.width, 8 * len(title)), body.height + _TITLE_H), (255, 255, 255))
    out.paste(backdrop.convert("RGB"), ((out.width - body.width) // 2, _TITLE_H))
    ImageDraw.Draw(out).text((2, 3), title, fill=(0, 0, 0))
    return out


def heatmap_image(mat: th.Tensor, title: str = "", xlabel: str = "", ylabel: str = "", cell: Optional[int] = None
                  ) -> Image.Image:
    """Square matrix in [0,1] as a colour map (dark blue -> yellow), one cell per entry."""
    m = mat.detach().to(th.float32).cpu().clamp(0, 1).numpy()
    n = m.shape[0]
    cell = cell or max(4, 320 // max(1, n))
    stops = np.array([[13, 8, 135], [156, 23, 158], [237, 121, 83], [240, 249, 33]], dtype=np.float32)
    x = m * (len(stops) - 1)
    lo = np.clip(np.floor(x).astype(int), 0, len(stops) - 2)
    frac = (x - lo)[..., None]
    rgb = (stops[lo] * (1 - frac) + stops[lo + 1] * frac).astype(np.uint8)
    body = Image.fromarray(rgb, "RGB").resize((m.shape[1] * cell, n * cell), Image.NEAREST)
    out = Image.new("RGB", (max(body.width + 2 * _TITLE_H, 8 * len(title)), body.height + 2 * _TITLE_H), (255, 255, 255))
    out.paste(body, (_TITLE_H, _TITLE_H))
    d = ImageDraw.Draw(out)
    d.text((2, 3), title, fill=(0, 0, 0))
    d.text((_TITLE_H, body.height + _TITLE_H + 3), f"x: {xlabel}   y: {ylabel}", fill=(0, 0, 0))
    return out


def visualize_steps(episode_sampler: EpisodeSampler, img: th.Tensor, img_ori: th.Tensor, window_size: int,
                    output_dir: str, class_map: Mapping[Any, int]) -> None:
    """``img`` f32[C,H,W] goes through one episode (batch of 1); ``img_ori`` is what gets drawn."""
    idx_to_class = {class_map[k]: k for k in class_map}
    with th.no_grad():
        output = episode_sampler.run_episode(img.unsqueeze(0))
    nb_steps, nb_agents = output.step_preds.size(0), output.step_preds.size(1)
    preds = output.step_preds.mean(dim=1).cpu()  # vote: mean over agents, visualization.py:35
    pos = output.step_pos.cpu()
    ori = img_ori.detach().to(th.float32).cpu().permute(1, 2, 0)
    if ori.size(2) == 1:
        ori = ori.repeat(1, 1, 3)
    h, w, _ = ori.shape
    ori_u8 = (ori.clamp(0, 1) * 255).round().to(th.uint8).numpy()

    frames: List[Image.Image] = []
    first = _frame(np.concatenate([ori_u8, np.full((h, w, 1), 255, np.uint8)], axis=2), "Original")
    first.save(join(output_dir, "pred_original.png"))
    frames.extend([first] * 5)  # 5 x 200 ms = 1 s on the original

    seen = np.zeros((h, w, 4), np.uint8)
    for t in range(nb_steps):
        for a in range(nb_agents):
            y, x = int(pos[t, a, 0, 0]), int(pos[t, a, 0, 1])
            seen[y:y + window_size, x:x + window_size, :3] = ori_u8[y:y + window_size, x:x + window_size]
            seen[y:y + window_size, x:x + window_size, 3] = 255
        proba = th.softmax(preds[t, 0], dim=-1)
        best = int(proba.argmax())
        fr = _frame(seen, f"Step = {t}, step_pred_class = {idx_to_class[best]} ({proba[best].item() * 100.:.1f}%)")
        fr.save(join(output_dir, f"pred_step_{t}.png"))
        frames.append(fr)

    width, height = max(f.width for f in frames), max(f.height for f in frames)
    same = []
    for f in frames:  # GIF frames must share one size
        canvas = Image.new("RGB", (width, height), (255, 255, 255))
        canvas.paste(f, (0, 0))
        same.append(canvas)
    same[0].save(join(output_dir, "animated_gif.gif"), save_all=True, append_images=same[1:], duration=200, loop=0)
