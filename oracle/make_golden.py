"""Generate golden fixtures from the UNMODIFIED reference (test infrastructure).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.pt

For every case it
  1. builds the reference's ModelsWrapper/MultiAgent/Environment
     (config.py:95-104 ``build_marl``), perturbs the weights with a seeded
     generator so biases / norm affines are non-trivial (or loads the shipped
     trained MNIST checkpoint),
  2. runs ONE iteration of the reference's own ``Trainer.train_epoch``
     (trainer.py:55-162) on a one-batch loader while RECORDING the three random
     sites (torch.randint x2, torch.randn x4, torch.multinomial xT -- SURVEY
     section 8c); the gradients are read from ``param.grad`` afterwards and the
     loss parts from the reference's metric-logger callback,
  3. restores the pre-step weights and REPLAYS the recorded draws through the
     reference's ``EpisodeSampler.run_episode`` (episode.py:84-85) to dump
     step_preds / step_log_probas / step_values / step_pos and the
     observations o_0..o_T (environment.py:47-54).

Nothing from the reference is copied: it is imported from where it lies.
matplotlib / mlflow are absent from the image, so an empty stub
``matplotlib.pyplot`` is put on sys.path for the import of
``marl_classification.metrics`` (only used for meters, never for numbers).
"""

from __future__ import annotations

import json
import os
import sys
import tempfile
from contextlib import contextmanager

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _import_reference():
    stub = tempfile.mkdtemp(prefix="mpl_stub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    with open(os.path.join(stub, "matplotlib", "__init__.py"), "w") as fh:
        fh.write("")
    with open(os.path.join(stub, "matplotlib", "pyplot.py"), "w") as fh:
        fh.write("def __getattr__(name):\n    return lambda *a, **k: None\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    import marl_classification  # noqa: F401

    assert marl_classification.__file__.startswith(REF), marl_classification.__file__


class Tape:
    """Record or replay the path's three random sites."""

    def __init__(self, replay=None):
        self.randint, self.randn, self.multinomial = [], [], []
        self.replay = replay
        self.i = [0, 0, 0]


@contextmanager
def rng_tape(tape: Tape):
    o_randint, o_randn, o_mult = torch.randint, torch.randn, torch.multinomial

    def randint(*a, **k):
        if tape.replay is not None:
            out = tape.replay.randint[tape.i[0]].clone()
            tape.i[0] += 1
            return out
        out = o_randint(*a, **k)
        tape.randint.append(out.clone())
        return out

    def randn(*a, **k):
        if tape.replay is not None:
            out = tape.replay.randn[tape.i[1]].clone()
            tape.i[1] += 1
            return out
        out = o_randn(*a, **k)
        tape.randn.append(out.clone())
        return out

    def multinomial(*a, **k):
        if tape.replay is not None:
            out = tape.replay.multinomial[tape.i[2]].clone()
            tape.i[2] += 1
            return out
        out = o_mult(*a, **k)
        tape.multinomial.append(out.clone())
        return out

    torch.randint, torch.randn, torch.multinomial = randint, randn, multinomial
    try:
        yield tape
    finally:
        torch.randint, torch.randn, torch.multinomial = o_randint, o_randn, o_mult


CASES = {
    # shipped trained weights (resources/trained_models/mnist), RGB input like
    # data/datasets.py:22; CNN reads channel 0 (vision.py:64)
    "mnist_ckpt": dict(
        marl_json=f"{REF}/resources/trained_models/mnist/marl.json",
        ckpt=f"{REF}/resources/trained_models/mnist/nn_models_epoch_49.pt",
        na=3, nb=6, T=5, C=3, H=28, W=28, gamma=0.99, seed=11,
    ),
    # the reference's own test fixture sizes (tests/conftest.py:16-79): ragged dims
    "conftest_odd": dict(
        cfg=dict(ft_extr_str="mnist", window_size=12, hidden_size_belief=23, hidden_size_action=22,
                 hidden_size_msg=21, hidden_size_msg_output=20, hidden_size_state=19, state_dim=2,
                 actions=[[1, 0], [-1, 0], [0, 1], [0, -1]], nb_class=10,
                 hidden_size_linear_belief=24, hidden_size_linear_action=25),
        na=5, nb=7, T=7, C=1, H=28, W=28, gamma=0.99, seed=12,
    ),
    # RESISC45 CNN (3 conv blocks, RGB), non-square image, reduced widths
    "resisc_small": dict(
        cfg=dict(ft_extr_str="resisc45", window_size=12, hidden_size_belief=48, hidden_size_action=40,
                 hidden_size_msg=16, hidden_size_msg_output=24, hidden_size_state=8, state_dim=2,
                 actions=[[1, 0], [-1, 0], [0, 1], [0, -1]], nb_class=45,
                 hidden_size_linear_belief=64, hidden_size_linear_action=56),
        na=4, nb=3, T=6, C=3, H=40, W=56, gamma=0.95, seed=13,
    ),
    # AID CNN (4 conv blocks), moves of +-3 and a "stay" action, single agent
    # edge: Na=1 in aid_single exercises aggregate_messages' zero branch
    "aid_small": dict(
        cfg=dict(ft_extr_str="aid", window_size=24, hidden_size_belief=32, hidden_size_action=32,
                 hidden_size_msg=8, hidden_size_msg_output=12, hidden_size_state=4, state_dim=2,
                 actions=[[3, 0], [-3, 0], [0, 3], [0, -3], [0, 0]], nb_class=30,
                 hidden_size_linear_belief=40, hidden_size_linear_action=36),
        na=3, nb=2, T=5, C=3, H=64, W=48, gamma=0.9, seed=14,
    ),
    "single_agent": dict(
        cfg=dict(ft_extr_str="mnist", window_size=6, hidden_size_belief=16, hidden_size_action=16,
                 hidden_size_msg=8, hidden_size_msg_output=8, hidden_size_state=4, state_dim=2,
                 actions=[[1, 0], [-1, 0], [0, 1], [0, -1]], nb_class=10,
                 hidden_size_linear_belief=16, hidden_size_linear_action=16),
        na=1, nb=4, T=4, C=1, H=12, W=12, gamma=0.99, seed=15,
    ),
}


def run_case(name: str, spec: dict) -> dict:
    from marl_classification.config import ModelConfig
    from marl_classification.core import EpisodeSampler
    from marl_classification.training.trainer import Trainer

    torch.manual_seed(spec["seed"])
    if "marl_json" in spec:
        mcfg = ModelConfig.load_marl_config(spec["marl_json"])
    else:
        mcfg = ModelConfig(**spec["cfg"])
    model, marl, env = mcfg.build_marl(spec["na"])
    if "ckpt" in spec:
        model.load_state_dict(torch.load(spec["ckpt"], map_location="cpu"))
    else:
        g = torch.Generator().manual_seed(spec["seed"] + 1000)
        with torch.no_grad():
            for prm in model.parameters():
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}

    g = torch.Generator().manual_seed(spec["seed"] + 2000)
    img = torch.rand(spec["nb"], spec["C"], spec["H"], spec["W"], generator=g)
    targets = torch.randint(mcfg.nb_class, (spec["nb"],), generator=g)

    logged = {}

    def logger(step, metrics):
        logged.update(metrics)

    sampler = EpisodeSampler(marl, env, spec["T"])
    trainer = Trainer(model, mcfg.nb_class, 1e-4, spec["gamma"], metric_logger=logger, log_interval=1)
    rec = Tape()
    with rng_tape(rec):
        trainer.train_epoch([(img, targets)], 0, sampler)
    assert (len(rec.randint), len(rec.randn), len(rec.multinomial)) == (2, 4, spec["T"]), (
        len(rec.randint), len(rec.randn), len(rec.multinomial))
    grads = {
        k: (prm.grad.detach().clone() if prm.grad is not None else torch.zeros_like(prm))
        for k, prm in model.named_parameters()
    }  # None only for decode_msg when Na == 1 (message.py:14-15 cuts the graph)

    # replay on the pre-step weights for the forward outputs
    model.load_state_dict(state0)
    obs_log = []
    orig_observe = env.observe

    def observe_spy():
        o = orig_observe()
        obs_log.append(o.detach().clone())
        return o

    env.observe = observe_spy
    with torch.no_grad(), rng_tape(Tape(replay=rec)):
        out = sampler.run_episode(img)
    env.observe = orig_observe
    assert len(obs_log) == spec["T"] + 1

    na, nb = spec["na"], spec["nb"]
    fx = dict(
        name=name,
        model_config=json.loads(mcfg.model_dump_json()),
        na=na, nb=nb, T=spec["T"], gamma=spec["gamma"],
        state_dict=state0,
        img=img, targets=targets,
        pos0=torch.stack(rec.randint, dim=-1),  # environment.py:33-43
        hidden0=[t for t in rec.randn],  # h, c, h^, c^ (models.py:151-159)
        actions=torch.stack([m.view(na, nb) for m in rec.multinomial]),  # agent.py:53-55
        step_preds=out.step_preds, step_log_probas=out.step_log_probas,
        step_values=out.step_values, step_pos=out.step_pos,
        obs=[obs_log[0], obs_log[1], obs_log[-1]],  # o_0, o_1, o_T
        logged={k: float(v) for k, v in logged.items()},
        grads=grads,
        torch_version=torch.__version__,
    )
    return fx


def run_input_pipeline_case() -> dict:
    """The reference's host-side image pipeline (registry.py:56-57: Compose([ToTensor()])) applied,
    as its DataLoader does, per PIL image (datasets.py:17-22 converts to RGB first); batches are
    the default collate (stack).  Sizes: MNIST-like, odd, W not a multiple of 4, single channel."""
    import numpy as np
    from PIL import Image

    from marl_classification.registry import default_image_pipeline

    pipe = default_image_pipeline()
    g = torch.Generator().manual_seed(7)
    fx = {"cases": []}
    for (b, h, w, c) in [(4, 28, 28, 3), (3, 7, 9, 3), (2, 12, 16, 3), (2, 10, 6, 1), (1, 64, 64, 3)]:
        u8 = torch.randint(0, 256, (b, h, w, c), dtype=torch.uint8, generator=g)
        u8[0].view(-1)[:2] = torch.tensor([0, 255], dtype=torch.uint8)  # both ends of the range
        imgs = [Image.fromarray(u8[i].numpy().squeeze(-1) if c == 1 else u8[i].numpy()) for i in range(b)]
        out = torch.stack([pipe(im) for im in imgs])
        fx["cases"].append({"u8_hwc": u8, "f32_chw": out})
    return fx


def main() -> None:
    _import_reference()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    if not only or "input_pipeline" in only:
        fx = run_input_pipeline_case()
        path = os.path.join(OUT, "input_pipeline.pt")
        torch.save(fx, path)
        print(f"input_pipeline: wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        fx = run_case(name, spec)
        path = os.path.join(OUT, f"{name}.pt")
        torch.save(fx, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB) loss={fx['logged']['loss']:.6f}")


if __name__ == "__main__":
    main()
