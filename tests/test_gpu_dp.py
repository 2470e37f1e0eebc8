"""Two-rank NCCL run of the real train step: sharded batch + stats exchange + ONE flat
gradient all-reduce + fused Adam  ==  the single-GPU full-batch step (needs >= 2 GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(fx, dev):
    from marlclassification_b200.config import ModelConfig
    from marlclassification_b200.core import EpisodeSampler

    model, marl, env = ModelConfig(**fx["model_config"]).build_marl(fx["na"])
    model.use_tc = False
    model.load_state_dict(fx["state_dict"])
    model.to(dev)
    return model, EpisodeSampler(marl, env, fx["T"], gamma=fx["gamma"])


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from marlclassification_b200.parallel import DataParallelContext
    from marlclassification_b200.training import Trainer

    fx = load_golden("conftest_odd")
    nb = fx["nb"] - fx["nb"] % world
    dp = DataParallelContext()
    model, sampler = _build(fx, dev)
    trainer = Trainer(model, fx["model_config"]["nb_class"], 1e-3, fx["gamma"], dp=dp, cuda_graph=False)
    sl = lambda t, dim: dp.shard(t[(slice(None),) * dim + (slice(0, nb),)].transpose(0, dim)).transpose(0, dim).contiguous().to(dev)  # noqa: E731
    img, y = sl(fx["img"], 0), sl(fx["targets"], 0)
    inject = dict(pos0=sl(fx["pos0"], 1), hidden0=[sl(h, 1) for h in fx["hidden0"]], actions=sl(fx["actions"], 2))
    for _ in range(2):
        out = trainer.train_step(img, y, sampler, **inject)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"params": model.flat_params.cpu(), "loss": out.cpu()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_matches_single_gpu(tmp_path):
    from marlclassification_b200.training import Trainer

    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    fx = load_golden("conftest_odd")
    nb = fx["nb"] - fx["nb"] % 2
    dev = torch.device("cuda", 0)
    model, sampler = _build(fx, dev)
    trainer = Trainer(model, fx["model_config"]["nb_class"], 1e-3, fx["gamma"], cuda_graph=False)
    inject = dict(pos0=fx["pos0"][:, :nb].contiguous().to(dev), hidden0=[h[:, :nb].contiguous().to(dev) for h in fx["hidden0"]],
                  actions=fx["actions"][:, :, :nb].contiguous().to(dev))
    for _ in range(2):
        trainer.train_step(fx["img"][:nb].to(dev), fx["targets"][:nb].to(dev), sampler, **inject)
    torch.cuda.synchronize()
    assert rel_l2(res["params"], model.flat_params.cpu()) < 1e-5
