"""Data-parallel protocol on CPU (gloo, world_size 2): sharding the image batch,
exchanging the advantage statistics {sum, sumsq, n} and averaging the flat gradient
bucket reproduces the single-process full-batch gradient.  The
arithmetic here is the CPU oracle; the collectives and the sharding are the product's
``DataParallelContext`` -- exactly what runs over NCCL on the GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_golden, oracle_config, rel_l2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_path: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from marlclassification_b200.parallel import DataParallelContext
    from oracle import marl_oracle as O

    torch.set_num_threads(1)
    fx = load_golden("resisc_small")
    cfg = oracle_config(fx["model_config"])
    nb = fx["nb"] - fx["nb"] % world  # equal shards
    dp = DataParallelContext()
    assert dp.enabled and dp.world_size == world and dp.rank == rank
    img, y = fx["img"][:nb], fx["targets"][:nb]
    pos0, hidden0, actions = fx["pos0"][:, :nb], [h[:, :nb] for h in fx["hidden0"]], fx["actions"][:, :, :nb]
    # batch dim of every per-image tensor is sharded the same way (all agents of an image stay together)
    sl = lambda t, dim: dp.shard(t.transpose(0, dim)).transpose(0, dim).contiguous()  # noqa: E731
    img_s, y_s = dp.shard(img), dp.shard(y)
    pos_s, hid_s, act_s = sl(pos0, 1), [sl(h, 1) for h in hidden0], sl(actions, 2)
    leaf = {k: v.clone().requires_grad_(True) for k, v in fx["state_dict"].items()}
    ro = O.rollout(leaf, cfg, img_s, pos_s, hid_s, act_s, fx["T"])
    with torch.no_grad():
        adv = O.discounted_returns(O.classification_rewards(ro.step_preds, y_s), fx["gamma"]) - ro.step_values
        stats = torch.zeros(16, dtype=torch.float64)
        stats[0], stats[1], stats[2] = adv.double().sum(), (adv.double() ** 2).sum(), adv.numel()
    dp.all_reduce_stats(stats)  # phase A -> phase B exchange
    n, mean = stats[2].item(), (stats[0] / stats[2]).item()
    std = ((stats[1] - n * mean * mean) / (n - 1)).clamp(min=0).sqrt().item()
    parts = O.a2c_loss(ro.step_preds, ro.step_log_probas, ro.step_values, y_s, fx["gamma"], adv_stats=(mean, std))
    names = list(leaf)
    grads = torch.autograd.grad(parts.loss, [leaf[k] for k in names])
    flat = torch.cat([g.flatten() for g in grads])
    dp.all_reduce_grads(flat)  # ONE collective over the whole bucket, then 1/world
    if rank == 0:
        _, parts_full, grads_full = O.loss_and_grads(fx["state_dict"], cfg, img, y, pos0, hidden0, actions, fx["T"], fx["gamma"])
        flat_full = torch.cat([grads_full[k].flatten() for k in names])
        torch.save({"err": rel_l2(flat, flat_full), "n": n, "expected_n": fx["T"] * fx["na"] * nb}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_protocol_matches_full_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["n"] == res["expected_n"]
    assert res["err"] < 1e-4, res


def test_shard_requires_divisible_batch():
    from marlclassification_b200.parallel import DataParallelContext

    dp = DataParallelContext(enabled=False)
    x = torch.arange(6)
    assert torch.equal(dp.shard(x), x) and dp.world_size == 1
