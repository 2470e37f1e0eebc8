"""Reward / return helpers with the reference's signatures (training/functions.py).

The Trainer does not call these on its hot path: the fused CUDA loss
(csrc/loss.cu) computes the same quantities in four small kernels.  They are
kept as the public, tensor-level API of the reference."""
from __future__ import annotations

from math import log

import torch as th


def classification_rewards(step_preds: th.Tensor, targets: th.Tensor) -> th.Tensor:
    """[Ns,Na,Nb,Nc], [Nb] -> [Ns,Na,Nb]: (ln Nc - CE) / ln Nc (functions.py:7-32)."""
    nb_class = step_preds.size(-1)
    logp = th.log_softmax(step_preds, dim=-1)
    picked = logp.gather(-1, targets.view(1, 1, -1, 1).expand(*step_preds.shape[:3], 1)).squeeze(-1)
    return (log(nb_class) + picked) / log(nb_class)


def discounted_returns(rewards: th.Tensor, gamma: float) -> th.Tensor:
    """G_t = r_t + gamma G_{t+1} along dim 0 (functions.py:35-51)."""
    out = th.empty_like(rewards)
    acc = th.zeros_like(rewards[0])
    for t in range(rewards.size(0) - 1, -1, -1):
        acc = rewards[t] + gamma * acc
        out[t] = acc
    return out


def standardize(values: th.Tensor, eps: float = 1e-8) -> th.Tensor:
    """(x - mean) / (unbiased std + eps) over ALL elements (functions.py:54-55)."""
    return (values - values.mean()) / (values.std() + eps)
