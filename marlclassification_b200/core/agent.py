"""MultiAgent: the team's networks, what they carry from step to step (both LSTM states and the
last broadcast message) and the policy draw (reference: core/agent.py:8-79)."""
from __future__ import annotations

from typing import NamedTuple, Optional, Tuple

import torch as th

from ..networks.models import ModelsWrapper, RecurrentOutput


class AgentOutput(NamedTuple):
    """What one decision step yields, all shaped ``[Na, Nb, ...]`` (agent.py:8-14)."""
    actions: th.Tensor            # int64 indices into the environment's move table
    actions_log_probs: th.Tensor  # log pi(a | state)
    predictions: th.Tensor        # class logits
    values: th.Tensor             # critic


def draw_actions(probs: th.Tensor) -> Tuple[th.Tensor, th.Tensor]:
    """Sample one move per (agent, image) from ``probs[Na, Nb, nA]`` and return it with its
    log-probability.  The draw goes through ``th.multinomial`` on the ``[Na*Nb, nA]`` matrix with the
    reference's keywords (agent.py:53-55), so tests that patch ``torch.multinomial`` to inject the
    reference's samples work on this class exactly as on the reference's."""
    na = probs.shape[0]
    picked = th.multinomial(probs.flatten(0, 1), num_samples=1, replacement=True).view(na, -1)
    chosen = probs.gather(-1, picked.unsqueeze(-1)).squeeze(-1)
    return picked, chosen.log()


class MultiAgent:
    def __init__(self, nb_agents: int, model: ModelsWrapper) -> None:
        self._team, self._net = nb_agents, model
        self._carry: Optional[Tuple[RecurrentOutput, th.Tensor]] = None  # (recurrent state, last message)

    def __len__(self) -> int:
        return self._team

    # read-only views the episode sampler and the trainer use
    model = property(lambda self: self._net)
    nb_class = property(lambda self: self._net.nb_class)
    device = property(lambda self: self._net.device)

    def reset(self, batch_size: int) -> None:
        """Fresh episode: h, c, h^, c^ ~ N(0,1), message 0 (agent.py:33-38; models.py:148-162)."""
        self._carry = (self._net.random_first_state(self._team, batch_size),
                       self._net.zero_first_message(self._team, batch_size))

    def act(self, observation: th.Tensor, norm_pos: th.Tensor) -> AgentOutput:
        """One decision step (agent.py:40-68): networks, then the policy draw."""
        assert self._carry is not None, "reset() must be called before act()"
        state, heard = self._carry
        out, state = self._net(observation, heard, norm_pos, state)
        self._carry = (state, out.messages)
        moves, log_p = draw_actions(out.actions_probabilities)
        return AgentOutput(moves, log_p, out.predictions, out.values)

    def _adopt(self, hidden: RecurrentOutput, last_msg: th.Tensor) -> None:
        """Take over the state a fused episode left behind (EpisodeSampler)."""
        self._carry = (hidden, last_msg)
