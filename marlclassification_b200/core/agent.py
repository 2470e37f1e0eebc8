"""MultiAgent: networks + recurrent state + policy sampling (reference: core/agent.py)."""
from __future__ import annotations

from dataclasses import dataclass

import torch as th

from ..networks.models import ModelsWrapper, RecurrentOutput


@dataclass
class AgentOutput:
    actions: th.Tensor  # [Na, Nb] indices into the environment's action set
    actions_log_probs: th.Tensor
    predictions: th.Tensor
    values: th.Tensor


class MultiAgent:
    def __init__(self, nb_agents: int, model: ModelsWrapper) -> None:
        self.__nb_agents = nb_agents
        self.__model = model
        self.__hidden: RecurrentOutput | None = None
        self.__last_msg: th.Tensor | None = None

    def reset(self, batch_size: int) -> None:
        """h, c, h^, c^ ~ N(0,1) and a zero message (agent.py:33-38)."""
        self.__hidden = self.__model.random_first_state(len(self), batch_size)
        self.__last_msg = self.__model.zero_first_message(len(self), batch_size)

    def act(self, observation: th.Tensor, norm_pos: th.Tensor) -> AgentOutput:
        """One decision step (agent.py:40-68): networks -> ``th.multinomial`` -> log p[a]."""
        assert self.__hidden is not None, "reset() must be called before act()"
        output, hidden = self.__model(observation, self.__last_msg, norm_pos, self.__hidden)
        self.__hidden, self.__last_msg = hidden, output.messages
        probs = output.actions_probabilities
        # same call form as the reference so RNG injection by patching torch.multinomial works
        action_indices = th.multinomial(probs.flatten(0, 1), num_samples=1, replacement=True).view(self.__nb_agents, -1)
        log_probs = th.gather(probs, -1, action_indices.unsqueeze(-1)).squeeze(-1).log()
        return AgentOutput(actions=action_indices, actions_log_probs=log_probs, predictions=output.predictions,
                           values=output.values)

    @property
    def model(self) -> ModelsWrapper:
        return self.__model

    @property
    def nb_class(self) -> int:
        return self.__model.nb_class

    @property
    def device(self) -> th.device:
        return self.__model.device

    def __len__(self) -> int:
        return self.__nb_agents

    def _adopt(self, hidden: RecurrentOutput, last_msg: th.Tensor) -> None:
        self.__hidden, self.__last_msg = hidden, last_msg
