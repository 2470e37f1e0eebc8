"""Model configuration with the reference's ``marl.json`` schema (config.py:18-104)
so reference run directories load unchanged, plus the CLI option groups (config.py:11-15,
107-131)."""
from __future__ import annotations

import json
from os.path import exists, isfile

from pydantic import BaseModel, create_model

from .core import Environment, MultiAgent
from .networks import ModelsWrapper
from .registry import get_dataset_spec

_FIELDS = (
    "ft_extr_str", "window_size", "hidden_size_belief", "hidden_size_action", "hidden_size_msg",
    "hidden_size_msg_output", "hidden_size_state", "state_dim", "actions", "nb_class",
    "hidden_size_linear_belief", "hidden_size_linear_action",
)


def _option_group(name: str, **fields: type) -> type[BaseModel]:
    """A validated bag of required CLI options (the reference spells each one out as a pydantic class)."""
    return create_model(name, **{key: (kind, ...) for key, kind in fields.items()})


# config.py:11-15 -- options shared by every mode
MainConfig = _option_group("MainConfig", step=int, run_id=str, cuda=bool, nb_agent=int)
# config.py:107-131 -- one group per mode
TrainConfig = _option_group("TrainConfig", img_size=int, nb_epoch=int, learning_rate=float, batch_size=int,
                            resources_dir=str, output_dir=str, gamma=float)
EvalConfig = _option_group("EvalConfig", img_size=int, state_dict_path=str, batch_size=int, json_path=str,
                           dataset_path=str, output_dir=str)
InferConfig = _option_group("InferConfig", state_dict_path=str, json_path=str, images_path=list[str], output_dir=str,
                            class_to_idx=str)


class ModelConfig(BaseModel):
    ft_extr_str: str
    window_size: int
    hidden_size_belief: int
    hidden_size_action: int
    hidden_size_msg: int
    hidden_size_msg_output: int
    hidden_size_state: int
    state_dim: int
    actions: list[list[int]]
    nb_class: int
    hidden_size_linear_belief: int
    hidden_size_linear_action: int

    def save_marl_config(self, out_json_path: str) -> None:
        with open(out_json_path, "w", encoding="utf-8") as fh:
            json.dump({k: getattr(self, k) for k in _FIELDS}, fh)

    @classmethod
    def load_marl_config(cls, json_path: str) -> "ModelConfig":
        assert exists(json_path) and isfile(json_path), f'"{json_path}" does not exist or is not a file'
        with open(json_path, "r", encoding="utf-8") as fh:
            raw = json.load(fh)
        return cls(**{k: raw[k] for k in _FIELDS})

    def build_networks(self) -> ModelsWrapper:
        spec = get_dataset_spec(self.ft_extr_str)
        return ModelsWrapper(
            spec.cnn_constructor(self.window_size), self.hidden_size_belief, self.hidden_size_action,
            self.hidden_size_msg, self.hidden_size_msg_output, self.hidden_size_state, self.state_dim,
            len(self.actions), self.nb_class, self.hidden_size_linear_belief, self.hidden_size_linear_action,
        )

    def build_environment(self) -> Environment:
        return Environment(self.actions, self.window_size)

    def build_marl(self, nb_agents: int) -> tuple[ModelsWrapper, MultiAgent, Environment]:
        networks = self.build_networks()
        return networks, MultiAgent(nb_agents, networks), self.build_environment()


    @classmethod
    def load_trained(cls, json_path: str, state_dict_path: str, nb_agents: int, device) -> tuple:
        """``marl.json`` + a saved ``state_dict`` -> (config, networks in eval mode on ``device``, agents,
        environment): the common opening of the ``test`` and ``infer`` modes (eval.py:44-61, infer.py:44-57)."""
        import torch as th

        for path, what in ((json_path, "JSON path"), (state_dict_path, "State dict path")):
            assert exists(path), f'{what} "{path}" does not exist'
            assert isfile(path), f'"{path}" is not a file'
        config = cls.load_marl_config(json_path)
        networks, agents, env = config.build_marl(nb_agents)
        networks.load_state_dict(th.load(state_dict_path, map_location="cpu"))
        networks.eval()
        networks.to(device)
        return config, networks, agents, env
