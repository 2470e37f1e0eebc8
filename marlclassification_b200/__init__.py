"""marlclassification_b200: the MARLClassification episode rollout + actor-critic
update on B200 (hand-written sm_100a CUDA behind the reference's Python API)."""
from . import _lib  # noqa: F401
from .core import Environment, EpisodeSampler, MultiAgent  # noqa: F401
from .networks import ModelsWrapper  # noqa: F401

__all__ = ["Environment", "EpisodeSampler", "MultiAgent", "ModelsWrapper"]
