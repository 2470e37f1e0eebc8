from marlclassification_b200.core import *  # noqa: F401,F403
from marlclassification_b200.core import AgentOutput, Environment, EpisodeDetailedOutput, EpisodeOutput, EpisodeSampler, MultiAgent  # noqa: F401
