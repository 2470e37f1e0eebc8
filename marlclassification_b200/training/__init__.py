from .functions import classification_rewards, discounted_returns, standardize
from .trainer import MetricLogger, Trainer
