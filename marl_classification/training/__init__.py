from marlclassification_b200.training import MetricLogger, Trainer, classification_rewards, discounted_returns, standardize  # noqa: F401
