from marlclassification_b200.config import EvalConfig, InferConfig, MainConfig, ModelConfig, TrainConfig  # noqa: F401
