from marlclassification_b200.config import ModelConfig  # noqa: F401
