from marlclassification_b200.train import train_main  # noqa: F401
