from marlclassification_b200.infer import infer_main  # noqa: F401
