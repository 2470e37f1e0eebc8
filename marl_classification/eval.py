from marlclassification_b200.eval import eval_main  # noqa: F401
