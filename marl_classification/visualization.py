from marlclassification_b200.visualization import visualize_steps  # noqa: F401
