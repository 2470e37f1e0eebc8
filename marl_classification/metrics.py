from marlclassification_b200.metrics import ConfusionMeter, LossMeter, format_metric  # noqa: F401
