"""Drop-in alias: the reference's import paths (``marl_classification.core`` ...) served by
the B200 implementation in ``marlclassification_b200``."""
