from marlclassification_b200.registry import DATASET_REGISTRY, DatasetSpec, get_dataset_spec  # noqa: F401
