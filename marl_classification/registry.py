from marlclassification_b200.registry import DATASET_REGISTRY, DatasetSpec, default_image_pipeline, get_dataset_spec  # noqa: F401
