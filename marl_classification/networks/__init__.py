from marlclassification_b200.networks import ModelOutput, ModelsWrapper, RecurrentOutput, VisionCnnModule  # noqa: F401
