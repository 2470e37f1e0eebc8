from marlclassification_b200.networks.vision import *  # noqa: F401,F403
from marlclassification_b200.networks.vision import AidCnn, KneeMriCnn, MnistCnn, Resisc45Cnn, SkinCancerCnn, VisionCnnModule, WorldStratCnn  # noqa: F401
