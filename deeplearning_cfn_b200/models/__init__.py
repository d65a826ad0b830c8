from .resnet import ResNetV2, resnet, to_b200_layout  # noqa: F401
