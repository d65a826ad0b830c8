from .bert import Bert, bert_base, bert_large, bert_pretrain_loss  # noqa: F401
from .inception import InceptionV3  # noqa: F401
from .resnet import ResNetV2, resnet, to_b200_layout  # noqa: F401
from .small import AlexNet, Cifar10CNN, LeNet, MLP  # noqa: F401
from .vgg import VGG  # noqa: F401


def get_model(name: str, num_classes: int = 1000, image_size: int = 224, num_layers: int | None = None, **kw):
    """Model zoo lookup used by the examples: `resnet50_v2`, `resnet` + num_layers, `vgg11`, `vgg16_bn`, `alexnet`,
    `inception-v3`, `cifar10_cnn`, `lenet`, `mlp`, `bert-large`, `bert-base`."""
    n = name.lower().replace("_", "-")
    if n.startswith("resnet"):
        digits = "".join(c for c in n.split("-")[0][6:] if c.isdigit())
        layers = int(digits) if digits else (num_layers or 50)
        return resnet(layers, num_classes, image_size)
    if n.startswith("vgg"):
        digits = "".join(c for c in n[3:5] if c.isdigit())
        return VGG(int(digits) if digits else (num_layers or 11), num_classes, batch_norm="bn" in n, image_size=image_size)
    if n == "alexnet":
        return AlexNet(num_classes, image_size)
    if n.startswith("inception"):
        return InceptionV3(num_classes)
    if n in ("cifar10-cnn", "cifar10"):
        return Cifar10CNN(num_classes, image_size)
    if n == "lenet":
        return LeNet(num_classes, image_size, kw.get("in_ch", 1))
    if n == "mlp":
        return MLP(num_classes, kw.get("in_features", 784))
    if n == "bert-large":
        return bert_large(**kw)
    if n == "bert-base":
        return bert_base(**kw)
    raise ValueError(f"unknown model {name!r}")
