from .bert import Bert, bert_base, bert_large, bert_pretrain_loss  # noqa: F401
from .inception import InceptionV3  # noqa: F401
from .resnet import ResNetV2, resnet, to_b200_layout  # noqa: F401
from .small import AlexNet, Cifar10CNN, LeNet, MLP  # noqa: F401
from .vgg import VGG  # noqa: F401
from .zoo import (DenseNet, GoogLeNet, InceptionBN, InceptionResNetV2, InceptionV4, MobileNet, MobileNetV2,  # noqa: F401
                  ResNetV1, SqueezeNet)


def get_model(name: str, num_classes: int = 1000, image_size: int = 224, num_layers: int | None = None, **kw):
    """Model zoo lookup used by the examples.  Symbol-file names of `train_imagenet.py --network` (resnet, resnet-v1,
    resnext, vgg, alexnet, googlenet, inception-bn, inception-v3, inception-v4, inception-resnet-v2, mobilenet,
    mobilenetv2, lenet, mlp; depth from `num_layers`) and gluon model-zoo names of `image_classification.py --model`
    (resnet50_v1, resnet50_v2, vgg16_bn, squeezenet1.0, densenet121, mobilenet0.5, mobilenetv2_1.0, inceptionv3, ...),
    plus `cifar10_cnn`, `bert-large`, `bert-base`."""
    n = name.lower().replace("_", "-")
    digits_of = lambda t: "".join(c for c in t if c.isdigit())     # noqa: E731
    if n.startswith("resnext"):
        return ResNetV1(int(digits_of(n.split("-")[0]) or num_layers or 50), num_classes, image_size, groups=kw.get("num_group", 32))
    if n.startswith("resnet") and ("v1" in n.split("-")[1:] or n.startswith("resnetv1")):
        head = n.split("-")[0].replace("resnetv1", "")
        return ResNetV1(int(digits_of(head) or num_layers or 50), num_classes, image_size)
    if n.startswith("mobilenetv2"):
        mult = n[len("mobilenetv2"):].strip("-")
        return MobileNetV2(num_classes, float(mult) if mult else kw.get("multiplier", 1.0))
    if n.startswith("mobilenet"):
        mult = n[len("mobilenet"):].strip("-")
        return MobileNet(num_classes, float(mult) if mult else kw.get("alpha", 1.0))
    if n == "googlenet":
        return GoogLeNet(num_classes)
    if n in ("inception-bn", "inceptionbn"):
        return InceptionBN(num_classes)
    if n in ("inception-v4", "inceptionv4"):
        return InceptionV4(num_classes)
    if n in ("inception-resnet-v2", "inceptionresnetv2"):
        return InceptionResNetV2(num_classes)
    if n.startswith("squeezenet"):
        return SqueezeNet(n[len("squeezenet"):].strip("-") or "1.0", num_classes)
    if n.startswith("densenet"):
        return DenseNet(int(digits_of(n) or num_layers or 121), num_classes)
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
