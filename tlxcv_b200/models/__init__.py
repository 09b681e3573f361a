"""``tlxcv.models`` constructors on the B200 hot path.

Same constructor names and keyword arguments as the reference, same module
paths and parameter manifests (checked against tests/golden/manifests.json,
which is minted from the reference's own files), built on ``tlxcv_b200.nn``.
"""
from .resnet import (ResNet, resnet18, resnet34, resnet50, resnet101, resnet152,  # noqa: F401
                     wide_resnet50_2, wide_resnet101_2)
from .resnext import (ResNeXt, resnext50_32x4d, resnext50_64x4d, resnext101_32x4d,  # noqa: F401
                      resnext101_64x4d, resnext152_32x4d, resnext152_64x4d)
from .mobilenetv1 import MobileNetV1  # noqa: F401
from .mobilenetv2 import MobileNetV2, mobilenet_v2  # noqa: F401
from .darknet53 import DarkNet53, darknet53  # noqa: F401
from .darknet import DarkNet  # noqa: F401
from .det_mobilenet import MobileNet  # noqa: F401
from .yolov3 import YOLOv3, YOLOv3FPN, YOLOv3Head, YoloDetBlock  # noqa: F401
from .resnet_vd import ResNet_vd  # noqa: F401
from .resnest import ResNeSt, resnest50, resnest101  # noqa: F401

# name -> constructor, keyed like tlxcv_b200.testing.RECIPES / tests/golden
REGISTRY = {
    "resnet18": resnet18, "resnet34": resnet34, "resnet50": resnet50, "resnet101": resnet101,
    "resnet152": resnet152, "wide_resnet50_2": wide_resnet50_2, "wide_resnet101_2": wide_resnet101_2,
    "resnext50_32x4d": resnext50_32x4d, "resnext50_64x4d": resnext50_64x4d,
    "resnext101_32x4d": resnext101_32x4d,
    "mobilenet_v1": MobileNetV1, "mobilenet_v2": mobilenet_v2,
    "darknet53_cls": darknet53, "darknet53_det": DarkNet,
    "mobilenet_v1_det": MobileNet, "yolov3_darknet53": YOLOv3,
    "resnest50": resnest50, "resnest101": resnest101,
    "resnet50_vd": ResNet_vd, "resnet18_vd": lambda **kw: ResNet_vd(layers=18, **kw),
}
