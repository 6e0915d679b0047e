"""build_backbone — mirrors model/modules/backbone/__init__.py:3-7 of the reference."""
from . import resnet


def build_backbone(backbone, output_stride, BatchNorm):
    if backbone != 'resnet':
        raise NotImplementedError
    return resnet.ResNet101(output_stride, BatchNorm)
