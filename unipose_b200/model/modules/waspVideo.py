"""WASP module of the video model: mirrors model/modules/waspVideo.py — identical to wasp.py except that the
global-average-pool branch has no BatchNorm (waspVideo.py:56-59; the video model trains at batch 1) and the
dead drn / mobilenet inplanes switch (waspVideo.py:36-41)."""
from .wasp import wasp as _wasp_image


class wasp(_wasp_image):
    _gap_has_bn = False

    @staticmethod
    def _inplanes(backbone):
        return {'drn': 512, 'mobilenet': 320}.get(backbone, 2048)


def build_wasp(backbone, output_stride, BatchNorm):
    return wasp(backbone, output_stride, BatchNorm)
