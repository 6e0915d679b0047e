"""GPU input / label pipeline - the steps either side of the hot path.

* `gaussian_heatmaps` / `center_map`: the ground-truth synthesis of the reference's datasets
  (utils/mpii_data.py:165-181, same code in lsp_lspet_data.py / bbc_data.py / penn_action_data.py) on the device,
  bit-for-bit (float64 Gaussian, clipping, fp32 store, background channel).
* uint8 images go straight into the network through `unipose.forward_uint8` (model/unipose.py): normalisation
  `(x - 128) / 256` (mpii_data.py:184-185) is fused into the stem's input packing.
"""
from __future__ import annotations

import torch

from . import _lib, ops


def _kpts(t: torch.Tensor) -> torch.Tensor:
    ops.require_cuda(t, "key-points")
    t = t.detach().float().contiguous()
    if t.dim() != 3 or t.shape[2] < 2:
        raise ValueError("key-points must be [n, k, 2(+)] (x, y in input-image pixels); got %s" % (tuple(t.shape),))
    return t[:, :, :2].contiguous()


def gaussian_heatmaps(kpts: torch.Tensor, height: int, width: int, stride: int = 8, sigma: float = 3.0,
                      background: bool = True) -> torch.Tensor:
    """[n, k(+1), height/stride, width/stride] fp32 heat-maps for key-points given in input-image pixels
    (height x width = the network input, 368 x 368 in the reference)."""
    k = _kpts(kpts)
    n, nk = k.shape[:2]
    h, w = int(height / stride), int(width / stride)
    out = torch.empty((n, nk + (1 if background else 0), h, w), dtype=torch.float32, device=k.device)
    with torch.cuda.device(k.device):
        _lib.call("up_gaussian_labels", ops._ptr(k), ops._ptr(out), n, nk, h, w, float(stride), float(sigma),
                  1 if background else 0, 1, ops._stream())
    return out


def center_map(center: torch.Tensor, height: int, width: int, stride: int = 8, sigma: float = 3.0) -> torch.Tensor:
    """[n, 1, height/stride, width/stride] centre map (mpii_data.py:177-181: Gaussian at int(center / stride))."""
    c = center.detach().float().reshape(center.shape[0], 1, -1)
    k = _kpts(c)
    n = k.shape[0]
    h, w = int(height / stride), int(width / stride)
    out = torch.empty((n, 1, h, w), dtype=torch.float32, device=k.device)
    with torch.cuda.device(k.device):
        _lib.call("up_gaussian_labels", ops._ptr(k), ops._ptr(out), n, 1, h, w, float(stride), float(sigma), 0, 2,
                  ops._stream())
    return out
