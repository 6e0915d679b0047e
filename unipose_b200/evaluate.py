"""GPU evaluation path — drop-in for utils/evaluate.py of the reference (same function names, argument
meaning and return types: numpy arrays) plus utils/utils.py:get_kpts.

The heat-maps stay on the device: `get_max_preds` is one arg-max kernel per call (first-occurrence tie-break,
identical integer indices to numpy), `calc_dists` / `dist_acc` run on the [N, K, 2] predictions in float64 like
the reference; only O(N*K) numbers ever cross PCIe.  Inputs may be CUDA tensors (preferred) or numpy arrays
(uploaded, for call-compatibility with `accuracy(heat.detach().cpu().numpy(), ...)`, unipose.py:161).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, ops

# (a0, a1, b0, b1, mul): length = mul * || (t[a0]+t[a1])/2 - (t[b0]+t[b1])/2 ||  (utils/evaluate.py:93-110)
_HEAD = {
    "LSP": (14, 14, 13, 13, 1.0), "COCO": (4, 4, 5, 5, 1.0), "Penn_Action": (0, 0, 1, 2, 1.0),
    "NTID": (4, 4, 3, 3, 2.0), "PoseTrack": (1, 1, 2, 2, 2.0), "BBC": (1, 1, 6, 7, 1.0), "MPII": (9, 9, 10, 10, 1.0),
}
# torso (utils/evaluate.py:130-156); MPII and BBC have their own quirks, handled below
_TORSO = {
    "COCO": (13, 13, 12, 13, 1.0), "Penn_Action": (1, 2, 7, 8, 1.0), "NTID": (3, 3, 1, 1, 1.0),
    "PoseTrack": (12, 13, 6, 7, 1.0), "LSP": (13, 13, 3, 4, 1.0),
}


def _device_tensor(a) -> torch.Tensor:
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    ops.require_cuda(a, "heat-maps")
    return a.detach().float().contiguous()


def _max_preds_device(heat: torch.Tensor):
    n, k, h, w = heat.shape
    idx = torch.empty((n, k), dtype=torch.int32, device=heat.device)
    preds = torch.empty((n, k, 2), dtype=torch.float32, device=heat.device)
    maxvals = torch.empty((n, k), dtype=torch.float32, device=heat.device)
    with torch.cuda.device(heat.device):     # launches go to the current device's stream
        _lib.call("up_argmax2d", ops._ptr(heat), ops._ptr(idx), ops._ptr(preds), ops._ptr(maxvals), n, k, h, w,
                  ops._stream())
    return idx, preds, maxvals


def get_max_preds(batch_heatmaps):
    """utils/evaluate.py:32-54 -> (preds [N,K,2] float32, maxvals [N,K,1] float32) as numpy arrays."""
    heat = _device_tensor(batch_heatmaps)
    _, preds, maxvals = _max_preds_device(heat)
    return preds.cpu().numpy(), maxvals.cpu().numpy()[..., None]


def argmax_indices(batch_heatmaps) -> torch.Tensor:
    """Flat first-occurrence arg-max index per (n, joint): int32 CUDA tensor [N, K]."""
    return _max_preds_device(_device_tensor(batch_heatmaps))[0]


def _calc_dists_device(preds: torch.Tensor, target: torch.Tensor, norm_xy):
    n, k, _ = preds.shape
    dists = torch.empty((k, n), dtype=torch.float64, device=preds.device)
    with torch.cuda.device(preds.device):
        _lib.call("up_calc_dists", ops._ptr(preds), ops._ptr(target), ops._ptr(dists), n, k,
                  ctypes.c_double(norm_xy[0]), ctypes.c_double(norm_xy[1]), ops._stream())
    return dists


def calc_dists(preds, target, normalize):
    """utils/evaluate.py:5-19 (normalize: [N,2], identical rows as built by accuracy())."""
    p = torch.as_tensor(np.asarray(preds, dtype=np.float32)).cuda()
    t = torch.as_tensor(np.asarray(target, dtype=np.float32)).cuda()
    return _calc_dists_device(p, t, (float(normalize[0][0]), float(normalize[0][1]))).cpu().numpy()


def _dist_acc_device(dists: torch.Tensor, threshold: float) -> torch.Tensor:
    k, n = dists.shape
    acc = torch.empty((k,), dtype=torch.float64, device=dists.device)
    with torch.cuda.device(dists.device):
        _lib.call("up_dist_acc", ops._ptr(dists), ops._ptr(acc), n, k, ctypes.c_double(threshold), ops._stream())
    return acc


def dist_acc(dists, threshold=0.5):
    """utils/evaluate.py:22-29 for one joint's distances."""
    d = torch.as_tensor(np.asarray(dists, dtype=np.float64)).cuda().view(1, -1)
    return float(_dist_acc_device(d, float(threshold)).cpu()[0])


def _length(t0: np.ndarray, rule) -> np.float32:
    a0, a1, b0, b1, mul = rule
    a = (t0[a0] + t0[a1]) / 2
    b = (t0[b0] + t0[b1]) / 2
    return np.float32(mul) * np.linalg.norm(a - b)


def accuracy(output, target, thr_PCK, thr_PCKh, dataset, hm_type='gaussian', threshold=0.5):
    """utils/evaluate.py:58-172 -> (acc, PCK, PCKh, cnt, pred, visible)."""
    out_d, tgt_d = _device_tensor(output), _device_tensor(target)
    n, k, h, w = out_d.shape
    _, pred_d, _ = _max_preds_device(out_d)
    _, tgt_p_d, _ = _max_preds_device(tgt_d)
    dists = _calc_dists_device(pred_d, tgt_p_d, (h / 10.0, w / 10.0))
    # sample-0 target joints fix the head / torso scale (utils/evaluate.py:91-156): K*2 floats to the host
    t0 = tgt_p_d[0].cpu().numpy()
    if dataset == "MPII":
        torso = np.linalg.norm(t0[7, 0] - t0[8, 0])
    elif dataset == "BBC":
        neck = [(t0[6, 0] + t0[7, 0]) / 2, (t0[6, 1] + t0[7, 1]) / 2]
        torso = np.linalg.norm(3 * (t0[1, 0] - neck))
    else:
        torso = _length(t0, _TORSO[dataset])
    head = _length(t0, _HEAD[dataset])
    raw = torch.stack([_dist_acc_device(dists, float(threshold)),
                       _dist_acc_device(dists, float(thr_PCK * torso)),
                       _dist_acc_device(dists, float(thr_PCKh * head))]).cpu().numpy()
    visible = (raw[0] >= 0).astype(np.float64)
    cnt = int(visible.sum())

    def finish(v):
        res = np.where(v >= 0, v, 0.0)
        if cnt != 0:
            res[0] = v[v >= 0].sum() / cnt
        return res

    return finish(raw[0]), finish(raw[1]), finish(raw[2]), cnt, pred_d.cpu().numpy(), visible


def get_kpts(maps, img_h=368.0, img_w=368.0):
    """utils/utils.py:94-106: key-points of sample 0 (channel 0 skipped) scaled to image coordinates."""
    heat = _device_tensor(maps)[:1]
    idx = _max_preds_device(heat)[0][0].cpu().numpy()
    hh, ww = heat.shape[2], heat.shape[3]
    return [[int((i % ww) * img_w / ww), int((i // ww) * img_h / hh)] for i in idx[1:]]
