"""CPU oracle for the evaluation path (utils/evaluate.py, utils/utils.py:get_kpts) — TEST INFRASTRUCTURE ONLY.

Vectorised numpy restatement; pinned against the reference functions themselves (loaded by file path
from /root/reference/utils/evaluate.py, pure numpy) in tests/test_oracle.py and through the golden
fixtures written by oracle/make_golden.py.
"""
from __future__ import annotations

import numpy as np

# joint pairs the reference hard-codes per dataset (utils/evaluate.py:93-110 head, :130-156 torso)
HEAD_RULES = {
    "LSP": ("pair", 14, 13, 1.0),
    "COCO": ("pair", 4, 5, 1.0),
    "Penn_Action": ("mid", 0, (1, 2), 1.0),
    "NTID": ("pair", 4, 3, 2.0),
    "PoseTrack": ("pair", 1, 2, 2.0),
    "BBC": ("mid", 1, (6, 7), 1.0),
    "MPII": ("pair", 9, 10, 1.0),
}


def get_max_preds(heat: np.ndarray):
    """utils/evaluate.py:32-54: first-occurrence flat argmax per (n, joint); (x, y) zeroed where max <= 0."""
    n, k, h, w = heat.shape
    flat = heat.reshape(n, k, h * w)
    idx = flat.argmax(axis=2)
    maxvals = np.take_along_axis(flat, idx[..., None], axis=2)
    preds = np.empty((n, k, 2), dtype=np.float32)
    preds[..., 0] = (idx % w).astype(np.float32)
    preds[..., 1] = np.floor(idx.astype(np.float32) / w)
    preds *= (maxvals > 0.0).astype(np.float32)
    return preds, maxvals


def calc_dists(preds, target, normalize):
    """utils/evaluate.py:5-19: dists[c, n]; -1 unless target x > 1 and y > 1.  float64 like the reference
    (float32 operands divided by a float64 `normalize`)."""
    p = preds.astype(np.float32).astype(np.float64) / normalize[:, None, :]
    t = target.astype(np.float32).astype(np.float64) / normalize[:, None, :]
    d = np.sqrt(((p - t) ** 2).sum(-1))                       # [n, k]
    ok = (target[..., 0] > 1) & (target[..., 1] > 1)
    return np.where(ok, d, -1.0).T.copy()                     # [k, n]


def dist_acc(dists, threshold=0.5):
    """utils/evaluate.py:22-29."""
    valid = dists != -1
    cnt = valid.sum()
    if cnt > 0:
        return (dists[valid] < threshold).sum() * 1.0 / cnt
    return -1


def _head_length(target, dataset):
    kind, a, b, mul = HEAD_RULES[dataset]
    if kind == "pair":
        return mul * np.linalg.norm(target[0, a, :] - target[0, b, :])
    neck = [(target[0, b[0], 0] + target[0, b[1], 0]) / 2, (target[0, b[0], 1] + target[0, b[1], 1]) / 2]
    return mul * np.linalg.norm(target[0, a, :] - neck)


def _torso(target, dataset):
    """utils/evaluate.py:130-156 (kept quirk for quirk, e.g. MPII uses only the x coordinates)."""
    t = target
    if dataset == "COCO":
        pelvis = [(t[0, 12, 0] + t[0, 13, 0]) / 2, (t[0, 12, 1] + t[0, 13, 1]) / 2]
        return np.linalg.norm(t[0, 13, :] - pelvis)
    if dataset == "Penn_Action":
        return np.linalg.norm((t[0, 1, :] + t[0, 2, :]) / 2 - (t[0, 7, :] + t[0, 8, :]) / 2)
    if dataset == "NTID":
        return np.linalg.norm(t[0, 3, :] - t[0, 1, :])
    if dataset == "PoseTrack":
        return np.linalg.norm((t[0, 12, :] + t[0, 13, :]) / 2 - (t[0, 6, :] + t[0, 7, :]) / 2)
    if dataset == "BBC":
        neck = [(t[0, 6, 0] + t[0, 7, 0]) / 2, (t[0, 6, 1] + t[0, 7, 1]) / 2]
        return np.linalg.norm(3 * (t[0, 1, 0] - neck))
    if dataset == "LSP":
        pelvis = [(t[0, 3, 0] + t[0, 4, 0]) / 2, (t[0, 3, 1] + t[0, 4, 1]) / 2]
        return np.linalg.norm(t[0, 13, :] - pelvis)
    if dataset == "MPII":
        return np.linalg.norm(t[0, 7, 0] - t[0, 8, 0])
    raise KeyError(dataset)


def accuracy(output, target, thr_PCK, thr_PCKh, dataset):
    """utils/evaluate.py:58-172 -> (acc, PCK, PCKh, cnt, pred, visible)."""
    k = output.shape[1]
    pred, _ = get_max_preds(output)
    tgt, _ = get_max_preds(target)
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    dists = calc_dists(pred, tgt, norm)

    def per_joint(thr):
        vals = np.array([dist_acc(dists[i], thr) for i in range(k)], dtype=np.float64)
        return vals

    raw = per_joint(0.5)
    visible = (raw >= 0).astype(np.float64)
    cnt = int(visible.sum())

    def finish(vals):
        s = vals[vals >= 0].sum()
        out = np.where(vals >= 0, vals, 0.0)
        if cnt != 0:
            out[0] = s / cnt
        return out

    acc = finish(raw)
    PCKh = finish(per_joint(thr_PCKh * _head_length(tgt, dataset)))
    PCK = finish(per_joint(thr_PCK * _torso(tgt, dataset)))
    return acc, PCK, PCKh, cnt, pred, visible


def get_kpts(maps: np.ndarray, img_h: float = 368.0, img_w: float = 368.0):
    """utils/utils.py:94-106: per-joint argmax of sample 0 (channel 0 skipped), scaled with int() truncation."""
    out = []
    for m in maps[0][1:]:
        hh, ww = np.unravel_index(m.argmax(), m.shape)
        out.append([int(ww * img_w / m.shape[1]), int(hh * img_h / m.shape[0])])
    return out


# --------------------------------------------------------------------------------------------
# seeded synthetic inputs (shared by make_golden.py, the tests and the bench)
# --------------------------------------------------------------------------------------------
def gaussian_heatmaps(n, k, h, w, seed, sigma=3.0):
    """Synthetic GT like utils/mpii_data.py:165-181: K gaussians (values < 0.0099 clipped to 0) at
    seeded integer centres, background channel 0 = 1 - max over joints."""
    rng = np.random.RandomState(seed)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((n, k + 1, h, w), dtype=np.float32)
    for b in range(n):
        for j in range(k):
            cx, cy = rng.randint(4, w - 4), rng.randint(4, h - 4)
            g = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sigma * sigma))
            g[g < 0.0099] = 0
            out[b, j + 1] = g
        out[b, 0] = 1.0 - out[b, 1:].max(axis=0)
    return out


def synth_eval_inputs(n, k, hw, seed=11, noise=0.25):
    """(gt, pred) heat-map pair: pred = gt + N(0, noise) with an all-zero map and an exact tie planted."""
    gt = gaussian_heatmaps(n, k, hw, hw, seed=seed)
    rng = np.random.RandomState(seed + 1)
    pred = (gt + rng.normal(0, noise, gt.shape)).astype(np.float32)
    pred[0, 3] = 0.0                        # all-zero map: argmax 0, masked by maxval > 0
    pred[1, 5, 7, 9] = pred[1, 5].max() + 1.0
    pred[1, 5, 20, 2] = pred[1, 5, 7, 9]    # tie: the first occurrence (row-major) must win
    return gt, pred


# --------------------------------------------------------------------------------------------
# ground-truth label synthesis of the datasets (utils/mpii_data.py:62-65,165-181; identical code in
# lsp_lspet_data.py:65-68, bbc_data.py:17-20)
# --------------------------------------------------------------------------------------------
def guassian_kernel(size_w, size_h, center_x, center_y, sigma):
    """mpii_data.py:62-65 (the reference's spelling)."""
    gridy, gridx = np.mgrid[0:size_h, 0:size_w]
    D2 = (gridx - center_x) ** 2 + (gridy - center_y) ** 2
    return np.exp(-D2 / 2.0 / sigma / sigma)


def reference_labels(kpt, center, height, width, stride, sigma):
    """mpii_data.py:165-181 for ONE sample: kpt [K, 2+] float32 (x, y in input pixels), center [2] float32 ->
    (heatmap [K+1, h, w] fp32 with the background channel first, centermap [1, h, w] fp32)."""
    h, w = int(height / stride), int(width / stride)
    heatmap = np.zeros((h, w, len(kpt) + 1), dtype=np.float32)
    for i in range(len(kpt)):
        x = int(kpt[i][0]) * 1.0 / stride
        y = int(kpt[i][1]) * 1.0 / stride
        heat_map = guassian_kernel(size_h=h, size_w=w, center_x=x, center_y=y, sigma=sigma)
        heat_map[heat_map > 1] = 1
        heat_map[heat_map < 0.0099] = 0
        heatmap[:, :, i + 1] = heat_map
    heatmap[:, :, 0] = 1.0 - np.max(heatmap[:, :, 1:], axis=2)
    cm = guassian_kernel(size_h=h, size_w=w, center_x=int(np.float32(center[0]) / np.float32(stride)),
                         center_y=int(np.float32(center[1]) / np.float32(stride)), sigma=3)
    cm[cm > 1] = 1
    cm[cm < 0.0099] = 0
    return heatmap.transpose(2, 0, 1).copy(), cm.astype(np.float32)[None]


def synth_keypoints(n, k, height, width, seed):
    """Seeded key-points / centres in input pixels, with a few on the border and fractional coordinates."""
    rng = np.random.RandomState(seed)
    kpts = (rng.rand(n, k, 2) * [width - 1, height - 1]).astype(np.float32)
    kpts[0, 0] = [0.0, 0.0]
    kpts[0, 1] = [width - 1, height - 1]
    kpts[1 % n, 2] = [7.99, 8.01]
    center = (rng.rand(n, 2) * [width - 1, height - 1]).astype(np.float32)
    return kpts, center
