"""GPU parity of the evaluation path (unipose_b200.evaluate) against the oracle and the golden fixtures:
integer joint indices bit-exact, PCK / PCKh identical."""
import os

import numpy as np
import pytest
import torch

from oracle import evaluate_oracle as E

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dataset,k,hw,n", [("mpii", "MPII", 16, 48, 8), ("lsp", "LSP", 14, 32, 4),
                                                 ("penn", "Penn_Action", 13, 46, 4)])
def test_accuracy_matches_reference_fixture(name, dataset, k, hw, n):
    from unipose_b200 import evaluate as ev
    g = np.load(os.path.join(GOLDEN, "evaluate.npz"))
    gt, pred = E.synth_eval_inputs(n, k, hw)
    acc, PCK, PCKh, cnt, p, vis = ev.accuracy(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 0.2, 0.5,
                                              dataset)
    assert np.array_equal(p, g[name + "_preds"])     # bit-exact integer joint coordinates
    assert cnt == int(g[name + "_cnt"])
    np.testing.assert_allclose(acc, g[name + "_acc"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(PCK, g[name + "_PCK"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(PCKh, g[name + "_PCKh"], rtol=0, atol=1e-12)   # north star: within 0.1
    np.testing.assert_array_equal(vis, g[name + "_visible"])
    preds, maxvals = ev.get_max_preds(pred)          # numpy input path
    assert np.array_equal(preds, g[name + "_preds"])
    assert np.array_equal(maxvals, g[name + "_maxvals"])


def test_argmax_edge_cases_vs_numpy():
    from unipose_b200 import evaluate as ev
    rng = np.random.RandomState(0)
    heat = rng.randn(5, 18, 64, 64).astype(np.float32)
    heat[0, 0] = 0.0                       # all equal -> index 0, masked (max <= 0)
    heat[0, 1] = -1.0
    heat[1, 2, 63, 63] = 50.0              # last element
    heat[1, 3, 10, 5] = 7.0
    heat[1, 3, 10, 6] = 7.0                # tie -> first
    heat[2, 4, 3, 3] = np.nan              # numpy: NaN is the arg-max
    heat[2, 5].fill(np.float32(-np.inf))
    idx = ev.argmax_indices(torch.from_numpy(heat).cuda()).cpu().numpy()
    ref = heat.reshape(5, 18, -1).argmax(axis=2)
    assert np.array_equal(idx, ref.astype(np.int32))
    p, m = ev.get_max_preds(heat)
    rp, rm = E.get_max_preds(heat)
    assert np.array_equal(p, rp)
    assert np.array_equal(m[~np.isnan(rm)], rm[~np.isnan(rm)])


def test_full_size_c2_heatmaps_and_kpts():
    from unipose_b200 import evaluate as ev
    gt, pred = E.synth_eval_inputs(32, 16, 48, seed=3)
    a = ev.accuracy(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 0.2, 0.5, "MPII")
    b = E.accuracy(pred, gt, 0.2, 0.5, "MPII")
    for u, v in zip(a, b):
        np.testing.assert_allclose(np.asarray(u, np.float64), np.asarray(v, np.float64), rtol=0, atol=1e-12)
    assert ev.get_kpts(torch.from_numpy(pred[:1]).cuda(), 384.0, 384.0) == E.get_kpts(pred[:1], 384.0, 384.0)


def test_calc_dists_and_dist_acc_wrappers():
    from unipose_b200 import evaluate as ev
    gt, pred = E.synth_eval_inputs(4, 14, 32, seed=5)
    p, _ = E.get_max_preds(pred)
    t, _ = E.get_max_preds(gt)
    norm = np.ones((4, 2)) * np.array([32, 32]) / 10
    d = ev.calc_dists(p, t, norm)
    np.testing.assert_allclose(d, E.calc_dists(p, t, norm), rtol=0, atol=1e-15)
    for j in range(14):
        assert ev.dist_acc(d[j], 0.5) == E.dist_acc(d[j], 0.5)
