"""mAP evaluation: mAP1 / mAP of the reference (Vision.py:1696-1800).
CPU: the numpy oracle vs the reference-generated golden; GPU: the matching kernel and the integration kernel
(rn_map_match + rn_map_ap through metrics.mAP) vs both.  The table is float64 arithmetic over integer counts, so it is compared
exactly (nan where a category has no ground truth, as the reference returns)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THRESHOLDS = {"coco": [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95], "pascal": [0.5], "odd": [0.3, 0.62]}


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.map_cases()


def _oracle_flags(predictions, targets, thresholds, device=None):
    """match_flags with the matching done by the oracle's float32 IoU (CPU stand-in for rn_map_match in the host test)."""
    pb = [b for p in predictions for b in p[0]]
    pc = np.array([int(c) for p in predictions for c in p[1]], np.int32)
    ps = np.array([s for p in predictions for s in p[2]], np.float32)
    tc = np.array([int(c) for t in targets for _, c in t], np.int32)
    flags = np.zeros((len(thresholds), len(pb)), np.uint8)
    off = 0
    for p, t in zip(predictions, targets):
        n = len(p[0])
        for b, c in t:
            sel = [j for j in range(n) if int(p[1][j]) == c]
            if sel:
                jac = orc.jaccard_f32(np.array([b]), np.array([p[0][j] for j in sel]))[0]
                k = int(np.argmax(jac))
                for ti, th in enumerate(thresholds):
                    if jac[k] > np.float32(th):
                        flags[ti, off + sel[k]] = 1
        off += n
    return flags, pc, ps, tc


def test_oracle_matches_reference_map(golden_dir):
    g = np.load(os.path.join(golden_dir, "map_scores.npz"))
    for k, (predictions, targets, categories) in enumerate(_cases()):
        for name, th in THRESHOLDS.items():
            table = orc.map_table(predictions, targets, len(categories), th)
            assert np.array_equal(table, g["case%d_%s_table" % (k, name)], equal_nan=True)
            assert np.array_equal(np.mean(table), g["case%d_%s_mean" % (k, name)], equal_nan=True)


def test_no_host_integration_left():
    """The precision/recall integration lives in rn_map_ap (device); the package holds no NumPy restatement of it."""
    from neuralnetworklibrary_b200 import metrics
    assert not hasattr(metrics, "average_precision")
    assert metrics.COCO_thresholds == THRESHOLDS["coco"] and metrics.Pascal_thresholds == THRESHOLDS["pascal"]


def test_map_requires_the_cuda_library():
    """No CPU fallback: without a device the metric raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from neuralnetworklibrary_b200 import metrics
    predictions, targets, categories = _cases()[0]
    with pytest.raises(Exception):
        metrics.mAP(predictions, targets, categories, verbose=False)


@pytest.mark.gpu
def test_kernel_matches_reference_map(golden_dir, capsys):
    from neuralnetworklibrary_b200 import metrics
    g = np.load(os.path.join(golden_dir, "map_scores.npz"))
    for k, (predictions, targets, categories) in enumerate(_cases()):
        for name, th in THRESHOLDS.items():
            flags, pc, ps, tc = metrics.match_flags(predictions, targets, th)
            oflags = _oracle_flags(predictions, targets, th)[0]
            assert np.array_equal(flags, oflags), "is_correct flags differ from the oracle"
            table = metrics.mAP_table(predictions, targets, len(categories), th)
            assert np.array_equal(table, g["case%d_%s_table" % (k, name)], equal_nan=True)
            mean = metrics.mAP(predictions, targets, categories, th)
            assert np.array_equal(mean, g["case%d_%s_mean" % (k, name)], equal_nan=True)
    assert "Overall mAP" in capsys.readouterr().out


@pytest.mark.gpu
def test_map_large_validation_set_vs_oracle():
    """500 images x 80 categories x 10 thresholds: the reference's triple Python loop takes minutes here; the oracle's
    table is the checker."""
    rng = np.random.RandomState(5)
    N, C = 500, 80
    predictions, targets = [], []
    for i in range(N):
        nt = int(rng.randint(0, 8))
        xy = rng.uniform(0, 600, size=(nt, 2)); wh = rng.uniform(20, 300, size=(nt, 2))
        tb = np.concatenate([xy, xy + wh], 1); tc = rng.randint(0, C, size=nt)
        targets.append([(tb[j], int(tc[j])) for j in range(nt)])
        pb, pc, ps = [], [], []
        for j in range(nt):
            for _ in range(int(rng.randint(0, 4))):
                pb.append((tb[j] + rng.normal(0, 10, size=4)).astype(np.float32)); pc.append(np.int64(tc[j]))
                ps.append(np.float32(rng.uniform(0.05, 1)))
        predictions.append([pb, pc, ps])
    from neuralnetworklibrary_b200 import metrics
    th = metrics.COCO_thresholds
    assert np.array_equal(metrics.mAP_table(predictions, targets, C, th), orc.map_table(predictions, targets, C, th),
                          equal_nan=True)


@pytest.mark.gpu
def test_map_empty_inputs():
    from neuralnetworklibrary_b200 import metrics
    cats = {0: "a", 1: "b"}
    t = metrics.mAP_table([[[], [], []]], [[(np.array([0., 0., 10., 10.]), 0)]], 2, [0.5])
    assert t[0, 0] == 0.0 and np.isnan(t[0, 1])
    t = metrics.mAP_table([[[np.array([0, 0, 10, 10], np.float32)], [np.int64(0)], [np.float32(0.9)]]], [[]], 2, [0.5])
    assert np.isnan(t).all()
    assert np.isnan(metrics.mAP([], [], cats, [0.5], verbose=False))


@pytest.mark.gpu
def test_map_many_predictions_per_category_and_score_ties():
    """Two categories with ~6 000 predictions each: the rank-by-counting sort runs over several shared-memory tiles, the sum over
    the correct positions goes through NumPy's pairwise recursion (n > 128), and quantised scores produce long runs of equal
    scores whose order is decided by is_correct (sorted(zip(Scores, IsCorrect), reverse=True), Vision.py:1730)."""
    from neuralnetworklibrary_b200 import metrics
    rng = np.random.RandomState(9)
    N, C = 1500, 2
    predictions, targets = [], []
    for i in range(N):
        nt = int(rng.randint(1, 6))
        xy = rng.uniform(0, 400, size=(nt, 2)); wh = rng.uniform(20, 200, size=(nt, 2))
        tb = np.concatenate([xy, xy + wh], 1); tcs = rng.randint(0, C, size=nt)
        targets.append([(tb[j], int(tcs[j])) for j in range(nt)])
        pb, pc, ps = [], [], []
        for j in range(nt):
            for _ in range(int(rng.randint(1, 5))):
                pb.append((tb[j] + rng.normal(0, 12, size=4)).astype(np.float32)); pc.append(np.int64(tcs[j]))
                ps.append(np.float32(np.round(rng.uniform(0.05, 1), 2)))       # 96 distinct scores for ~12 000 predictions
        predictions.append([pb, pc, ps])
    th = [0.5, 0.75, 0.9]
    got = metrics.mAP_table(predictions, targets, C, th)
    want = orc.map_table(predictions, targets, C, th)
    assert got.shape == (3, 2) and np.array_equal(got, want, equal_nan=True)
    assert min(sum(1 for p in predictions for c in p[1] if int(c) == k) for k in range(C)) > 4096
