"""The classifier (models/gat.py:66-72, tf.layers.dense) and the masked cross-entropy (models/base_gattn.py:41-48) on
their own kernels (han_b200/csrc/dense_ce.cu) against the fp64 oracle restatement: outputs and every gradient."""
import numpy as np
import pytest
import torch

from oracle import han_oracle as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,D,C", [(1000, 64, 3), (777, 16, 7), (513, 64, 349), (64, 64, 8), (3025, 64, 4), (130, 32, 33)])
def test_dense_and_masked_ce_match_oracle(n, D, C):
    from han_b200 import ops
    import han_b200 as hb
    rng = np.random.default_rng(n + D + C)
    X = torch.from_numpy(rng.normal(size=(n, D)))
    W = torch.from_numpy(rng.normal(size=(D, C)) * 0.3)
    b = torch.from_numpy(rng.normal(size=C) * 0.1)
    y = rng.integers(0, C, size=n)
    labels = torch.zeros(n, C, dtype=torch.float64)
    labels[torch.arange(n), torch.from_numpy(y)] = 1.0
    mask = torch.from_numpy((rng.random(n) < 0.3).astype(np.float64))
    Xo, Wo, bo = (t.clone().requires_grad_(True) for t in (X, W, b))
    lo = Xo @ Wo + bo
    ce_o = O.masked_softmax_cross_entropy(lo, labels, mask)
    (ce_o * 1.7).backward()
    dev = torch.device("cuda")
    Xp, Wp, bp = (t.float().to(dev).requires_grad_(True) for t in (X, W, b))
    lp = ops.dense(Xp, Wp, bp)
    ce_p = hb.BaseGAttN.masked_softmax_cross_entropy(lp, labels.float().to(dev), mask.float().to(dev))
    (ce_p * 1.7).backward()
    assert_close(lp, lo, "logits")
    assert_close(ce_p, ce_o, "loss")
    assert_close(Xp.grad, Xo.grad, "dX")
    assert_close(Wp.grad, Wo.grad, "dW")
    assert_close(bp.grad, bo.grad, "db")


def test_masked_ce_with_soft_labels_and_no_grad():
    """tf.nn.softmax_cross_entropy_with_logits takes any label distribution; evaluation runs without dlogits."""
    import han_b200 as hb
    rng = np.random.default_rng(5)
    n, C = 300, 5
    logits = torch.from_numpy(rng.normal(size=(n, C)) * 3)
    labels = torch.from_numpy(rng.random((n, C)))
    mask = torch.from_numpy((rng.random(n) < 0.5).astype(np.float64))
    ref = O.masked_softmax_cross_entropy(logits, labels, mask)
    with torch.no_grad():
        got = hb.BaseGAttN.masked_softmax_cross_entropy(logits.float().cuda(), labels.float().cuda(), mask.float().cuda())
    assert_close(got, ref, "loss")
