"""K-C / K-F parity: fused semantic attention forward and backward against the fp64 oracle
(utils/layers.py:132-164 restated), reference (per-node beta) and paper (node-mean beta) modes."""
import numpy as np
import pytest
import torch

from oracle import han_oracle as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _run(n, P, D, A, mode, seed):
    import han_b200 as hb
    rng = np.random.default_rng(seed)
    Z64 = torch.from_numpy(rng.normal(size=(n, P, D)))
    sp64 = {"w_omega": torch.from_numpy(rng.normal(size=(D, A)) * 0.3), "b_omega": torch.from_numpy(rng.normal(size=A) * 0.3),
            "u_omega": torch.from_numpy(rng.normal(size=A))}
    up = torch.from_numpy(rng.normal(size=(n, D)))
    Zo = Z64.clone().requires_grad_(True)
    spo = {k: v.clone().requires_grad_(True) for k, v in sp64.items()}
    out_o, al_o = O.SimpleAttLayer(Zo, A, spo, return_alphas=True, mode=mode)
    (out_o * up).sum().backward()

    dev = torch.device("cuda")
    Zp = Z64.float().to(dev).requires_grad_(True)
    spp = {k: v.float().to(dev).requires_grad_(True) for k, v in sp64.items()}
    out_p, al_p = hb.layers.SimpleAttLayer(Zp, A, time_major=False, return_alphas=True, params=spp, mode=mode)
    (out_p * up.float().to(dev)).sum().backward()
    torch.cuda.synchronize()
    assert_close(out_p, out_o, "out")
    assert_close(al_p, al_o, "alphas")
    assert_close(Zp.grad, Zo.grad, "dZ")
    for k in sp64:
        assert_close(spp[k].grad, spo[k].grad, "d" + k)


@pytest.mark.parametrize("mode", ["reference", "paper"])
@pytest.mark.parametrize("n,P,D,A", [(300, 2, 64, 128), (257, 3, 64, 128), (1000, 4, 64, 128), (65, 1, 64, 128),
                                     (130, 2, 32, 64), (99, 5, 16, 32), (50, 3, 32, 32), (70, 2, 64, 64),
                                     (40, 2, 128, 128), (33, 7, 8, 32)])
def test_semantic_fwd_bwd_parity(mode, n, P, D, A):
    _run(n, P, D, A, mode, seed=n + 7 * P + D + A)


def test_p1_beta_is_exactly_one():
    import han_b200 as hb
    dev = torch.device("cuda")
    Z = torch.randn(100, 1, 64, device=dev)
    sp = {"w_omega": torch.randn(64, 128, device=dev) * 0.1, "b_omega": torch.randn(128, device=dev) * 0.1,
          "u_omega": torch.randn(128, device=dev) * 0.1}
    out, al = hb.layers.SimpleAttLayer(Z, 128, return_alphas=True, params=sp)
    assert torch.equal(al, torch.ones_like(al)) and torch.equal(out, Z[:, 0])


def test_time_major_and_no_alphas():
    import han_b200 as hb
    dev = torch.device("cuda")
    Z = torch.randn(3, 40, 64, device=dev)   # (P,N,D) time-major
    sp = {"w_omega": torch.randn(64, 128, device=dev) * 0.1, "b_omega": torch.randn(128, device=dev) * 0.1,
          "u_omega": torch.randn(128, device=dev) * 0.1}
    a = hb.layers.SimpleAttLayer(Z, 128, time_major=True, params=sp)
    b, _ = hb.layers.SimpleAttLayer(Z.transpose(0, 1).contiguous(), 128, return_alphas=True, params=sp)
    assert torch.equal(a, b)


@pytest.mark.parametrize("tc", [True, False], ids=["tcgen05", "mma_sync"])
@pytest.mark.parametrize("mode", ["reference", "paper"])
@pytest.mark.parametrize("n,P", [(300, 2), (257, 3), (40000, 4), (65, 1), (5000, 5)])
def test_semantic_64x128_on_both_kernel_families(monkeypatch, tc, mode, n, P):
    """(D, A) = (64, 128) runs on the tcgen05 kernels of semantic_tc.cu by default (persistent CTAs, TMA ring,
    TMEM accumulators; forward AND backward) and on the mma.sync kernels of semantic.cu with HAN_SEM_TC=0.
    Several tiles per CTA (n*P / 128 > 148), partial last tiles, P that does not divide 128."""
    from han_b200 import ops
    monkeypatch.setattr(ops, "SEM_TC", tc)
    _run(n, P, 64, 128, mode, seed=n + 11 * P)


@pytest.mark.parametrize("eg", [1, 2, 4])
def test_semantic_forward_epilogue_groups(monkeypatch, eg):
    """HAN_SEM_TC_EG: 1 / 2 / 4 epilogue warp groups of the tcgen05 forward split the 128 columns of a tile."""
    from han_b200 import ops
    monkeypatch.setattr(ops, "SEM_TC", True)
    monkeypatch.setattr(ops, "SEM_TC_EG", eg)
    _run(20000, 4, 64, 128, "reference", seed=500 + eg)

