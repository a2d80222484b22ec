"""Memory-safety evidence without compute-sanitizer (it is closed on this GPU pool: "runs under it have left GPUs needing
a reset"): every buffer the ops allocate is embedded between two guard regions filled with a sentinel byte, and its
interior is pre-filled with NaN.  After a complete forward + backward
  * every guard byte must be intact (no kernel wrote outside a buffer it was given), and
  * the results must still match the fp64 oracle (no kernel consumed a location that nothing had written: a NaN would
    have propagated into the outputs or gradients).
Covered: ragged sizes (N, F not multiples of the tile sizes), the FFMA and the tcgen05 projection, the tcgen05 and the
mma.sync semantic kernels, heavy rows through the virtual-row kernels, and a dropout step."""
import numpy as np
import pytest
import torch

from han_b200 import synth
from oracle import han_oracle as O
from tests.util import compare_step, oracle_step, product_step

pytestmark = pytest.mark.gpu

GUARD_BYTES = 4096
SENTINEL = 0xA5


class GuardedAllocator:
    def __init__(self):
        self.blocks = []

    def __call__(self, shape, device, dtype=torch.float32):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
        n = int(np.prod(shape)) if len(shape) else 1
        esz = torch.empty(0, dtype=dtype).element_size()
        g = GUARD_BYTES // esz
        flat = torch.empty(n + 2 * g, dtype=dtype, device=device)
        flat.view(torch.uint8).fill_(SENTINEL)
        body = flat[g:g + n]
        if dtype.is_floating_point:
            body.fill_(float("nan"))
        self.blocks.append((flat, g * esz, n * esz, shape))
        return body.view(shape)

    def check(self):
        assert self.blocks, "the patched allocator was never used"
        for flat, gb, nb, shape in self.blocks:
            raw = flat.view(torch.uint8)
            head_ok = bool((raw[:gb] == SENTINEL).all())
            tail_ok = bool((raw[gb + nb:] == SENTINEL).all())
            assert head_ok and tail_ok, f"guard of a buffer of shape {shape} was overwritten (head ok: {head_ok}, tail ok: {tail_ok})"
        return len(self.blocks)


@pytest.fixture
def guarded(monkeypatch):
    from han_b200 import ops
    alloc = GuardedAllocator()
    monkeypatch.setattr(ops, "_empty", alloc)
    return alloc


@pytest.mark.parametrize("project_mode", [0, 1])
def test_full_step_writes_only_inside_its_buffers(guarded, project_mode):
    cfg = synth.tiny(seed=201, n=333, f=44 if project_mode else 37, p=3, deg=6.0)      # nothing is a multiple of 128 / 32
    params = O.init_params(np.random.default_rng(202), [cfg.F] * cfg.P, cfg.C)
    out_o, grads_o = oracle_step(cfg, params)
    out_p, grads_p, _ = product_step(cfg, params, project_mode=project_mode)
    n = guarded.check()
    assert n >= 20
    compare_step(out_o, grads_o, out_p, grads_p)


def test_mma_sync_semantic_kernels_write_only_inside_their_buffers(guarded, monkeypatch):
    from han_b200 import ops
    monkeypatch.setattr(ops, "SEM_TC", False)
    cfg = synth.tiny(seed=203, n=257, f=24, p=2, deg=4.0)
    params = O.init_params(np.random.default_rng(204), [cfg.F] * cfg.P, cfg.C)
    out_o, grads_o = oracle_step(cfg, params)
    out_p, grads_p, _ = product_step(cfg, params)
    guarded.check()
    compare_step(out_o, grads_o, out_p, grads_p)


def test_heavy_row_kernels_write_only_inside_their_buffers(guarded, monkeypatch):
    import han_b200 as hb
    from han_b200 import graph as hg
    cfg = synth.tiny(seed=205, n=260, f=20, p=2, deg=5.0)
    rng = np.random.default_rng(206)
    for m in cfg.masks:
        for h in rng.choice(cfg.N, size=3, replace=False):
            sel = rng.random(cfg.N) < 0.9
            m[h, sel] = True
            m[sel, h] = True
    params = O.init_params(np.random.default_rng(207), [cfg.F] * cfg.P, cfg.C)
    out_o, grads_o = oracle_step(cfg, params)
    monkeypatch.setattr(hg, "SPLIT_ROW_EDGES", 32)
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    assert graphs[0].split_view() is not None
    out_p, grads_p, _ = product_step(cfg, params, graphs=graphs)
    guarded.check()
    compare_step(out_o, grads_o, out_p, grads_p)


def test_dropout_step_writes_only_inside_its_buffers(guarded):
    import han_b200 as hb
    cfg = synth.tiny(seed=208, n=301, f=29, p=2, deg=5.0)
    params = O.init_params(np.random.default_rng(209), [cfg.F] * cfg.P, cfg.C)
    dev = torch.device("cuda")
    hp = hb.HANParams([cfg.F] * cfg.P, cfg.C, device=dev).load_dict(params)
    X = torch.from_numpy(cfg.X).to(dev).unsqueeze(0)
    graphs = [hb.process.adj_to_bias(a, [cfg.N]) for a in cfg.adjs()]
    logits, final_embed, att_val = hb.HeteGAT_multi.inference([X] * cfg.P, cfg.C, cfg.N, True, 0.6, 0.6, graphs, [8], [8, 1], params=hp)
    labels = torch.from_numpy(cfg.labels).to(dev)
    m = torch.from_numpy(cfg.train_mask.astype(np.float32)).to(dev)
    loss = hb.BaseGAttN.masked_softmax_cross_entropy(logits.reshape(-1, cfg.C), labels, m)
    loss.backward()
    torch.cuda.synchronize()
    guarded.check()
    assert torch.isfinite(loss) and torch.isfinite(final_embed).all()
    for g in hp.grad_dict().values():
        for t in (g if isinstance(g, (list, tuple)) else [g]):
            assert torch.isfinite(t).all()
