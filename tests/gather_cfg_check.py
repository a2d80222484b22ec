"""Run by tests/test_gpu_attn.py::test_gather_ring_geometries in a subprocess with HAN_GATHER_CFG set (the library reads
it once per process): a complete step on a graph with empty-ish, short and long rows against the fp64 oracle."""
import sys

import numpy as np

from han_b200 import synth
from oracle import han_oracle as O
from tests.util import compare_step, oracle_step, product_step


def main():
    cfg = synth.tiny(seed=611, n=700, f=33, p=2, deg=23.0)
    params = O.init_params(np.random.default_rng(612), [cfg.F] * cfg.P, cfg.C)
    out_o, grads_o = oracle_step(cfg, params)
    out_p, grads_p, _ = product_step(cfg, params)
    compare_step(out_o, grads_o, out_p, grads_p)
    print("gather cfg ok")


if __name__ == "__main__":
    sys.exit(main())
