"""Times the semantic layer (D=64, A=128) forward and backward on the 2M x 4 shape with CUDA events.
python tools/sem_time.py [n] [P] [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import han_b200 as hb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    D, A = 64, 128
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    Z = torch.randn(n, P, D, device=dev, generator=g).requires_grad_(True)
    sp = {"w_omega": (torch.randn(D, A, device=dev, generator=g) * 0.1).requires_grad_(True),
          "b_omega": (torch.randn(A, device=dev, generator=g) * 0.1).requires_grad_(True),
          "u_omega": (torch.randn(A, device=dev, generator=g) * 0.1).requires_grad_(True)}
    up = torch.randn(n, D, device=dev, generator=g)
    tf, tb = [], []
    for it in range(iters + 2):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        o, _ = hb.layers.SimpleAttLayer(Z, A, return_alphas=True, params=sp)
        e[1].record()
        o.backward(up)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
        Z.grad = None
    print(f"n={n} P={P}: forward {sum(tf) / len(tf):.3f} ms, backward {sum(tb) / len(tb):.3f} ms", flush=True)


if __name__ == "__main__":
    main()
