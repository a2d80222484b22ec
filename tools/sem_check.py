"""Semantic layer (D=64, A=128) on the tcgen05 kernels vs torch fp64 on the GPU: per-tensor max-norm relative errors.
python tools/sem_check.py [n] [P]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import han_b200 as hb  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mode = sys.argv[3] if len(sys.argv) > 3 else "reference"
    D, A = 64, 128
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(n + P)
    Z = torch.randn(n, P, D, device=dev, generator=g)
    w = torch.randn(D, A, device=dev, generator=g) * 0.3
    b = torch.randn(A, device=dev, generator=g) * 0.3
    u = torch.randn(A, device=dev, generator=g)
    up = torch.randn(n, D, device=dev, generator=g)
    Zr, wr, br, ur = (t.double().requires_grad_(True) for t in (Z, w, b, u))
    v = torch.tanh(Zr @ wr + br)
    s = v @ ur
    al = torch.softmax(s, -1) if mode == "reference" else torch.softmax(s.mean(0), -1).expand_as(s)
    out = (Zr * al.unsqueeze(-1)).sum(1)
    (out * up.double()).sum().backward()
    Zp = Z.clone().requires_grad_(True)
    sp = {"w_omega": w.clone().requires_grad_(True), "b_omega": b.clone().requires_grad_(True),
          "u_omega": u.clone().requires_grad_(True)}
    o, a = hb.layers.SimpleAttLayer(Zp, A, return_alphas=True, params=sp, mode=mode)
    (o * up).sum().backward()
    torch.cuda.synchronize()
    print(f"n={n} P={P} {mode}: out {rel(o, out):.2e} alphas {rel(a, al):.2e} dZ {rel(Zp.grad, Zr.grad):.2e} "
          f"dw {rel(sp['w_omega'].grad, wr.grad):.2e} db {rel(sp['b_omega'].grad, br.grad):.2e} "
          f"du {rel(sp['u_omega'].grad, ur.grad):.2e}", flush=True)


if __name__ == "__main__":
    main()
