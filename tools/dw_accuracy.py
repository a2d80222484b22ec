"""Accuracy of dW = X^T dS at the 2M-node scale: tcgen05 3xTF32 split-K kernel and the FFMA kernel against torch fp64.
Run on a GPU box: python tools/dw_accuracy.py [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from han_b200._lib import call, ptr, query, stream_ptr  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    F, G, D = 256, 4, 64
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(n, F, device=dev, generator=g)
    for tag, dS in (("dense dS", torch.randn(G, n, D, device=dev, generator=g)),
                    ("sparse dS (18% rows)", torch.randn(G, n, D, device=dev, generator=g) *
                     (torch.rand(G, n, 1, device=dev, generator=g) < 0.18))):
        ref = torch.empty(F, G * D, dtype=torch.float64, device=dev)
        for gg in range(G):
            acc = torch.zeros(F, D, dtype=torch.float64, device=dev)
            for c0 in range(0, n, 1 << 18):
                acc += X[c0:c0 + (1 << 18)].double().t() @ dS[gg, c0:c0 + (1 << 18)].double()
            ref[:, gg * D:(gg + 1) * D] = acc
        for mode, name in ((1, "tcgen05 3xTF32"), (0, "FFMA fp32")):
            dW = torch.empty(F, G * D, device=dev)
            if mode:
                wsb = query("han_project_bwd_tc_workspace_bytes", n, F, G)
                ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
                call("han_project_bwd_tc", ptr(X), n, F, F, ptr(dS), G, ptr(dW), ptr(ws), wsb, mode, stream_ptr())
            else:
                wsb = query("han_project_bwd_workspace_bytes", n, F, G, D)
                ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
                call("han_project_bwd", ptr(X), n, F, F, ptr(dS), G, D, ptr(dW), ptr(ws), wsb, 0, stream_ptr())
            torch.cuda.synchronize()
            err = (dW.double() - ref).abs().max().item() / ref.abs().max().item()
            print(f"n={n} {tag}: {name}: max-norm rel err {err:.3e}", flush=True)


if __name__ == "__main__":
    main()
