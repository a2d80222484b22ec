"""Times MetaPathGraph.transpose() (the by-source view build) on one meta-path of the 2M-node bench workload."""
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import han_b200 as hb  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = bench.build_workload("syn2m", dev, None)
    g0 = wl["graphs"][0]
    times = []
    for rep in range(4):
        g = hb.MetaPathGraph.from_csr(g0.indptr.clone(), g0.indices.clone(), n_cols=wl["N"], device=dev)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.transpose()
        e.record()
        torch.cuda.synchronize()
        times.append(round(s.elapsed_time(e), 3))
    print("transpose ms per 100M-edge meta-path:", times)


if __name__ == "__main__":
    main()
