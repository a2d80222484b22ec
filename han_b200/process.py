"""Drop-in counterpart of the reference's ``utils/process.adj_to_bias`` (utils/process.py:14-25).

The reference returns a dense float64 (G,N,N) array holding 0 on edges (self-loops included) and
-1e9 elsewhere, which the driver re-feeds to the device every step (ex_acm3025.py:118,180-181).
Here the same call returns device-resident CSR handles carrying exactly that mask.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np
import torch

from .graph import MetaPathGraph


def adj_to_bias(adj, sizes: Sequence[int], nhood: int = 1, device=None) -> Union[MetaPathGraph, List[MetaPathGraph]]:
    """``adj`` (G,N,N) dense (numpy or torch, any float dtype; the reference passes
    ``metapath_matrix - I``, ex_acm3025.py:61,110); ``sizes[g]`` must equal N (the only use in
    the reference, ex_acm3025.py:118).  Returns one ``MetaPathGraph`` when G == 1 (it reports
    ``shape == (1,N,N)`` like the reference array), else a list.
    """
    a = adj if isinstance(adj, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(adj))
    if a.dim() == 2:
        a = a.unsqueeze(0)
    G, N = a.shape[0], a.shape[1]
    if len(sizes) != G:
        raise ValueError("len(sizes) must equal the number of graphs")
    graphs = []
    for g in range(G):
        if int(sizes[g]) != N:
            # process.py:21-24 leaves entries outside sizes[g]^2 un-thresholded (biases like
            # -1e9*(1-3) = +2e9 would result); the reference never does this
            raise ValueError("sizes[g] != nb_nodes is not supported")
        graphs.append(MetaPathGraph.from_dense_adj(a[g], nhood=nhood, device=device))
    return graphs[0] if G == 1 else graphs
