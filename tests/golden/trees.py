"""Nested dict / list of tensors <-> flat npz keys ('p/deep/0/W/1'), shared by the fixture generators and the tests."""
import numpy as np
import torch


def flatten_tree(tree, prefix):
    out = {}
    if isinstance(tree, dict):
        for k, v in tree.items():
            out.update(flatten_tree(v, f"{prefix}/{k}"))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            out.update(flatten_tree(v, f"{prefix}/{i}"))
    else:
        out[prefix] = tree.detach().numpy() if isinstance(tree, torch.Tensor) else np.asarray(tree)
    return out


def unflatten_tree(d, prefix):
    """Inverse of flatten_tree for the keys under ``prefix`` (integer path components become lists)."""
    root = {}
    for key in d.keys():
        if not key.startswith(prefix + "/"):
            continue
        parts = key[len(prefix) + 1:].split("/")
        node = root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = torch.from_numpy(np.asarray(d[key]))

    def fix(node):
        if not isinstance(node, dict):
            return node
        node = {k: fix(v) for k, v in node.items()}
        if node and all(k.isdigit() for k in node):
            return [node[str(i)] for i in range(len(node))]
        return node
    return fix(root)


