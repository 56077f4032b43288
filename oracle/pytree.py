"""Minimal ravel_pytree (jax.flatten_util) restatement for nested dict/list/tuple of torch tensors.

ORACLE / TEST INFRASTRUCTURE ONLY.  Dict keys are visited in sorted order, sequences in
order, exactly as jax.tree_util does (reference use: mcdboundingmachine.py:122).
"""
import torch


def _leaves(tree, out):
    if isinstance(tree, dict):
        for k in sorted(tree):
            _leaves(tree[k], out)
    elif isinstance(tree, (list, tuple)):
        for v in tree:
            _leaves(v, out)
    elif tree is None:
        pass
    else:
        out.append(tree)
    return out


def _rebuild(tree, it):
    if isinstance(tree, dict):
        return {k: _rebuild(tree[k], it) for k in sorted(tree)}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_rebuild(v, it) for v in tree)
    if tree is None:
        return None
    return next(it)


def ravel_pytree(tree, dtype=torch.float32):
    leaves = [torch.as_tensor(l, dtype=dtype) for l in _leaves(tree, [])]
    shapes = [tuple(l.shape) for l in leaves]
    sizes = [l.numel() for l in leaves]
    flat = torch.cat([l.reshape(-1) for l in leaves]) if leaves else torch.zeros(0, dtype=dtype)

    def unflatten(vec):
        parts = torch.split(vec, sizes)
        return _rebuild(tree, iter(p.reshape(s) for p, s in zip(parts, shapes)))

    return flat, unflatten
