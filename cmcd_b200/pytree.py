"""ravel_pytree for nested dict/list/tuple of torch tensors.

Mirrors jax.flatten_util.ravel_pytree as used by mcdboundingmachine.initialize
(/root/reference/src/mcdboundingmachine.py:122): dict keys in sorted order, sequences in order;
``unflatten`` is differentiable (split + reshape views of the flat vector).
"""
import torch


def tree_leaves(tree, out=None):
    out = [] if out is None else out
    if isinstance(tree, dict):
        for k in sorted(tree):
            tree_leaves(tree[k], out)
    elif isinstance(tree, (list, tuple)):
        for v in tree:
            tree_leaves(v, out)
    elif tree is not None:
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


def tree_map(fn, tree):
    return _rebuild(tree, iter([fn(l) for l in tree_leaves(tree)]))


class Unflatten:
    """Hashable callable (the reference passes ``unflatten`` as a static jit argument, main.py:174-177)."""

    def __init__(self, tree, shapes, sizes):
        self._tree, self._shapes, self._sizes = tree, shapes, sizes

    def __call__(self, vec):
        parts = torch.split(vec, self._sizes)
        return _rebuild(self._tree, iter(p.reshape(s) for p, s in zip(parts, self._shapes)))

    def leaf_offsets(self):
        """The pytree with every leaf replaced by (offset into the flat vector, shape) -- what the fused O(K) chain kernels
        (csrc/chain.cu) need to read the parameters from / write the gradient into the flat vector directly."""
        offs, o = [], 0
        for n in self._sizes:
            offs.append(o)
            o += n
        return _rebuild(self._tree, iter(list(zip(offs, self._shapes))))

    @property
    def size(self):
        return sum(self._sizes)


def ravel_pytree(tree, dtype=torch.float32, device=None):
    leaves = [torch.as_tensor(l, dtype=dtype, device=device) for l in tree_leaves(tree)]
    shapes = [tuple(l.shape) for l in leaves]
    sizes = [l.numel() for l in leaves]
    flat = torch.cat([l.reshape(-1) for l in leaves]) if leaves else torch.zeros(0, dtype=dtype, device=device)
    skeleton = tree_map(lambda l: 0, tree)
    return flat, Unflatten(skeleton, shapes, sizes)
