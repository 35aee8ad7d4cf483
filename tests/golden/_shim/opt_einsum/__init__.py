"""Minimal stand-in for the `opt_einsum` package, used ONLY by tests/golden/make_golden.py.

The reference package (emdgroup/tnmf, /root/reference) imports `opt_einsum`, which is not installed in
this container and cannot be fetched (no network).  tnmf uses exactly two entry points of it:

  * contract(op, labels, op, labels, out_labels, optimize=...)  -- interleaved form with string labels
    (tnmf/backends/NumPy.py:82-90,106-119,128-131)
  * contract_expression('nm...,mc...->nc...', shapeA, shapeB)   -- returns a callable
    (tnmf/backends/_NumPyFFTBackend.py:59,74,87)

Both are mapped onto numpy.einsum here.  This file is test infrastructure for generating golden vectors
from the unmodified reference; nothing in the product imports it.
"""
import numpy as np

from . import contract as _contract_module  # noqa: F401  (exposes opt_einsum.contract.ContractExpression)
from .contract import ContractExpression


def contract(*operands, optimize=True, **_ignored):
    if isinstance(operands[0], str):
        return np.einsum(*operands, optimize=True)
    # interleaved form: (array, labels, array, labels, ..., out_labels); labels may be arbitrary strings
    arrays, label_lists = [], []
    ops = list(operands)
    out = ops.pop() if len(ops) % 2 == 1 else None
    for i in range(0, len(ops), 2):
        arrays.append(ops[i])
        label_lists.append(list(ops[i + 1]))
    table = {}
    def idx(lbl):
        return table.setdefault(lbl, len(table))
    args = []
    for a, ls in zip(arrays, label_lists):
        args += [a, [idx(x) for x in ls]]
    if out is not None:
        args.append([idx(x) for x in out])
    return np.einsum(*args, optimize=True)


def contract_expression(subscripts, *shapes, **_ignored):
    return ContractExpression(subscripts)
