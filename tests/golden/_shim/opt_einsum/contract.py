"""`opt_einsum.contract` submodule stand-in: only the ContractExpression type is needed."""
import numpy as np


class ContractExpression:
    def __init__(self, subscripts):
        self._subscripts = subscripts

    def __call__(self, *arrays):
        return np.einsum(self._subscripts, *arrays, optimize=True)
