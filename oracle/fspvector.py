"""TEST INFRASTRUCTURE (oracle): literal restatement of the marginal sum of the reference's FspVectorSparse.

Reference: src/fspvector/fspvector.jl:66-99 -- ``sum(p, dims)`` walks the states in order, keeps the species not in
``dims``, appends a reduced state the first time it is seen and accumulates the values in state order.
Only tests/ may import this module."""
import numpy as np


def marginal_sum(states, values, dims):
    states = np.asarray(states, dtype=np.int64)
    ns = states.shape[1]
    dims = sorted(set(int(d) for d in dims))
    if not (dims[0] >= 1 and dims[-1] <= ns):
        raise ValueError(f"Input dimensions must be between 1 and {ns}.")
    keep = [k for k in range(ns) if (k + 1) not in dims]
    idx = {}
    rstates, rvals = [], []
    for i in range(states.shape[0]):
        key = tuple(int(v) for v in states[i, keep])
        j = idx.get(key)
        if j is None:
            idx[key] = len(rstates)
            rstates.append(key)
            rvals.append(float(values[i]))
        else:
            rvals[j] += float(values[i])
    return np.asarray(rstates, dtype=np.int64).reshape(len(rstates), len(keep)), np.asarray(rvals, dtype=np.float64)
