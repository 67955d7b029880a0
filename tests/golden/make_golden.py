"""Generates the committed golden vectors from the oracle (the reference itself is Julia and cannot run here, so
the goldens are outputs of the oracle *after* it was pinned on the reference's KATs in tests/test_oracle_*.py).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from fixtures import FSPMAT_THETA, SENS_THETA, TELEGRAPH_S, TOGGLE_S, fspmat_propensities, sens_telegraph  # noqa: E402
from oracle.fspmatrix import FspMatrixOracle  # noqa: E402
from oracle.sensmatrix import SensFspMatrixOracle  # noqa: E402
from oracle.statespace import StateSpaceOracle, StateSpaceOracleFast  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    # 1. telegraph, expand!(12): states + connectivity + matvec at three times for the three formulations
    sp = StateSpaceOracle(TELEGRAPH_S, [1, 0, 0])
    sp.expand(12)
    v = rng.random(sp.get_state_count() + 4)
    out = {"states": sp.states_array(), "state_conn": sp.state_connectivity_array(), "sink_conn": sp.sink_connectivity_array(),
           "v": v, "times": np.array([0.0, 0.37, 1.0])}
    for kind in ("ti", "tv", "tvj"):
        A = FspMatrixOracle(sp, fspmat_propensities(kind), FSPMAT_THETA)
        out[f"w_{kind}"] = np.stack([A.matvec(t, v) for t in out["times"]])
    np.savez_compressed(os.path.join(HERE, "telegraph_expand12.npz"), **out)
    # 2. toggle insertion order after expand!(15), delete, expand!(2)
    sp = StateSpaceOracle(TOGGLE_S, [0, 0])
    sp.expand(15)
    ids = np.random.default_rng(1).choice(sp.get_state_count(), size=sp.get_state_count() // 3, replace=False) + 1
    sp.deleteat(ids)
    sp.expand(2)
    np.savez_compressed(os.path.join(HERE, "toggle_expand_delete.npz"), ids=ids, states=sp.states_array(),
                        state_conn=sp.state_connectivity_array(), sink_conn=sp.sink_connectivity_array())
    # 3. sensitivity matvec on the 1002-state rectangular fixture (test/sensmat/telegraph.jl:29-31)
    props, grads, pattern, states = sens_telegraph()
    SA = SensFspMatrixOracle(StateSpaceOracleFast(TELEGRAPH_S, states), props, grads, pattern, SENS_THETA)
    vs = rng.random(6 * SA.fspmatrix.rowcount)
    np.savez_compressed(os.path.join(HERE, "sens_telegraph.npz"), vs=vs, times=np.array([10.0, 30.0]),
                        out=np.stack([SA.matvec(t, vs) for t in (10.0, 30.0)]))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
