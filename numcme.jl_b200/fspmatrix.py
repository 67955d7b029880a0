"""FspMatrixSparse behind the reference's API; all arithmetic is the fused CUDA matvec of libncme.

Reference: src/fspmatrix/sparse/fspsparsematrix.jl -- constructor :47-108, ``matvec!`` :196-217,
``matvecadd!`` :226-247, ``matvec`` :254-258, ``*`` :262-264, ``size`` :174-186, getters :29-37.
Julia's ``matvec!`` / ``matvecadd!`` are ``matvec_`` / ``matvecadd_`` here.

Vectors may be numpy arrays (host path: H2D + kernel + D2H inside the call, what a Julia
``Vector{Float64}`` binding does) or device-resident (``DeviceVector`` / torch CUDA float64 tensor:
kernel only, asynchronous on the context's stream).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .cmemodel import JOINT_TV, SEPARABLE_TV, Propensity, eval_over_columns, eval_over_states
from .device import DeviceVector, device_ptr, is_device, vec_len
from .statespace import StateSpaceSparse


# probe times of the rank-1 (separability) test of joint time-varying propensities: irregular, spread over several
# decades so that switch-like time factors are sampled on both sides when they switch inside [0, 1e4]
_PROBE_TIMES = (0.0, 0.7310585786300049, 19.098300562505255, 738.90560989306495, 5459.8150033144236, 28813.3)


# below this many states the incremental rebuild (one more kernel + host round trip) costs more than it saves
INCREMENTAL_MIN_STATES = 2048
# up to this many states a matrix keeps an eager host copy of the state list (the reference's deepcopy, :97)
EAGER_STATES_MAX = 1 << 21
# from this many states on, the reactions' state factors are evaluated concurrently on host threads
PARALLEL_EVAL_MIN_STATES = 1 << 20


def _rank1_info(f, states, parameters, probe_times=_PROBE_TIMES, rtol=1e-12):
    """Worker of ``detect_rank1``: returns ``None`` or a dict with the state factor ``g = f(t_ref, .)``, the sentinel
    state indices, ``t_ref`` and the ratios ``c(t_k) / c(t_ref)`` at the other probe times (what an incremental rebuild
    needs to validate states added later against the same factorisation)."""
    n = states.shape[0]
    if n == 0:
        return None
    g, t_ref, ratios = None, None, []
    for t in probe_times:
        v = eval_over_states(f, states, parameters, t=t)
        if not np.all(np.isfinite(v)):
            return None
        if g is None:
            if np.any(v != 0.0):
                g, t_ref = v, t
            else:
                ratios.append((t, 0.0))        # vanishes identically at this probe time
            continue
        supp = g != 0.0
        if np.any(v[~supp] != 0.0):
            return None
        ratio = v[supp] / g[supp]
        if np.abs(ratio - ratio[0]).max() > rtol * max(abs(ratio[0]), 1e-300) and np.abs(ratio).max() > 0.0:
            return None
        ratios.append((t, float(ratio[0])))
    if g is None:
        return None            # vanishes at every probe time: leave it on the exact (joint) path
    order = np.argsort(-np.abs(g), kind="stable")
    nz = order[: int(np.count_nonzero(g))]
    sent = [int(nz[0])] + [int(nz[k]) for k in (len(nz) // 3, 2 * len(nz) // 3, len(nz) - 1) if 0 < k < len(nz)]
    # one more sentinel OUTSIDE the support of g (the state with the largest molecule count): f must stay 0 there
    off = np.nonzero(g == 0.0)[0]
    zero_sent = [int(off[np.argmax(states[off].sum(axis=1))])] if off.size else []
    return {"g": g, "sent": list(dict.fromkeys(sent)), "zero_sent": zero_sent, "t_ref": t_ref, "ratios": ratios, "rtol": rtol}


def _rank1_extend(f, info, new_states, parameters):
    """State factor of ``new_states`` under the factorisation found earlier, or ``None`` if they do not follow it."""
    if new_states.shape[0] == 0:
        return np.zeros(0)
    g = eval_over_states(f, new_states, parameters, t=info["t_ref"])
    if not np.all(np.isfinite(g)):
        return None
    for t, ratio in info["ratios"]:
        v = eval_over_states(f, new_states, parameters, t=t)
        if np.abs(v - ratio * g).max() > 1e-11 * max(np.abs(g).max() * abs(ratio), 1e-300) and np.abs(v - ratio * g).max() > 0.0:
            return None
    return g


def detect_rank1(f, states, parameters, probe_times=_PROBE_TIMES, rtol=1e-12):
    """is the joint propensity ``f(t, x, p)`` a product ``c(t) g(x)`` on ``states``?
    Evaluates f over all states at a few probe times; separable iff every evaluation is a scalar multiple of the first
    non-vanishing one (same support, ratio constant to ``rtol``).  Returns ``(g, sentinels)`` -- the state factor
    ``g = f(t_ref, .)`` and up to four state indices on which the time factor ``c(t) = f(t, x*, p) / g(x*)`` is
    evaluated (the first) and cross-checked (the others) at run time -- or ``None``."""
    info = _rank1_info(f, states, parameters, probe_times, rtol)
    return None if info is None else (info["g"], info["sent"])


class FspMatrixSparse:
    def __init__(self, space: StateSpaceSparse, propensity_functions, parameters=(), comm=None, detect_separable=True,
                 previous: "FspMatrixSparse | None" = None):
        """``comm`` (a parallel.Comm) restricts the matrix to this rank's row block (K8); vectors are then the
        local slices ``[rows | R sinks]`` and matvec inputs need halo margins (see parallel.ShardedVector).

        ``detect_separable``: joint time-varying propensities (what the Catalyst import produces,
        catalyst_interface.jl:19-27) need n host closure calls and an upload per distinct ``t``
        (``_update_sparsematrix!``, fspsparsematrix.jl:154-166).  When such a propensity is numerically a product
        c(t) g(x) on the current states (``detect_rank1``) it is put on the separable path instead -- one host call per
        right-hand side, nothing uploaded -- and cross-checked on sentinel states at every ``t``.

        ``previous``: the matrix this space was assembled into before its last prune / expand (the rebuild after an
        ``adapt!``, fspsolve.jl:176).  Only the states added since are evaluated on the host and uploaded; the factors
        of the surviving states are carried over on the device (``ncme_matrix_create_incremental``).  Any
        mismatch (other propensities, a joint propensity that stops being separable on the new states, a space that was
        assembled into another matrix in between) silently falls back to the full build."""
        self.ctx = space.ctx
        self.comm = comm
        self.parameters = parameters
        self.propensities = list(propensity_functions)
        if len(self.propensities) != space.nr:
            raise L.ArgumentError("one propensity per reaction is required")
        for a in self.propensities:
            if not isinstance(a, Propensity):
                raise L.ArgumentError("propensities must be Propensity instances (see propensity())")
        n = space.get_state_count()
        self.n = n
        self._space, self._space_version, self._states = space, space.version, None
        if n <= EAGER_STATES_MAX:                 # host copy, like `deepcopy(space.states)` (:97); lazy for huge spaces
            self._states = space.get_states()
        self.nr = space.nr
        self.rowcount = self.colcount = n + space.get_sink_count()   # global size, as size(A) in the reference
        self.timeinvariant_propensity_ids = [i + 1 for i, a in enumerate(self.propensities) if a.kind == "ti"]
        self.separabletv_propensity_ids = [i + 1 for i, a in enumerate(self.propensities) if a.kind == "sep"]
        self.jointtv_propensity_ids = [i + 1 for i, a in enumerate(self.propensities) if a.kind == "joint"]
        self.incremental = False
        self.build_window = None
        h = None
        if previous is not None and getattr(previous, "_h", None) and n >= max(1, INCREMENTAL_MIN_STATES):
            h = self._build_incremental(space, previous, comm)
        if h is None:
            h = self._build_full(space, comm, detect_separable)
        self._h = h
        self.device_separable_ids = sorted(self._tfactor)
        self.device_joint_ids = [r for r in self.jointtv_propensity_ids if r not in self._rank1]
        info = self.shard_info()
        self.local_len = info["row_hi"] - info["row_lo"] + self.nr
        self.t_cache = -np.inf
        self._coef = np.ones(self.nr, dtype=np.float64)

    @property
    def states(self) -> np.ndarray:
        """The states this matrix was built on (n x NS).  Above EAGER_STATES_MAX states the copy is made on first use
        (the matrix itself only needs the species columns of its own rows); it is an error to ask for it after the
        space has been mutated."""
        if self._states is None:
            if self._space.version != self._space_version:
                raise L.NcmeError("the state space was modified after this matrix was built; its state list is gone")
            self._states = self._space.get_states()
        return self._states

    def _build_full(self, space, comm, detect_separable):
        n, parameters = self.n, self.parameters
        self.kinds = np.array([a.kind_code for a in self.propensities], dtype=np.int32)
        # what the device matrix does with each reaction (differs from the user's classification only for joint
        # propensities found to be rank-1 separable)
        self._tfactor = {}                 # 1-based reaction id -> callable t -> c(t)
        self._rank1 = {}                   # 1-based reaction id -> factorisation info of detected reactions
        lib = L.load()
        # Row-sharded build: every rank evaluates the state factors of its own rows + predecessor window
        # only -- host evaluation and upload shrink with the number of ranks.  Joint propensities are refreshed over the
        # same window at every distinct t (_refresh_joint).
        lo, hi = 0, n
        windowed = comm is not None and comm.nranks > 1 and n > 0
        if windowed:
            w = (C.c_int64 * 4)()
            L.check(lib.ncme_matrix_shard_window(space.handle, comm.handle, w))
            lo, hi = int(w[2]), int(w[3])
            self.build_window = (lo, hi)
        nw = hi - lo
        # species columns of the window straight from the device keys: one contiguous float64 array per species
        cols = space.get_state_columns(lo, nw) if nw else [np.zeros(0) for _ in range(space.ns)]
        propvals = np.empty((self.nr, max(nw, 1)), dtype=np.float64)

        def fill(r):
            a = self.propensities[r]
            if a.kind == "ti":
                propvals[r, :nw] = eval_over_columns(a.f, cols, parameters)
            elif a.kind == "sep":
                propvals[r, :nw] = eval_over_columns(a.statefactor, cols, parameters)
            else:
                propvals[r, :nw] = 0.0
        if nw >= PARALLEL_EVAL_MIN_STATES and self.nr > 1:
            # numpy releases the GIL inside its loops: the reactions are evaluated concurrently on the host cores
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(self.nr, os.cpu_count() or 1, 16)) as pool:
                list(pool.map(fill, range(self.nr)))
        else:
            for r in range(self.nr):
                fill(r)
        for r, a in enumerate(self.propensities):
            if a.kind == "sep":
                self._tfactor[r + 1] = (lambda t, a=a: float(a.tfactor(t, self.parameters)))
            elif a.kind == "joint" and detect_separable:
                info = _rank1_info(a.f, self.states, parameters)
                if info is not None:
                    g = info.pop("g")
                    propvals[r, :nw] = g[lo:hi]
                    self.kinds[r] = SEPARABLE_TV
                    info["sent_states"] = [[int(v) for v in self.states[i]] for i in info["sent"]]
                    info["zero_states"] = [[int(v) for v in self.states[i]] for i in info["zero_sent"]]
                    info["sent_g"] = [float(g[i]) for i in info["sent"]]
                    self._rank1[r + 1] = info
                    self._tfactor[r + 1] = self._make_rank1_tfactor(r + 1, a.f, info)
        propvals = np.ascontiguousarray(propvals[:, :nw]) if nw else propvals
        h = L.p_void()
        if windowed:
            L.check(lib.ncme_matrix_create_window(space.handle, comm.handle, L.ptr(self.kinds, C.c_int32),
                                                  L.ptr(propvals, C.c_double), lo, hi, C.byref(h)))
        else:
            L.check(lib.ncme_matrix_create_sharded(space.handle, comm.handle if comm is not None else None,
                                                   L.ptr(self.kinds, C.c_int32), L.ptr(propvals, C.c_double), C.byref(h)))
        return h

    def _build_incremental(self, space, previous, comm):
        """Rebuild after an adapt!: evaluate the propensities of the appended states only (see ``previous``)."""
        if len(previous.propensities) != len(self.propensities) or any(
                a is not b for a, b in zip(previous.propensities, self.propensities)) or previous.comm is not comm:
            return None
        lib = L.load()
        nk, nn = C.c_int64(), C.c_int64()
        L.check(lib.ncme_space_new_count(space.handle, C.byref(nk), C.byref(nn)))
        n_kept, n_new = nk.value, nn.value
        if n_kept == 0:
            return None
        new_states = space.get_states(n_kept, n_new) if self._states is None else self._states[n_kept:]
        th = self.parameters
        propvals = np.zeros((self.nr, max(n_new, 1)), dtype=np.float64)
        for r, a in enumerate(self.propensities):
            if a.kind == "ti":
                propvals[r, :n_new] = eval_over_states(a.f, new_states, th)
            elif a.kind == "sep":
                propvals[r, :n_new] = eval_over_states(a.statefactor, new_states, th)
            elif (r + 1) in previous._rank1:
                g = _rank1_extend(a.f, previous._rank1[r + 1], new_states, th)
                if g is None:
                    return None            # the new states do not follow the old factorisation: full build
                propvals[r, :n_new] = g
        self.kinds = previous.kinds.copy()
        propvals = np.ascontiguousarray(propvals[:, :n_new]) if n_new else propvals
        h = L.p_void()
        st = lib.ncme_matrix_create_incremental(space.handle, comm.handle if comm is not None else None, previous._h,
                                                L.ptr(self.kinds, C.c_int32), L.ptr(propvals, C.c_double), C.byref(h))
        if st != 0:
            return None                    # e.g. the space was assembled into another matrix in between
        self._rank1 = dict(previous._rank1)
        self._tfactor = {}
        for r, a in enumerate(self.propensities):
            if a.kind == "sep":
                self._tfactor[r + 1] = (lambda t, a=a: float(a.tfactor(t, self.parameters)))
            elif (r + 1) in self._rank1:
                self._tfactor[r + 1] = self._make_rank1_tfactor(r + 1, a.f, self._rank1[r + 1])
        self.incremental = True
        self.new_state_count = n_new
        return h

    @property
    def handle(self):
        return self._h

    # -- accessors (fspsparsematrix.jl:29-37)
    def size(self, dim=None):
        if dim is None:
            return (self.rowcount, self.colcount)
        if dim not in (1, 2):
            raise L.ArgumentError("Second argument must be either 1 or 2.")
        return self.rowcount if dim == 1 else self.colcount

    def stats(self) -> dict:
        nt = C.c_int()
        nnz = (C.c_int64 * 40)()
        ab = C.c_int64()
        db = C.c_int64()
        L.check(L.load().ncme_matrix_stats(self._h, C.byref(nt), nnz, C.byref(ab), C.byref(db)))
        return {"nterms": nt.value, "nnz_per_term": [nnz[k] for k in range(nt.value)],
                "algorithmic_bytes": ab.value, "device_bytes": db.value}

    def set_pipe(self, rows: int, stages: int):
        L.check(L.load().ncme_matrix_set_pipe(self._h, int(rows), int(stages)))

    def compression_info(self) -> dict:
        info = (C.c_int64 * 4)()
        L.check(L.load().ncme_matrix_compression_info(self._h, info))
        return {"chunk_slots": int(info[0]), "wide_chunk_slots": int(info[1]), "enabled": bool(info[2]), "slots": int(info[3])}

    def shard_info(self) -> dict:
        info = (C.c_int64 * 8)()
        L.check(L.load().ncme_matrix_shard_info(self._h, info))
        keys = ["row_lo", "row_hi", "halo_lo", "halo_hi", "n_global", "interior_begin", "interior_end", "nranks"]
        return dict(zip(keys, [int(v) for v in info]))

    def set_tuning(self, rows_per_thread: int):
        L.check(L.load().ncme_matrix_set_tuning(self._h, int(rows_per_thread)))

    def _make_rank1_tfactor(self, rid, f, info):
        xs, gs, zs = info["sent_states"], info["sent_g"], info.get("zero_states", [])

        def tfactor(t):
            c = float(f(t, xs[0], self.parameters)) / gs[0]
            bad = any(float(f(t, z, self.parameters)) != 0.0 for z in zs)
            for x, gv in zip(xs[1:], gs[1:]):          # sentinels: the product form must hold at every t actually used
                ck = float(f(t, x, self.parameters)) / gv
                bad |= abs(ck - c) > 1e-9 * max(abs(c), abs(ck), 1e-300)
            if bad:                                    # solve() catches this and repeats the segment on the joint path
                raise L.SeparabilityError(f"propensity {rid} was classified as c(t) g(x) on the probe times but is not "
                                          f"separable at t = {t!r}; build the matrix with detect_separable=False")
            return c
        return tfactor

    # -- time-dependent pieces evaluated on the host (they are opaque closures, as in the reference)
    def coefficients(self, t: float) -> np.ndarray:
        for r in self.device_separable_ids:
            self._coef[r - 1] = self._tfactor[r](t)
        return self._coef

    def _refresh_joint(self, t: float):
        if t == self.t_cache:
            return
        self.t_cache = t
        if not self.device_joint_ids:
            return
        states = self.states
        if self.comm is not None and self.comm.nranks > 1:
            # row-sharded: this rank's rows + predecessor window [row_lo - halo_lo, row_hi + halo_hi)
            info = self.shard_info()
            states = states[info["row_lo"] - info["halo_lo"]: info["row_hi"] + info["halo_hi"]]
        for r in self.device_joint_ids:
            vals = np.ascontiguousarray(eval_over_states(self.propensities[r - 1].f, states, self.parameters, t=t),
                                        dtype=np.float64)
            L.check(L.load().ncme_matrix_set_joint_values(self._h, r, L.ptr(vals, C.c_double)))

    def _apply(self, out, t, v, beta: float):
        N = self.local_len
        if vec_len(out) != N or vec_len(v) != N:
            raise L.ArgumentError(f"DimensionMismatch: matrix is {N}x{N}, got vectors of length {vec_len(v)} and {vec_len(out)}")
        coef = self.coefficients(float(t))
        self._refresh_joint(float(t))
        lib = L.load()
        if is_device(out) and is_device(v):
            L.check(lib.ncme_matvec(self._h, L.ptr(coef, C.c_double), C.c_void_p(device_ptr(v)), C.c_void_p(device_ptr(out)),
                                    beta))
            return
        if isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous:
            vv = np.ascontiguousarray(v, dtype=np.float64)
            L.check(lib.ncme_matvec_host(self._h, L.ptr(coef, C.c_double), vv.ctypes.data_as(C.c_void_p),
                                         out.ctypes.data_as(C.c_void_p), beta))
            return
        raise L.ArgumentError("out/v must both be device vectors or `out` a contiguous float64 numpy array")

    def matvec_local_(self, out, t, v):
        """Sharded matvec that leaves the sink entries as per-rank partial sums (see ncme_matvec_local)."""
        coef = self.coefficients(float(t))
        L.check(L.load().ncme_matvec_local(self._h, L.ptr(coef, C.c_double), C.c_void_p(device_ptr(v)),
                                           C.c_void_p(device_ptr(out))))

    def close(self):
        if getattr(self, "_h", None):
            L.load().ncme_matrix_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.close()
        except Exception:
            pass

    def __matmul__(self, v):
        return matvec(0.0, self, v)

    def __mul__(self, v):
        return matvec(0.0, self, v)


def matvec_(out, t, A, v):
    """matvec!(out, t, A, v):  out = A(t) v   (A: FspMatrixSparse or ForwardSensFspMatrixSparse)"""
    if not isinstance(A, FspMatrixSparse):
        return A.matvec_(out, t, v)
    A._apply(out, t, v, 0.0)


def matvecadd_(out, t, A: FspMatrixSparse, v):
    """matvecadd!(out, t, A, v):  out = out + A(t) v"""
    A._apply(out, t, v, 1.0)


def matvec(t, A: FspMatrixSparse, v):
    """w = matvec(t, A, v)"""
    if is_device(v):
        w = DeviceVector(A.ctx, vec_len(v))
    else:
        w = np.empty(vec_len(v), dtype=np.float64)
    matvec_(w, t, A, v)
    return w


def get_rowcount(A):
    return A.rowcount


def get_colcount(A):
    return A.colcount


def get_propensities_of(A):
    return A.propensities
