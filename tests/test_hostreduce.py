"""The host-side shared-memory all-reduce of the sharded integrators' step scalars (numcme.jl_b200/csrc/hostreduce.h,
used by comm.cu) between REAL processes on the CPU: every rank must obtain bitwise the same sums -- the rank-ordered sum
of what all ranks contributed -- over thousands of back-to-back reductions (double-buffered slots, sequence numbers)."""
import ctypes as C
import multiprocessing as mp
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmpdir):
    out = os.path.join(tmpdir, "libhr.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "numcme.jl_b200", "csrc"),
                           os.path.join(ROOT, "tests", "harness", "hostreduce_harness.cpp"), "-o", out, "-lrt"])
    return out


def _lib(path):
    lib = C.CDLL(path)
    lib.h_open.restype = C.c_void_p
    lib.h_open.argtypes = [C.c_char_p, C.c_int]
    lib.h_close.argtypes = [C.c_void_p]
    lib.h_unlink.argtypes = [C.c_char_p]
    lib.h_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    return lib


def _worker(path, name, rank, nranks, rounds, count, q):
    lib = _lib(path)
    hr = lib.h_open(name, 0)
    out = np.zeros(rounds * count)
    rc = lib.h_run(hr, rank, nranks, rounds, count, out.ctypes.data_as(C.POINTER(C.c_double))) if hr else -1
    lib.h_close(hr)
    q.put((rank, rc, out))


def _expected(nranks, rounds, count):
    want = np.zeros((rounds, count))
    for e in range(1, rounds + 1):
        acc = np.zeros(count)
        for rank in range(nranks):                       # rank order, like hr_sum
            v = np.empty(count)
            for k in range(count):
                z = (0x9E3779B97F4A7C15 * (rank * 1000003 + e * 7919 + k + 1)) & 0xFFFFFFFFFFFFFFFF
                z ^= z >> 31
                v[k] = (rank + 1) * 1e-3 * (k + 1) + e * 0.5 + float(z % 1000003) * 1e-9
            acc = acc + v
        want[e - 1] = acc
    return want.reshape(-1)


@pytest.mark.parametrize("nranks,rounds,count", [(2, 3000, 26), (4, 1500, 289), (3, 2000, 1)])
def test_hostreduce_between_processes(tmp_path, nranks, rounds, count):
    path = _build(str(tmp_path))
    lib = _lib(path)
    name = f"/ncme_hr_test_{os.getpid()}_{nranks}".encode()
    lib.h_unlink(name)
    hr0 = lib.h_open(name, 1)
    assert hr0, "shm_open failed"
    try:
        ctx = mp.get_context("fork")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(path, name, r, nranks, rounds, count, q)) for r in range(nranks)]
        for p in procs:
            p.start()
        res = {}
        for _ in procs:
            rank, rc, out = q.get(timeout=120)
            res[rank] = (rc, out)
        for p in procs:
            p.join(timeout=30)
    finally:
        lib.h_close(hr0)
        lib.h_unlink(name)
    want = _expected(nranks, min(rounds, 40), count)          # the python restatement is slow: check the first rounds exactly
    for r in range(nranks):
        rc, out = res[r]
        assert rc == 0
        assert np.array_equal(out, res[0][1])                 # identical bits on every rank, every round
        assert np.array_equal(out[: want.size], want)         # = the sum in rank order
    # later rounds: plausible values (monotone in the round index through the 0.5 * e term)
    last = res[0][1].reshape(rounds, count)
    assert np.all(np.diff(last[:, 0]) > 0)
