"""julia/NumCMEB200.jl cannot be executed here (no Julia in the image): check statically that every `ccall` in it names a
symbol of include/ncme.h and passes the argument list the ctypes table (numcme.jl_b200/_lib.py, itself checked against
the header and the shared library by tests/test_abi.py) declares -- same count, same scalar / pointer category."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _jl_category(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring",):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "Int64": "i64", "Float64": "f64", "Csize_t": "size", "Cvoid": "void"}[t]


def _ct_category(t):
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "_type_") and isinstance(t._type_, type):
        return "ptr"
    return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_double: "f64", C.c_size_t: "size"}[t]


def test_julia_ccalls_match_the_abi(pkg):
    from numcme_jl_b200._lib import SIGNATURES
    src = open(os.path.join(ROOT, "julia", "NumCMEB200.jl")).read()
    header = open(os.path.join(ROOT, "include", "ncme.h")).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*libncme\),\s*(\w+),\s*\(", src):
        name, ret = m.group(1), m.group(2)
        i, depth = m.end(), 1
        while depth:                      # matching parenthesis of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        calls.append((name, ret, _split_top(src[m.end():i - 1])))
    assert len(calls) >= 30
    used = set()
    for name, ret, args in calls:
        assert re.search(r"\b%s\s*\(" % name, header), f"{name} is not declared in include/ncme.h"
        res, argtypes = SIGNATURES[name]
        assert len(args) == len(argtypes), f"{name}: Julia passes {len(args)} arguments, the ABI takes {len(argtypes)}"
        assert _jl_category(ret) == _ct_category(res), f"{name}: return type {ret}"
        for k, (ja, ca) in enumerate(zip(args, argtypes)):
            assert _jl_category(ja) == _ct_category(ca), f"{name}: argument {k + 1} is {ja} in the Julia glue"
        used.add(name)
    # the entry points of the hot path are all bound
    for must in ("ncme_space_create", "ncme_space_expand", "ncme_space_delete", "ncme_matrix_create", "ncme_matvec",
                 "ncme_matvec_host", "ncme_solve_segment", "ncme_space_prune_by_mass", "ncme_space_compact_vector",
                 "ncme_space_marginal", "ncme_matrix_create_incremental", "ncme_space_new_count", "ncme_sensmatrix_create",
                 "ncme_sens_matvec", "ncme_sensmatrix_set_joint_values"):
        assert must in used, f"the Julia glue does not bind {must}"
