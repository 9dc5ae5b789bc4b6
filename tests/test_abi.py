"""CPU: the C-ABI library loads, exports every symbol include/cgvae_b200.h declares, and the ctypes
prototypes in coarsegrainingvae_b200/_lib.py agree with the header argument by argument.
No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from coarsegrainingvae_b200 import _lib


def _prototypes():
    with open(_lib.HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(cgvae_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, arglist)
    return protos


def _ctype_of(decl):
    decl = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", decl.strip()).strip() if not decl.strip().endswith("*") else decl.strip()
    if "*" in decl or decl == "cgvae_stream_t":
        return ctypes.c_void_p
    return {"int64_t": ctypes.c_int64, "int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
            "int32_t": ctypes.c_int32}[decl.replace("const ", "")]


def test_library_exports_every_header_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from coarsegrainingvae_b200.build import build
        build()
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name
    assert lib.cgvae_abi_version() == 1
    assert _lib.last_error() == ""


def test_ctypes_prototypes_match_header():
    protos = _prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (ret, args) in protos.items():
        res, argtypes = _lib.SIGNATURES[name]
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
        for i, (decl, ct) in enumerate(zip(args, argtypes)):
            assert _ctype_of(decl) is ct, (name, i, decl, ct)
        if ret == "int":
            assert res is ctypes.c_int, name
        elif ret == "size_t":
            assert res is ctypes.c_size_t, name


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.cgvae_gemm(7, None, 0, None, 0, None, 0, 1, 1, 1, None, 0, None, None, 0, None, None, 0, None, 0, None)
    assert rc < 0 and "bad form" in _lib.last_error()
    rc = lib.cgvae_message_fwd(5, None, None, None, None, None, None, None, None, None, 1, 1, 8, 4, 8, None, None, 0,
                               None, None, None, None)
    assert rc < 0 and "n_split" in _lib.last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "message_fwd")


def test_wgrad_problem_table_layout_matches_header_struct():
    """ops builds the problem table of cgvae_wgrad_grouped as a numpy struct array: it must be byte-compatible with
    `struct cgvae_wgrad_problem` of the header (4 pointers, 6 int32, 2 int64 = 72 bytes, natural alignment)."""
    import ctypes
    import re
    from coarsegrainingvae_b200 import _lib, ops

    class Problem(ctypes.Structure):
        _fields_ = [("gy", ctypes.c_void_p), ("x", ctypes.c_void_p), ("dW", ctypes.c_void_p), ("db", ctypes.c_void_p),
                    ("rows", ctypes.c_int32), ("n_out", ctypes.c_int32), ("n_in", ctypes.c_int32), ("ldg", ctypes.c_int32),
                    ("ldx", ctypes.c_int32), ("seg_rows", ctypes.c_int32), ("seg_stride_g", ctypes.c_int64),
                    ("seg_stride_x", ctypes.c_int64)]

    with open(_lib.HEADER_PATH) as fh:
        text = fh.read()
    body = re.search(r"typedef struct cgvae_wgrad_problem \{(.*?)\} cgvae_wgrad_problem;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip().lstrip("*") for n in decl.split(",")]
            names[-len(decl.split(",")):] = [re.split(r"[\s\*]+", n.strip())[-1] for n in decl.split(",")]
    assert names == [f[0] for f in Problem._fields_]
    dt = ops._wgrad_dtype()
    assert dt.itemsize == ctypes.sizeof(Problem) == 72
    for name, _ in Problem._fields_:
        assert dt.fields[name][1] == getattr(Problem, name).offset, name
