"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (velox_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from velox_b200.vector import (BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR, CColumn, CTable, Column, FLAT,
                               RowVector, NP_DTYPES, pack_bits)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleUserError(Exception):
    """The oracle's analogue of VeloxUserError (arithmetic overflow, division by zero...)."""


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_run_plan.restype = C.c_void_p
        L.orc_run_plan.argtypes = [C.c_char_p, C.c_int32, C.POINTER(CTable), C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        L.orc_result_rows.restype = C.c_int64
        L.orc_result_rows.argtypes = [C.c_void_p]
        L.orc_result_cols.argtypes = [C.c_void_p]
        L.orc_result_type.argtypes = [C.c_void_p, C.c_int32]
        L.orc_result_copy.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_result_str_bytes.restype = C.c_int64
        L.orc_result_str_bytes.argtypes = [C.c_void_p, C.c_int32]
        L.orc_result_copy_str.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_hash_columns.argtypes = [C.POINTER(CColumn), C.c_int32, C.c_int64, C.c_void_p]
        L.orc_partition.argtypes = [C.POINTER(CColumn), C.c_int32, C.c_int64, C.c_int32, C.c_void_p]
        L.orc_twang_mix64.restype = C.c_uint64
        L.orc_twang_mix64.argtypes = [C.c_uint64]
        L.orc_jenkins_rev_mix32.restype = C.c_uint32
        L.orc_jenkins_rev_mix32.argtypes = [C.c_uint32]
        L.orc_hash_mix.restype = C.c_uint64
        L.orc_hash_mix.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_hash_bytes.restype = C.c_uint64
        L.orc_hash_bytes.argtypes = [C.c_uint64, C.c_char_p, C.c_int64]
        L.orc_hash_f64.restype = C.c_uint64
        L.orc_hash_f64.argtypes = [C.c_double]
        L.orc_compare_f64.argtypes = [C.c_int32, C.c_double, C.c_double]
        L.orc_checked_i64.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        _LIB = L
    return _LIB


def result_to_rowvector(L, prefix: str, res, names=None) -> RowVector:
    """Shared by the oracle and product bindings (same accessor shapes, different prefixes)."""
    rows = getattr(L, prefix + "result_rows")(res)
    ncols = getattr(L, prefix + "result_cols")(res)
    cols = []
    for c in range(ncols):
        t = getattr(L, prefix + "result_type")(res, c)
        nulls = np.zeros(max(rows, 1), dtype=np.uint8)
        if t == VARCHAR:
            nbytes = getattr(L, prefix + "result_str_bytes")(res, c)
            off = np.zeros(rows + 1, dtype=np.int32)
            chars = np.zeros(max(nbytes, 1), dtype=np.uint8)
            getattr(L, prefix + "result_copy_str")(res, c, off.ctypes.data, chars.ctypes.data, nulls.ctypes.data)
            col = Column(VARCHAR, FLAT, rows, off, None, chars=chars)
        elif t == BOOLEAN:
            vals = np.zeros(max(rows, 1), dtype=np.uint8)
            getattr(L, prefix + "result_copy")(res, c, vals.ctypes.data, nulls.ctypes.data)
            col = Column(BOOLEAN, FLAT, rows, pack_bits(vals[:rows].astype(bool)))
            col._bool_count = rows
        else:
            vals = np.zeros(max(rows, 1), dtype=NP_DTYPES[t])
            getattr(L, prefix + "result_copy")(res, c, vals.ctypes.data, nulls.ctypes.data)
            col = Column(t, FLAT, rows, vals[:rows])
        nb = nulls[:rows].astype(bool)
        col.nulls = nb if nb.any() else None
        cols.append(col)
    return RowVector(list(names) if names else [f"c{i}" for i in range(ncols)], cols)


def run_plan(plan, sources, threads: int = 1, batch_rows: int = 10000) -> RowVector:
    """plan: velox_b200.plan._Node (or plan text); sources: list of RowVector by source id."""
    L = lib()
    text = plan if isinstance(plan, str) else plan.sexpr
    names = None if isinstance(plan, str) else plan.names
    tabs = [s.to_c() for s in sources]
    arr = (CTable * len(tabs))(*tabs)
    err = C.create_string_buffer(1024)
    res = L.orc_run_plan(text.encode(), len(tabs), arr, threads, batch_rows, err, 1024)
    if not res:
        msg = err.value.decode()
        if msg.startswith("VeloxUserError"):
            raise OracleUserError(msg)
        raise RuntimeError(msg)
    try:
        return result_to_rowvector(L, "orc_", res, names)
    finally:
        L.orc_result_free(res)


def hash_columns(columns) -> np.ndarray:
    L = lib()
    n = columns[0].size
    arr = (CColumn * len(columns))(*[c.to_c() for c in columns])
    out = np.zeros(n, dtype=np.uint64)
    assert L.orc_hash_columns(arr, len(columns), n, out.ctypes.data) == 0
    return out


def partition(columns, num_partitions: int) -> np.ndarray:
    L = lib()
    n = columns[0].size
    arr = (CColumn * len(columns))(*[c.to_c() for c in columns])
    out = np.zeros(n, dtype=np.uint32)
    assert L.orc_partition(arr, len(columns), n, num_partitions, out.ctypes.data) == 0
    return out
