"""Host-side column builders for the C ABI (`vb2_column`, include/velox_b200.h).

The layout is the vector data contract of the reference: flat values buffer,
LSB-first validity bitmap (1 = not null, velox/common/base/Nulls.h:26-27), int32 dictionary
indices (velox/vector/DictionaryVector.h:275-278), bit-packed BOOLEAN values. VARCHAR travels as
int32 offsets + chars (the device cannot follow StringView pointers, velox/type/StringView.h:76-77).

Names follow the reference's test helper `VectorMaker` (velox/vector/tests/utils/VectorMaker.h):
flat_vector / dictionary_vector / constant_vector / row_vector.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

BOOLEAN, INTEGER, BIGINT, DOUBLE, VARCHAR = 0, 3, 4, 6, 7
DATE = INTEGER  # DATE is int32 days since epoch (velox/type/Type.h:1305)
FLAT, DICTIONARY, CONSTANT = 0, 1, 2

TYPE_NAMES = {BOOLEAN: "BOOLEAN", INTEGER: "INTEGER", BIGINT: "BIGINT", DOUBLE: "DOUBLE", VARCHAR: "VARCHAR"}
NP_DTYPES = {INTEGER: np.int32, BIGINT: np.int64, DOUBLE: np.float64}


class CColumn(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("encoding", C.c_int32),
        ("size", C.c_int64),
        ("values", C.c_void_p),
        ("nulls", C.c_void_p),
        ("indices", C.c_void_p),
        ("dict_size", C.c_int64),
        ("dict_nulls", C.c_void_p),
        ("aux", C.c_void_p),
    ]


class CTable(C.Structure):
    _fields_ = [("ncols", C.c_int32), ("reserved", C.c_int32), ("rows", C.c_int64), ("cols", C.POINTER(CColumn))]


def pack_bits(flags: np.ndarray) -> np.ndarray:
    """bool[n] -> u64 words, LSB first."""
    flags = np.asarray(flags, dtype=bool)
    n = flags.size
    words = (n + 63) // 64
    padded = np.zeros(words * 64, dtype=bool)
    padded[:n] = flags
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint64).copy()


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(np.asarray(words).view(np.uint8), bitorder="little")[:n].astype(bool)


def encode_strings(strings: Sequence[Optional[str]]):
    offsets = np.zeros(len(strings) + 1, dtype=np.int32)
    chunks = []
    pos = 0
    for i, s in enumerate(strings):
        b = b"" if s is None else (s if isinstance(s, bytes) else s.encode())
        chunks.append(b)
        pos += len(b)
        offsets[i + 1] = pos
    chars = np.frombuffer(b"".join(chunks) or b"\0", dtype=np.uint8).copy()
    return offsets, chars


@dataclass
class Column:
    """One host column. `nulls` is a bool array with True = NULL (or None)."""

    type: int
    encoding: int
    size: int
    values: Optional[np.ndarray] = None  # flat values / dictionary base values / 1 constant
    nulls: Optional[np.ndarray] = None
    indices: Optional[np.ndarray] = None
    dict_nulls: Optional[np.ndarray] = None
    chars: Optional[np.ndarray] = None  # VARCHAR: values holds int32 offsets
    _keep: list = field(default_factory=list, repr=False)

    @property
    def dict_size(self) -> int:
        if self.encoding != DICTIONARY:
            return 0
        return (self.values.size - 1) if self.type == VARCHAR else (
            self._bool_count if self.type == BOOLEAN else self.values.size)

    def to_c(self) -> CColumn:
        c = CColumn()
        c.type, c.encoding, c.size = self.type, self.encoding, self.size
        keep = self._keep
        keep.clear()

        def ptr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a)
            keep.append(a)
            return a.ctypes.data

        c.values = ptr(self.values)
        c.aux = ptr(self.chars)
        c.nulls = ptr(pack_bits(~self.nulls)) if self.nulls is not None else None
        c.indices = ptr(self.indices)
        c.dict_size = self.dict_size
        c.dict_nulls = ptr(pack_bits(~self.dict_nulls)) if self.dict_nulls is not None else None
        return c

    # Logical python values (None for null) — used by tests to compare result sets.
    def to_pylist(self):
        base = self._base_pylist()
        if self.encoding == FLAT:
            out = base
        elif self.encoding == DICTIONARY:
            out = [base[i] for i in self.indices]
        else:
            out = [base[0]] * self.size
        if self.nulls is not None and self.encoding != CONSTANT:
            out = [None if n else v for v, n in zip(out, self.nulls)]
        return out

    def _base_pylist(self):
        bn = self.dict_nulls if self.encoding == DICTIONARY else (self.nulls if self.encoding != FLAT else None)
        if self.type == VARCHAR:
            off = self.values
            raw = self.chars.tobytes()
            vals = [raw[off[i]:off[i + 1]].decode() for i in range(off.size - 1)]
        elif self.type == BOOLEAN:
            vals = [bool(x) for x in unpack_bits(self.values, self._bool_count)]
        else:
            vals = self.values.tolist()
        if bn is not None:
            vals = [None if n else v for v, n in zip(vals, bn)]
        return vals

    _bool_count: int = 0


def _values_of(type_: int, data, nulls):
    """Returns (values ndarray, chars, bool_count)."""
    if type_ == VARCHAR:
        off, chars = encode_strings(data)
        return off, chars, 0
    if type_ == BOOLEAN:
        arr = np.array([bool(x) if x is not None else False for x in data], dtype=bool)
        return pack_bits(arr), None, arr.size
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data.astype(NP_DTYPES[type_], copy=False)), None, 0
    return np.array([0 if x is None else x for x in data], dtype=NP_DTYPES[type_]), None, 0


def _nulls_of(data, nulls):
    if nulls is not None:
        nulls = np.asarray(nulls, dtype=bool)
        return nulls if nulls.any() else None
    if isinstance(data, np.ndarray):
        return None
    flags = np.array([x is None for x in data], dtype=bool)
    return flags if flags.any() else None


def flat_vector(type_: int, data, nulls=None) -> Column:
    n = len(data)
    v, chars, bc = _values_of(type_, data, nulls)
    col = Column(type_, FLAT, n, v, _nulls_of(data, nulls), chars=chars)
    col._bool_count = bc
    return col


def dictionary_vector(type_: int, indices, base, index_nulls=None, base_nulls=None) -> Column:
    """DictionaryVector: row i = base[indices[i]]; `index_nulls` adds nulls on the wrapper."""
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int32))
    v, chars, bc = _values_of(type_, base, base_nulls)
    wn = None
    if index_nulls is not None:
        wn = np.asarray(index_nulls, dtype=bool)
        wn = wn if wn.any() else None
    col = Column(type_, DICTIONARY, idx.size, v, wn, idx, _nulls_of(base, base_nulls), chars)
    col._bool_count = bc
    return col


def constant_vector(type_: int, value, size: int) -> Column:
    v, chars, bc = _values_of(type_, [value], None)
    col = Column(type_, CONSTANT, size, v, np.array([True]) if value is None else None, chars=chars)
    col._bool_count = bc
    return col


@dataclass
class RowVector:
    """A batch: equally sized named columns (velox/vector/ComplexVector.h RowVector)."""

    names: list
    columns: list

    @property
    def size(self) -> int:
        return self.columns[0].size if self.columns else 0

    @property
    def types(self):
        return [c.type for c in self.columns]

    def to_c(self):
        arr = (CColumn * len(self.columns))(*[c.to_c() for c in self.columns])
        t = CTable(len(self.columns), 0, self.size, arr)
        t._arr = arr
        return t

    def rows(self):
        cols = [c.to_pylist() for c in self.columns]
        return list(zip(*cols)) if cols else []


def row_vector(names, columns) -> RowVector:
    assert len(names) == len(columns)
    assert len({c.size for c in columns}) <= 1, "children must have equal sizes"
    return RowVector(list(names), list(columns))
