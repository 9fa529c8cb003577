"""Arrow interop for host batches (the scan-side step in front of the path).

pyarrow arrays map onto the same buffers the C ABI describes (`vb2_column`): fixed-width values are
borrowed zero-copy, strings are Arrow's int32 offsets + data exactly as VARCHAR is laid out here, and
dictionary arrays become DictionaryVector-style columns. Reference counterpart:
velox/vector/arrow/Bridge.h (importFromArrow / exportToArrow)."""
from __future__ import annotations

import numpy as np

from .vector import (BIGINT, BOOLEAN, DICTIONARY, DOUBLE, FLAT, INTEGER, VARCHAR, Column, RowVector, pack_bits, row_vector, unpack_bits)


def _arrow():
    import pyarrow as pa
    return pa


def _type_of(t) -> int:
    pa = _arrow()
    if pa.types.is_int64(t):
        return BIGINT
    if pa.types.is_int32(t) or pa.types.is_date32(t):
        return INTEGER
    if pa.types.is_float64(t):
        return DOUBLE
    if pa.types.is_boolean(t):
        return BOOLEAN
    if pa.types.is_string(t):
        return VARCHAR
    raise TypeError(f"Arrow type {t} has no counterpart on the B200 path (BOOLEAN, INTEGER/DATE, BIGINT, DOUBLE, VARCHAR)")


def _nulls(arr):
    if arr.null_count == 0:
        return None
    return np.asarray(arr.is_null().to_numpy(zero_copy_only=False), dtype=bool)


def _flat(arr) -> Column:
    """One Arrow array (no chunks) -> flat Column. Fixed-width values are views of the Arrow buffers."""
    t = _type_of(arr.type)
    n = len(arr)
    bufs = arr.buffers()
    if t == VARCHAR:
        off = np.frombuffer(bufs[1], dtype=np.int32)[arr.offset:arr.offset + n + 1]
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, dtype=np.uint8)
        if n and off[0] != 0:  # sliced array: rebase the offsets
            data = data[off[0]:off[-1]]
            off = off - off[0]
        if data.size == 0:
            data = np.zeros(1, dtype=np.uint8)
        return Column(VARCHAR, FLAT, n, np.ascontiguousarray(off), _nulls(arr), chars=np.ascontiguousarray(data))
    if t == BOOLEAN:
        flags = np.asarray(arr.fill_null(False).to_numpy(zero_copy_only=False), dtype=bool)
        col = Column(BOOLEAN, FLAT, n, pack_bits(flags), _nulls(arr))
        col._bool_count = n
        return col
    dtype = {INTEGER: np.int32, BIGINT: np.int64, DOUBLE: np.float64}[t]
    values = np.frombuffer(bufs[1], dtype=dtype)[arr.offset:arr.offset + n] if n else np.zeros(0, dtype=dtype)
    return Column(t, FLAT, n, values, _nulls(arr))


def column_from_arrow(arr) -> Column:
    pa = _arrow()
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
    if pa.types.is_dictionary(arr.type):
        base = _flat(arr.dictionary)
        idx = arr.indices.cast(pa.int32())
        wrapper_nulls = _nulls(idx)
        indices = np.ascontiguousarray(np.asarray(idx.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.int32))
        col = Column(base.type, DICTIONARY, len(arr), base.values, wrapper_nulls, indices, base.nulls, base.chars)
        col._bool_count = base._bool_count
        return col
    return _flat(arr)


def row_vector_from_arrow(table) -> RowVector:
    """pyarrow Table / RecordBatch -> RowVector of host columns (zero-copy where the layouts agree)."""
    names = list(table.schema.names)
    return row_vector(names, [column_from_arrow(table.column(i)) for i in range(len(names))])


def column_to_arrow(col: Column):
    pa = _arrow()
    n = col.size
    arrow_type = {BOOLEAN: pa.bool_(), INTEGER: pa.int32(), BIGINT: pa.int64(), DOUBLE: pa.float64(), VARCHAR: pa.string()}[col.type]

    def base_array(values, nulls, count):
        mask = None if nulls is None else np.asarray(nulls, dtype=bool)
        if col.type == VARCHAR:
            off, raw = values, col.chars.tobytes()
            strings = [None if (mask is not None and mask[i]) else raw[off[i]:off[i + 1]].decode() for i in range(count)]
            return pa.array(strings, type=arrow_type)
        if col.type == BOOLEAN:
            return pa.array(unpack_bits(values, count), type=arrow_type, mask=mask)
        return pa.array(values, type=arrow_type, mask=mask)

    if col.encoding == FLAT:
        return base_array(col.values, col.nulls, n)
    if col.encoding == DICTIONARY:
        count = col.dict_size
        dictionary = base_array(col.values, col.dict_nulls, count)
        idx = pa.array(col.indices, type=pa.int32(), mask=None if col.nulls is None else np.asarray(col.nulls, dtype=bool))
        return pa.DictionaryArray.from_arrays(idx, dictionary)
    one = base_array(col.values, col.nulls, 1)
    return pa.repeat(one[0], n)


def row_vector_to_arrow(rv: RowVector):
    pa = _arrow()
    return pa.table([column_to_arrow(c) for c in rv.columns], names=list(rv.names))
