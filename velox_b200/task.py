"""ctypes binding of the operator-level C ABI (include/velox_b200.h): a Velox plan fragment run
by the shim Driver with the B200 operators installed. Mirrors how the reference's tests drive a
Task through `AssertQueryBuilder` (velox/exec/tests/utils/AssertQueryBuilder.h)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from ._lib import VB2_ERR_USER, VeloxRuntimeError, VeloxUserError, lib
from .vector import (BOOLEAN, CColumn, Column, FLAT, NP_DTYPES, RowVector, VARCHAR, pack_bits)

HOST, DEVICE = 0, 1


def _bind():
    L = lib()
    if getattr(L, "_task_bound", False):
        return L
    L.vb2_task_create.restype = C.c_void_p
    L.vb2_task_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]
    L.vb2_task_add_input.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CColumn), C.c_int32, C.c_int64, C.c_int32, C.c_char_p, C.c_int32]
    L.vb2_task_run.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    L.vb2_result_rows.restype = C.c_int64
    L.vb2_result_rows.argtypes = [C.c_void_p]
    L.vb2_result_cols.argtypes = [C.c_void_p]
    L.vb2_result_type.argtypes = [C.c_void_p, C.c_int32]
    L.vb2_result_copy.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.vb2_result_str_bytes.restype = C.c_int64
    L.vb2_result_str_bytes.argtypes = [C.c_void_p, C.c_int32]
    L.vb2_result_copy_str.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.vb2_result_layout.restype = C.c_int64
    L.vb2_result_layout.argtypes = [C.c_void_p, C.c_void_p]
    L.vb2_result_copy_all.argtypes = [C.c_void_p, C.c_void_p]
    L.vb2_task_stats.restype = C.c_char_p
    L.vb2_task_stats.argtypes = [C.c_void_p]
    L.vb2_task_free.argtypes = [C.c_void_p]
    L.vb2_upload_cache_create.restype = C.c_void_p
    L.vb2_upload_cache_free.argtypes = [C.c_void_p]
    L.vb2_task_set_upload_cache.argtypes = [C.c_void_p, C.c_void_p]
    L.vb2_task_set_comm.argtypes = [C.c_void_p, C.c_void_p]
    L._task_bound = True
    return L


def _raise(code: int, err) -> None:
    msg = err.value.decode(errors="replace")
    if code == VB2_ERR_USER:
        raise VeloxUserError(msg)
    raise VeloxRuntimeError(msg)


class UploadCache:
    """Device copies of host buffers shared by several tasks (vb2_upload_cache): a host table that
    two queries read crosses PCIe once. The host buffers must stay unchanged while the cache lives."""

    def __init__(self):
        self.L = _bind()
        self.h = self.L.vb2_upload_cache_create()

    def close(self):
        if self.h:
            self.L.vb2_upload_cache_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Task:
    def set_comm(self, comm) -> None:
        """Exchange transport for plans with exchange nodes: velox_b200.comm.Comm of the ranks
        running this plan (vb2_task_set_comm)."""
        self.L.vb2_task_set_comm(self.h, C.c_void_p(comm.h) if comm is not None else None)
        self._comm = comm

    def set_upload_cache(self, cache: Optional["UploadCache"]) -> None:
        self.L.vb2_task_set_upload_cache(self.h, cache.h if cache is not None else None)
        self._cache = cache  # keep it alive for the run

    def __init__(self, plan, config: Optional[Dict[str, str]] = None):
        self.L = _bind()
        self.plan = plan
        text = plan if isinstance(plan, str) else plan.sexpr
        cfg = ";".join(f"{k}={v}" for k, v in (config or {}).items())
        err = C.create_string_buffer(2048)
        self.h = self.L.vb2_task_create(text.encode(), cfg.encode(), err, 2048)
        if not self.h:
            _raise(2, err)
        self._keep = []

    def add_input(self, source_id: int, batch) -> None:
        """batch: RowVector of host columns, or a list of velox_b200.kernels.DeviceColumn."""
        err = C.create_string_buffer(2048)
        if isinstance(batch, RowVector):
            cols = [c.to_c() for c in batch.columns]
            rows, loc = batch.size, HOST
            self._keep.append(batch)
        else:
            cols = [c.to_c() for c in batch]
            rows, loc = batch[0].size, DEVICE
            self._keep.append(batch)
        arr = (CColumn * len(cols))(*cols)
        self._keep.append(arr)
        rc = self.L.vb2_task_add_input(self.h, source_id, arr, len(cols), rows, loc, err, 2048)
        if rc:
            _raise(rc, err)

    def add_arrow(self, source_id: int, batch) -> None:
        """Adds a pyarrow RecordBatch (or single-chunk Table) through the Arrow C data interface: the
        ArrowArray / ArrowSchema pair is exported by pyarrow and imported in C++ (vb2_task_add_arrow ->
        importFromArrowAsOwner, csrc/host/arrow_bridge.cpp); the buffers are viewed, not copied, and
        released by the task."""
        import pyarrow as pa
        if isinstance(batch, pa.Table):
            batches = batch.combine_chunks().to_batches()
        else:
            batches = [batch]
        self.L.vb2_task_add_arrow.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int32]
        for b in batches:
            if b.num_rows == 0:
                continue
            # struct ArrowArray is 80 bytes, struct ArrowSchema 72: the C++ side moves them out (release = NULL)
            arr = C.create_string_buffer(80)
            sch = C.create_string_buffer(72)
            b._export_to_c(C.addressof(arr), C.addressof(sch))
            err = C.create_string_buffer(2048)
            rc = self.L.vb2_task_add_arrow(self.h, source_id, C.addressof(arr), C.addressof(sch), err, 2048)
            if rc:
                # not consumed on failure: release what pyarrow exported
                pa.RecordBatch._import_from_c(C.addressof(arr), C.addressof(sch))
                _raise(rc, err)

    def device_result(self):
        """With config {"b200.result_on_device": "true"}: the result batches as lists of zero-copy torch
        views over the library's device buffers (flat fixed-width columns), valid until close()."""
        import torch
        from .vector import BIGINT, DOUBLE, INTEGER
        L = self.L
        L.vb2_result_device_columns.restype = C.c_int64
        L.vb2_result_device_columns.argtypes = [C.c_void_p, C.c_int32, C.POINTER(CColumn), C.c_int32]
        ncols = L.vb2_result_cols(self.h)
        typestr = {BIGINT: "<i8", DOUBLE: "<f8", INTEGER: "<i4"}

        class _View:
            def __init__(self, ptr, n, ts):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": ts, "data": (ptr, False), "version": 2}

        out = []
        for b in range(L.vb2_result_device_batches(self.h)):
            arr = (CColumn * ncols)()
            n = L.vb2_result_device_columns(self.h, b, arr, ncols)
            cols = []
            for c in range(ncols):
                if arr[c].encoding != FLAT or arr[c].type not in typestr:
                    raise VeloxRuntimeError("device_result(): flat BIGINT / DOUBLE / INTEGER columns only")
                cols.append(torch.as_tensor(_View(arr[c].values, n, typestr[arr[c].type]), device="cuda") if n else torch.empty(0, device="cuda"))
            out.append(cols)
        return out

    def stats(self, only: Optional[str] = None) -> Dict[str, int]:
        """Runtime stats of the task's operators; `only` keeps the entries whose name contains it."""
        out = {}
        text = self.L.vb2_task_stats(self.h)
        needle = only.encode() if only else None
        for line in text.splitlines():
            if needle is not None and needle not in line:
                continue
            k, _, v = line.decode().partition("=")
            out[k] = int(v)
        return out

    def close(self):
        if self.h:
            self.L.vb2_task_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _result(L, h, names=None) -> RowVector:
    """Result columns through the bulk calls (vb2_result_layout + vb2_result_copy_all): one blob, numpy views."""
    rows = L.vb2_result_rows(h)
    ncols = L.vb2_result_cols(h)
    layout = np.zeros(max(1, ncols) * 4, dtype=np.int64)
    total = L.vb2_result_layout(h, layout.ctypes.data)
    blob = np.zeros(max(8, total), dtype=np.uint8)
    L.vb2_result_copy_all(h, blob.ctypes.data)
    pad = lambda b: (b + 7) // 8 * 8
    cols, at = [], 0
    for c in range(ncols):
        t, vbytes, obytes, cbytes = (int(x) for x in layout[c * 4:c * 4 + 4])
        vals = blob[at:at + vbytes]
        at += pad(vbytes)
        if t == VARCHAR:
            off = blob[at:at + obytes].view(np.int32)
            at += pad(obytes)
            chars = blob[at:at + cbytes] if cbytes else np.zeros(1, dtype=np.uint8)
            at += pad(cbytes)
            col = Column(VARCHAR, FLAT, rows, off, None, chars=chars)
        elif t == BOOLEAN:
            col = Column(BOOLEAN, FLAT, rows, pack_bits(vals[:rows].astype(bool)))
            col._bool_count = rows
        else:
            col = Column(t, FLAT, rows, vals.view(NP_DTYPES[t])[:rows])
        nb = blob[at:at + rows].astype(bool)
        at += pad(rows)
        col.nulls = nb if nb.any() else None
        cols.append(col)
    return RowVector(list(names) if names else [f"c{i}" for i in range(ncols)], cols)


def _result_per_column(L, h, names=None) -> RowVector:
    rows = L.vb2_result_rows(h)
    ncols = L.vb2_result_cols(h)
    cols = []
    for c in range(ncols):
        t = L.vb2_result_type(h, c)
        nulls = np.zeros(max(rows, 1), dtype=np.uint8)
        if t == VARCHAR:
            nbytes = L.vb2_result_str_bytes(h, c)
            off = np.zeros(rows + 1, dtype=np.int32)
            chars = np.zeros(max(nbytes, 1), dtype=np.uint8)
            L.vb2_result_copy_str(h, c, off.ctypes.data, chars.ctypes.data, nulls.ctypes.data)
            col = Column(VARCHAR, FLAT, rows, off, None, chars=chars)
        elif t == BOOLEAN:
            vals = np.zeros(max(rows, 1), dtype=np.uint8)
            L.vb2_result_copy(h, c, vals.ctypes.data, nulls.ctypes.data)
            col = Column(BOOLEAN, FLAT, rows, pack_bits(vals[:rows].astype(bool)))
            col._bool_count = rows
        else:
            vals = np.zeros(max(rows, 1), dtype=NP_DTYPES[t])
            L.vb2_result_copy(h, c, vals.ctypes.data, nulls.ctypes.data)
            col = Column(t, FLAT, rows, vals[:rows])
        nb = nulls[:rows].astype(bool)
        col.nulls = nb if nb.any() else None
        cols.append(col)
    return RowVector(list(names) if names else [f"c{i}" for i in range(ncols)], cols)


def _run(self) -> RowVector:
    err = C.create_string_buffer(2048)
    rc = self.L.vb2_task_run(self.h, err, 2048)
    if rc:
        _raise(rc, err)
    names = None if isinstance(self.plan, str) else self.plan.names
    return _result(self.L, self.h, names)


def _run_only(self) -> None:
    """vb2_task_run without fetching result columns (device-resident results: see device_result())."""
    err = C.create_string_buffer(2048)
    rc = self.L.vb2_task_run(self.h, err, 2048)
    if rc:
        _raise(rc, err)


Task.run = _run
Task._run_only = _run_only


def run_tasks(tasks: Sequence["Task"]):
    """vb2_tasks_run: the prepared tasks run concurrently, one host thread each inside the library.
    Returns their result RowVectors."""
    L = tasks[0].L
    L.vb2_tasks_run.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_char_p, C.c_int32]
    arr = (C.c_void_p * len(tasks))(*[t.h for t in tasks])
    err = C.create_string_buffer(2048)
    rc = L.vb2_tasks_run(arr, len(tasks), err, 2048)
    if rc:
        _raise(rc, err)
    return [_result(L, t.h, None if isinstance(t.plan, str) else t.plan.names) for t in tasks]


def run_plan(plan, sources: Sequence, config: Optional[Dict[str, str]] = None, batch_rows: Optional[int] = None):
    """Runs `plan` over sources[i] (RowVector per source id; optionally split into batches of
    batch_rows). Returns (result RowVector, stats)."""
    t = Task(plan, config)
    try:
        for sid, src in enumerate(sources):
            if isinstance(src, RowVector) and batch_rows and src.size > batch_rows:
                for part in split_rowvector(src, batch_rows):
                    t.add_input(sid, part)
            else:
                t.add_input(sid, src)
        out = t.run()
        return out, t.stats()
    finally:
        t.close()


def split_rowvector(rv: RowVector, batch_rows: int, max_batches: Optional[int] = None):
    """Slices host columns into batches (FLAT / DICTIONARY / CONSTANT aware); at most `max_batches` of them when given."""
    from .vector import CONSTANT, DICTIONARY
    out = []
    for r0 in range(0, rv.size, batch_rows):
        if max_batches is not None and len(out) >= max_batches:
            break
        r1 = min(rv.size, r0 + batch_rows)
        cols = []
        for c in rv.columns:
            n = r1 - r0
            if c.encoding == FLAT:
                if c.type == VARCHAR:
                    off = c.values[r0:r1 + 1]
                    chars = c.chars[off[0]:off[-1]] if off[-1] > off[0] else np.zeros(1, dtype=np.uint8)
                    nc = Column(VARCHAR, FLAT, n, (off - off[0]).astype(np.int32), None if c.nulls is None else c.nulls[r0:r1], chars=np.ascontiguousarray(chars))
                elif c.type == BOOLEAN:
                    from .vector import unpack_bits
                    bits_ = unpack_bits(c.values, c._bool_count)[r0:r1]
                    nc = Column(BOOLEAN, FLAT, n, pack_bits(bits_), None if c.nulls is None else c.nulls[r0:r1])
                    nc._bool_count = n
                else:
                    nc = Column(c.type, FLAT, n, c.values[r0:r1], None if c.nulls is None else c.nulls[r0:r1])
            elif c.encoding == DICTIONARY:
                nc = Column(c.type, DICTIONARY, n, c.values, None if c.nulls is None else c.nulls[r0:r1], c.indices[r0:r1], c.dict_nulls, c.chars)
                nc._bool_count = c._bool_count
            else:
                nc = Column(c.type, CONSTANT, n, c.values, c.nulls, chars=c.chars)
                nc._bool_count = c._bool_count
            if nc.nulls is not None and not nc.nulls.any():
                nc.nulls = None
            cols.append(nc)
        out.append(RowVector(rv.names, cols))
    return out
