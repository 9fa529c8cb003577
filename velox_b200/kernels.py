"""ctypes binding of the kernel-level C ABI (include/velox_b200_kernels.h) over torch device
memory. torch is plumbing here (allocation, streams); every computation is one of our kernels."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from ._lib import check, lib
from .vector import (BIGINT, BOOLEAN, CColumn, CONSTANT, DICTIONARY, DOUBLE, FLAT, INTEGER, VARCHAR, Column)

MAX_COLS, MAX_PARAMS, MAX_KEYS = 8, 12, 2


class FusedArgs(C.Structure):
    _fields_ = [
        ("cols", C.c_void_p * MAX_COLS),
        ("pf", C.c_double * MAX_PARAMS),
        ("pl", C.c_int64 * MAX_PARAMS),
        ("pi", C.c_int32 * MAX_PARAMS),
        ("rows", C.c_int64),
        ("nkeys", C.c_int32),
        ("ngroups", C.c_int32),
        ("key", C.c_void_p * MAX_KEYS),
        ("key_is64", C.c_int32 * MAX_KEYS),
        ("key_mult", C.c_int32 * MAX_KEYS),
        ("key_min", C.c_int64 * MAX_KEYS),
        ("key_lut", C.c_void_p * MAX_KEYS),
        ("join_slot_flags", C.c_void_p),
        ("join_min", C.c_int64),
        ("join_range", C.c_int64),
    ]


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class DeviceColumn:
    """A column resident in HBM (torch tensors own the buffers)."""

    def __init__(self, type_: int, encoding: int, size: int, values=None, nulls=None, indices=None,
                 dict_size: int = 0, dict_nulls=None, aux=None):
        self.type, self.encoding, self.size = type_, encoding, size
        self.values, self.nulls, self.indices = values, nulls, indices
        self.dict_size, self.dict_nulls, self.aux = dict_size, dict_nulls, aux

    @staticmethod
    def from_host(col: Column, device="cuda") -> "DeviceColumn":
        c = col.to_c()  # packs nulls; keeps numpy buffers alive in col._keep

        def up(a):
            return None if a is None else torch.from_numpy(a).to(device)

        import numpy as np
        from .vector import pack_bits
        return DeviceColumn(
            col.type, col.encoding, col.size,
            values=up(col.values), nulls=up(pack_bits(~col.nulls).view(np.int64)) if col.nulls is not None else None,
            indices=up(col.indices), dict_size=col.dict_size,
            dict_nulls=up(pack_bits(~col.dict_nulls).view(np.int64)) if col.dict_nulls is not None else None,
            aux=up(col.chars))

    def to_c(self) -> CColumn:
        c = CColumn()
        c.type, c.encoding, c.size = self.type, self.encoding, self.size
        c.values, c.nulls, c.indices = _ptr(self.values), _ptr(self.nulls), _ptr(self.indices)
        c.dict_size, c.dict_nulls, c.aux = self.dict_size, _ptr(self.dict_nulls), _ptr(self.aux)
        return c


def flat_device(type_: int, t: torch.Tensor) -> DeviceColumn:
    return DeviceColumn(type_, FLAT, t.numel(), values=t)


def hash_columns(cols: Sequence[DeviceColumn]) -> torch.Tensor:
    n = cols[0].size
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    arr = (CColumn * len(cols))(*[c.to_c() for c in cols])
    check(lib().vb2k_hash_columns(arr, len(cols), C.c_int64(n), C.c_void_p(out.data_ptr()), _stream()))
    return out  # uint64 bit pattern in int64


def partition_ids(hashes: torch.Tensor, num_partitions: int) -> torch.Tensor:
    out = torch.empty(hashes.numel(), dtype=torch.int32, device="cuda")
    check(lib().vb2k_partition_ids(C.c_void_p(hashes.data_ptr()), C.c_int64(hashes.numel()), num_partitions,
                                   C.c_void_p(out.data_ptr()), _stream()))
    return out


def partition_scatter_order(ids: torch.Tensor, num_partitions: int):
    counts = torch.zeros(num_partitions, dtype=torch.int64, device="cuda")
    order = torch.empty(ids.numel(), dtype=torch.int32, device="cuda")
    check(lib().vb2k_partition_scatter_order(C.c_void_p(ids.data_ptr()), C.c_int64(ids.numel()), num_partitions,
                                             C.c_void_p(counts.data_ptr()), C.c_void_p(order.data_ptr()), _stream()))
    return counts, order


SENTINEL_KEY = -0x7F7F7F7F7F7F7F80  # VB2_SENTINEL_KEY (bytes 0x80)


def partition_segments(keys: torch.Tensor, cols: Sequence[torch.Tensor], rows: int, rows_dev: Optional[torch.Tensor], num_partitions: int,
                       segcap: int, overflow: torch.Tensor):
    """Sync-free hash partitioning into fixed-capacity per-destination segments (vb2k_partition_segments).
    Returns (seg_keys[P * segcap], [seg_col...], counts[P]); tails hold SENTINEL_KEY."""
    total = num_partitions * segcap
    seg_keys = torch.empty(total, dtype=torch.int64, device="cuda")
    seg_cols = [torch.empty(total, dtype=t.dtype, device="cuda") for t in cols]
    counts = torch.empty(num_partitions, dtype=torch.int64, device="cuda")
    n = len(cols)
    ins = (C.c_void_p * max(1, n))(*[t.data_ptr() for t in cols])
    outs = (C.c_void_p * max(1, n))(*[t.data_ptr() for t in seg_cols])
    eb = (C.c_int32 * max(1, n))(*[t.element_size() for t in cols])
    check(lib().vb2k_partition_segments(C.c_void_p(keys.data_ptr()), ins, eb, n, C.c_int64(rows), C.c_void_p(_ptr(rows_dev)), num_partitions,
                                        C.c_int64(segcap), C.c_void_p(seg_keys.data_ptr()), outs, C.c_void_p(counts.data_ptr()),
                                        C.c_void_p(overflow.data_ptr()), _stream()))
    return seg_keys, seg_cols, counts


def key_range_check(keys: torch.Tensor, lo: int, hi: int, flag: torch.Tensor):
    check(lib().vb2k_key_range_check(C.c_void_p(keys.data_ptr()), C.c_int64(keys.numel()), C.c_int64(lo), C.c_int64(hi),
                                     C.c_void_p(flag.data_ptr()), _stream()))


def gather(src: torch.Tensor, order: torch.Tensor) -> torch.Tensor:
    out = torch.empty(order.numel(), dtype=src.dtype, device="cuda")
    check(lib().vb2k_gather(C.c_void_p(src.data_ptr()), C.c_void_p(order.data_ptr()), C.c_int64(order.numel()),
                            src.element_size(), C.c_void_p(out.data_ptr()), _stream()))
    return out


def join_slot_flags(head: torch.Tensor, codes: Optional[torch.Tensor], flag: Optional[torch.Tensor]) -> torch.Tensor:
    out = torch.empty(head.numel(), dtype=torch.uint8, device="cuda")
    check(lib().vb2k_join_slot_flags(C.c_void_p(head.data_ptr()), C.c_void_p(_ptr(codes)), C.c_void_p(_ptr(flag)),
                                     C.c_int64(head.numel()), C.c_void_p(out.data_ptr()), _stream()))
    return out


def fused_find(signature: str) -> int:
    return lib().vb2k_fused_find(signature.encode())


def fused_signatures():
    L = lib()
    L.vb2k_fused_signature.restype = C.c_char_p
    return [L.vb2k_fused_signature(i).decode() for i in range(L.vb2k_fused_count())]


class FusedScanAgg:
    """Persistent accumulators + workspace for one fused pipeline instance."""

    def __init__(self, signature: str, ngroups: int = 1):
        L = lib()
        self.id = L.vb2k_fused_find(signature.encode())
        if self.id < 0:
            raise KeyError(f"no fused pipeline for {signature}")
        self.nproj = L.vb2k_fused_nproj(self.id)
        self.ngroups = max(1, ngroups)
        L.vb2k_fused_workspace_bytes.restype = C.c_size_t
        self.ws_bytes = L.vb2k_fused_workspace_bytes(self.id, self.ngroups)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device="cuda")
        self.sums = torch.zeros(self.ngroups * self.nproj, dtype=torch.float64, device="cuda")
        self.counts = torch.zeros(self.ngroups, dtype=torch.int64, device="cuda")

    def reset(self):
        self.sums.zero_()
        self.counts.zero_()

    def add_batch(self, cols: Sequence[torch.Tensor], rows: int, pf=(), pl=(), pi=(), keys=(), key_min=(),
                  key_mult=(), key_lut=(), join=None):
        a = FusedArgs()
        for i, t in enumerate(cols):
            a.cols[i] = t.data_ptr()
        for i, v in enumerate(pf):
            a.pf[i] = v
        for i, v in enumerate(pl):
            a.pl[i] = v
        for i, v in enumerate(pi):
            a.pi[i] = v
        a.rows = rows
        a.nkeys = len(keys)
        a.ngroups = self.ngroups
        for k, t in enumerate(keys):
            a.key[k] = t.data_ptr()
            a.key_is64[k] = 1 if t.dtype == torch.int64 else 0
            a.key_min[k] = key_min[k] if key_min else 0
            a.key_mult[k] = key_mult[k]
            a.key_lut[k] = key_lut[k].data_ptr() if key_lut and key_lut[k] is not None else None
        if join is not None:
            a.join_slot_flags = join["slot_flags"].data_ptr()
            a.join_min = join["min"]
            a.join_range = join["slot_flags"].numel()
        check(lib().vb2k_fused_scan_agg(self.id, C.byref(a), C.c_void_p(self.sums.data_ptr()),
                                        C.c_void_p(self.counts.data_ptr()), C.c_void_p(self.ws.data_ptr()),
                                        C.c_size_t(self.ws_bytes), _stream()))


# ---- additional kernel-level wrappers used by bench.py / the multi-GPU path -------------------------
class JoinTable(C.Structure):
    _fields_ = [("mode", C.c_int32), ("pad", C.c_int32), ("key_min", C.c_int64), ("capacity", C.c_int64),
                ("keys", C.c_void_p), ("head", C.c_void_p), ("next", C.c_void_p), ("build_rows", C.c_int64)]


class Instr(C.Structure):
    _fields_ = [("op", C.c_int32), ("type", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32)]


class Const(C.Structure):
    _fields_ = [("type", C.c_int32), ("is_null", C.c_int32), ("i", C.c_int64), ("d", C.c_double), ("str", C.c_void_p),
                ("len", C.c_int32), ("pad", C.c_int32)]


class Program(C.Structure):
    _fields_ = [("instrs", C.POINTER(Instr)), ("n_instrs", C.c_int32), ("n_filter_instrs", C.c_int32), ("filter_reg", C.c_int32),
                ("n_regs", C.c_int32), ("consts", C.POINTER(Const)), ("n_consts", C.c_int32), ("pad", C.c_int32)]


class Output(C.Structure):
    _fields_ = [("reg", C.c_int32), ("type", C.c_int32), ("values", C.c_void_p), ("nulls", C.c_void_p)]


VB2_OP_LIKE = 22


def column_minmax(col: DeviceColumn):
    """(min, max, non-null count) of an integer column; synchronises."""
    out = torch.empty(3, dtype=torch.int64, device="cuda")
    c = col.to_c()
    check(lib().vb2k_column_minmax(C.byref(c), C.c_int64(col.size), C.c_void_p(out.data_ptr()), _stream()))
    lo, hi, nn = out.tolist()
    return lo, hi, nn


def normalize_keys(cols: Sequence[DeviceColumn], mins, mults, ranges=None, nulls_invalid=False, n=None):
    n = cols[0].size if n is None else n
    arr = (CColumn * len(cols))(*[c.to_c() for c in cols])
    keys = torch.empty(n, dtype=torch.int64, device="cuda")
    valid = torch.empty((n + 63) // 64, dtype=torch.int64, device="cuda") if (ranges is not None or nulls_invalid) else None
    k = len(cols)
    mins_a = (C.c_int64 * k)(*mins)
    mults_a = (C.c_uint64 * k)(*mults)
    ranges_a = (C.c_uint64 * k)(*ranges) if ranges is not None else None
    check(lib().vb2k_normalize_keys(arr, k, mins_a, mults_a, ranges_a, 1 if nulls_invalid else 0, None, C.c_int64(n),
                                    C.c_void_p(keys.data_ptr()), C.c_void_p(_ptr(valid)), _stream()))
    return keys, valid


def join_build_array(keys: torch.Tensor, valid: Optional[torch.Tensor], capacity: int):
    """Array-mode join table over normalized keys: head[slot] = build row + 1. Returns (head, next, has_duplicates)."""
    n = keys.numel()
    head = torch.zeros(capacity, dtype=torch.int32, device="cuda")
    nxt = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    t = JoinTable(0, 0, 0, capacity, None, head.data_ptr(), nxt.data_ptr(), n)
    check(lib().vb2k_join_build(C.byref(t), C.c_void_p(keys.data_ptr()), C.c_void_p(_ptr(valid)), C.c_int64(n),
                                C.c_void_p(flags.data_ptr()), _stream()))
    return head, nxt, flags


class LikeOnAlphabet:
    """`column LIKE pattern` evaluated by the expression VM over a (small) VARCHAR alphabet."""

    def __init__(self, strings: Sequence[str], pattern: str):
        from .vector import flat_vector, VARCHAR as _V
        self.col = DeviceColumn.from_host(flat_vector(_V, list(strings)))
        self.n = len(strings)
        self.pat = torch.tensor(list(pattern.encode()), dtype=torch.uint8, device="cuda")
        self.instr = (Instr * 1)(Instr(VB2_OP_LIKE, BOOLEAN, 0, 0, 0, 0))
        self.const = (Const * 1)(Const(VARCHAR, 0, 0, 0.0, self.pat.data_ptr(), len(pattern), 0))
        self.prog = Program(self.instr, 1, 0, -1, 1, self.const, 1, 0)
        self.flags = torch.zeros(self.n + 8, dtype=torch.uint8, device="cuda")
        self.nulls = torch.zeros((self.n + 63) // 64 + 1, dtype=torch.int64, device="cuda")
        self.err = torch.zeros(2, dtype=torch.int32, device="cuda")
        self.out = (Output * 1)(Output(0, BOOLEAN, self.flags.data_ptr(), self.nulls.data_ptr()))
        self.ccol = self.col.to_c()

    def run(self) -> torch.Tensor:
        check(lib().vb2k_eval_project(C.byref(self.prog), C.byref(self.ccol), 1, None, C.c_int64(self.n), self.out, 1,
                                      C.c_void_p(self.err.data_ptr()), _stream()))
        return self.flags


class FusedScanCompact:
    """Scan -> filter -> project -> compact pipeline (signature "F:...;C:...")."""

    def __init__(self, signature: str, capacity: int):
        L = lib()
        self.id = L.vb2k_fused_find(signature.encode())
        if self.id < 0:
            raise KeyError(f"no fused pipeline for {signature}")
        self.nout = L.vb2k_fused_nproj(self.id)
        self.capacity = capacity
        self.widths = [L.vb2k_fused_output_width(self.id, i) for i in range(self.nout)]
        self.outs = [torch.empty(capacity * w, dtype=torch.uint8, device="cuda") for w in self.widths]
        self.count = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.err = torch.zeros(2, dtype=torch.int32, device="cuda")
        self.out_ptrs = (C.c_void_p * self.nout)(*[t.data_ptr() for t in self.outs])

    def run(self, cols: Sequence[torch.Tensor], rows: int, pf=(), pl=(), pi=()):
        a = FusedArgs()
        for i, t in enumerate(cols):
            a.cols[i] = t.data_ptr()
        for i, v in enumerate(pf):
            a.pf[i] = v
        for i, v in enumerate(pl):
            a.pl[i] = v
        for i, v in enumerate(pi):
            a.pi[i] = v
        a.rows = rows
        self.count.zero_()
        check(lib().vb2k_fused_scan_compact(self.id, C.byref(a), self.out_ptrs, self.nout, C.c_int64(self.capacity),
                                            C.c_void_p(self.count.data_ptr()), C.c_void_p(self.err.data_ptr()), _stream()))

    def result(self, dtypes):
        """Synchronises: returns the compacted columns as typed views of length count."""
        n = int(self.count.item())
        if int(self.err[0].item()) != 0 or n > self.capacity:
            raise RuntimeError("scan-compact output capacity exceeded")
        return n, [t.view(dt)[:n] for t, dt in zip(self.outs, dtypes)]
