"""NCCL communicator of the exchange path (include/velox_b200.h vb2_comm_*): one process per GPU;
the ncclUniqueId travels over torch.distributed (plumbing), the data path is ours."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from ._lib import VeloxRuntimeError, lib


class Comm:
    def __init__(self):
        assert dist.is_initialized()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        L = lib()
        L.vb2_comm_create.restype = C.c_void_p
        L.vb2_comm_create.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            if L.vb2_comm_unique_id(buf) != 0:
                raise VeloxRuntimeError("ncclGetUniqueId failed")
            uid.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().tolist())
        err = C.create_string_buffer(512)
        self.h = L.vb2_comm_create(raw, self.world, self.rank, err, 512)
        if not self.h:
            raise VeloxRuntimeError(err.value.decode())
        self.L = L

    @property
    def peer_memory(self) -> bool:
        """True when rows move between the ranks through CUDA-IPC mapped peer memory (exchange_p2p.cu)."""
        return bool(self.L.vb2_comm_peer_memory(C.c_void_p(self.h)))

    def exchanges(self):
        self.L.vb2_comm_exchanges.restype = C.c_int64
        return {"peer_memory": int(self.L.vb2_comm_exchanges(C.c_void_p(self.h), 1)), "nccl": int(self.L.vb2_comm_exchanges(C.c_void_p(self.h), 0))}

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def exchange_counts(self, send_counts):
        w = self.world
        s = (C.c_int64 * w)(*send_counts)
        r = (C.c_int64 * w)()
        rc = self.L.vb2_comm_exchange_counts(C.c_void_p(self.h), s, r, self._st())
        if rc:
            raise VeloxRuntimeError("count exchange failed")
        return list(r)

    def exchange_counts_dev(self, send_counts_dev: torch.Tensor):
        """send counts on the device -> (send_counts, recv_counts) host lists, one synchronisation."""
        w = self.world
        s = (C.c_int64 * w)()
        r = (C.c_int64 * w)()
        rc = self.L.vb2_comm_exchange_counts_dev(C.c_void_p(self.h), C.c_void_p(send_counts_dev.data_ptr()), s, r, self._st())
        if rc:
            raise VeloxRuntimeError("count exchange failed")
        return list(s), list(r)

    def all_to_all(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        """send: elements grouped by destination rank (counts in elements)."""
        w = self.world
        out = torch.empty(max(1, sum(recv_counts)), dtype=send.dtype, device="cuda")
        s = (C.c_int64 * w)(*send_counts)
        r = (C.c_int64 * w)(*recv_counts)
        rc = self.L.vb2_comm_all_to_all(C.c_void_p(self.h), C.c_void_p(send.data_ptr()), s, C.c_void_p(out.data_ptr()), r,
                                        send.element_size(), self._st())
        if rc:
            raise VeloxRuntimeError("all-to-all failed")
        return out[:sum(recv_counts)]

    def all_to_all_columns(self, sends, send_counts, recv_counts):
        """Exchanges several columns of the same row partitioning in one NCCL group."""
        w, n = self.world, len(sends)
        total = sum(recv_counts)
        outs = [torch.empty(max(1, total), dtype=t.dtype, device="cuda") for t in sends]
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in sends])
        rp = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
        eb = (C.c_int32 * n)(*[t.element_size() for t in sends])
        s = (C.c_int64 * w)(*send_counts)
        r = (C.c_int64 * w)(*recv_counts)
        rc = self.L.vb2_comm_all_to_all_columns(C.c_void_p(self.h), n, sp, rp, eb, s, r, self._st())
        if rc:
            raise VeloxRuntimeError("all-to-all failed")
        return [o[:total] for o in outs]

    def all_reduce_(self, t: torch.Tensor):
        fn = self.L.vb2_comm_all_reduce_f64 if t.dtype == torch.float64 else self.L.vb2_comm_all_reduce_i64
        rc = fn(C.c_void_p(self.h), C.c_void_p(t.data_ptr()), C.c_int64(t.numel()), self._st())
        if rc:
            raise VeloxRuntimeError("all-reduce failed")
        return t

    def all_reduce_max_(self, t: torch.Tensor):
        assert t.dtype == torch.int64
        if self.L.vb2_comm_all_reduce_max_i64(C.c_void_p(self.h), C.c_void_p(t.data_ptr()), C.c_int64(t.numel()), self._st()):
            raise VeloxRuntimeError("all-reduce failed")
        return t

    def close(self):
        if self.h:
            self.L.vb2_comm_free.argtypes = [C.c_void_p]
            self.L.vb2_comm_free(self.h)
            self.h = None
