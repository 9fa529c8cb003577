"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/libtpchref.so: the reference's own vendored
TPC-H dbgen (velox/tpch/gen/dbgen/*.cpp) compiled where it lies (oracle/build_ref.sh) behind the
wrapper oracle/dbgen_wrap.cpp. Produces exactly the lineitem / part columns Velox's TPC-H
connector produces (velox/tpch/gen/TpchGen.cpp:402-534)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libtpchref.so")
ORDERS_PER_SF = 1_500_000
PARTS_PER_SF = 200_000


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _lib():
    L = C.CDLL(LIB_PATH)
    L.ref_gen_lineitem.restype = C.c_int64
    return L


def gen_lineitem(scale: float, order_offset: int = 0, n_orders: int = None):
    L = _lib()
    if n_orders is None:
        n_orders = int(ORDERS_PER_SF * scale) - order_offset
    cap = 7 * n_orders
    cols = {"l_orderkey": np.zeros(cap, np.int64), "l_partkey": np.zeros(cap, np.int64), "l_quantity": np.zeros(cap), "l_extendedprice": np.zeros(cap),
            "l_discount": np.zeros(cap), "l_tax": np.zeros(cap), "l_returnflag": np.zeros(cap, np.uint8), "l_linestatus": np.zeros(cap, np.uint8),
            "l_shipdate": np.zeros(cap, np.int32)}
    n = L.ref_gen_lineitem(C.c_double(scale), C.c_int64(order_offset), C.c_int64(n_orders), *[C.c_void_p(a.ctypes.data) for a in cols.values()])
    return {k: v[:n].copy() for k, v in cols.items()}


def gen_part(scale: float, offset: int = 0, n: int = None):
    L = _lib()
    if n is None:
        n = int(PARTS_PER_SF * scale) - offset
    pk = np.zeros(n, np.int64)
    ty = np.zeros(n * 26, np.uint8)
    L.ref_gen_part(C.c_double(scale), C.c_int64(offset), C.c_int64(n), C.c_void_p(pk.ctypes.data), C.c_void_p(ty.ctypes.data))
    types = [bytes(ty[i * 26:(i + 1) * 26]).split(b"\0")[0].decode() for i in range(n)]
    return {"p_partkey": pk, "p_type": types}
