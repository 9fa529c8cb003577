"""Generates tests/golden/tpch_sf001.npz: TPC-H lineitem (first 15 000 orders = SF0.01) and part
(2 000 rows) columns from the reference's own dbgen (oracle/_ref/libtpchref.so, built by
oracle/build_ref.sh from /root/reference). Run in the authoring container; the GPU box only
reads the committed fixture."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tpch_ref  # noqa: E402

li = tpch_ref.gen_lineitem(0.01)
pt = tpch_ref.gen_part(0.01)
types, codes = np.unique(np.array(pt["p_type"]), return_inverse=True)
out = os.path.join(ROOT, "tests", "golden", "tpch_sf001.npz")
np.savez_compressed(out, p_partkey=pt["p_partkey"], p_type_codes=codes.astype(np.int32), p_type_dict=types, **li)
print(out, len(li["l_orderkey"]), "lineitem rows,", len(pt["p_partkey"]), "part rows,", os.path.getsize(out), "bytes")
