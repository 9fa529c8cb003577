"""Seeded synthetic TPC-H lineitem / part columns with the value distributions of TPC-H 4.2.3
(schema velox/tpch/gen/TpchGen.cpp:278-317). Works on CPU and CUDA torch devices.

Flags are dictionary-encoded the way a Parquet/DWRF reader hands them to the operators:
int32 indices over a tiny VARCHAR alphabet (DictionaryVector, 4 B/row)."""
from __future__ import annotations

import datetime

import torch

EPOCH = datetime.date(1970, 1, 1)


def days(s: str) -> int:
    return (datetime.date.fromisoformat(s) - EPOCH).days


SHIP_LO, SHIP_HI = days("1992-01-02"), days("1998-12-01")
CURRENT = days("1995-06-17")
RETURNFLAG_DICT = ["A", "N", "R"]
LINESTATUS_DICT = ["F", "O"]
_S1 = ["STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"]
_S2 = ["ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"]
_S3 = ["TIN", "NICKEL", "BRASS", "STEEL", "COPPER"]
PTYPE_DICT = [f"{a} {b} {c}" for a in _S1 for b in _S2 for c in _S3]  # 150 values

LINEITEM_ROWS_PER_SF = 6_000_379  # SF100 = 600 037 902
PART_ROWS_PER_SF = 200_000


def gen_lineitem(rows: int, nparts: int, seed: int = 42, device="cpu", chunk: int = 1 << 26):
    """Returns dict of tensors. Generated in chunks so SF100 fits comfortably while generating."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = {
        "l_quantity": torch.empty(rows, dtype=torch.float64, device=device),
        "l_extendedprice": torch.empty(rows, dtype=torch.float64, device=device),
        "l_discount": torch.empty(rows, dtype=torch.float64, device=device),
        "l_tax": torch.empty(rows, dtype=torch.float64, device=device),
        "l_shipdate": torch.empty(rows, dtype=torch.int32, device=device),
        "l_returnflag": torch.empty(rows, dtype=torch.int32, device=device),
        "l_linestatus": torch.empty(rows, dtype=torch.int32, device=device),
        "l_partkey": torch.empty(rows, dtype=torch.int64, device=device),
    }
    for r0 in range(0, rows, chunk):
        n = min(chunk, rows - r0)
        s = slice(r0, r0 + n)
        qty = torch.randint(1, 51, (n,), generator=g, device=device).to(torch.float64)
        price = torch.randint(90000, 210001, (n,), generator=g, device=device).to(torch.float64) / 100.0
        out["l_quantity"][s] = qty
        out["l_extendedprice"][s] = torch.round(qty * price * 100.0) / 100.0
        out["l_discount"][s] = torch.randint(0, 11, (n,), generator=g, device=device).to(torch.float64) / 100.0
        out["l_tax"][s] = torch.randint(0, 9, (n,), generator=g, device=device).to(torch.float64) / 100.0
        ship = torch.randint(SHIP_LO, SHIP_HI + 1, (n,), generator=g, device=device, dtype=torch.int32)
        receipt = ship + torch.randint(1, 31, (n,), generator=g, device=device, dtype=torch.int32)
        ra = torch.randint(0, 2, (n,), generator=g, device=device, dtype=torch.int32) * 2  # A=0 / R=2
        out["l_shipdate"][s] = ship
        out["l_returnflag"][s] = torch.where(receipt <= CURRENT, ra, torch.ones_like(ra))
        out["l_linestatus"][s] = (ship > CURRENT).to(torch.int32)
        out["l_partkey"][s] = torch.randint(1, nparts + 1, (n,), generator=g, device=device)
    return out


def gen_part(nparts: int, seed: int = 43, device="cpu"):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return {
        "p_partkey": torch.arange(1, nparts + 1, dtype=torch.int64, device=device),
        "p_type": torch.randint(0, len(PTYPE_DICT), (nparts,), generator=g, device=device, dtype=torch.int32),
    }


# Query constants (velox/exec/tests/utils/TpchQueryBuilder.cpp:218,764-776,1650-1651)
Q1_SHIPDATE_LT = days("1998-09-03")
Q6_SHIP_LO, Q6_SHIP_HI = days("1994-01-01"), days("1994-12-31")
Q14_SHIP_LO, Q14_SHIP_HI = days("1995-09-01"), days("1995-09-30")

Q6_SIG = "F:and(between(i0,pi0,pi1),between(f1,pf0,pf1),lt(f2,pf2));P:multiply(f3,f1)"
Q1_SIG = ("F:lt(i0,pi0);P:f1|f2|multiply(f2,minus(pf0,f3))|"
          "multiply(multiply(f2,minus(pf1,f3)),plus(pf2,f4))|f3")
Q14_SIG = ("F:between(i0,pi0,pi1);P:multiply(f2,minus(pf0,f3))|"
           "switch(joinflag,multiply(f2,minus(pf1,f3)),pf2);J:l1")

# Algorithmic bytes per lineitem row: each referenced column once at stored width.
Q6_BYTES_PER_ROW = 4 + 8 + 8 + 8
Q1_BYTES_PER_ROW = 4 * 8 + 4 + 2 * 4
Q14_BYTES_PER_ROW = 8 + 8 + 8 + 4
