"""Test/bench front-end: a PlanBuilder that mirrors the reference's test helper
(velox/exec/tests/utils/PlanBuilder.h:673-1370) and emits the plan text understood by both the
product library (`vb2_task_create`) and the CPU oracle.

The reference parses SQL expression strings with DuckDB (velox/parse, velox/duckdb — absent here
and out of scope); this module holds a small recursive-descent parser for the subset the hot
path's plans use (TPC-H Q1/Q6/Q14: arithmetic, comparisons, BETWEEN, LIKE, AND/OR/NOT,
CASE WHEN, CAST, IS NULL, DATE literals). It is a front-end only: no data flows through it.

Plan text grammar (S-expressions)
  plan := (values SRC (TYPE ...)) | (filter EXPR plan) | (project (EXPR ...) plan)
        | (aggregation STEP (keys I ...) (aggs AGG ...) plan)      STEP: single|partial|intermediate|final
        | (hashjoin TYPE (probekeys I ...) (buildkeys I ...) EXPR|nil (out (p I)|(b I) ...) probe build)
        | (orderby ((I asc|desc first|last) ...) plan) | (topn N ((I asc|desc first|last) ...) plan)
        | (exchange partitioned|broadcast|gather (keys I ...) plan)  PartitionedOutput -> Exchange across the ranks
  AGG  := (sum I [(mask I)]) | (avg I) | (count [I]) | (min I) | (max I)
  EXPR := (field I) | (f64 X) | (i64 N) | (i32 N) | (bool true|false) | (str "s") | (null TYPE)
        | (cast TYPE e) | (and e ...) | (or e ...) | (switch c1 v1 ... [else]) | (NAME e ...)
  NAME follows the reference's function names: plus minus multiply divide modulus negate
        lt lte gt gte eq neq between like not is_null.
"""
from __future__ import annotations

import datetime
import re
from dataclasses import dataclass
from typing import List, Optional, Sequence

from .vector import BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR, TYPE_NAMES

_EPOCH = datetime.date(1970, 1, 1)


def date_to_days(s: str) -> int:
    return (datetime.date.fromisoformat(s) - _EPOCH).days


@dataclass
class E:
    """Typed expression node."""
    sexpr: str
    type: int
    is_literal: bool = False
    literal: object = None
    field: Optional[int] = None


_TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+|\d+[eE][-+]?\d+)|(\d+)|'((?:[^']|'')*)'|"
                    r"(<>|!=|<=|>=|::|[-+*/%(),<>=])|([A-Za-z_][A-Za-z_0-9.]*))")

_KEYWORDS = {"and", "or", "not", "between", "like", "is", "null", "case", "when", "then", "else", "end",
             "cast", "as", "true", "false", "date", "if"}
_TYPES = {"boolean": BOOLEAN, "integer": INTEGER, "int": INTEGER, "date": INTEGER, "bigint": BIGINT,
          "double": DOUBLE, "varchar": VARCHAR}
_CMP = {"=": "eq", "<>": "neq", "!=": "neq", "<": "lt", "<=": "lte", ">": "gt", ">=": "gte"}
_RANK = {INTEGER: 0, BIGINT: 1, DOUBLE: 2}


def _lit(type_: int, value) -> E:
    if value is None:
        return E(f"(null {TYPE_NAMES[type_]})", type_, True, None)
    if type_ == DOUBLE:
        return E(f"(f64 {float(value)!r})", DOUBLE, True, float(value))
    if type_ == BIGINT:
        return E(f"(i64 {int(value)})", BIGINT, True, int(value))
    if type_ == INTEGER:
        return E(f"(i32 {int(value)})", INTEGER, True, int(value))
    if type_ == BOOLEAN:
        return E(f"(bool {'true' if value else 'false'})", BOOLEAN, True, bool(value))
    s = str(value).replace("\\", "\\\\").replace('"', '\\"')
    return E(f'(str "{s}")', VARCHAR, True, str(value))


def _cast(e: E, to: int) -> E:
    if e.type == to:
        return e
    if e.is_literal and e.literal is not None or (e.is_literal and e.literal is None):
        if e.literal is None:
            return _lit(to, None)
        if to in _RANK and e.type in _RANK:
            return _lit(to, e.literal)
    return E(f"(cast {TYPE_NAMES[to]} {e.sexpr})", to)


def _unify(args: List[E]) -> List[E]:
    """Implicit numeric widening the way the reference's SQL front-end resolves signatures:
    literals adopt the column's type, otherwise widen INTEGER -> BIGINT -> DOUBLE."""
    types = {a.type for a in args}
    if len(types) == 1:
        return args
    if not all(t in _RANK for t in types):
        # VARCHAR literal compared with DATE column: 'yyyy-mm-dd' -> days
        out = []
        non_str = [a.type for a in args if a.type != VARCHAR]
        if non_str and all(t == INTEGER for t in non_str):
            for a in args:
                out.append(_lit(INTEGER, date_to_days(a.literal)) if a.type == VARCHAR and a.is_literal else a)
            return out
        raise ValueError(f"cannot unify types {[TYPE_NAMES[t] for t in types]}")
    non_lit = [a.type for a in args if not a.is_literal]
    if non_lit:
        target = max(non_lit, key=lambda t: _RANK[t])
        # a fractional literal forces DOUBLE
        for a in args:
            if a.is_literal and a.type == DOUBLE:
                target = DOUBLE
    else:
        target = max(types, key=lambda t: _RANK[t])
    return [_cast(a, target) for a in args]


class _Parser:
    def __init__(self, text: str, names: Sequence[str], types: Sequence[int]):
        self.toks = []
        pos = 0
        text = text.strip()
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if not m:
                raise ValueError(f"cannot tokenize {text[pos:]!r}")
            pos = m.end()
            if m.group(1) is not None:
                self.toks.append(("float", m.group(1)))
            elif m.group(2) is not None:
                self.toks.append(("int", m.group(2)))
            elif m.group(3) is not None:
                self.toks.append(("str", m.group(3).replace("''", "'")))
            elif m.group(4) is not None:
                self.toks.append(("op", m.group(4)))
            else:
                w = m.group(5)
                self.toks.append(("kw", w.lower()) if w.lower() in _KEYWORDS else ("id", w))
        self.i = 0
        self.names = list(names)
        self.types = list(types)

    def peek(self, k=0):
        return self.toks[self.i + k] if self.i + k < len(self.toks) else ("eof", "")

    def take(self):
        t = self.peek()
        self.i += 1
        return t

    def accept(self, kind, val=None):
        t = self.peek()
        if t[0] == kind and (val is None or t[1] == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        if not self.accept(kind, val):
            raise ValueError(f"expected {val or kind}, got {self.peek()}")

    # precedence climbing
    def parse(self) -> E:
        e = self.p_or()
        return e

    def p_or(self):
        args = [self.p_and()]
        while self.accept("kw", "or"):
            args.append(self.p_and())
        return args[0] if len(args) == 1 else E("(or " + " ".join(a.sexpr for a in args) + ")", BOOLEAN)

    def p_and(self):
        args = [self.p_not()]
        while self.accept("kw", "and"):
            args.append(self.p_not())
        return args[0] if len(args) == 1 else E("(and " + " ".join(a.sexpr for a in args) + ")", BOOLEAN)

    def p_not(self):
        if self.accept("kw", "not"):
            return E(f"(not {self.p_not().sexpr})", BOOLEAN)
        return self.p_cmp()

    def p_cmp(self):
        left = self.p_add()
        t = self.peek()
        if t[0] == "op" and t[1] in _CMP:
            self.take()
            right = self.p_add()
            a, b = _unify([left, right])
            return E(f"({_CMP[t[1]]} {a.sexpr} {b.sexpr})", BOOLEAN)
        negate = False
        if t == ("kw", "not") and self.peek(1) in (("kw", "between"), ("kw", "like")):
            self.take()
            negate = True
            t = self.peek()
        if t == ("kw", "between"):
            self.take()
            lo = self.p_add()
            self.expect("kw", "and")
            hi = self.p_add()
            a, b, c = _unify([left, lo, hi])
            e = E(f"(between {a.sexpr} {b.sexpr} {c.sexpr})", BOOLEAN)
            return E(f"(not {e.sexpr})", BOOLEAN) if negate else e
        if t == ("kw", "like"):
            self.take()
            pat = self.p_add()
            e = E(f"(like {left.sexpr} {pat.sexpr})", BOOLEAN)
            return E(f"(not {e.sexpr})", BOOLEAN) if negate else e
        if t == ("kw", "is"):
            self.take()
            neg = self.accept("kw", "not")
            self.expect("kw", "null")
            e = E(f"(is_null {left.sexpr})", BOOLEAN)
            return E(f"(not {e.sexpr})", BOOLEAN) if neg else e
        return left

    def p_add(self):
        left = self.p_mul()
        while self.peek()[0] == "op" and self.peek()[1] in "+-":
            op = self.take()[1]
            right = self.p_mul()
            a, b = _unify([left, right])
            left = E(f"({'plus' if op == '+' else 'minus'} {a.sexpr} {b.sexpr})", a.type)
        return left

    def p_mul(self):
        left = self.p_unary()
        while self.peek()[0] == "op" and self.peek()[1] in "*/%":
            op = self.take()[1]
            right = self.p_unary()
            a, b = _unify([left, right])
            name = {"*": "multiply", "/": "divide", "%": "modulus"}[op]
            left = E(f"({name} {a.sexpr} {b.sexpr})", a.type)
        return left

    def p_unary(self):
        if self.accept("op", "-"):
            e = self.p_unary()
            if e.is_literal and e.literal is not None:
                return _lit(e.type, -e.literal)
            return E(f"(negate {e.sexpr})", e.type)
        self.accept("op", "+")
        return self.p_postfix()

    def p_postfix(self):
        e = self.p_primary()
        while self.accept("op", "::"):
            t = self.take()
            ty = _TYPES[t[1].lower()]
            if t[1].lower() == "date" and e.is_literal and e.type == VARCHAR:
                e = _lit(INTEGER, date_to_days(e.literal))
            else:
                e = _cast(e, ty)
        return e

    def p_type(self):
        t = self.take()
        return t[1].lower()

    def p_primary(self):
        t = self.take()
        if t[0] == "float":
            return _lit(DOUBLE, float(t[1]))
        if t[0] == "int":
            return _lit(BIGINT, int(t[1]))
        if t[0] == "str":
            return _lit(VARCHAR, t[1])
        if t == ("op", "("):
            e = self.p_or()
            self.expect("op", ")")
            return e
        if t == ("kw", "true") or t == ("kw", "false"):
            return _lit(BOOLEAN, t[1] == "true")
        if t == ("kw", "null"):
            return _lit(BIGINT, None)
        if t == ("kw", "date"):
            s = self.take()
            return _lit(INTEGER, date_to_days(s[1]))
        if t == ("kw", "cast"):
            self.expect("op", "(")
            e = self.p_or()
            self.expect("kw", "as")
            tn = self.p_type()
            self.expect("op", ")")
            if tn == "date" and e.is_literal and e.type == VARCHAR:
                return _lit(INTEGER, date_to_days(e.literal))
            return _cast(e, _TYPES[tn])
        if t == ("kw", "case"):
            parts = []
            while self.accept("kw", "when"):
                c = self.p_or()
                self.expect("kw", "then")
                parts.append((c, self.p_or()))
            els = self.p_or() if self.accept("kw", "else") else None
            self.expect("kw", "end")
            vals = [v for _, v in parts] + ([els] if els is not None else [])
            vals = _unify(vals)
            out = []
            for (c, _), v in zip(parts, vals):
                out += [c.sexpr, v.sexpr]
            if els is not None:
                out.append(vals[-1].sexpr)
            return E("(switch " + " ".join(out) + ")", vals[0].type)
        if t == ("kw", "if"):
            self.expect("op", "(")
            c = self.p_or()
            self.expect("op", ",")
            a = self.p_or()
            self.expect("op", ",")
            b = self.p_or()
            self.expect("op", ")")
            a, b = _unify([a, b])
            return E(f"(switch {c.sexpr} {a.sexpr} {b.sexpr})", a.type)
        if t[0] == "id":
            if self.peek() == ("op", "("):
                if t[1] not in USER_FUNCTIONS:
                    raise ValueError(f"unsupported function {t[1]}")
                ret, arg_types = USER_FUNCTIONS[t[1]]
                self.take()
                args = []
                while True:
                    args.append(self.p_or())
                    if self.accept("op", ")"):
                        break
                    self.expect("op", ",")
                if len(args) != len(arg_types):
                    raise ValueError(f"{t[1]} takes {len(arg_types)} arguments")
                args = [_cast(a, ty) for a, ty in zip(args, arg_types)]
                return E(f"({t[1]} " + " ".join(a.sexpr for a in args) + ")", ret)
            if t[1] not in self.names:
                raise ValueError(f"unknown column {t[1]} (have {self.names})")
            i = self.names.index(t[1])
            return E(f"(field {i})", self.types[i], field=i)
        raise ValueError(f"unexpected token {t}")


def parse_expr(text: str, names, types):
    """Returns (E, alias or None)."""
    m = re.match(r"^(.*?)\s+[aA][sS]\s+([A-Za-z_][A-Za-z_0-9]*)\s*$", text, re.S)
    alias = None
    if m and not re.search(r"cast\s*\([^)]*$", m.group(1), re.I):
        text, alias = m.group(1), m.group(2)
    p = _Parser(text, names, types)
    e = p.parse()
    if p.peek()[0] != "eof":
        raise ValueError(f"trailing tokens in {text!r}: {p.peek()}")
    return e, alias


# Functions the application registered with the engine (register_scalar_function /
# register_aggregate_function below): name -> (return type, argument types) and
# name -> (family, input function, final function). The front-end types calls from these.
USER_FUNCTIONS: dict = {}
USER_AGGREGATES: dict = {}
_BUILTIN_BOOL = {"lt", "lte", "gt", "gte", "eq", "neq", "between", "like", "not", "is_null"}


def register_scalar_function(name: str, ret_type: int, arg_types: Sequence[int], cuda_source: str, entry: Optional[str] = None) -> None:
    """exec::registerVectorFunction for a device function given as CUDA source text
    (`__device__ RET entry(ARGS...)`; include/velox_b200.h vb2_register_scalar_function)."""
    import ctypes as C
    from ._lib import check, lib
    at = (C.c_int32 * len(arg_types))(*arg_types)
    err = C.create_string_buffer(1024)
    rc = lib().vb2_register_scalar_function(name.encode(), (entry or name).encode(), cuda_source.encode(), int(ret_type), at, len(arg_types), err, 1024)
    if rc:
        from ._lib import VeloxRuntimeError
        raise VeloxRuntimeError(err.value.decode(errors="replace"))
    USER_FUNCTIONS[name] = (int(ret_type), [int(t) for t in arg_types])


def _scalar_return_type(name: str, arg_type: int) -> int:
    if name in USER_FUNCTIONS:
        return USER_FUNCTIONS[name][0]
    return BOOLEAN if name in _BUILTIN_BOOL else arg_type


def register_aggregate_function(name: str, family: str, input_function: str = "", final_function: str = "") -> None:
    """exec::registerAggregateFunction: name(x) = final_function(FAMILY(input_function(x))), FAMILY one of
    sum avg count min max (include/velox_b200.h vb2_register_aggregate_function)."""
    import ctypes as C
    from ._lib import lib, VeloxRuntimeError
    err = C.create_string_buffer(1024)
    rc = lib().vb2_register_aggregate_function(name.encode(), family.encode(), input_function.encode(), final_function.encode(), err, 1024)
    if rc:
        raise VeloxRuntimeError(err.value.decode(errors="replace"))
    USER_AGGREGATES[name] = (family, input_function, final_function)


_AGG = re.compile(r"^\s*([A-Za-z_][A-Za-z_0-9]*)\s*\(\s*((?:[dD][iI][sS][tT][iI][nN][cC][tT]\s+)?)([A-Za-z_0-9*]*)\s*\)\s*(?:[aA][sS]\s+([A-Za-z_][A-Za-z_0-9]*))?\s*$")


@dataclass
class _Node:
    sexpr: str
    names: List[str]
    types: List[int]
    # for finalAggregation(): how to merge the partial below
    partial: Optional[dict] = None


class PlanBuilder:
    """Mirrors velox/exec/tests/utils/PlanBuilder.h. `values()` declares a source (default id 0;
    join build sides name theirs explicitly) fed at run time through the `sources` list."""

    def __init__(self):
        self.node: Optional[_Node] = None
        self.sources: List[int] = []

    def values(self, names, types, source: int = 0) -> "PlanBuilder":
        self.sources.append(source)
        ts = " ".join(TYPE_NAMES[t] for t in types)
        self.node = _Node(f"(values {source} ({ts}))", list(names), list(types))
        return self

    # tableScan with filters pushed into the scan is, result-wise, values + filter.
    def filter(self, text: str) -> "PlanBuilder":
        n = self.node
        e, _ = parse_expr(text, n.names, n.types)
        if e.type != BOOLEAN:
            raise ValueError("filter must be BOOLEAN")
        self.node = _Node(f"(filter {e.sexpr} {n.sexpr})", n.names, n.types)
        return self

    def project(self, exprs: Sequence[str]) -> "PlanBuilder":
        n = self.node
        out, names, types = [], [], []
        for i, text in enumerate(exprs):
            e, alias = parse_expr(text, n.names, n.types)
            out.append(e.sexpr)
            names.append(alias or (n.names[e.field] if e.field is not None else f"p{i}"))
            types.append(e.type)
        self.node = _Node(f"(project ({' '.join(out)}) {n.sexpr})", names, types)
        return self

    def _aggregation(self, step: str, keys, aggs, masks=None) -> "PlanBuilder":
        n = self.node
        key_idx = [n.names.index(k) for k in keys]
        names = list(keys)
        types = [n.types[i] for i in key_idx]
        specs, partial_specs = [], []
        for j, a in enumerate(aggs):
            m = _AGG.match(a)
            if not m:
                raise ValueError(f"bad aggregate {a!r}")
            name, distinct, arg, alias = m.group(1), bool(m.group(2)), m.group(3), m.group(4)
            if distinct and (step != "single" or name not in ("sum", "avg", "count", "min", "max") or arg in ("", "*") or arg.isdigit()):
                raise ValueError(f"{a!r}: DISTINCT applies to sum / avg / count / min / max over a column in a single aggregation")
            if name not in ("sum", "avg", "count", "min", "max") and name not in USER_AGGREGATES:
                raise ValueError(f"unknown aggregate {name!r} (register_aggregate_function)")
            fn, in_fn, fin_fn = USER_AGGREGATES.get(name, (name, "", ""))  # the plan carries the registered name, typing follows its family
            alias = alias or f"a{j}"
            mask = masks[j] if masks else None
            mask_s = (f" (mask {n.names.index(mask)})" if mask else "") + (" (distinct)" if distinct else "")
            if fn == "count" and (arg in ("", "*") or arg.isdigit()):
                specs.append(f"({name}{mask_s})")
                in_type = BIGINT
            else:
                col = n.names.index(arg)
                in_type = _scalar_return_type(in_fn, n.types[col]) if in_fn else n.types[col]
                specs.append(f"({name} {col}{mask_s})")
            partial_specs.append((fn, alias, in_type, name))
            if fin_fn and step in ("single", "final"):
                family_type = DOUBLE if fn == "avg" else BIGINT if fn == "count" else (DOUBLE if in_type == DOUBLE else BIGINT) if fn == "sum" else in_type
                names.append(alias); types.append(_scalar_return_type(fin_fn, family_type))
            elif fn == "avg":
                if step in ("single", "final"):
                    names.append(alias); types.append(DOUBLE)
                else:
                    names += [alias + "_sum", alias + "_count"]; types += [DOUBLE, BIGINT]
            elif fn == "count":
                names.append(alias); types.append(BIGINT)
            elif fn == "sum":
                names.append(alias); types.append(DOUBLE if in_type == DOUBLE else BIGINT)
            else:
                names.append(alias); types.append(in_type)
        sexpr = (f"(aggregation {step} (keys {' '.join(map(str, key_idx))}) "
                 f"(aggs {' '.join(specs)}) {n.sexpr})")
        self.node = _Node(sexpr, names, types, partial={"keys": list(keys), "aggs": partial_specs})
        return self

    def singleAggregation(self, keys, aggs, masks=None):
        return self._aggregation("single", keys, aggs, masks)

    def partialAggregation(self, keys, aggs, masks=None):
        return self._aggregation("partial", keys, aggs, masks)

    def _merge(self, step: str) -> "PlanBuilder":
        n = self.node
        if not n.partial:
            raise ValueError(f"{step}Aggregation() must follow a partial/intermediate aggregation")
        keys = n.partial["keys"]
        nk = len(keys)
        specs, names, types = [], list(keys), n.types[:nk]
        c = nk
        for fn, alias, _, name in n.partial["aggs"]:
            specs.append(f"({name} {c})")
            fin_fn = USER_AGGREGATES.get(name, ("", "", ""))[2]
            if fin_fn and step == "final":
                family_type = DOUBLE if fn == "avg" else n.types[c]
                names.append(alias); types.append(_scalar_return_type(fin_fn, family_type))
                c += 2 if fn == "avg" else 1
            elif fn == "avg":
                if step == "final":
                    names.append(alias); types.append(DOUBLE)
                else:
                    names += [alias + "_sum", alias + "_count"]; types += [DOUBLE, BIGINT]
                c += 2
            else:
                names.append(alias); types.append(n.types[c])
                c += 1
        sexpr = (f"(aggregation {step} (keys {' '.join(map(str, range(nk)))}) "
                 f"(aggs {' '.join(specs)}) {n.sexpr})")
        self.node = _Node(sexpr, names, types, partial=n.partial)
        return self

    def finalAggregation(self):
        return self._merge("final")

    def intermediateAggregation(self):
        return self._merge("intermediate")

    def localPartition(self, keys=()):
        """Gathers the outputs of the drivers of the pipeline below (exec/LocalPartition.cpp, gather form).
        With task.max_drivers = 1 (the default) one driver runs both sides and the node is the identity;
        above 1 it is the boundary between the N-driver pipeline below and the single consumer above."""
        n = self.node
        self.node = _Node(f"(localpartition {n.sexpr})", n.names, n.types, partial=n.partial)
        return self

    def _exchange(self, kind: str, keys=()) -> "PlanBuilder":
        n = self.node
        idx = " ".join(str(n.names.index(k)) for k in keys)
        self.node = _Node(f"(exchange {kind} (keys {idx}) {n.sexpr})", n.names, n.types, partial=n.partial)
        return self

    def partitionedOutput(self, keys) -> "PlanBuilder":
        """PartitionedOutput (hash(keys) % world, HashPartitionFunction) followed by the Exchange that
        reads this rank's partition (PlanBuilder::partitionedOutput + exchange of the reference's
        multi-fragment tests, velox/exec/tests/MultiFragmentTest.cpp)."""
        return self._exchange("partitioned", keys)

    def partitionedOutputBroadcast(self) -> "PlanBuilder":
        return self._exchange("broadcast")

    def gatherExchange(self) -> "PlanBuilder":
        """Every rank's rows to partition 0 (a PartitionedOutput with one partition)."""
        return self._exchange("gather")

    def hashJoin(self, leftKeys, rightKeys, build: "PlanBuilder", filter: str, output: Sequence[str],
                 joinType: str = "inner") -> "PlanBuilder":
        p, b = self.node, build.node
        pk = [p.names.index(k) for k in leftKeys]
        bk = [b.names.index(k) for k in rightKeys]
        filt = "nil"
        if filter:
            e, _ = parse_expr(filter, p.names + b.names, p.types + b.types)
            filt = e.sexpr
        outs, names, types = [], [], []
        for o in output:
            if o in p.names:
                i = p.names.index(o)
                outs.append(f"(p {i})"); types.append(p.types[i])
            else:
                i = b.names.index(o)
                outs.append(f"(b {i})"); types.append(b.types[i])
            names.append(o)
        sexpr = (f"(hashjoin {joinType} (probekeys {' '.join(map(str, pk))}) (buildkeys {' '.join(map(str, bk))}) "
                 f"{filt} (out {' '.join(outs)}) {p.sexpr} {b.sexpr})")
        self.sources += build.sources
        self.node = _Node(sexpr, names, types)
        return self

    def orderBy(self, keys: Sequence[str], limit: Optional[int] = None) -> "PlanBuilder":
        """ORDER BY (exec/OrderBy.cpp; PlanBuilder::orderBy of the reference's test utilities): keys
        like "c0", "c1 DESC", "c2 ASC NULLS FIRST". Default NULLS LAST for both directions, as
        core::kAscNullsLast / kDescNullsLast. Runs as B200OrderBy (csrc/host/orderby.cpp, csrc/sort.cu)."""
        n = self.node
        parts = []
        for k in keys:
            words = k.split()
            col = n.names.index(words[0])
            rest = [w.upper() for w in words[1:]]
            asc = "DESC" not in rest
            nulls_first = "FIRST" in rest
            parts.append(f"({col} {'asc' if asc else 'desc'} {'first' if nulls_first else 'last'})")
        head = f"topn {int(limit)}" if limit is not None else "orderby"
        self.node = _Node(f"({head} ({' '.join(parts)}) {n.sexpr})", n.names, n.types)
        return self

    def topN(self, keys: Sequence[str], count: int) -> "PlanBuilder":
        """ORDER BY keys LIMIT count (exec/TopN.cpp; PlanBuilder::topN of the reference's test utilities)."""
        return self.orderBy(keys, limit=count)

    def planNode(self) -> _Node:
        return self.node
