"""The shim headers under velox_b200/abi/ restate the reference's operator / plan-node / registry
interface; this test diffs the restated declarations against the reference's own headers.

For every row of SIGNATURES the parameter TYPE list of the declaration is extracted from the
reference header and from the shim (names, defaults, const and & dropped, namespace qualifiers
dropped) and the two lists must be equal. STRUCTS does the same for the member lists of plain
structs. Rows in DEVIATIONS are the documented differences; each must be described in
INTEGRATION.md under its key, so a deviation cannot exist silently.

/root/reference only exists in the build container: the comparison is skipped elsewhere, the
shim-side extraction still runs (a renamed or dropped declaration fails the test anywhere).
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/velox"
SHIM_EXEC = os.path.join(ROOT, "velox_b200", "abi", "exec_abi.h")
SHIM_VEC = os.path.join(ROOT, "velox_b200", "abi", "vector_abi.h")


def _strip_comments(t):
    t = re.sub(r"/\*.*?\*/", "", t, flags=re.S)
    return re.sub(r"//[^\n]*", "", t)


_cache = {}


def _text(path):
    if path not in _cache:
        _cache[path] = _strip_comments(open(path).read())
    return _cache[path]


def _balanced(text, i):
    """text[i] is an opening bracket; returns the index of its partner."""
    d = 0
    j = i
    while True:
        c = text[j]
        if c in "(<[{":
            d += 1
        elif c in ")>]}":
            d -= 1
            if d == 0:
                return j
        elif c == "-" and text[j + 1] == ">":
            j += 1  # '->' is not a bracket
        j += 1


def _split_top(body, sep):
    parts, d, cur = [], 0, ""
    for c in body:
        if c in "(<[{":
            d += 1
        elif c in ")>]}":
            d -= 1
        if c == sep and d == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += c
    if cur.strip():
        parts.append(cur)
    return parts


def _norm_type(p):
    p = re.sub(r"\s+", " ", p.strip())
    p = re.sub(r"\bconst\b", "", p)
    p = re.sub(r"(\w+::)+", "", p)
    return p.replace("&", "").replace(" ", "")


def _param_types(text, anchor, nth=0):
    ms = list(re.finditer(anchor, text))
    if len(ms) <= nth:
        return None
    i = ms[nth].end() - 1
    assert text[i] == "(", (anchor, text[i - 30 : i + 5])
    body = text[i + 1 : _balanced(text, i)]
    out = []
    for p in _split_top(body, ","):
        defaulted = "=" in p
        p = p.split("=")[0].strip()
        m = re.match(r"^(.*?[\s&*>])([A-Za-z_]\w*)$", p, flags=re.S)
        if m:
            p = m.group(1)
        out.append((_norm_type(p), defaulted))
    return out


def _struct_members(text, anchor):
    """(type, name) of the data members of the struct whose opening brace `anchor` ends at; member
    functions and constructors (with their bodies and init lists) are skipped."""
    m = re.search(anchor, text)
    if not m:
        return None
    i = text.index("{", m.end() - 1)
    body = text[i + 1 : _balanced(text, i)]
    stmts, cur, d, k = [], "", 0, 0
    while k < len(body):
        c = body[k]
        if c == "{" and d == 0 and re.search(r"\)\s*(const)?\s*(:[^{]*)?$", cur, flags=re.S):
            k = _balanced(body, k) + 1  # a function body: drop the whole definition
            cur = ""
            continue
        if c in "(<[{":
            d += 1
        elif c in ")>]}":
            d -= 1
        if c == ";" and d == 0:
            stmts.append(cur)
            cur = ""
        else:
            cur += c
        k += 1
    out = []
    for stmt in stmts:
        stmt = re.sub(r"^\s*(public|private|protected):", "", stmt.strip()).strip()
        if not stmt or stmt.startswith(("using ", "static ", "friend ", "struct ", "class ", "enum ")):
            continue
        decl = re.sub(r"\{.*\}\s*$", "", stmt.split("=")[0].strip(), flags=re.S).strip()
        # a declaration whose parentheses are not inside template brackets is a function
        flat = re.sub(r"<[^<>]*(<[^<>]*>[^<>]*)*>", "", decl)
        if "(" in flat:
            continue
        m2 = re.match(r"^(.*?[\s&*>])([A-Za-z_]\w*)$", decl, flags=re.S)
        if m2:
            out.append((_norm_type(m2.group(1)), m2.group(2)))
    return out


# (key, reference header, anchor in the reference, nth match, shim header, anchor in the shim, nth match)
SIGNATURES = [
    ("Operator::Operator", "exec/Operator.h", r"\n  Operator\(", 0, SHIM_EXEC, r"\n  Operator\(", 0),
    ("Operator::addInput", "exec/Operator.h", r"virtual void addInput\(", 0, SHIM_EXEC, r"virtual void addInput\(", 0),
    ("Operator::isBlocked", "exec/Operator.h", r"virtual BlockingReason isBlocked\(", 0, SHIM_EXEC, r"virtual BlockingReason isBlocked\(", 0),
    ("Operator::getOutput", "exec/Operator.h", r"virtual RowVectorPtr getOutput\(", 0, SHIM_EXEC, r"virtual RowVectorPtr getOutput\(", 0),
    ("Operator::needsInput", "exec/Operator.h", r"virtual bool needsInput\(", 0, SHIM_EXEC, r"virtual bool needsInput\(", 0),
    ("Operator::noMoreInput", "exec/Operator.h", r"virtual void noMoreInput\(", 0, SHIM_EXEC, r"virtual void noMoreInput\(", 0),
    ("Operator::isFinished", "exec/Operator.h", r"virtual bool isFinished\(", 0, SHIM_EXEC, r"virtual bool isFinished\(", 0),
    ("DriverFactory::registerAdapter", "exec/Driver.h", r"static void registerAdapter\(", 0, SHIM_EXEC, r"static void registerAdapter\(", 0),
    ("HashJoinBridge::setHashTable", "exec/HashJoinBridge.h", r"void setHashTable\(", 1, SHIM_EXEC, r"void setHashTable\(", 0),
    ("HashJoinBridge::tableOrFuture", "exec/HashJoinBridge.h", r"tableOrFuture\(", 0, SHIM_EXEC, r"tableOrFuture\(", 0),
    ("FieldAccessTypedExpr", "core/Expressions.h", r"\n  FieldAccessTypedExpr\(", 0, SHIM_EXEC, r"\n  FieldAccessTypedExpr\(", 0),
    ("CallTypedExpr", "core/Expressions.h", r"\n  CallTypedExpr\(", 0, SHIM_EXEC, r"\n  CallTypedExpr\(", 0),
    ("CastTypedExpr", "core/Expressions.h", r"\n  CastTypedExpr\(", 0, SHIM_EXEC, r"\n  CastTypedExpr\(", 0),
    ("FilterNode", "core/PlanNode.h", r"\n  FilterNode\(", 0, SHIM_EXEC, r"\n  FilterNode\(", 0),
    ("ProjectNode", "core/PlanNode.h", r"\n  ProjectNode\(", 1, SHIM_EXEC, r"\n  ProjectNode\(", 0),
    ("AggregationNode", "core/PlanNode.h", r"\n  AggregationNode\(", 0, SHIM_EXEC, r"\n  AggregationNode\(", 0),
    ("HashJoinNode", "core/PlanNode.h", r"\n  HashJoinNode\(", 0, SHIM_EXEC, r"\n  HashJoinNode\(", 0),
    ("PartitionedOutputNode", "core/PlanNode.h", r"\n  PartitionedOutputNode\(", 2, SHIM_EXEC, r"\n  PartitionedOutputNode\(", 0),
    ("ExchangeNode", "core/PlanNode.h", r"\n  ExchangeNode\(", 0, SHIM_EXEC, r"\n  ExchangeNode\(", 0),
    ("VectorFunction::apply", "expression/VectorFunction.h", r"virtual void apply\(", 0, SHIM_EXEC, r"virtual void apply\(", 0),
    ("registerVectorFunction", "expression/VectorFunction.h", r"bool registerVectorFunction\(", 0, SHIM_EXEC, r"bool registerVectorFunction\(", 0),
    ("registerAggregateFunction", "exec/Aggregate.h", r"AggregateRegistrationResult registerAggregateFunction\(", 0, SHIM_EXEC, r"AggregateRegistrationResult registerAggregateFunction\(", 0),
    ("Aggregate::create", "exec/Aggregate.h", r"static std::unique_ptr<Aggregate> create\(", 0, SHIM_EXEC, r"static std::unique_ptr<Aggregate> create\(", 0),
    ("RowVector", "vector/ComplexVector.h", r"\n  RowVector\((?=\s*velox)", 0, SHIM_VEC, r"\n  RowVector\(", 0),
    ("FlatVector", "vector/FlatVector.h", r"\n  FlatVector\((?=\s*velox)", 0, SHIM_VEC, r"\n  FlatVector\(", 0),
    ("DictionaryVector", "vector/DictionaryVector.h", r"\n  DictionaryVector\((?=\s*velox)", 0, SHIM_VEC, r"\n  DictionaryVector\(", 0),
    ("ConstantVector", "vector/ConstantVector.h", r"\n  ConstantVector\((?=\s*velox)", 0, SHIM_VEC, r"\n  ConstantVector\(", 0),
    ("BaseVector::wrapInDictionary", "vector/BaseVector.h", r"static VectorPtr wrapInDictionary\(", 0, SHIM_VEC, r"static VectorPtr wrapInDictionary\(", 0),
    ("HashPartitionFunctionSpec", "exec/HashPartitionFunction.h", r"\n  HashPartitionFunctionSpec\(", 0, SHIM_EXEC, r"\n  HashPartitionFunctionSpec\(", 0),
    ("SortOrder", "core/PlanNode.h", r"\n  SortOrder\(", 0, SHIM_EXEC, r"\n  SortOrder\(", 0),
]

STRUCTS = [
    ("DriverAdapter", "exec/Driver.h", r"struct DriverAdapter\s*\{", SHIM_EXEC, r"struct DriverAdapter\s*\{"),
    ("AggregationNode::Aggregate", "core/PlanNode.h", r"struct Aggregate\s*\{", SHIM_EXEC, r"struct Aggregate\s*\{"),
    ("HashBuildResult", "exec/HashJoinBridge.h", r"struct HashBuildResult\s*\{", SHIM_EXEC, r"struct HashBuildResult\s*\{"),
    ("AggregateRegistrationResult", "exec/AggregateUtil.h", r"struct AggregateRegistrationResult\s*\{", SHIM_EXEC, r"struct AggregateRegistrationResult\s*\{"),
    ("AggregateFunctionMetadata", "exec/Aggregate.h", r"struct AggregateFunctionMetadata\s*\{", SHIM_EXEC, r"struct AggregateFunctionMetadata\s*\{"),
    ("AggregateFunctionEntry", "exec/Aggregate.h", r"struct AggregateFunctionEntry\s*\{", SHIM_EXEC, r"struct AggregateFunctionEntry\s*\{"),
    ("IdentityProjection", "exec/Operator.h", r"struct IdentityProjection\s*\{", SHIM_EXEC, r"struct IdentityProjection\s*\{"),
]

# key -> (what differs, the parameter positions / member names excused)
DEVIATIONS = {
    "HashBuildResult": ("the CPU table and the spill bookkeeping are not carried", ("table", "restoredPartitionId", "spillPartitionIds")),
}


def _have_reference():
    return os.path.isdir(REF)


@pytest.mark.parametrize("row", SIGNATURES, ids=[r[0] for r in SIGNATURES])
def test_signature_matches_reference(row):
    key, ref_file, ref_anchor, ref_nth, shim_file, shim_anchor, shim_nth = row
    shim = _param_types(_text(shim_file), shim_anchor, shim_nth)
    assert shim is not None, f"{key}: not declared in {os.path.basename(shim_file)}"
    if not _have_reference():
        pytest.skip("reference tree not present on this machine")
    ref = _param_types(_text(os.path.join(REF, ref_file)), ref_anchor, ref_nth)
    assert ref is not None, f"{key}: anchor not found in the reference's {ref_file}"
    excused = DEVIATIONS.get(key, (None, ()))[1]
    ref_cmp = [t for i, (t, _) in enumerate(ref) if i not in excused]
    shim_cmp = [t for i, (t, _) in enumerate(shim) if i not in excused]
    # Code written against the shim must compile against the reference: the same leading parameter
    # types, and whatever the reference declares beyond them has a default value.
    assert shim_cmp == ref_cmp[: len(shim_cmp)], f"{key}: shim {shim} vs reference {ref}"
    assert all(d for _, d in ref[len(shim) :]), f"{key}: the reference has further required parameters {ref[len(shim):]}"


@pytest.mark.parametrize("row", STRUCTS, ids=[r[0] for r in STRUCTS])
def test_struct_members_match_reference(row):
    key, ref_file, ref_anchor, shim_file, shim_anchor = row
    shim = _struct_members(_text(shim_file), shim_anchor)
    assert shim, f"{key}: not declared in {os.path.basename(shim_file)}"
    if not _have_reference():
        pytest.skip("reference tree not present on this machine")
    ref = _struct_members(_text(os.path.join(REF, ref_file)), ref_anchor)
    assert ref, f"{key}: anchor not found in the reference's {ref_file}"
    excused = set(DEVIATIONS.get(key, (None, ()))[1])
    ref_cmp = [m for m in ref if m[1] not in excused]
    shim_cmp = [m for m in shim if m[1] not in excused]
    assert shim_cmp == ref_cmp, f"{key}: shim {shim} vs reference {ref}"


def test_deviations_are_documented():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for key, (what, _) in DEVIATIONS.items():
        assert key in doc, f"deviation {key} ({what}) is not described in INTEGRATION.md"
