#include "expr_compiler.h"
#include "operators.h"
#include "plan_resolve.h"

#include <cstring>
#include <sstream>

namespace velox_b200 {

namespace {

// B200 scalar function: resolved by name through the registry; the device body is an opcode of
// the expression VM, so whole expression trees fuse into one kernel launch. Node-at-a-time
// apply() is not how the device engine runs.
class B200ScalarFunction : public B200VectorFunction {
 public:
  explicit B200ScalarFunction(int opcode) : opcode_(opcode) {}
  int opcode() const { return opcode_; }

 private:
  int opcode_;
};

// Name a function object is registered under (apply() compiles a call by name).
std::string registeredNameOf(const exec::VectorFunction* fn) {
  std::lock_guard<std::mutex> l(exec::vectorFunctionMutex());
  for (auto& [name, entry] : exec::vectorFunctionFactories())
    if (entry.function.get() == fn) return name;
  VELOX_FAIL("VectorFunction::apply on a function object that is not in the registry");
}

exec::FunctionSignaturePtr sig(std::string ret, std::vector<std::string> args) {
  auto s = std::make_shared<exec::FunctionSignature>();
  s->returnType = std::move(ret);
  s->argTypes = std::move(args);
  return s;
}

std::string constKey(const Variant& v) {
  if (v.isNull) return "null:" + std::to_string(static_cast<int>(v.kind));
  std::ostringstream os;
  os.precision(17);
  switch (v.kind) {
    case TypeKind::BOOLEAN: os << "b:" << std::get<bool>(v.value); break;
    case TypeKind::INTEGER: os << "i:" << std::get<int32_t>(v.value); break;
    case TypeKind::BIGINT: os << "l:" << std::get<int64_t>(v.value); break;
    case TypeKind::DOUBLE: {
      double d = std::get<double>(v.value);
      uint64_t u;
      std::memcpy(&u, &d, 8);
      os << "d:" << u;
      break;
    }
    default: os << "s:" << std::get<std::string>(v.value);
  }
  return os.str();
}

struct Compiler {
  const RowTypePtr& inputType;
  CompiledProgram& out;
  std::map<std::string, int> cse;  // canonical text -> register

  int newReg() {
    VELOX_CHECK(out.nRegs < 64, "expression needs more than 64 registers");
    return out.nRegs++;
  }
  int emit(int op, int type, int a = 0, int b = 0, int c = 0) {
    const int dst = newReg();
    VELOX_CHECK(out.instrs.size() < 256, "expression program above 256 instructions");
    out.instrs.push_back(vb2_instr{op, type, dst, a, b, c});
    return dst;
  }
  int addConst(const Variant& v, const TypePtr& type) {
    vb2_const c{};
    c.type = veloxTypeToVb2(type);
    c.is_null = v.isNull;
    if (!v.isNull) {
      switch (v.kind) {
        case TypeKind::BOOLEAN: c.i = std::get<bool>(v.value); break;
        case TypeKind::INTEGER: c.i = std::get<int32_t>(v.value); break;
        case TypeKind::BIGINT: c.i = std::get<int64_t>(v.value); break;
        case TypeKind::DOUBLE: c.d = std::get<double>(v.value); break;
        default:
          out.constStrings.push_back(std::get<std::string>(v.value));
          c.len = static_cast<int32_t>(out.constStrings.back().size());
          c.pad = static_cast<int32_t>(out.constStrings.size());  // 1-based index, resolved at upload
      }
    }
    VELOX_CHECK(out.consts.size() < 32, "expression uses more than 32 constants");
    out.consts.push_back(c);
    return static_cast<int>(out.consts.size()) - 1;
  }

  static bool isVarchar(const core::TypedExprPtr& e) { return e->type()->kind() == TypeKind::VARCHAR; }

  std::string key(const core::TypedExprPtr& e) {
    if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) return "#" + f->name();
    if (auto c = dynamic_cast<const core::ConstantTypedExpr*>(e.get())) return constKey(c->value());
    std::string s;
    if (auto call = dynamic_cast<const core::CallTypedExpr*>(e.get())) s = call->name();
    else if (dynamic_cast<const core::CastTypedExpr*>(e.get())) s = "cast:" + e->type()->toString();
    s += "(";
    for (auto& in : e->inputs()) s += key(in) + ",";
    return s + ")";
  }

  int compile(const core::TypedExprPtr& e) {
    const std::string k = key(e);
    auto it = cse.find(k);
    if (it != cse.end()) return it->second;
    const int r = compileNew(e);
    cse[k] = r;
    return r;
  }

  int compileNew(const core::TypedExprPtr& e) {
    const int t = veloxTypeToVb2(e->type());
    if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) {
      const int32_t channel = channelOf(inputType, *f);
      if (t == VB2_VARCHAR) VELOX_UNSUPPORTED("VARCHAR values in expressions other than LIKE / comparison with a constant");
      return emit(VB2_OP_LOAD, t, channel);
    }
    if (auto c = dynamic_cast<const core::ConstantTypedExpr*>(e.get())) {
      if (t == VB2_VARCHAR) VELOX_UNSUPPORTED("VARCHAR constant outside LIKE / comparison");
      return emit(VB2_OP_CONST, t, addConst(c->value(), e->type()));
    }
    if (dynamic_cast<const core::CastTypedExpr*>(e.get())) {
      const int a = compile(e->inputs()[0]);
      const int from = veloxTypeToVb2(e->inputs()[0]->type());
      if (from != t && t != VB2_BOOLEAN) out.canRaise = out.canRaise || from == VB2_DOUBLE || (from == VB2_BIGINT && t == VB2_INTEGER);
      return emit(VB2_OP_CAST, t, a, from);
    }
    auto call = dynamic_cast<const core::CallTypedExpr*>(e.get());
    VELOX_CHECK(call != nullptr, "unknown expression node");
    const std::string& name = call->name();
    const auto& in = e->inputs();
    if (name == "and" || name == "or") {
      VELOX_CHECK(!in.empty(), "and/or without arguments");
      int r = compile(in[0]);
      for (size_t i = 1; i < in.size(); ++i) r = emit(name == "and" ? VB2_OP_AND : VB2_OP_OR, VB2_BOOLEAN, r, compile(in[i]));
      return r;
    }
    if (name == "switch" || name == "if") {
      VELOX_CHECK(in.size() >= 2, "switch needs a condition and a value");
      int acc = -1;
      size_t n = in.size();
      if (n % 2 == 1) acc = compile(in[n - 1]), --n;
      for (size_t i = n; i >= 2; i -= 2) {
        const int cond = compile(in[i - 2]);
        const int val = compile(in[i - 1]);
        acc = emit(VB2_OP_SELECT, t, cond, val, acc);
      }
      return acc;
    }
    registerB200Functions();
    if (auto fn = exec::getVectorFunction(name)) {
      if (auto* user = dynamic_cast<const B200DeviceFunction*>(fn.get())) {
        // registered device function: one CALL instruction, fused into the kernel of this ExprSet
        VELOX_CHECK(in.size() == user->argTypes().size(), name + ": takes " + std::to_string(user->argTypes().size()) + " arguments");
        VELOX_CHECK(user->returnType()->kind() == e->type()->kind(), name + ": returns " + user->returnType()->toString());
        int regs[3] = {-1, -1, -1};
        for (size_t i = 0; i < in.size(); ++i) {
          VELOX_CHECK(in[i]->type()->kind() == user->argTypes()[i]->kind(), name + ": argument " + std::to_string(i) + " must be " + user->argTypes()[i]->toString());
          regs[i] = compile(in[i]);
        }
        return emit(VB2_OP_CALL, t | (user->id() << 8), regs[0], regs[1], regs[2]);
      }
    }
    const int opcode = opcodeForFunction(name);
    if (opcode < 0) VELOX_UNSUPPORTED("scalar function '" + name + "' is not registered for the B200 engine");
    // string predicates: column vs constant only
    if (opcode == VB2_OP_LIKE || (!in.empty() && isVarchar(in[0]))) {
      const core::FieldAccessTypedExpr* field = nullptr;
      const core::ConstantTypedExpr* cst = nullptr;
      bool swapped = false;
      if (in.size() == 2) {
        field = dynamic_cast<const core::FieldAccessTypedExpr*>(in[0].get());
        cst = dynamic_cast<const core::ConstantTypedExpr*>(in[1].get());
        if (!field || !cst) {
          field = dynamic_cast<const core::FieldAccessTypedExpr*>(in[1].get());
          cst = dynamic_cast<const core::ConstantTypedExpr*>(in[0].get());
          swapped = true;
        }
      }
      if (!field || !cst || (opcode == VB2_OP_LIKE && swapped)) VELOX_UNSUPPORTED("VARCHAR predicate must compare a column with a constant");
      const int k = addConst(cst->value(), cst->type());
      if (opcode == VB2_OP_LIKE) return emit(VB2_OP_LIKE, VB2_BOOLEAN, channelOf(inputType, *field), k);
      VELOX_CHECK(opcode >= VB2_OP_LT && opcode <= VB2_OP_NEQ, "unsupported VARCHAR function " + name);
      int cmp = opcode - VB2_OP_LT;  // 0 lt 1 lte 2 gt 3 gte 4 eq 5 neq
      if (swapped) { static const int mirror[] = {2, 3, 0, 1, 4, 5}; cmp = mirror[cmp]; }
      return emit(VB2_OP_STRCMP, VB2_BOOLEAN, channelOf(inputType, *field), k, cmp);
    }
    std::vector<int> regs;
    for (auto& a : in) regs.push_back(compile(a));
    const int argType = in.empty() ? t : veloxTypeToVb2(in[0]->type());
    for (auto& a : in) VELOX_CHECK(veloxTypeToVb2(a->type()) == argType || opcode == VB2_OP_IS_NULL, name + ": argument types differ");
    switch (opcode) {
      case VB2_OP_ADD: case VB2_OP_SUB: case VB2_OP_MUL: case VB2_OP_DIV: case VB2_OP_MOD:
        VELOX_CHECK(regs.size() == 2, name + " takes two arguments");
        if (argType != VB2_DOUBLE) out.canRaise = true;
        return emit(opcode, argType, regs[0], regs[1]);
      case VB2_OP_NEG:
        VELOX_CHECK(regs.size() == 1, "negate takes one argument");
        if (argType != VB2_DOUBLE) out.canRaise = true;
        return emit(opcode, argType, regs[0]);
      case VB2_OP_LT: case VB2_OP_LTE: case VB2_OP_GT: case VB2_OP_GTE: case VB2_OP_EQ: case VB2_OP_NEQ:
        VELOX_CHECK(regs.size() == 2, name + " takes two arguments");
        return emit(opcode, argType, regs[0], regs[1]);
      case VB2_OP_BETWEEN:
        VELOX_CHECK(regs.size() == 3, "between takes three arguments");
        return emit(opcode, argType, regs[0], regs[1], regs[2]);
      case VB2_OP_NOT: case VB2_OP_IS_NULL:
        VELOX_CHECK(regs.size() == 1, name + " takes one argument");
        return emit(opcode, VB2_BOOLEAN, regs[0]);
      default: VELOX_UNSUPPORTED("function " + name);
    }
  }
};

}  // namespace

B200DeviceFunction::B200DeviceFunction(std::string entry, std::string cudaSource, TypePtr returnType, std::vector<TypePtr> argTypes)
    : returnType_(std::move(returnType)), argTypes_(std::move(argTypes)) {
  VELOX_CHECK(!argTypes_.empty() && argTypes_.size() <= 3, "device functions take one to three arguments");
  std::vector<int32_t> at;
  for (auto& a : argTypes_) at.push_back(veloxTypeToVb2(a));
  id_ = vb2k_register_device_function(entry.c_str(), cudaSource.c_str(), veloxTypeToVb2(returnType_), at.data(), static_cast<int32_t>(at.size()));
  VELOX_CHECK(id_ >= 0, "device function '" + entry + "': BIGINT / INTEGER / DOUBLE / BOOLEAN arguments and result only");
}

void B200VectorFunction::apply(const SelectivityVector& rows, std::vector<VectorPtr>& args, const TypePtr& outputType, exec::EvalCtx&,
                               VectorPtr& result) const {
  if (!rows.hasSelections()) return;
  const std::string name = registeredNameOf(this);
  const vector_size_t n = rows.end();
  std::vector<std::string> names;
  std::vector<TypePtr> types;
  std::vector<core::TypedExprPtr> fields;
  for (size_t i = 0; i < args.size(); ++i) {
    VELOX_CHECK(args[i] && args[i]->size() >= n, name + ": argument vector shorter than the selected rows");
    names.push_back("a" + std::to_string(i));
    types.push_back(args[i]->type());
    fields.push_back(std::make_shared<core::FieldAccessTypedExpr>(args[i]->type(), names.back()));
  }
  auto rowType = ROW(names, types);
  auto pool = args.empty() ? nullptr : args[0]->pool();
  auto host = std::make_shared<RowVector>(pool, rowType, nullptr, n, args);
  auto dev = driverDeviceContext(nullptr);  // node-at-a-time calls share one stream per process
  auto call = std::make_shared<core::CallTypedExpr>(outputType, fields, name);
  CompiledProgram program = compileExprs({call}, false, rowType);
  program.uploadConstants(dev->stream);
  auto in = toDevice(host, dev->stream);
  auto flag = allocDeviceZeroed(8, dev->stream);
  auto out = evalProjections(program, in, nullptr, n, dev->stream, flag, ROW({"r"}, {outputType}), pool);
  checkDeviceError(flag, dev->stream, name.c_str());
  VectorPtr computed = toHost(out)->childAt(0);
  if (!result || rows.countSelected() == n) {
    // nothing to preserve (no pre-allocated result, or every row selected): hand the computed vector over
    if (!result || result->size() <= n) { result = computed; return; }
  }
  // result-reuse rule (VectorFunction.h:44-80): only the selected rows may be overwritten
  VELOX_CHECK(result->isFlatEncoding() && result->size() >= n, name + ": pre-allocated result must be a flat vector covering the selected rows");
  auto copy = [&](auto tag) {
    using T = decltype(tag);
    auto* dst = result->as<FlatVector<T>>();
    auto* src = computed->as<FlatVector<T>>();
    rows.applyToSelected([&](vector_size_t i) {
      const bool null = src->isNullAt(i);
      result->setNull(i, null);
      if (null) return;
      if constexpr (std::is_same_v<T, bool>) bits::setBit(dst->values()->template asMutable<uint64_t>(), i, src->valueAt(i));
      else dst->mutableRawValues()[i] = src->valueAt(i);
    });
  };
  switch (outputType->kind()) {
    case TypeKind::BOOLEAN: copy(bool{}); break;
    case TypeKind::INTEGER: copy(int32_t{}); break;
    case TypeKind::BIGINT: copy(int64_t{}); break;
    case TypeKind::DOUBLE: copy(double{}); break;
    default: VELOX_UNSUPPORTED(name + ": result type " + outputType->toString());
  }
}

void registerB200Functions() {
  static bool done = false;
  if (done) return;
  done = true;
  auto reg = [](const char* name, int op, exec::FunctionSignaturePtr s, bool defaultNull = true) {
    exec::VectorFunctionMetadata md;
    md.defaultNullBehavior = defaultNull;
    exec::registerVectorFunction(name, {std::move(s)}, std::make_unique<B200ScalarFunction>(op), md);
  };
  // functions/prestosql/registration/MathematicalOperatorsRegistration.cpp:25-82
  reg("plus", VB2_OP_ADD, sig("T", {"T", "T"}));
  reg("minus", VB2_OP_SUB, sig("T", {"T", "T"}));
  reg("multiply", VB2_OP_MUL, sig("T", {"T", "T"}));
  reg("divide", VB2_OP_DIV, sig("T", {"T", "T"}));
  reg("modulus", VB2_OP_MOD, sig("T", {"T", "T"}));
  reg("mod", VB2_OP_MOD, sig("T", {"T", "T"}));
  reg("negate", VB2_OP_NEG, sig("T", {"T"}));
  // functions/prestosql/Comparisons.h:24-160
  reg("lt", VB2_OP_LT, sig("boolean", {"T", "T"}));
  reg("lte", VB2_OP_LTE, sig("boolean", {"T", "T"}));
  reg("gt", VB2_OP_GT, sig("boolean", {"T", "T"}));
  reg("gte", VB2_OP_GTE, sig("boolean", {"T", "T"}));
  reg("eq", VB2_OP_EQ, sig("boolean", {"T", "T"}));
  reg("neq", VB2_OP_NEQ, sig("boolean", {"T", "T"}));
  reg("between", VB2_OP_BETWEEN, sig("boolean", {"T", "T", "T"}));
  reg("like", VB2_OP_LIKE, sig("boolean", {"varchar", "varchar"}));
  reg("not", VB2_OP_NOT, sig("boolean", {"boolean"}));
  reg("is_null", VB2_OP_IS_NULL, sig("boolean", {"T"}), false);
}

int opcodeForFunction(const std::string& name) {
  registerB200Functions();
  auto fn = exec::getVectorFunction(name);
  auto* b = dynamic_cast<B200ScalarFunction*>(fn.get());
  return b ? b->opcode() : -1;
}

CompiledProgram compileExprs(const std::vector<core::TypedExprPtr>& exprs, bool hasFilter, const RowTypePtr& inputType) {
  CompiledProgram p;
  Compiler c{inputType, p, {}};
  size_t first = 0;
  if (hasFilter) {
    VELOX_CHECK(!exprs.empty() && exprs[0]->type()->kind() == TypeKind::BOOLEAN, "filter must be BOOLEAN");
    p.filterReg = c.compile(exprs[0]);
    p.nFilterInstrs = static_cast<int>(p.instrs.size());
    first = 1;
  }
  for (size_t i = first; i < exprs.size(); ++i) {
    CompiledProgram::Output o;
    o.type = exprs[i]->type();
    if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(exprs[i].get())) {
      o.identityField = channelOf(inputType, *f);  // zero-copy: wrapped or passed through, never evaluated
    } else {
      if (o.type->kind() == TypeKind::VARCHAR) VELOX_UNSUPPORTED("computed VARCHAR projections");
      o.reg = c.compile(exprs[i]);
    }
    p.outputs.push_back(o);
  }
  return p;
}

std::vector<bool> CompiledProgram::nullability(const std::vector<bool>& columnMayBeNull) const {
  std::vector<bool> n(nRegs > 0 ? nRegs : 1, false);
  for (auto& in : instrs) {
    bool r = false;
    switch (in.op) {
      case VB2_OP_LOAD: r = columnMayBeNull.at(in.a); break;
      case VB2_OP_CONST: r = consts[in.a].is_null != 0; break;
      case VB2_OP_NULL: r = true; break;
      case VB2_OP_IS_NULL: r = false; break;
      case VB2_OP_SELECT: r = n[in.b] || (in.c < 0 ? true : n[in.c]); break;
      case VB2_OP_LIKE: case VB2_OP_STRCMP: r = columnMayBeNull.at(in.a) || consts[in.b].is_null; break;
      case VB2_OP_NEG: case VB2_OP_NOT: case VB2_OP_CAST: r = n[in.a]; break;
      case VB2_OP_BETWEEN: r = n[in.a] || n[in.b] || n[in.c]; break;
      case VB2_OP_CALL: r = n[in.a] || (in.b >= 0 && n[in.b]) || (in.c >= 0 && n[in.c]); break;
      default: r = n[in.a] || n[in.b];
    }
    n[in.dst] = r;
  }
  return n;
}

void CompiledProgram::uploadConstants(cudaStream_t stream) {
  if (constStrings.empty() || constChars) return;
  std::string all;
  std::vector<size_t> offs;
  for (auto& s : constStrings) { offs.push_back(all.size()); all += s; }
  constChars = allocDevice(all.size() + 1, stream);
  VB2_CU(cudaMemcpyAsync(constChars->data(), all.data(), all.size(), cudaMemcpyHostToDevice, stream));
  VB2_CU(cudaStreamSynchronize(stream));
  for (auto& c : consts)
    if (c.type == VB2_VARCHAR && !c.is_null && c.pad > 0) c.str = constChars->as<char>() + offs[c.pad - 1];
}

core::TypedExprPtr substituteFields(const core::TypedExprPtr& expr, const std::vector<core::TypedExprPtr>& fields, const RowTypePtr& type) {
  if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(expr.get())) return fields.at(channelOf(type, *f));
  if (dynamic_cast<const core::ConstantTypedExpr*>(expr.get())) return expr;
  std::vector<core::TypedExprPtr> in;
  for (auto& i : expr->inputs()) in.push_back(substituteFields(i, fields, type));
  if (auto c = dynamic_cast<const core::CallTypedExpr*>(expr.get())) return std::make_shared<core::CallTypedExpr>(expr->type(), std::move(in), c->name());
  if (auto c = dynamic_cast<const core::CastTypedExpr*>(expr.get())) return std::make_shared<core::CastTypedExpr>(expr->type(), in[0], c->nullOnFailure());
  VELOX_UNSUPPORTED("substituteFields: unknown node");
}

namespace {

struct SigBuilder {
  const RowTypePtr& inputType;
  FusedBinding& b;
  const core::ITypedExpr* joinFlag;
  std::map<int, int> colIndex;  // input column -> renumbered

  int column(int inputCol) {
    auto it = colIndex.find(inputCol);
    if (it != colIndex.end()) return it->second;
    const int id = static_cast<int>(b.columns.size());
    b.columns.push_back(inputCol);
    colIndex[inputCol] = id;
    return id;
  }
  bool print(const core::TypedExprPtr& e, std::string& out) {
    if (joinFlag && e.get() == joinFlag) { out += "joinflag"; return true; }
    if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) {
      char k;
      switch (e->type()->kind()) {
        case TypeKind::DOUBLE: k = 'f'; break;
        case TypeKind::INTEGER: k = 'i'; break;
        case TypeKind::BIGINT: k = 'l'; break;
        default: return false;
      }
      out += k + std::to_string(column(channelOf(inputType, *f)));
      return true;
    }
    if (auto c = dynamic_cast<const core::ConstantTypedExpr*>(e.get())) {
      const Variant& v = c->value();
      if (v.isNull) return false;
      switch (v.kind) {
        case TypeKind::DOUBLE: out += "pf" + std::to_string(b.pf.size()); b.pf.push_back(std::get<double>(v.value)); return true;
        case TypeKind::INTEGER: out += "pi" + std::to_string(b.pi.size()); b.pi.push_back(std::get<int32_t>(v.value)); return true;
        case TypeKind::BIGINT: out += "pl" + std::to_string(b.pl.size()); b.pl.push_back(std::get<int64_t>(v.value)); return true;
        default: return false;
      }
    }
    auto call = dynamic_cast<const core::CallTypedExpr*>(e.get());
    if (!call) return false;
    std::string name = call->name() == "if" ? "switch" : call->name();
    out += name + "(";
    bool first = true;
    for (auto& in : e->inputs()) {
      if (!first) out += ",";
      first = false;
      if (!print(in, out)) return false;
    }
    out += ")";
    return true;
  }
};

}  // namespace

FusedBinding fusedSignature(const core::TypedExprPtr& filter, const std::vector<core::TypedExprPtr>& projections,
                            const RowTypePtr& inputType, int joinKeyColumn, const core::ITypedExpr* joinFlagExpr) {
  FusedBinding b;
  SigBuilder sb{inputType, b, joinFlagExpr, {}};
  std::string s = "F:";
  if (filter) {
    if (!sb.print(filter, s)) return b;
  } else {
    s += "true";
  }
  int joinCol = -1;
  if (joinKeyColumn >= 0) {
    if (inputType->childAt(joinKeyColumn)->kind() != TypeKind::BIGINT) return b;
    joinCol = sb.column(joinKeyColumn);
  }
  s += ";P:";
  for (size_t i = 0; i < projections.size(); ++i) {
    if (i) s += "|";
    if (projections[i]->type()->kind() != TypeKind::DOUBLE) return b;
    if (!sb.print(projections[i], s)) return b;
  }
  if (joinCol >= 0) s += ";J:l" + std::to_string(joinCol);
  if (b.columns.size() > VB2_FUSED_MAX_COLS || b.pf.size() > VB2_FUSED_MAX_PARAMS || b.pl.size() > VB2_FUSED_MAX_PARAMS ||
      b.pi.size() > VB2_FUSED_MAX_PARAMS)
    return b;
  b.signature = s;
  b.ok = true;
  return b;
}

}  // namespace velox_b200
