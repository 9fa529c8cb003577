// Reader for the plan text of the C ABI (vb2_task_create): builds core::PlanNode / ITypedExpr
// trees — the structures a Velox application hands to Task::create. Grammar in DESIGN.md and
// velox_b200/plan.py. (The CPU oracle has its own, independent reader.)
#include "plan_text.h"
#include "expr_compiler.h"
#include "operators.h"

#include <cctype>

namespace velox_b200 {

namespace {

struct Tok {
  enum Kind { LP, RP, ATOM, STR, END } kind;
  std::string text;
};

class Lexer {
 public:
  explicit Lexer(const std::string& s) : sp_(&s) { advance(); }
  const Tok& peek() const { return cur_; }
  Tok take() {
    Tok t = cur_;
    advance();
    return t;
  }
  void expect(Tok::Kind k, const char* what) {
    if (cur_.kind != k) throw VeloxRuntimeError(std::string("plan text: expected ") + what + " near '" + cur_.text + "'");
    advance();
  }
  std::string atom(const char* what) {
    if (cur_.kind != Tok::ATOM) throw VeloxRuntimeError(std::string("plan text: expected ") + what + " near '" + cur_.text + "'");
    return take().text;
  }

 private:
  void advance() {
    const std::string& s_ = *sp_;
    while (p_ < s_.size() && std::isspace(static_cast<unsigned char>(s_[p_]))) ++p_;
    if (p_ >= s_.size()) { cur_ = {Tok::END, ""}; return; }
    const char c = s_[p_];
    if (c == '(') { ++p_; cur_ = {Tok::LP, "("}; return; }
    if (c == ')') { ++p_; cur_ = {Tok::RP, ")"}; return; }
    if (c == '"') {
      std::string out;
      ++p_;
      while (p_ < s_.size() && s_[p_] != '"') {
        if (s_[p_] == '\\' && p_ + 1 < s_.size()) ++p_;
        out.push_back(s_[p_++]);
      }
      if (p_ >= s_.size()) throw VeloxRuntimeError("plan text: unterminated string literal");
      ++p_;
      cur_ = {Tok::STR, out};
      return;
    }
    size_t b = p_;
    while (p_ < s_.size() && !std::isspace(static_cast<unsigned char>(s_[p_])) && s_[p_] != '(' && s_[p_] != ')') ++p_;
    cur_ = {Tok::ATOM, s_.substr(b, p_ - b)};
  }
  const std::string* sp_;
  size_t p_ = 0;
  Tok cur_{Tok::END, ""};
};

TypePtr typeFromName(const std::string& n) {
  if (n == "BOOLEAN") return BOOLEAN();
  if (n == "INTEGER") return INTEGER();
  if (n == "DATE") return DATE();
  if (n == "BIGINT") return BIGINT();
  if (n == "DOUBLE") return DOUBLE();
  if (n == "VARCHAR") return VARCHAR();
  throw VeloxRuntimeError("plan text: unknown type " + n);
}

bool isComparison(const std::string& f) { return f == "lt" || f == "lte" || f == "gt" || f == "gte" || f == "eq" || f == "neq"; }
bool isArithmetic(const std::string& f) { return f == "plus" || f == "minus" || f == "multiply" || f == "divide" || f == "modulus"; }

class Parser {
 public:
  explicit Parser(const std::string& text) : lex_(text) {}

  core::PlanNodePtr plan() {
    auto p = node();
    if (lex_.peek().kind != Tok::END) throw VeloxRuntimeError("plan text: trailing input");
    return p;
  }

 private:
  std::string nextId() { return std::to_string(nextId_++); }

  std::vector<int32_t> intList(const char* head) {
    lex_.expect(Tok::LP, "(");
    if (lex_.atom(head) != head) throw VeloxRuntimeError(std::string("plan text: expected (") + head + " ...)");
    std::vector<int32_t> out;
    while (lex_.peek().kind == Tok::ATOM) out.push_back(std::stoi(lex_.take().text));
    lex_.expect(Tok::RP, ")");
    return out;
  }

  core::TypedExprPtr expr(const RowTypePtr& in) {
    lex_.expect(Tok::LP, "(");
    const std::string h = lex_.atom("expression head");
    core::TypedExprPtr out;
    if (h == "field") {
      const int i = std::stoi(lex_.atom("field index"));
      if (i < 0 || i >= static_cast<int>(in->size())) throw VeloxRuntimeError("plan text: field index out of range");
      out = std::make_shared<core::FieldAccessTypedExpr>(in->childAt(i), in->nameOf(i));
    } else if (h == "f64") {
      out = std::make_shared<core::ConstantTypedExpr>(DOUBLE(), Variant::of<double>(TypeKind::DOUBLE, std::stod(lex_.atom("number"))));
    } else if (h == "i64") {
      out = std::make_shared<core::ConstantTypedExpr>(BIGINT(), Variant::of<int64_t>(TypeKind::BIGINT, std::stoll(lex_.atom("number"))));
    } else if (h == "i32" || h == "date") {
      out = std::make_shared<core::ConstantTypedExpr>(h == "date" ? DATE() : INTEGER(),
                                                      Variant::of<int32_t>(TypeKind::INTEGER, static_cast<int32_t>(std::stol(lex_.atom("number")))));
    } else if (h == "bool") {
      out = std::make_shared<core::ConstantTypedExpr>(BOOLEAN(), Variant::of<bool>(TypeKind::BOOLEAN, lex_.atom("true|false") == "true"));
    } else if (h == "str") {
      if (lex_.peek().kind != Tok::STR) throw VeloxRuntimeError("plan text: (str \"...\") expected");
      out = std::make_shared<core::ConstantTypedExpr>(VARCHAR(), Variant::of<std::string>(TypeKind::VARCHAR, lex_.take().text));
    } else if (h == "null") {
      TypePtr t = typeFromName(lex_.atom("type"));
      out = std::make_shared<core::ConstantTypedExpr>(t, Variant::null(t->kind()));
    } else if (h == "cast") {
      TypePtr t = typeFromName(lex_.atom("type"));
      out = std::make_shared<core::CastTypedExpr>(t, expr(in));
    } else {
      std::vector<core::TypedExprPtr> args;
      while (lex_.peek().kind == Tok::LP) args.push_back(expr(in));
      TypePtr t;
      auto need = [&](size_t n) { if (args.size() != n) throw VeloxRuntimeError("plan text: " + h + ": wrong argument count"); };
      if (h == "and" || h == "or" || h == "not" || h == "is_null" || h == "like" || h == "between" || isComparison(h)) t = BOOLEAN();
      else if (h == "switch" || h == "if") { if (args.size() < 2) throw VeloxRuntimeError("plan text: switch: too few arguments"); t = args[1]->type(); }
      else if (isArithmetic(h)) { need(2); t = args[0]->type(); }
      else if (h == "negate") { need(1); t = args[0]->type(); }
      else if (auto* user = dynamic_cast<const B200DeviceFunction*>(exec::getVectorFunction(h).get())) t = user->returnType();  // registered by the application
      else throw VeloxRuntimeError("plan text: unknown function " + h);
      out = std::make_shared<core::CallTypedExpr>(t, std::move(args), h);
    }
    lex_.expect(Tok::RP, ")");
    return out;
  }

  core::PlanNodePtr node() {
    lex_.expect(Tok::LP, "(");
    const std::string h = lex_.atom("plan node");
    core::PlanNodePtr out;
    if (h == "values") {
      const int src = std::stoi(lex_.atom("source id"));
      lex_.expect(Tok::LP, "(");
      std::vector<std::string> names;
      std::vector<TypePtr> types;
      while (lex_.peek().kind == Tok::ATOM) {
        types.push_back(typeFromName(lex_.take().text));
        names.push_back("s" + std::to_string(src) + "c" + std::to_string(names.size()));  // unique across the plan: joins resolve output columns by name
      }
      lex_.expect(Tok::RP, ")");
      out = std::make_shared<core::ValuesNode>(nextId(), ROW(names, types), src);
    } else if (h == "filter") {
      // child follows the expression in the text but the expression needs the child's type:
      // remember the position of the expression and parse the child first.
      const Lexer saved = lex_;
      skip();
      auto child = node();
      Lexer after = lex_;
      lex_ = saved;
      auto f = expr(child->outputType());
      lex_ = after;
      out = std::make_shared<core::FilterNode>(nextId(), f, child);
    } else if (h == "project") {
      const Lexer saved = lex_;
      skip();
      auto child = node();
      Lexer after = lex_;
      lex_ = saved;
      lex_.expect(Tok::LP, "(");
      std::vector<core::TypedExprPtr> exprs;
      std::vector<std::string> names;
      const std::string id = nextId();
      while (lex_.peek().kind == Tok::LP) {
        exprs.push_back(expr(child->outputType()));
        names.push_back("n" + id + "p" + std::to_string(names.size()));
      }
      lex_.expect(Tok::RP, ")");
      lex_ = after;
      out = std::make_shared<core::ProjectNode>(id, names, exprs, child);
    } else if (h == "aggregation") {
      const std::string stepName = lex_.atom("step");
      core::AggregationNode::Step step;
      if (stepName == "single") step = core::AggregationNode::Step::kSingle;
      else if (stepName == "partial") step = core::AggregationNode::Step::kPartial;
      else if (stepName == "final") step = core::AggregationNode::Step::kFinal;
      else if (stepName == "intermediate") step = core::AggregationNode::Step::kIntermediate;
      else throw VeloxRuntimeError("plan text: unknown aggregation step " + stepName);
      auto keys = intList("keys");
      // (aggs (fn [col] [(mask col)] [(distinct)]) ...)
      struct RawAgg { std::string fn; int col = -1; int mask = -1; bool distinct = false; };
      std::vector<RawAgg> raws;
      lex_.expect(Tok::LP, "(");
      if (lex_.atom("aggs") != "aggs") throw VeloxRuntimeError("plan text: expected (aggs ...)");
      while (lex_.peek().kind == Tok::LP) {
        lex_.take();
        RawAgg a;
        a.fn = lex_.atom("aggregate name");
        while (lex_.peek().kind != Tok::RP) {
          if (lex_.peek().kind == Tok::LP) {
            lex_.take();
            const std::string what = lex_.atom("mask | distinct");
            if (what == "distinct") a.distinct = true;
            else if (what == "mask") a.mask = std::stoi(lex_.atom("mask column"));
            else throw VeloxRuntimeError("plan text: expected (mask col) or (distinct)");
            lex_.expect(Tok::RP, ")");
          } else {
            a.col = std::stoi(lex_.atom("column"));
          }
        }
        lex_.take();
        raws.push_back(a);
      }
      lex_.expect(Tok::RP, ")");
      auto child = node();
      const auto& in = child->outputType();
      const bool raw = step == core::AggregationNode::Step::kSingle || step == core::AggregationNode::Step::kPartial;
      const bool fin = step == core::AggregationNode::Step::kSingle || step == core::AggregationNode::Step::kFinal;
      const std::string id = nextId();
      auto field = [&](int c) {
        if (c < 0 || c >= static_cast<int>(in->size())) throw VeloxRuntimeError("plan text: aggregate column out of range");
        return std::make_shared<core::FieldAccessTypedExpr>(in->childAt(c), in->nameOf(c));
      };
      std::vector<core::FieldAccessTypedExprPtr> keyExprs;
      std::vector<std::string> names;
      std::vector<TypePtr> types;
      for (int k : keys) { keyExprs.push_back(field(k)); names.push_back(in->nameOf(k)); types.push_back(in->childAt(k)); }
      std::vector<core::AggregationNode::Aggregate> aggs;
      std::vector<std::string> aggNames;
      for (auto& r : raws) {
        core::AggregationNode::Aggregate a;
        std::vector<core::TypedExprPtr> args;
        TypePtr rawType;
        // registered aggregates (exec::registerAggregateFunction): the call keeps its registered name, typing
        // follows the B200Aggregate's accumulator family and transforms
        const std::string callName = r.fn;
        std::string inputFn, finalFn;
        if (r.fn != "count" && r.fn != "sum" && r.fn != "min" && r.fn != "max" && r.fn != "avg") {
          registerB200Aggregates();
          if (!exec::getAggregateFunctionEntry(r.fn)) throw VeloxRuntimeError("plan text: unknown aggregate " + r.fn);
          auto created = exec::Aggregate::create(r.fn, step, {}, nullptr, core::QueryConfig());
          auto* agg = dynamic_cast<B200Aggregate*>(created.get());
          if (!agg) throw VeloxRuntimeError("plan text: aggregate " + r.fn + " has no B200 implementation");
          r.fn = agg->family();
          inputFn = agg->inputFunction();
          finalFn = agg->finalFunction();
        }
        if (r.col >= 0) {
          args.push_back(field(r.col));
          rawType = in->childAt(r.col);
          a.rawInputTypes.push_back(rawType);
          if (!raw && r.fn == "avg") args.push_back(field(r.col + 1));  // the flattened (sum, count) pair
          if (raw && !inputFn.empty()) rawType = scalarFunctionReturnType(inputFn, rawType);  // the accumulator sees the transformed input
        }
        if (r.mask >= 0) a.mask = field(r.mask);
        a.distinct = r.distinct;
        const std::string n = "n" + id + "a" + std::to_string(aggs.size());
        TypePtr resultType;
        if (r.fn == "count") { resultType = BIGINT(); names.push_back(n); types.push_back(BIGINT()); }
        else if (r.fn == "sum") { resultType = raw ? (rawType->kind() == TypeKind::DOUBLE ? DOUBLE() : BIGINT()) : rawType; names.push_back(n); types.push_back(resultType); }
        else if (r.fn == "min" || r.fn == "max") { resultType = rawType; names.push_back(n); types.push_back(rawType); }
        else if (r.fn == "avg") {
          // intermediate avg is ROW(DOUBLE, BIGINT) in the reference (AverageAggregateBase.h:66-69);
          // the C ABI carries it flattened as two columns
          resultType = DOUBLE();
          if (fin) { names.push_back(n); types.push_back(DOUBLE()); }
          else { names.push_back(n + "_sum"); types.push_back(DOUBLE()); names.push_back(n + "_count"); types.push_back(BIGINT()); }
        } else throw VeloxRuntimeError("plan text: unknown aggregate " + r.fn);
        if (fin && !finalFn.empty()) {
          resultType = scalarFunctionReturnType(finalFn, resultType);
          types.back() = resultType;
        }
        a.call = std::make_shared<core::CallTypedExpr>(resultType, std::move(args), callName);
        aggs.push_back(std::move(a));
        aggNames.push_back(n);
      }
      auto agg = std::make_shared<core::AggregationNode>(id, step, keyExprs, std::vector<core::FieldAccessTypedExprPtr>{}, aggNames, aggs, false, false, child);
      agg->setOutputType(ROW(names, types));
      out = agg;
    } else if (h == "localpartition") {
      // (localpartition plan): gather of the drivers of the pipeline below (exec/LocalPartition.cpp)
      auto child = node();
      out = std::make_shared<core::LocalPartitionNode>(nextId(), core::LocalPartitionNode::Type::kGather, false, nullptr, std::vector<core::PlanNodePtr>{child});
    } else if (h == "orderby" || h == "topn") {
      // (orderby ((I asc|desc first|last) ...) plan) | (topn N ((I asc|desc first|last) ...) plan)
      int32_t count = 0;
      if (h == "topn") count = static_cast<int32_t>(std::stol(lex_.atom("row count")));
      struct K { int col; bool asc, first; };
      std::vector<K> ks;
      lex_.expect(Tok::LP, "(");
      while (lex_.peek().kind == Tok::LP) {
        lex_.take();
        K k;
        k.col = std::stoi(lex_.atom("sort column"));
        const std::string dir = lex_.atom("asc|desc"), nulls = lex_.atom("first|last");
        if ((dir != "asc" && dir != "desc") || (nulls != "first" && nulls != "last")) throw VeloxRuntimeError("plan text: sort key must be (I asc|desc first|last)");
        k.asc = dir == "asc";
        k.first = nulls == "first";
        lex_.expect(Tok::RP, ")");
        ks.push_back(k);
      }
      lex_.expect(Tok::RP, ")");
      auto child = node();
      const auto& in = child->outputType();
      std::vector<core::FieldAccessTypedExprPtr> keys;
      std::vector<core::SortOrder> orders;
      for (auto& k : ks) {
        if (k.col < 0 || k.col >= static_cast<int>(in->size())) throw VeloxRuntimeError("plan text: sort column out of range");
        keys.push_back(std::make_shared<core::FieldAccessTypedExpr>(in->childAt(k.col), in->nameOf(k.col)));
        orders.emplace_back(k.asc, k.first);
      }
      if (h == "topn") out = std::make_shared<core::TopNNode>(nextId(), keys, orders, count, false, child);
      else out = std::make_shared<core::OrderByNode>(nextId(), keys, orders, false, child);
    } else if (h == "exchange") {
      // (exchange partitioned|broadcast|gather (keys I ...) plan): PartitionedOutputNode on top of the
      // producing fragment, ExchangeNode as the leaf of the consuming one (core/PlanNode.h:2712,2182)
      const std::string kindName = lex_.atom("exchange kind");
      auto keyIdx = intList("keys");
      auto child = node();
      const auto& in = child->outputType();
      std::vector<core::TypedExprPtr> keys;
      for (int k : keyIdx) {
        if (k < 0 || k >= static_cast<int>(in->size())) throw VeloxRuntimeError("plan text: exchange key out of range");
        keys.push_back(std::make_shared<core::FieldAccessTypedExpr>(in->childAt(k), in->nameOf(k)));
      }
      core::PartitionedOutputNode::Kind kind;
      int parts = 0;  // one partition per rank
      if (kindName == "partitioned") kind = core::PartitionedOutputNode::Kind::kPartitioned;
      else if (kindName == "broadcast") kind = core::PartitionedOutputNode::Kind::kBroadcast;
      else if (kindName == "gather") { kind = core::PartitionedOutputNode::Kind::kPartitioned; parts = 1; }
      else throw VeloxRuntimeError("plan text: unknown exchange kind " + kindName);
      core::PartitionFunctionSpecPtr spec;
      if (kindName == "partitioned") spec = std::make_shared<exec::HashPartitionFunctionSpec>(in, std::vector<exec::column_index_t>(keyIdx.begin(), keyIdx.end()));
      auto po = std::make_shared<core::PartitionedOutputNode>(nextId(), kind, keys, parts, false, spec, in, "B200Columnar", child);
      auto ex = std::make_shared<core::ExchangeNode>(nextId(), in, "B200Columnar");
      ex->setUpstream(po);
      out = ex;
    } else if (h == "hashjoin") {
      const std::string jt = lex_.atom("join type");
      JoinTypeText type{jt};
      auto pk = intList("probekeys");
      auto bk = intList("buildkeys");
      // filter: nil or expression over probe ++ build columns — needs both children first
      const Lexer filterPos = lex_;
      if (lex_.peek().kind == Tok::ATOM) lex_.take(); else skip();
      // outputs
      lex_.expect(Tok::LP, "(");
      if (lex_.atom("out") != "out") throw VeloxRuntimeError("plan text: expected (out ...)");
      struct Out { bool fromProbe; int column; };
      std::vector<Out> outs;
      while (lex_.peek().kind == Tok::LP) {
        lex_.take();
        const std::string side = lex_.atom("p|b");
        const int col = std::stoi(lex_.atom("column"));
        lex_.expect(Tok::RP, ")");
        outs.push_back({side == "p", col});
      }
      lex_.expect(Tok::RP, ")");
      auto probe = node();
      auto build = node();
      Lexer after = lex_;
      core::TypedExprPtr filter;
      lex_ = filterPos;
      if (lex_.peek().kind == Tok::LP) {
        std::vector<std::string> names = probe->outputType()->names();
        std::vector<TypePtr> types = probe->outputType()->children();
        for (uint32_t i = 0; i < build->outputType()->size(); ++i) {
          names.push_back(build->outputType()->nameOf(i));
          types.push_back(build->outputType()->childAt(i));
        }
        filter = expr(ROW(names, types));
      }
      lex_ = after;
      std::vector<std::string> names;
      std::vector<TypePtr> types;
      for (auto& o : outs) {
        const auto& t = o.fromProbe ? probe->outputType() : build->outputType();
        if (o.column < 0 || o.column >= static_cast<int>(t->size())) throw VeloxRuntimeError("plan text: join output column out of range");
        names.push_back(t->nameOf(o.column));  // unique plan-wide: the operators find the column's side by its name
        types.push_back(t->childAt(o.column));
      }
      auto keyExprs = [&](const std::vector<int32_t>& cols, const RowTypePtr& t) {
        std::vector<core::FieldAccessTypedExprPtr> ks;
        for (int32_t c : cols) {
          if (c < 0 || c >= static_cast<int>(t->size())) throw VeloxRuntimeError("plan text: join key out of range");
          ks.push_back(std::make_shared<core::FieldAccessTypedExpr>(t->childAt(c), t->nameOf(c)));
        }
        return ks;
      };
      out = std::make_shared<core::HashJoinNode>(nextId(), type.parse(), false, keyExprs(pk, probe->outputType()), keyExprs(bk, build->outputType()), filter, probe,
                                                 build, ROW(names, types));
    } else {
      throw VeloxRuntimeError("plan text: unknown plan node " + h);
    }
    lex_.expect(Tok::RP, ")");
    return out;
  }

  struct JoinTypeText {
    std::string s;
    core::JoinType parse() const {
      if (s == "inner") return core::JoinType::kInner;
      if (s == "left") return core::JoinType::kLeft;
      if (s == "semi") return core::JoinType::kLeftSemiFilter;
      if (s == "anti") return core::JoinType::kAnti;
      throw VeloxRuntimeError("plan text: unknown join type " + s);
    }
  };

  // skips one balanced S-expression (or atom)
  void skip() {
    if (lex_.peek().kind != Tok::LP) { lex_.take(); return; }
    int depth = 0;
    do {
      const Tok t = lex_.take();
      if (t.kind == Tok::LP) ++depth;
      else if (t.kind == Tok::RP) --depth;
      else if (t.kind == Tok::END) throw VeloxRuntimeError("plan text: unbalanced parentheses");
    } while (depth > 0);
  }

  Lexer lex_;
  int nextId_ = 0;
};

}  // namespace

core::PlanNodePtr parsePlanText(const std::string& text) { return Parser(text).plan(); }

}  // namespace velox_b200
