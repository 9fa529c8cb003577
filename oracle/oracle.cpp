// TEST INFRASTRUCTURE — CPU oracle. Not part of the shipped product path.
//
// CPU restatement of Velox's vectorized operator hot path, written to follow the reference's
// algorithm shape: batch-at-a-time drivers, node-at-a-time expression evaluation over a row
// selection, insertion-ordered groups, sequential accumulation. Citations are into
// /root/reference/velox.
//
//   expression walk ............ expression/Expr.cpp:801-930 (flat-no-nulls), :1235-1268 (default
//                                null rows removed), :1513-1566 (evalAll), :1787-1837 (apply)
//   AND / OR ................... expression/ConjunctExpr.cpp:93-179
//   IF / SWITCH ................ expression/SwitchExpr.cpp:71-180
//   arithmetic ................. functions/prestosql/Arithmetic.h:52-141, common/base/CheckedArithmetic.h:27-60
//   comparisons ................ functions/prestosql/Comparisons.h:24-160, type/FloatingPointUtil.h:52-98
//   LIKE ....................... functions/lib/Re2Functions.cpp:710-733 (prefix fast path; general % _ here)
//   filter result -> rows ...... exec/OperatorUtils.cpp:209-321
//   group by ................... exec/GroupingSet.cpp:288-365,810-884; exec/HashTable.cpp:1751-1838
//   aggregates ................. functions/lib/aggregates/{SumAggregateBase,AverageAggregateBase,
//                                SimpleNumericAggregate}.h, functions/prestosql/aggregates/CountAggregate.cpp:27-110
//   hash join .................. exec/HashBuild.cpp:442-598, exec/HashProbe.cpp:796-900,1189-1437
//   key hashing ................ exec/VectorHasher.cpp:62-126,567-594
//   partitioning ............... exec/HashPartitionFunction.cpp:75-118
#include "oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "hashing.hpp"
#include "sexpr.hpp"

namespace orc {

struct UserError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ----------------------------------------------------------------------------------------------
// Vectors
// ----------------------------------------------------------------------------------------------
inline int width_of(int type) {
  switch (type) {
    case ORC_BOOLEAN: return 1;  // unpacked to bytes internally
    case ORC_INTEGER: return 4;
    case ORC_BIGINT: return 8;
    case ORC_DOUBLE: return 8;
    default: return 0;
  }
}

struct Vec;
using VecPtr = std::shared_ptr<Vec>;

// Flat (possibly dictionary-wrapped) vector; borrowed or owned storage.
struct Vec {
  int type = ORC_BIGINT;
  int64_t n = 0;
  bool is_const = false;  // single value broadcast to n rows (index 0)
  const void* data = nullptr;
  const int32_t* off = nullptr;  // VARCHAR offsets[count+1]
  const char* chars = nullptr;
  const uint8_t* nulls = nullptr;  // byte per row, 1 = null
  const int32_t* idx = nullptr;    // dictionary wrap over base
  VecPtr base;
  std::shared_ptr<std::vector<uint8_t>> own_data, own_nulls;
  std::shared_ptr<std::vector<int32_t>> own_off, own_idx;
  std::shared_ptr<std::string> own_chars;

  template <class T>
  T* alloc(int64_t count) {
    own_data = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(count) * sizeof(T) + 8);
    data = own_data->data();
    return reinterpret_cast<T*>(own_data->data());
  }
  uint8_t* alloc_nulls(int64_t count) {
    own_nulls = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(count), 0);
    nulls = own_nulls->data();
    return own_nulls->data();
  }
  template <class T>
  const T* as() const { return reinterpret_cast<const T*>(data); }
  bool null_at(int64_t r) const { return nulls && nulls[is_const ? 0 : r]; }
};

struct Batch {
  int64_t n = 0;
  std::vector<VecPtr> cols;
};
using Table = std::vector<Batch>;

static bool bit_at(const uint64_t* bits, int64_t i) { return (bits[i >> 6] >> (i & 63)) & 1; }

// Slice rows [r0, r0+n) of a C column into an internal Vec (zero-copy when flat & null-free).
static VecPtr slice_column(const orc_column& c, int64_t r0, int64_t n, std::unordered_map<const void*, VecPtr>& base_cache) {
  auto v = std::make_shared<Vec>();
  v->type = c.type;
  v->n = n;
  auto import_values = [&](Vec& dst, const void* values, const void* aux, int64_t first, int64_t count) {
    if (c.type == ORC_VARCHAR) {
      dst.off = reinterpret_cast<const int32_t*>(values) + first;
      dst.chars = reinterpret_cast<const char*>(aux);
    } else if (c.type == ORC_BOOLEAN) {
      auto* out = dst.alloc<uint8_t>(count);
      auto* bits = reinterpret_cast<const uint64_t*>(values);
      for (int64_t i = 0; i < count; ++i) out[i] = bit_at(bits, first + i);
    } else {
      dst.data = reinterpret_cast<const uint8_t*>(values) + first * width_of(c.type);
    }
  };
  auto import_nulls = [&](Vec& dst, const uint64_t* valid, int64_t first, int64_t count) {
    if (!valid) return;
    bool any = false;
    for (int64_t i = 0; i < count && !any; ++i) any = !bit_at(valid, first + i);
    if (!any) return;
    auto* out = dst.alloc_nulls(count);
    for (int64_t i = 0; i < count; ++i) out[i] = !bit_at(valid, first + i);
  };
  if (c.encoding == ORC_FLAT) {
    import_values(*v, c.values, c.aux, r0, n);
    import_nulls(*v, c.nulls, r0, n);
  } else if (c.encoding == ORC_DICTIONARY) {
    auto it = base_cache.find(c.values);
    if (it == base_cache.end()) {
      auto b = std::make_shared<Vec>();
      b->type = c.type;
      b->n = c.dict_size;
      import_values(*b, c.values, c.aux, 0, c.dict_size);
      import_nulls(*b, c.dict_nulls, 0, c.dict_size);
      it = base_cache.emplace(c.values, b).first;
    }
    v->base = it->second;
    v->idx = c.indices + r0;
    import_nulls(*v, c.nulls, r0, n);
  } else {  // CONSTANT
    v->is_const = true;
    import_values(*v, c.values, c.aux, 0, 1);
    if (c.nulls && !bit_at(c.nulls, 0)) v->alloc_nulls(1)[0] = 1;
  }
  return v;
}

// Materialise dictionary / constant wrapping: result[i] = base[idx[i]] (equal to what the
// reference computes after peeling, expression/Expr.cpp:1135-1188).
static VecPtr flatten(const VecPtr& v) {
  if (!v->idx && !v->is_const) return v;
  auto out = std::make_shared<Vec>();
  out->type = v->type;
  out->n = v->n;
  const int64_t n = v->n;
  const Vec& b = v->idx ? *v->base : *v;
  auto src = [&](int64_t i) -> int64_t { return v->idx ? v->idx[i] : 0; };
  bool any_null = v->nulls || b.nulls;
  uint8_t* on = any_null ? out->alloc_nulls(n) : nullptr;
  if (any_null)
    for (int64_t i = 0; i < n; ++i) {
      bool wn = v->idx ? (v->nulls && v->nulls[i]) : false;
      on[i] = wn || (b.nulls && b.nulls[b.is_const ? 0 : src(i)]);
    }
  if (v->type == ORC_VARCHAR) {
    out->own_off = std::make_shared<std::vector<int32_t>>(n + 1);
    out->own_chars = std::make_shared<std::string>();
    auto& off = *out->own_off;
    off[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (!(on && on[i])) {
        int64_t s = src(i);
        out->own_chars->append(b.chars + b.off[s], b.off[s + 1] - b.off[s]);
      }
      off[i + 1] = static_cast<int32_t>(out->own_chars->size());
    }
    out->off = off.data();
    out->chars = out->own_chars->data();
  } else {
    int w = width_of(v->type);
    auto* o = out->alloc<uint8_t>(n * w);
    auto* s = reinterpret_cast<const uint8_t*>(b.data);
    for (int64_t i = 0; i < n; ++i) {
      if (on && on[i]) { std::memset(o + i * w, 0, w); continue; }
      std::memcpy(o + i * w, s + src(i) * w, w);
    }
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// Row selection (the SelectivityVector analogue: either all of [0,n) or an explicit ascending list)
// ----------------------------------------------------------------------------------------------
struct Rows {
  int64_t n = 0;  // batch size
  bool all = true;
  std::vector<int32_t> list;
  int64_t count() const { return all ? n : static_cast<int64_t>(list.size()); }
  template <class F>
  void for_each(F&& f) const {
    if (all) for (int64_t r = 0; r < n; ++r) f(r);
    else for (int32_t r : list) f(r);
  }
};

// ----------------------------------------------------------------------------------------------
// Expressions
// ----------------------------------------------------------------------------------------------
struct Expr;
using ExprPtr = std::shared_ptr<Expr>;
struct Expr {
  enum Kind { FIELD, CONST, CALL, AND, OR, SWITCH, CAST } kind = CALL;
  int type = ORC_BIGINT;
  int field = -1;
  std::string fn;
  std::vector<ExprPtr> args;
  // constant payload
  bool cnull = false;
  int64_t ci = 0;
  double cd = 0;
  std::string cs;
};

static int parse_type(const std::string& s) {
  if (s == "BOOLEAN") return ORC_BOOLEAN;
  if (s == "INTEGER" || s == "DATE") return ORC_INTEGER;
  if (s == "BIGINT") return ORC_BIGINT;
  if (s == "DOUBLE") return ORC_DOUBLE;
  if (s == "VARCHAR") return ORC_VARCHAR;
  throw std::runtime_error("unknown type " + s);
}

static bool is_cmp(const std::string& f) {
  return f == "lt" || f == "lte" || f == "gt" || f == "gte" || f == "eq" || f == "neq";
}
static bool is_arith(const std::string& f) {
  return f == "plus" || f == "minus" || f == "multiply" || f == "divide" || f == "modulus";
}

static ExprPtr parse_expr(const SNode& s, const std::vector<int>& schema) {
  auto e = std::make_shared<Expr>();
  const std::string& h = s.head();
  if (h == "field") {
    e->kind = Expr::FIELD;
    e->field = std::stoi(s.arg(0).atom);
    if (e->field < 0 || e->field >= static_cast<int>(schema.size())) throw std::runtime_error("field index out of range");
    e->type = schema[e->field];
  } else if (h == "f64") { e->kind = Expr::CONST; e->type = ORC_DOUBLE; e->cd = std::stod(s.arg(0).atom);
  } else if (h == "i64") { e->kind = Expr::CONST; e->type = ORC_BIGINT; e->ci = std::stoll(s.arg(0).atom);
  } else if (h == "i32" || h == "date") { e->kind = Expr::CONST; e->type = ORC_INTEGER; e->ci = std::stoll(s.arg(0).atom);
  } else if (h == "bool") { e->kind = Expr::CONST; e->type = ORC_BOOLEAN; e->ci = s.arg(0).atom == "true";
  } else if (h == "str") { e->kind = Expr::CONST; e->type = ORC_VARCHAR; e->cs = s.arg(0).atom;
  } else if (h == "null") { e->kind = Expr::CONST; e->type = parse_type(s.arg(0).atom); e->cnull = true;
  } else if (h == "cast") {
    e->kind = Expr::CAST;
    e->type = parse_type(s.arg(0).atom);
    e->args.push_back(parse_expr(s.arg(1), schema));
  } else {
    for (size_t i = 0; i < s.nargs(); ++i) e->args.push_back(parse_expr(s.arg(i), schema));
    auto need = [&](size_t k) { if (e->args.size() != k) throw std::runtime_error(h + ": wrong argument count"); };
    auto same = [&]() {
      for (auto& a : e->args)
        if (a->type != e->args[0]->type) throw std::runtime_error(h + ": argument types differ (insert a cast)");
    };
    e->fn = h;
    if (h == "and" || h == "or") {
      e->kind = h == "and" ? Expr::AND : Expr::OR;
      e->type = ORC_BOOLEAN;
      for (auto& a : e->args) if (a->type != ORC_BOOLEAN) throw std::runtime_error(h + ": BOOLEAN arguments expected");
    } else if (h == "switch" || h == "if") {
      e->kind = Expr::SWITCH;
      if (e->args.size() < 2) throw std::runtime_error("switch: too few arguments");
      e->type = e->args[1]->type;
    } else if (is_arith(h)) { need(2); same(); e->type = e->args[0]->type;
      if (e->type != ORC_DOUBLE && e->type != ORC_BIGINT && e->type != ORC_INTEGER) throw std::runtime_error(h + ": numeric arguments expected");
    } else if (h == "negate") { need(1); e->type = e->args[0]->type;
    } else if (is_cmp(h)) { need(2); same(); e->type = ORC_BOOLEAN;
    } else if (h == "between") { need(3); same(); e->type = ORC_BOOLEAN;
    } else if (h == "not") { need(1); e->type = ORC_BOOLEAN;
    } else if (h == "is_null") { need(1); e->type = ORC_BOOLEAN;
    } else if (h == "like") { need(2); e->type = ORC_BOOLEAN;
      if (e->args[1]->kind != Expr::CONST) throw std::runtime_error("like: constant pattern expected");
    } else {
      throw std::runtime_error("unknown function " + h);
    }
  }
  return e;
}

// SQL LIKE with % and _ (no escape). The reference special-cases prefix patterns
// (functions/lib/Re2Functions.cpp:710-733); results are identical.
static bool like_match(const char* s, int64_t sl, const char* p, int64_t pl) {
  int64_t si = 0, pi = 0, star = -1, mark = 0;
  while (si < sl) {
    if (pi < pl && (p[pi] == '_' || p[pi] == s[si])) { ++si; ++pi; }
    else if (pi < pl && p[pi] == '%') { star = pi++; mark = si; }
    else if (star >= 0) { pi = star + 1; si = ++mark; }
    else return false;
  }
  while (pi < pl && p[pi] == '%') ++pi;
  return pi == pl;
}

struct EvalCtx {
  const Batch* in;
  std::vector<VecPtr> flat_cache;  // flattened input columns
  // Per-row error capture inside AND / OR: an error on a row that a later (or earlier) conjunct
  // decides is dropped (expression/ConjunctExpr.cpp:98-99,167-168); outside it throws at once.
  int capture = 0;
  std::vector<uint8_t> row_err;
  std::string first_err;
  void fail(int64_t r, const char* msg) {
    if (capture == 0) throw UserError(msg);
    if (row_err.empty()) row_err.assign(in->n, 0);
    row_err[r] = 1;
    if (first_err.empty()) first_err = msg;
  }
};

static VecPtr input_flat(EvalCtx& ctx, int i) {
  if (ctx.flat_cache.size() < ctx.in->cols.size()) ctx.flat_cache.resize(ctx.in->cols.size());
  if (!ctx.flat_cache[i]) ctx.flat_cache[i] = flatten(ctx.in->cols[i]);
  return ctx.flat_cache[i];
}

static VecPtr make_result(int type, int64_t n) {
  auto v = std::make_shared<Vec>();
  v->type = type;
  v->n = n;
  return v;
}

template <class T>
struct Acc {
  const T* p;
  int64_t stride;
  explicit Acc(const Vec& v) : p(v.as<T>()), stride(v.is_const ? 0 : 1) {}
  T operator[](int64_t r) const { return p[r * stride]; }
};

static VecPtr eval(const Expr& e, EvalCtx& ctx, const Rows& rows);

// NaN-aware comparisons: type/FloatingPointUtil.h:52-98 (NaN is the largest value, NaN == NaN).
static inline bool cmp_f64(int op, double a, double b) {
  switch (op) {
    case 0: return (!std::isnan(a) && std::isnan(b)) ? true : a < b;
    case 1: return std::isnan(b) ? true : a <= b;
    case 2: return (std::isnan(a) && !std::isnan(b)) ? true : a > b;
    case 3: return std::isnan(a) ? true : a >= b;
    case 4: return (std::isnan(a) && std::isnan(b)) ? true : a == b;
    default: return !((std::isnan(a) && std::isnan(b)) ? true : a == b);
  }
}
template <class T>
static inline bool cmp_int(int op, T a, T b) {
  switch (op) {
    case 0: return a < b;
    case 1: return a <= b;
    case 2: return a > b;
    case 3: return a >= b;
    case 4: return a == b;
    default: return a != b;
  }
}
static int cmp_code(const std::string& f) {
  if (f == "lt") return 0;
  if (f == "lte") return 1;
  if (f == "gt") return 2;
  if (f == "gte") return 3;
  if (f == "eq") return 4;
  return 5;
}

template <class T>
static bool checked_arith(int op, T a, T b, T* out) {
  switch (op) {
    case 0: return !__builtin_add_overflow(a, b, out);
    case 1: return !__builtin_sub_overflow(a, b, out);
    case 2: return !__builtin_mul_overflow(a, b, out);
    case 3:
      if (b == 0) return false;
      if (a == std::numeric_limits<T>::min() && b == -1) return false;
      *out = a / b;
      return true;
    default:
      if (b == 0) return false;
      if (b == -1) { *out = 0; return true; }
      *out = a % b;
      return true;
  }
}
static int arith_code(const std::string& f) {
  if (f == "plus") return 0;
  if (f == "minus") return 1;
  if (f == "multiply") return 2;
  if (f == "divide") return 3;
  return 4;
}

// Default-null behaviour: rows where any argument is null produce null and the function is not
// called on them (expression/Expr.cpp:1235-1268).
template <class F>
static void for_non_null(const Rows& rows, const std::vector<VecPtr>& args, Vec& out, F&& f) {
  bool any = false;
  for (auto& a : args) any = any || a->nulls;
  if (!any) { rows.for_each(f); return; }
  uint8_t* on = out.alloc_nulls(out.n);
  rows.for_each([&](int64_t r) {
    for (auto& a : args)
      if (a->null_at(r)) { on[r] = 1; return; }
    f(r);
  });
}

static VecPtr eval_call(const Expr& e, EvalCtx& ctx, const Rows& rows) {
  const int64_t n = rows.n;
  std::vector<VecPtr> a;
  for (auto& x : e.args) a.push_back(eval(*x, ctx, rows));
  auto out = make_result(e.type, n);
  const std::string& f = e.fn;
  if (is_arith(f)) {
    int op = arith_code(f);
    if (e.type == ORC_DOUBLE) {
      auto* o = out->alloc<double>(n);
      Acc<double> x(*a[0]), y(*a[1]);
      // Plain IEEE-754 binary64, one rounding per operation (no FMA contraction):
      // functions/prestosql/Arithmetic.h:52-141.
      for_non_null(rows, a, *out, [&](int64_t r) {
        const double l = x[r], rr = y[r];  // built with -ffp-contract=off
        switch (op) {
          case 0: o[r] = l + rr; break;
          case 1: o[r] = l - rr; break;
          case 2: o[r] = l * rr; break;
          case 3: o[r] = l / rr; break;
          default: o[r] = std::fmod(l, rr); break;
        }
      });
    } else if (e.type == ORC_BIGINT) {
      auto* o = out->alloc<int64_t>(n);
      Acc<int64_t> x(*a[0]), y(*a[1]);
      for_non_null(rows, a, *out, [&](int64_t r) {
        if (!checked_arith<int64_t>(op, x[r], y[r], &o[r])) ctx.fail(r, op >= 3 && y[r] == 0 ? "division by zero" : "integer overflow");
      });
    } else {
      auto* o = out->alloc<int32_t>(n);
      Acc<int32_t> x(*a[0]), y(*a[1]);
      for_non_null(rows, a, *out, [&](int64_t r) {
        if (!checked_arith<int32_t>(op, x[r], y[r], &o[r])) ctx.fail(r, op >= 3 && y[r] == 0 ? "division by zero" : "integer overflow");
      });
    }
  } else if (f == "negate") {
    if (e.type == ORC_DOUBLE) {
      auto* o = out->alloc<double>(n);
      Acc<double> x(*a[0]);
      for_non_null(rows, a, *out, [&](int64_t r) { o[r] = -x[r]; });
    } else if (e.type == ORC_BIGINT) {
      auto* o = out->alloc<int64_t>(n);
      Acc<int64_t> x(*a[0]);
      for_non_null(rows, a, *out, [&](int64_t r) {
        if (x[r] == std::numeric_limits<int64_t>::min()) { ctx.fail(r, "integer overflow"); return; }
        o[r] = -x[r];
      });
    } else {
      auto* o = out->alloc<int32_t>(n);
      Acc<int32_t> x(*a[0]);
      for_non_null(rows, a, *out, [&](int64_t r) {
        if (x[r] == std::numeric_limits<int32_t>::min()) { ctx.fail(r, "integer overflow"); return; }
        o[r] = -x[r];
      });
    }
  } else if (is_cmp(f) || f == "between") {
    auto* o = out->alloc<uint8_t>(n);
    int t = e.args[0]->type;
    bool btw = f == "between";
    int op = btw ? 0 : cmp_code(f);
    if (t == ORC_DOUBLE) {
      Acc<double> x(*a[0]), y(*a[1]);
      if (btw) {
        Acc<double> z(*a[2]);
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = cmp_f64(3, x[r], y[r]) && cmp_f64(1, x[r], z[r]); });
      } else {
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = cmp_f64(op, x[r], y[r]); });
      }
    } else if (t == ORC_BIGINT) {
      Acc<int64_t> x(*a[0]), y(*a[1]);
      if (btw) {
        Acc<int64_t> z(*a[2]);
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = x[r] >= y[r] && x[r] <= z[r]; });
      } else {
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = cmp_int<int64_t>(op, x[r], y[r]); });
      }
    } else if (t == ORC_INTEGER) {
      Acc<int32_t> x(*a[0]), y(*a[1]);
      if (btw) {
        Acc<int32_t> z(*a[2]);
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = x[r] >= y[r] && x[r] <= z[r]; });
      } else {
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = cmp_int<int32_t>(op, x[r], y[r]); });
      }
    } else if (t == ORC_BOOLEAN) {
      Acc<uint8_t> x(*a[0]), y(*a[1]);
      if (btw) throw std::runtime_error("between on BOOLEAN unsupported");
      for_non_null(rows, a, *out, [&](int64_t r) { o[r] = cmp_int<int>(op, x[r], y[r]); });
    } else {  // VARCHAR: bytewise compare
      auto sv = [&](const Vec& v, int64_t r) {
        int64_t i = v.is_const ? 0 : r;
        return std::string_view(v.chars + v.off[i], v.off[i + 1] - v.off[i]);
      };
      if (btw) {
        for_non_null(rows, a, *out, [&](int64_t r) { o[r] = sv(*a[0], r) >= sv(*a[1], r) && sv(*a[0], r) <= sv(*a[2], r); });
      } else {
        for_non_null(rows, a, *out, [&](int64_t r) {
          int c = sv(*a[0], r).compare(sv(*a[1], r));
          o[r] = cmp_int<int>(op, c, 0);
        });
      }
    }
  } else if (f == "not") {
    auto* o = out->alloc<uint8_t>(n);
    Acc<uint8_t> x(*a[0]);
    for_non_null(rows, a, *out, [&](int64_t r) { o[r] = !x[r]; });
  } else if (f == "is_null") {
    auto* o = out->alloc<uint8_t>(n);
    rows.for_each([&](int64_t r) { o[r] = a[0]->null_at(r); });
  } else if (f == "like") {
    auto* o = out->alloc<uint8_t>(n);
    const std::string& pat = e.args[1]->cs;
    const Vec& s = *a[0];
    std::vector<VecPtr> only{a[0]};
    if (e.args[1]->cnull) {
      uint8_t* on = out->alloc_nulls(n);
      rows.for_each([&](int64_t r) { on[r] = 1; });
    } else {
      for_non_null(rows, only, *out, [&](int64_t r) {
        int64_t i = s.is_const ? 0 : r;
        o[r] = like_match(s.chars + s.off[i], s.off[i + 1] - s.off[i], pat.data(), pat.size());
      });
    }
  } else {
    throw std::runtime_error("unknown function " + f);
  }
  return out;
}

static VecPtr eval_const(const Expr& e, int64_t n) {
  auto v = make_result(e.type, n);
  v->is_const = true;
  switch (e.type) {
    case ORC_DOUBLE: v->alloc<double>(1)[0] = e.cd; break;
    case ORC_BIGINT: v->alloc<int64_t>(1)[0] = e.ci; break;
    case ORC_INTEGER: v->alloc<int32_t>(1)[0] = static_cast<int32_t>(e.ci); break;
    case ORC_BOOLEAN: v->alloc<uint8_t>(1)[0] = e.ci != 0; break;
    default: {
      v->own_off = std::make_shared<std::vector<int32_t>>(std::vector<int32_t>{0, static_cast<int32_t>(e.cs.size())});
      v->own_chars = std::make_shared<std::string>(e.cs);
      v->off = v->own_off->data();
      v->chars = v->own_chars->data();
    }
  }
  if (e.cnull) v->alloc_nulls(1)[0] = 1;
  return v;
}

// Three-valued AND / OR over a shrinking row set (expression/ConjunctExpr.cpp:93-179): a row
// leaves the active set as soon as it is decided (false for AND, true for OR); null is
// remembered and only wins if no later conjunct decides the row.
static VecPtr eval_conjunct(const Expr& e, EvalCtx& ctx, const Rows& rows, bool is_and) {
  const int64_t n = rows.n;
  auto out = make_result(ORC_BOOLEAN, n);
  auto* o = out->alloc<uint8_t>(n);
  std::vector<uint8_t> saw_null, saw_err;
  rows.for_each([&](int64_t r) { o[r] = is_and; });
  Rows active = rows;
  ++ctx.capture;
  for (auto& arg : e.args) {
    if (active.count() == 0) break;
    VecPtr v = eval(*arg, ctx, active);
    Acc<uint8_t> x(*v);
    std::vector<int32_t> keep;
    keep.reserve(active.count());
    active.for_each([&](int64_t r) {
      if (!ctx.row_err.empty() && ctx.row_err[r]) {  // errored on this conjunct: undecided so far
        ctx.row_err[r] = 0;
        if (saw_err.empty()) saw_err.assign(n, 0);
        saw_err[r] = 1;
        keep.push_back(static_cast<int32_t>(r));
      } else if (v->null_at(r)) {
        if (saw_null.empty()) saw_null.assign(n, 0);
        saw_null[r] = 1;
        keep.push_back(static_cast<int32_t>(r));
      } else if (static_cast<bool>(x[r]) != is_and) {
        o[r] = !is_and;  // decided
        if (!saw_null.empty()) saw_null[r] = 0;
        if (!saw_err.empty()) saw_err[r] = 0;
      } else {
        keep.push_back(static_cast<int32_t>(r));
      }
    });
    active.all = false;
    active.list.swap(keep);
  }
  --ctx.capture;
  if (!saw_err.empty()) {
    const std::string msg = ctx.first_err.empty() ? "arithmetic error" : ctx.first_err;
    rows.for_each([&](int64_t r) { if (saw_err[r]) ctx.fail(r, msg.c_str()); });  // propagates or throws
  }
  if (!saw_null.empty()) {
    uint8_t* on = out->alloc_nulls(n);
    rows.for_each([&](int64_t r) { on[r] = saw_null[r]; });
  }
  return out;
}

template <class T>
static void copy_rows(const Vec& src, Vec& dst, const Rows& rows, uint8_t*& dn) {
  T* d = const_cast<T*>(dst.as<T>());
  Acc<T> s(src);
  rows.for_each([&](int64_t r) {
    if (src.null_at(r)) {
      if (!dn) dn = dst.alloc_nulls(dst.n);
      dn[r] = 1;
    } else {
      d[r] = s[r];
    }
  });
}

// CASE: THEN evaluated only on rows whose condition is true; a null condition is false
// (expression/SwitchExpr.cpp:101-152).
static VecPtr eval_switch(const Expr& e, EvalCtx& ctx, const Rows& rows) {
  const int64_t n = rows.n;
  if (e.type == ORC_VARCHAR) throw std::runtime_error("switch over VARCHAR unsupported in oracle");
  auto out = make_result(e.type, n);
  int w = width_of(e.type);
  out->alloc<uint8_t>(n * w);
  uint8_t* dn = nullptr;
  Rows remaining = rows;
  auto assign = [&](const Expr& val, const Rows& sel) {
    if (sel.count() == 0) return;
    VecPtr v = eval(val, ctx, sel);
    switch (e.type) {
      case ORC_DOUBLE: copy_rows<double>(*v, *out, sel, dn); break;
      case ORC_BIGINT: copy_rows<int64_t>(*v, *out, sel, dn); break;
      case ORC_INTEGER: copy_rows<int32_t>(*v, *out, sel, dn); break;
      default: copy_rows<uint8_t>(*v, *out, sel, dn); break;
    }
  };
  size_t i = 0;
  for (; i + 1 < e.args.size(); i += 2) {
    if (remaining.count() == 0) break;
    VecPtr c = eval(*e.args[i], ctx, remaining);
    Acc<uint8_t> x(*c);
    Rows then_rows, else_rows;
    then_rows.n = else_rows.n = n;
    then_rows.all = else_rows.all = false;
    remaining.for_each([&](int64_t r) {
      if (!c->null_at(r) && x[r]) then_rows.list.push_back(static_cast<int32_t>(r));
      else else_rows.list.push_back(static_cast<int32_t>(r));
    });
    assign(*e.args[i + 1], then_rows);
    remaining = std::move(else_rows);
  }
  if (e.args.size() % 2 == 1) {
    assign(*e.args.back(), remaining);
  } else {
    remaining.for_each([&](int64_t r) {
      if (!dn) dn = out->alloc_nulls(n);
      dn[r] = 1;
    });
  }
  return out;
}

static VecPtr eval_cast(const Expr& e, EvalCtx& ctx, const Rows& rows) {
  VecPtr a = eval(*e.args[0], ctx, rows);
  int from = e.args[0]->type, to = e.type;
  if (from == to) return a;
  const int64_t n = rows.n;
  auto out = make_result(to, n);
  std::vector<VecPtr> args{a};
  auto num = [&](auto tag_from, auto tag_to) {
    using F = decltype(tag_from);
    using T = decltype(tag_to);
    T* o = out->alloc<T>(n);
    Acc<F> x(*a);
    for_non_null(rows, args, *out, [&](int64_t r) {
      F v = x[r];
      if constexpr (std::is_floating_point_v<F> && std::is_integral_v<T>) {
        if (std::isnan(v)) throw UserError("Cannot cast NaN to an integral value");
        double rounded = std::round(v);
        if (rounded < static_cast<double>(std::numeric_limits<T>::min()) ||
            rounded >= -static_cast<double>(std::numeric_limits<T>::min()))
          throw UserError("Cannot cast DOUBLE to integer: out of range");
        o[r] = static_cast<T>(rounded);
      } else if constexpr (std::is_integral_v<F> && std::is_integral_v<T> && sizeof(T) < sizeof(F)) {
        if (v < std::numeric_limits<T>::min() || v > std::numeric_limits<T>::max()) throw UserError("Cannot cast: out of range");
        o[r] = static_cast<T>(v);
      } else {
        o[r] = static_cast<T>(v);
      }
    });
  };
  if (from == ORC_BIGINT && to == ORC_DOUBLE) num(int64_t{}, double{});
  else if (from == ORC_INTEGER && to == ORC_DOUBLE) num(int32_t{}, double{});
  else if (from == ORC_INTEGER && to == ORC_BIGINT) num(int32_t{}, int64_t{});
  else if (from == ORC_BIGINT && to == ORC_INTEGER) num(int64_t{}, int32_t{});
  else if (from == ORC_DOUBLE && to == ORC_BIGINT) num(double{}, int64_t{});
  else if (from == ORC_DOUBLE && to == ORC_INTEGER) num(double{}, int32_t{});
  else if (from == ORC_BOOLEAN && to == ORC_BIGINT) num(uint8_t{}, int64_t{});
  else if (to == ORC_BOOLEAN && from != ORC_VARCHAR) {
    // velox/type/Conversions.h:158-207 (non-truncating policy): folly::to<bool>(v) == (v != 0); NaN != 0 is true
    uint8_t* o = out->alloc<uint8_t>(n);
    auto to_bool = [&](auto tag) {
      using F = decltype(tag);
      Acc<F> x(*a);
      for_non_null(rows, args, *out, [&](int64_t r) { o[r] = x[r] != F{} ? 1 : 0; });
    };
    if (from == ORC_DOUBLE) to_bool(double{});
    else if (from == ORC_BIGINT) to_bool(int64_t{});
    else to_bool(int32_t{});
  }
  else throw std::runtime_error("unsupported cast");
  return out;
}

static VecPtr eval(const Expr& e, EvalCtx& ctx, const Rows& rows) {
  switch (e.kind) {
    case Expr::FIELD: return input_flat(ctx, e.field);
    case Expr::CONST: return eval_const(e, rows.n);
    case Expr::AND: return eval_conjunct(e, ctx, rows, true);
    case Expr::OR: return eval_conjunct(e, ctx, rows, false);
    case Expr::SWITCH: return eval_switch(e, ctx, rows);
    case Expr::CAST: return eval_cast(e, ctx, rows);
    default: return eval_call(e, ctx, rows);
  }
}

// Gather rows of a (possibly wrapped) vector into a dense flat vector.
static VecPtr gather(const VecPtr& vin, const int32_t* sel, int64_t m) {
  // sel[i] < 0 produces null (used for outer-join misses).
  const Vec& v = *vin;
  auto out = make_result(v.type, m);
  const Vec& b = v.idx ? *v.base : v;
  uint8_t* on = nullptr;
  auto set_null = [&](int64_t i) {
    if (!on) on = out->alloc_nulls(m);
    on[i] = 1;
  };
  auto resolve = [&](int64_t i, int64_t& s) -> bool {  // returns false when null
    int64_t r = sel[i];
    if (r < 0) return false;
    if (v.is_const) { s = 0; return !(v.nulls && v.nulls[0]); }
    if (v.idx) {
      if (v.nulls && v.nulls[r]) return false;
      s = v.idx[r];
      return !(b.nulls && b.nulls[s]);
    }
    s = r;
    return !(v.nulls && v.nulls[r]);
  };
  if (v.type == ORC_VARCHAR) {
    out->own_off = std::make_shared<std::vector<int32_t>>(m + 1);
    out->own_chars = std::make_shared<std::string>();
    auto& off = *out->own_off;
    off[0] = 0;
    for (int64_t i = 0; i < m; ++i) {
      int64_t s;
      if (resolve(i, s)) out->own_chars->append(b.chars + b.off[s], b.off[s + 1] - b.off[s]);
      else set_null(i);
      off[i + 1] = static_cast<int32_t>(out->own_chars->size());
    }
    out->off = off.data();
    out->chars = out->own_chars->data();
  } else {
    int w = width_of(v.type);
    auto* o = out->alloc<uint8_t>(m * w);
    auto* src = reinterpret_cast<const uint8_t*>(b.data);
    for (int64_t i = 0; i < m; ++i) {
      int64_t s;
      if (resolve(i, s)) std::memcpy(o + i * w, src + s * w, w);
      else { std::memset(o + i * w, 0, w); set_null(i); }
    }
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// Plan nodes
// ----------------------------------------------------------------------------------------------
struct AggSpec {
  std::string fn;  // sum avg count min max
  int input = -1;  // input column (raw) or first intermediate column
  int mask = -1;
  int in_type = ORC_BIGINT;
  bool distinct = false;  // AggregationNode::Aggregate::distinct (core/PlanNode.h:1152)
};

struct Node;
using NodePtr = std::shared_ptr<Node>;
struct Node {
  enum Kind { VALUES, FILTER, PROJECT, AGG, JOIN, ORDERBY } kind = VALUES;
  std::vector<int> schema;  // output column types
  NodePtr child, build;
  int source = 0;
  ExprPtr filter;
  std::vector<ExprPtr> projections;
  // aggregation
  std::string step;
  std::vector<int> keys;
  std::vector<AggSpec> aggs;
  // join
  std::string join_type;
  std::vector<int> probe_keys, build_keys;
  std::vector<std::pair<char, int>> join_out;  // ('p'|'b', column)
  int64_t limit = -1;  // TopN: rows kept (exec/TopN.cpp); -1 = all
  // order by (exec/OrderBy.cpp; core::SortOrder{ascending, nullsFirst}): (column, ascending, nulls first)
  struct SortKey { int col; bool asc; bool nulls_first; };
  std::vector<SortKey> sort_keys;
};

static bool raw_input_step(const std::string& s) { return s == "single" || s == "partial"; }
static bool final_output_step(const std::string& s) { return s == "single" || s == "final"; }

static int sum_type(int t) { return t == ORC_DOUBLE ? ORC_DOUBLE : ORC_BIGINT; }

static NodePtr parse_plan(const SNode& s) {
  auto n = std::make_shared<Node>();
  const std::string& h = s.head();
  if (h == "values") {
    n->kind = Node::VALUES;
    n->source = std::stoi(s.arg(0).atom);
    for (auto& t : s.arg(1).kids) n->schema.push_back(parse_type(t.atom));
  } else if (h == "filter") {
    n->kind = Node::FILTER;
    n->child = parse_plan(s.arg(1));
    n->filter = parse_expr(s.arg(0), n->child->schema);
    if (n->filter->type != ORC_BOOLEAN) throw std::runtime_error("filter must be BOOLEAN");
    n->schema = n->child->schema;
  } else if (h == "project") {
    n->kind = Node::PROJECT;
    n->child = parse_plan(s.arg(1));
    for (auto& e : s.arg(0).kids) {
      n->projections.push_back(parse_expr(e, n->child->schema));
      n->schema.push_back(n->projections.back()->type);
    }
  } else if (h == "exchange") {
    // (exchange KIND (keys I ...) plan): PartitionedOutput -> Exchange between the tasks of a
    // distributed plan (exec/PartitionedOutput.cpp, exec/Exchange.cpp). The oracle runs the whole
    // plan over the whole data in one process, where the shuffle is the identity on the row multiset.
    return parse_plan(s.arg(2));
  } else if (h == "localpartition") {
    // (localpartition plan): gathers the drivers of a pipeline (exec/LocalPartition.cpp); the identity on the row multiset
    return parse_plan(s.arg(0));
  } else if (h == "orderby" || h == "topn") {
    // (orderby ((I asc|desc first|last) ...) plan) | (topn N ((I asc|desc first|last) ...) plan): exec/TopN.cpp keeps
    // the first N rows of the order
    n->kind = Node::ORDERBY;
    const size_t at = h == "topn" ? 1 : 0;
    if (h == "topn") {
      n->limit = std::stoll(s.arg(0).atom);
      if (n->limit <= 0) throw std::runtime_error("topn: count must be positive");
    }
    n->child = parse_plan(s.arg(at + 1));
    n->schema = n->child->schema;
    for (auto& k : s.arg(at).kids) {
      Node::SortKey sk;
      sk.col = std::stoi(k.head());
      sk.asc = k.arg(0).atom == "asc";
      sk.nulls_first = k.arg(1).atom == "first";
      if (sk.col < 0 || sk.col >= static_cast<int>(n->schema.size())) throw std::runtime_error("orderby: bad column");
      n->sort_keys.push_back(sk);
    }
  } else if (h == "aggregation") {
    n->kind = Node::AGG;
    n->step = s.arg(0).atom;
    n->child = parse_plan(s.arg(3));
    const auto& in = n->child->schema;
    for (size_t i = 0; i < s.arg(1).nargs(); ++i) n->keys.push_back(std::stoi(s.arg(1).arg(i).atom));
    for (int k : n->keys) n->schema.push_back(in.at(k));
    bool raw = raw_input_step(n->step), fin = final_output_step(n->step);
    for (size_t i = 0; i < s.arg(2).nargs(); ++i) {
      const SNode& a = s.arg(2).arg(i);
      AggSpec spec;
      spec.fn = a.head();
      for (size_t j = 0; j < a.nargs(); ++j) {
        if (a.arg(j).is_list) {
          if (a.arg(j).head() == "mask") spec.mask = std::stoi(a.arg(j).arg(0).atom);
          else if (a.arg(j).head() == "distinct") spec.distinct = true;
        } else {
          spec.input = std::stoi(a.arg(j).atom);
        }
      }
      if (spec.input >= 0) spec.in_type = in.at(spec.input);
      const std::string& f = spec.fn;
      if (f == "sum") n->schema.push_back(raw ? sum_type(spec.in_type) : spec.in_type);
      else if (f == "count") n->schema.push_back(ORC_BIGINT);
      else if (f == "min" || f == "max") n->schema.push_back(spec.in_type);
      else if (f == "avg") {
        if (fin) n->schema.push_back(ORC_DOUBLE);
        else { n->schema.push_back(ORC_DOUBLE); n->schema.push_back(ORC_BIGINT); }
      } else throw std::runtime_error("unknown aggregate " + f);
      n->aggs.push_back(spec);
    }
  } else if (h == "hashjoin") {
    n->kind = Node::JOIN;
    n->join_type = s.arg(0).atom;
    n->child = parse_plan(s.arg(5));
    n->build = parse_plan(s.arg(6));
    for (size_t i = 0; i < s.arg(1).nargs(); ++i) n->probe_keys.push_back(std::stoi(s.arg(1).arg(i).atom));
    for (size_t i = 0; i < s.arg(2).nargs(); ++i) n->build_keys.push_back(std::stoi(s.arg(2).arg(i).atom));
    // Join filter sees probe columns followed by build columns.
    std::vector<int> both = n->child->schema;
    both.insert(both.end(), n->build->schema.begin(), n->build->schema.end());
    if (s.arg(3).is_list) n->filter = parse_expr(s.arg(3), both);
    for (size_t i = 0; i < s.arg(4).nargs(); ++i) {
      const SNode& o = s.arg(4).arg(i);
      char side = o.head()[0];
      int c = std::stoi(o.arg(0).atom);
      n->join_out.emplace_back(side, c);
      n->schema.push_back(side == 'p' ? n->child->schema.at(c) : n->build->schema.at(c));
    }
    if (n->probe_keys.size() != n->build_keys.size() || n->probe_keys.empty()) throw std::runtime_error("hashjoin: bad keys");
  } else {
    throw std::runtime_error("unknown plan node " + h);
  }
  return n;
}

// ----------------------------------------------------------------------------------------------
// Key handling shared by group-by, join and partitioning
// ----------------------------------------------------------------------------------------------
// 64-bit open-addressing map (linear probing) from a key word to a dense id.
struct FlatMap64 {
  std::vector<uint64_t> keys;
  std::vector<int32_t> vals;
  std::vector<uint8_t> used;
  size_t mask = 0, count = 0;
  FlatMap64() { rehash(1024); }
  void rehash(size_t cap) {
    std::vector<uint64_t> ok;
    std::vector<int32_t> ov;
    std::vector<uint8_t> ou;
    ok.swap(keys); ov.swap(vals); ou.swap(used);
    keys.assign(cap, 0); vals.assign(cap, 0); used.assign(cap, 0);
    mask = cap - 1;
    for (size_t i = 0; i < ok.size(); ++i)
      if (ou[i]) {
        size_t p = twang_mix64(ok[i]) & mask;
        while (used[p]) p = (p + 1) & mask;
        keys[p] = ok[i]; vals[p] = ov[i]; used[p] = 1;
      }
  }
  // Returns id for key; assigns next_id if absent (is_new set).
  int32_t find_or_insert(uint64_t k, int32_t next_id, bool& is_new) {
    if ((count + 1) * 10 > keys.size() * 7) rehash(keys.size() * 2);  // load factor 0.7 (exec/HashTable.h:143)
    size_t p = twang_mix64(k) & mask;
    while (used[p]) {
      if (keys[p] == k) { is_new = false; return vals[p]; }
      p = (p + 1) & mask;
    }
    used[p] = 1; keys[p] = k; vals[p] = next_id; ++count;
    is_new = true;
    return next_id;
  }
  int32_t find(uint64_t k) const {
    size_t p = twang_mix64(k) & mask;
    while (used[p]) {
      if (keys[p] == k) return vals[p];
      p = (p + 1) & mask;
    }
    return -1;
  }
};

// Per-key-column value ids: every distinct (non-null) value gets a dense id in first-seen
// order; null is id 0 (exec/VectorHasher.h:523-585 reserves 0 for null likewise). Doubles are
// keyed by canonical bits (NaN == NaN, +0 == -0 as in NaNAwareHash / folly float_hasher).
struct ValueIds {
  int type;
  FlatMap64 ints;
  std::unordered_map<std::string, int32_t> strs;
  int32_t next = 1;
  // integer keys of a join build side over a dense range: ids follow from the value (no table), set by JoinTable::build
  bool range_mode = false;
  int64_t range_min = 0, range_max = -1;
  // cache for dictionary bases
  const Vec* cached_base = nullptr;
  std::vector<int32_t> base_ids;

  static uint64_t word(const Vec& v, int64_t s) {
    switch (v.type) {
      case ORC_BIGINT: return static_cast<uint64_t>(v.as<int64_t>()[s]);
      case ORC_INTEGER: return static_cast<uint64_t>(static_cast<int64_t>(v.as<int32_t>()[s]));
      case ORC_BOOLEAN: return v.as<uint8_t>()[s];
      default: {
        double d = v.as<double>()[s];
        if (std::isnan(d)) d = std::numeric_limits<double>::quiet_NaN();
        if (d == 0.0) d = 0.0;
        uint64_t u;
        std::memcpy(&u, &d, 8);
        return u;
      }
    }
  }
  int32_t id_of(const Vec& flat, int64_t s, bool insert) {
    if (flat.nulls && flat.nulls[s]) return 0;
    if (type == ORC_VARCHAR) {
      std::string k(flat.chars + flat.off[s], flat.off[s + 1] - flat.off[s]);
      auto it = strs.find(k);
      if (it != strs.end()) return it->second;
      if (!insert) return -1;
      strs.emplace(std::move(k), next);
      return next++;
    }
    uint64_t w = word(flat, s);
    if (range_mode) {  // id = v - min + 1 (VectorHasher range mode, exec/VectorHasher.h:523-585); values outside the range are unknown
      const int64_t v = static_cast<int64_t>(w);
      return (v < range_min || v > range_max) ? -1 : static_cast<int32_t>(v - range_min + 1);
    }
    if (!insert) return ints.find(w);
    bool is_new;
    int32_t id = ints.find_or_insert(w, next, is_new);
    if (is_new) ++next;
    return id;
  }
  // ids for all rows of a batch column. insert=false => unknown values give -1 and the hasher
  // is not modified (safe to share between probe drivers).
  void ids(const VecPtr& col, int64_t n, bool insert, std::vector<int32_t>& out) {
    out.resize(n);
    const Vec& v = *col;
    if (v.is_const) {
      int32_t id = id_of(v, 0, insert);
      std::fill(out.begin(), out.end(), id);
    } else if (v.idx) {
      const Vec& b = *v.base;
      std::vector<int32_t> local;
      std::vector<int32_t>* cache = &local;
      if (insert) {
        if (cached_base != &b) { base_ids.assign(b.n, -2); cached_base = &b; }
        cache = &base_ids;
      } else {
        local.assign(b.n, -2);
      }
      for (int64_t r = 0; r < n; ++r) {
        if (v.nulls && v.nulls[r]) { out[r] = 0; continue; }
        int32_t& c = (*cache)[v.idx[r]];
        if (c == -2) c = id_of(b, v.idx[r], insert);
        out[r] = c;
      }
    } else {
      for (int64_t r = 0; r < n; ++r) out[r] = id_of(v, r, insert);
    }
  }
};

// Builder for output columns.
struct ColBuilder {
  int type;
  std::vector<int64_t> i64;
  std::vector<double> f64;
  std::vector<int32_t> i32;
  std::vector<uint8_t> b8;
  std::vector<int32_t> off{0};
  std::string chars;
  std::vector<uint8_t> nulls;
  bool any_null = false;
  explicit ColBuilder(int t) : type(t) {}
  int64_t size() const { return static_cast<int64_t>(nulls.size()); }
  void push_null() {
    nulls.push_back(1); any_null = true;
    switch (type) {
      case ORC_BIGINT: i64.push_back(0); break;
      case ORC_DOUBLE: f64.push_back(0); break;
      case ORC_INTEGER: i32.push_back(0); break;
      case ORC_BOOLEAN: b8.push_back(0); break;
      default: off.push_back(off.back());
    }
  }
  void push_from(const Vec& flat, int64_t s) {  // flat (no idx) source
    if (flat.nulls && flat.nulls[flat.is_const ? 0 : s]) { push_null(); return; }
    if (flat.is_const) s = 0;
    nulls.push_back(0);
    switch (type) {
      case ORC_BIGINT: i64.push_back(flat.as<int64_t>()[s]); break;
      case ORC_DOUBLE: f64.push_back(flat.as<double>()[s]); break;
      case ORC_INTEGER: i32.push_back(flat.as<int32_t>()[s]); break;
      case ORC_BOOLEAN: b8.push_back(flat.as<uint8_t>()[s]); break;
      default:
        chars.append(flat.chars + flat.off[s], flat.off[s + 1] - flat.off[s]);
        off.push_back(static_cast<int32_t>(chars.size()));
    }
  }
  void push_i64(int64_t v) { nulls.push_back(0); i64.push_back(v); }
  void push_f64(double v) { nulls.push_back(0); f64.push_back(v); }
  VecPtr finish() {
    auto v = make_result(type, size());
    auto take = [&](auto& vec) {
      using T = typename std::remove_reference_t<decltype(vec)>::value_type;
      T* o = v->alloc<T>(vec.size());
      std::memcpy(o, vec.data(), vec.size() * sizeof(T));
    };
    switch (type) {
      case ORC_BIGINT: take(i64); break;
      case ORC_DOUBLE: take(f64); break;
      case ORC_INTEGER: take(i32); break;
      case ORC_BOOLEAN: take(b8); break;
      default:
        v->own_off = std::make_shared<std::vector<int32_t>>(off);
        v->own_chars = std::make_shared<std::string>(chars);
        v->off = v->own_off->data();
        v->chars = v->own_chars->data();
    }
    if (any_null) {
      uint8_t* on = v->alloc_nulls(size());
      std::memcpy(on, nulls.data(), nulls.size());
    }
    return v;
  }
};

// ----------------------------------------------------------------------------------------------
// Group by
// ----------------------------------------------------------------------------------------------
struct Accumulator {
  // One aggregate over all groups (SoA). sum/avg(double): dsum; sum(bigint): isum (checked);
  // count / avg count: cnt; has: saw a non-null input (functions/lib/aggregates/SumAggregateBase.h:143-150).
  std::vector<double> dsum;
  std::vector<int64_t> isum, cnt;
  std::vector<uint8_t> has;
  // DISTINCT aggregates (exec/DistinctAggregations.cpp: a set of the group's inputs, whose values reach the
  // function once each, in first-seen order): (group, canonical value bits) pairs seen so far.
  struct PairHash {
    size_t operator()(const std::pair<int32_t, uint64_t>& p) const { return twang_mix64(p.second ^ (static_cast<uint64_t>(p.first) * 0x9E3779B97F4A7C15ull)); }
  };
  std::unordered_set<std::pair<int32_t, uint64_t>, PairHash> seen;
  void grow(size_t g) { dsum.resize(g, 0); isum.resize(g, 0); cnt.resize(g, 0); has.resize(g, 0); }
};

struct GroupBy {
  const Node& node;
  bool raw, fin;
  std::vector<ValueIds> hashers;
  // group lookup: array over combined value ids while small (kArray, exec/HashTable.h:146),
  // hash map afterwards (kHash).
  static constexpr uint64_t kArrayMax = 2ull << 20;
  std::vector<uint64_t> caps;  // per-key id capacity in array mode
  std::vector<int32_t> array;  // combined id -> group+1
  bool array_mode = true;
  FlatMap64 combos;            // hash mode, <= 2 keys packed
  std::unordered_map<std::string, int32_t> wide;  // hash mode, > 2 keys
  std::vector<std::vector<int32_t>> group_ids;    // per key: value id of each group
  std::vector<ColBuilder> key_cols;
  std::vector<Accumulator> accs;
  int64_t ngroups = 0;

  explicit GroupBy(const Node& n) : node(n), raw(raw_input_step(n.step)), fin(final_output_step(n.step)) {
    const auto& in = n.child->schema;
    for (int k : n.keys) {
      hashers.emplace_back();
      hashers.back().type = in[k];
      key_cols.emplace_back(in[k]);
    }
    caps.assign(n.keys.size(), 4);
    group_ids.resize(n.keys.size());
    accs.resize(n.aggs.size());
    if (n.keys.empty()) new_group();
    else rebuild_array();
  }
  void new_group() {
    ++ngroups;
    for (auto& a : accs) a.grow(ngroups);
  }
  uint64_t array_size() const {
    uint64_t s = 1;
    for (auto c : caps) { s *= c; if (s > kArrayMax) return s; }
    return s;
  }
  void rebuild_array() {
    if (array_size() > kArrayMax) { to_hash_mode(); return; }
    array.assign(array_size(), 0);
    for (int64_t g = 0; g < ngroups; ++g) array[combined(g)] = static_cast<int32_t>(g + 1);
  }
  uint64_t combined(int64_t g) const {
    uint64_t c = 0;
    for (size_t k = 0; k < caps.size(); ++k) c = c * caps[k] + group_ids[k][g];
    return c;
  }
  std::string wide_key(const std::vector<int32_t>& ids) const {
    return std::string(reinterpret_cast<const char*>(ids.data()), ids.size() * 4);
  }
  void to_hash_mode() {
    array_mode = false;
    array.clear(); array.shrink_to_fit();
    std::vector<int32_t> ids(caps.size());
    for (int64_t g = 0; g < ngroups; ++g) {
      for (size_t k = 0; k < caps.size(); ++k) ids[k] = group_ids[k][g];
      bool is_new;
      if (caps.size() <= 2) combos.find_or_insert(pack(ids), static_cast<int32_t>(g), is_new);
      else wide.emplace(wide_key(ids), static_cast<int32_t>(g));
    }
  }
  static uint64_t pack(const std::vector<int32_t>& ids) {
    uint64_t w = static_cast<uint32_t>(ids[0]);
    if (ids.size() > 1) w |= static_cast<uint64_t>(static_cast<uint32_t>(ids[1])) << 32;
    return w;
  }

  void add(const Batch& b) {
    const int64_t n = b.n;
    std::vector<int32_t> groups(n, 0);
    const size_t nk = node.keys.size();
    if (nk > 0) {
      std::vector<std::vector<int32_t>> ids(nk);
      std::vector<VecPtr> flat_keys(nk);
      for (size_t k = 0; k < nk; ++k) {
        hashers[k].ids(b.cols[node.keys[k]], n, true, ids[k]);
        bool grew = false;
        while (static_cast<uint64_t>(hashers[k].next) > caps[k]) { caps[k] *= 2; grew = true; }
        if (grew && array_mode) rebuild_array();
      }
      std::vector<int32_t> row_ids(nk);
      for (int64_t r = 0; r < n; ++r) {
        int32_t g;
        if (array_mode) {
          uint64_t c = 0;
          for (size_t k = 0; k < nk; ++k) c = c * caps[k] + ids[k][r];
          int32_t& slot = array[c];
          if (!slot) { slot = static_cast<int32_t>(ngroups + 1); g = -1; } else g = slot - 1;
        } else {
          for (size_t k = 0; k < nk; ++k) row_ids[k] = ids[k][r];
          bool is_new = false;
          if (nk <= 2) {
            g = combos.find_or_insert(pack(row_ids), static_cast<int32_t>(ngroups), is_new);
          } else {
            auto it = wide.emplace(wide_key(row_ids), static_cast<int32_t>(ngroups));
            is_new = it.second;
            g = it.first->second;
          }
          if (is_new) g = -1;
        }
        if (g < 0) {  // first-seen order (exec/GroupingSet.cpp:826-840)
          g = static_cast<int32_t>(ngroups);
          for (size_t k = 0; k < nk; ++k) {
            group_ids[k].push_back(ids[k][r]);
            if (!flat_keys[k]) flat_keys[k] = flatten(b.cols[node.keys[k]]);
            key_cols[k].push_from(*flat_keys[k], r);
          }
          new_group();
        }
        groups[r] = g;
      }
    }
    for (size_t a = 0; a < node.aggs.size(); ++a) update(a, b, groups);
  }

  // Scatter-update in input order (functions/lib/aggregates/SimpleNumericAggregate.h:94-150).
  void update(size_t ai, const Batch& b, const std::vector<int32_t>& groups) {
    const AggSpec& s = node.aggs[ai];
    Accumulator& acc = accs[ai];
    const int64_t n = b.n;
    VecPtr in = s.input >= 0 ? flatten(b.cols[s.input]) : nullptr;
    VecPtr in2 = (!raw && s.fn == "avg") ? flatten(b.cols[s.input + 1]) : nullptr;
    VecPtr mask = s.mask >= 0 ? flatten(b.cols[s.mask]) : nullptr;
    // DISTINCT: rows whose (group, value) pair was seen before are skipped like masked rows
    std::vector<uint8_t> repeat;
    if (s.distinct) {
      if (!raw || !in) throw std::runtime_error("distinct aggregates take raw input over a column");
      repeat.assign(n, 0);
      const int64_t st = in->is_const ? 0 : 1;
      for (int64_t r = 0; r < n; ++r) {
        if ((mask && (mask->null_at(r) || !mask->as<uint8_t>()[r])) || in->null_at(r)) continue;
        uint64_t bits;
        switch (in->type) {
          case ORC_DOUBLE: {
            double d = in->as<double>()[r * st];
            if (std::isnan(d)) d = std::numeric_limits<double>::quiet_NaN();
            if (d == 0.0) d = 0.0;  // -0 and +0 are one value
            std::memcpy(&bits, &d, 8);
            if (std::isnan(d)) bits = 0x7ff8000000000000ull;
            break;
          }
          case ORC_BIGINT: bits = static_cast<uint64_t>(in->as<int64_t>()[r * st]); break;
          case ORC_INTEGER: bits = static_cast<uint64_t>(static_cast<int64_t>(in->as<int32_t>()[r * st])); break;
          case ORC_BOOLEAN: bits = in->as<uint8_t>()[r * st] ? 1 : 0; break;
          default: throw std::runtime_error("distinct aggregates over this input type are not restated");
        }
        if (!acc.seen.emplace(groups[r], bits).second) repeat[r] = 1;
      }
    }
    auto masked_out = [&](int64_t r) { return (mask && (mask->null_at(r) || !mask->as<uint8_t>()[r])) || (s.distinct && repeat[r]); };
    const std::string& f = s.fn;
    if (f == "count") {
      for (int64_t r = 0; r < n; ++r) {
        if (masked_out(r)) continue;
        if (raw) { if (in && in->null_at(r)) continue; acc.cnt[groups[r]] += 1; }
        else { if (in->null_at(r)) continue; acc.cnt[groups[r]] += in->as<int64_t>()[r]; }
      }
      return;
    }
    const bool no_skip = !mask && !in->nulls && !s.distinct;
    auto each = [&](auto&& body) {
      if (no_skip) { for (int64_t r = 0; r < n; ++r) body(r); }
      else { for (int64_t r = 0; r < n; ++r) { if (masked_out(r) || in->null_at(r)) continue; body(r); } }
    };
    const int64_t stride = in->is_const ? 0 : 1;
    const int32_t* g = groups.data();
    if (f == "sum" || f == "avg") {
      const bool as_double = f == "avg" || in->type == ORC_DOUBLE;
      if (as_double) {
        double* acc_d = acc.dsum.data();
        // one rounding per add, input order (functions/lib/aggregates/SumAggregateBase.h:71-142)
        if (in->type == ORC_DOUBLE) { const double* v = in->as<double>(); each([&](int64_t r) { acc_d[g[r]] += v[r * stride]; }); }
        else if (in->type == ORC_BIGINT) { const int64_t* v = in->as<int64_t>(); each([&](int64_t r) { acc_d[g[r]] += static_cast<double>(v[r * stride]); }); }
        else { const int32_t* v = in->as<int32_t>(); each([&](int64_t r) { acc_d[g[r]] += static_cast<double>(v[r * stride]); }); }
      } else {
        int64_t* acc_i = acc.isum.data();
        auto add = [&](int64_t r, int64_t v) { if (__builtin_add_overflow(acc_i[g[r]], v, &acc_i[g[r]])) throw UserError("integer overflow in sum"); };
        if (in->type == ORC_BIGINT) { const int64_t* v = in->as<int64_t>(); each([&](int64_t r) { add(r, v[r * stride]); }); }
        else { const int32_t* v = in->as<int32_t>(); each([&](int64_t r) { add(r, v[r * stride]); }); }
      }
      if (f == "avg") {
        int64_t* c = acc.cnt.data();
        if (raw) each([&](int64_t r) { c[g[r]] += 1; });
        else { const int64_t* v2 = in2->as<int64_t>(); const int64_t s2 = in2->is_const ? 0 : 1; each([&](int64_t r) { c[g[r]] += v2[r * s2]; }); }
      }
      uint8_t* has = acc.has.data();
      each([&](int64_t r) { has[g[r]] = 1; });
    } else {  // min / max (NaN is largest for doubles)
      const bool is_min = f == "min";
      each([&](int64_t r) {
        const int32_t gi = g[r];
        const int64_t s0 = r * stride;
        if (in->type == ORC_DOUBLE) {
          double v = in->as<double>()[s0];
          if (!acc.has[gi] || (is_min ? cmp_f64(0, v, acc.dsum[gi]) : cmp_f64(2, v, acc.dsum[gi]))) acc.dsum[gi] = v;
        } else {
          int64_t v = in->type == ORC_BIGINT ? in->as<int64_t>()[s0]
                    : in->type == ORC_INTEGER ? in->as<int32_t>()[s0] : in->as<uint8_t>()[s0];
          if (!acc.has[gi] || (is_min ? v < acc.isum[gi] : v > acc.isum[gi])) acc.isum[gi] = v;
        }
        acc.has[gi] = 1;
      });
    }
  }

  Batch output() {
    Batch out;
    out.n = ngroups;
    for (auto& k : key_cols) out.cols.push_back(k.finish());
    for (size_t a = 0; a < node.aggs.size(); ++a) {
      const AggSpec& s = node.aggs[a];
      Accumulator& acc = accs[a];
      const std::string& f = s.fn;
      if (f == "count") {
        ColBuilder c(ORC_BIGINT);
        for (int64_t g = 0; g < ngroups; ++g) c.push_i64(acc.cnt[g]);
        out.cols.push_back(c.finish());
      } else if (f == "avg") {
        if (fin) {
          ColBuilder c(ORC_DOUBLE);
          for (int64_t g = 0; g < ngroups; ++g) {
            if (!acc.has[g] || acc.cnt[g] == 0) c.push_null();
            else c.push_f64(acc.dsum[g] / static_cast<double>(acc.cnt[g]));  // AverageAggregateBase.h:86-107
          }
          out.cols.push_back(c.finish());
        } else {
          ColBuilder c(ORC_DOUBLE), k(ORC_BIGINT);
          for (int64_t g = 0; g < ngroups; ++g) {
            if (!acc.has[g]) { c.push_null(); k.push_null(); }
            else { c.push_f64(acc.dsum[g]); k.push_i64(acc.cnt[g]); }
          }
          out.cols.push_back(c.finish());
          out.cols.push_back(k.finish());
        }
      } else {
        int t = f == "sum" ? (raw ? sum_type(s.in_type) : s.in_type) : s.in_type;
        ColBuilder c(t);
        for (int64_t g = 0; g < ngroups; ++g) {
          if (!acc.has[g]) { c.push_null(); continue; }
          c.nulls.push_back(0);
          switch (t) {
            case ORC_DOUBLE: c.f64.push_back(acc.dsum[g]); break;
            case ORC_BIGINT: c.i64.push_back(acc.isum[g]); break;
            case ORC_INTEGER: c.i32.push_back(static_cast<int32_t>(acc.isum[g])); break;
            default: c.b8.push_back(static_cast<uint8_t>(acc.isum[g]));
          }
        }
        out.cols.push_back(c.finish());
      }
    }
    return out;
  }
};

// Runs fn(i) for i in [0, n) on up to `threads` threads (contiguous ranges per thread).
template <class F>
static void parallel_for(int64_t n, int threads, F&& fn) {
  const int T = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, n)));
  if (T <= 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
  std::vector<std::thread> ts;
  std::exception_ptr err;
  std::mutex m;
  for (int t = 0; t < T; ++t)
    ts.emplace_back([&, t] {
      try {
        for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i) fn(i);
      } catch (...) {
        std::lock_guard<std::mutex> l(m);
        if (!err) err = std::current_exception();
      }
    });
  for (auto& th : ts) th.join();
  if (err) std::rethrow_exception(err);
}

// ----------------------------------------------------------------------------------------------
// Hash join
// ----------------------------------------------------------------------------------------------
struct JoinTable {
  // Build rows concatenated into one batch; key -> first row, next[] chains duplicates
  // (exec/HashTable.cpp:1518 insertForJoin; null keys never inserted, exec/HashBuild.cpp:475-479).
  Batch rows;
  std::vector<VecPtr> flat_cols;
  std::vector<ValueIds> hashers;
  FlatMap64 combos;
  std::unordered_map<std::string, int32_t> wide;
  std::vector<int32_t> first, next_row;
  size_t nk = 0;

  static uint64_t pack(const std::vector<int32_t>& ids) { return GroupBy::pack(ids); }

  // `threads`: the build side is concatenated by all drivers' threads and, for one integer key over a dense range, ids
  // follow from the values — what the reference gets from per-driver row containers + parallelJoinBuild
  // (exec/HashBuild.cpp:819-993, exec/HashTable.cpp:1003) and VectorHasher's range mode; the chains themselves are
  // linked in row order either way.
  void build(const Node& n, Table&& t, int threads = 1) {
    nk = n.build_keys.size();
    const auto& schema = n.build->schema;
    const int64_t nb = static_cast<int64_t>(t.size());
    // concatenate: row offset of every batch, then every (column, batch) piece on its own
    std::vector<int64_t> row0(nb + 1, 0);
    for (int64_t i = 0; i < nb; ++i) row0[i + 1] = row0[i] + t[i].n;
    rows.n = row0[nb];
    const size_t ncols = schema.size();
    std::vector<std::vector<VecPtr>> flats(ncols, std::vector<VecPtr>(nb));
    parallel_for(static_cast<int64_t>(ncols) * nb, threads, [&](int64_t j) {
      const size_t c = static_cast<size_t>(j / nb);
      const int64_t i = j % nb;
      flats[c][i] = flatten(t[i].cols[c]);
    });
    for (size_t c = 0; c < ncols; ++c) {
      bool any_null = false, any_const = false;
      for (auto& f : flats[c]) { any_null = any_null || f->nulls; any_const = any_const || f->is_const; }
      if (any_const) {  // rare: value by value
        ColBuilder cb(schema[c]);
        for (int64_t i = 0; i < nb; ++i)
          for (int64_t r = 0; r < t[i].n; ++r) cb.push_from(*flats[c][i], r);
        rows.cols.push_back(cb.finish());
        continue;
      }
      auto v = make_result(schema[c], rows.n);
      uint8_t* on = any_null ? v->alloc_nulls(rows.n) : nullptr;
      if (schema[c] != ORC_VARCHAR) {
        const int w = width_of(schema[c]);
        uint8_t* o = v->alloc<uint8_t>(rows.n * w);
        parallel_for(nb, threads, [&](int64_t i) {
          const Vec& f = *flats[c][i];
          std::memcpy(o + row0[i] * w, f.data, static_cast<size_t>(t[i].n) * w);
          if (on && f.nulls) {
            std::memcpy(on + row0[i], f.nulls, static_cast<size_t>(t[i].n));
            for (int64_t r = 0; r < t[i].n; ++r)
              if (f.nulls[r]) std::memset(o + (row0[i] + r) * w, 0, w);  // NULL rows hold zeros, as ColBuilder leaves them
          }
        });
      } else {
        // chars of every batch (NULL rows contribute none), then offsets and bytes batch by batch
        std::vector<int64_t> char0(nb + 1, 0);
        parallel_for(nb, threads, [&](int64_t i) {
          const Vec& f = *flats[c][i];
          int64_t len = 0;
          if (!f.nulls) len = f.off[t[i].n] - f.off[0];
          else
            for (int64_t r = 0; r < t[i].n; ++r)
              if (!f.nulls[r]) len += f.off[r + 1] - f.off[r];
          char0[i + 1] = len;
        });
        for (int64_t i = 0; i < nb; ++i) char0[i + 1] += char0[i];
        if (char0[nb] >= (1ll << 31)) throw std::runtime_error("join build side: VARCHAR column above 2 GiB");
        v->own_off = std::make_shared<std::vector<int32_t>>(static_cast<size_t>(rows.n) + 1);
        v->own_chars = std::make_shared<std::string>(static_cast<size_t>(char0[nb]), '\0');
        int32_t* off = v->own_off->data();
        char* chars = v->own_chars->data();
        off[rows.n] = static_cast<int32_t>(char0[nb]);
        parallel_for(nb, threads, [&](int64_t i) {
          const Vec& f = *flats[c][i];
          int64_t at = char0[i];
          for (int64_t r = 0; r < t[i].n; ++r) {
            off[row0[i] + r] = static_cast<int32_t>(at);
            if (f.nulls && f.nulls[r]) continue;
            const int32_t len = f.off[r + 1] - f.off[r];
            std::memcpy(chars + at, f.chars + f.off[r], static_cast<size_t>(len));
            at += len;
          }
          if (on && f.nulls) std::memcpy(on + row0[i], f.nulls, static_cast<size_t>(t[i].n));
        });
        v->off = off;
        v->chars = chars;
      }
      if (on) {
        bool any = false;
        for (int64_t r = 0; r < rows.n && !any; ++r) any = on[r];
        if (!any) { v->own_nulls.reset(); v->nulls = nullptr; }
      }
      rows.cols.push_back(v);
    }
    flat_cols = rows.cols;
    for (size_t k = 0; k < nk; ++k) { hashers.emplace_back(); hashers.back().type = schema[n.build_keys[k]]; }
    // one BIGINT / INTEGER key over a dense range: range-mode ids
    bool ranged = false;
    if (nk == 1 && (schema[n.build_keys[0]] == ORC_BIGINT || schema[n.build_keys[0]] == ORC_INTEGER) && rows.n > 0) {
      const Vec& kv = *rows.cols[n.build_keys[0]];
      int64_t lo = INT64_MAX, hi = INT64_MIN;
      for (int64_t r = 0; r < rows.n; ++r) {
        if (kv.nulls && kv.nulls[r]) continue;
        const int64_t x = kv.type == ORC_BIGINT ? kv.as<int64_t>()[r] : kv.as<int32_t>()[r];
        lo = std::min(lo, x);
        hi = std::max(hi, x);
      }
      if (lo <= hi) {
        const unsigned __int128 span = static_cast<unsigned __int128>(static_cast<__int128>(hi) - lo) + 1;
        if (span <= static_cast<unsigned __int128>(std::max<int64_t>(1 << 16, rows.n * 4)) && span < (1u << 30)) {
          ranged = true;
          hashers[0].range_mode = true;
          hashers[0].range_min = lo;
          hashers[0].range_max = hi;
          first.assign(static_cast<size_t>(span), -1);
        }
      }
    }
    std::vector<std::vector<int32_t>> ids(nk);
    if (ranged) {
      const Vec& kv = *rows.cols[n.build_keys[0]];
      ids[0].resize(static_cast<size_t>(rows.n));
      const int64_t lo = hashers[0].range_min;
      parallel_for((rows.n + 65535) / 65536, threads, [&](int64_t blk) {
        for (int64_t r = blk * 65536; r < std::min<int64_t>(rows.n, (blk + 1) * 65536); ++r) {
          if (kv.nulls && kv.nulls[r]) { ids[0][r] = 0; continue; }
          const int64_t x = kv.type == ORC_BIGINT ? kv.as<int64_t>()[r] : kv.as<int32_t>()[r];
          ids[0][r] = static_cast<int32_t>(x - lo + 1);
        }
      });
    } else {
      for (size_t k = 0; k < nk; ++k) hashers[k].ids(rows.cols[n.build_keys[k]], rows.n, true, ids[k]);
    }
    next_row.assign(rows.n, -1);
    std::vector<int32_t> row_ids(nk);
    std::vector<int32_t> last;  // tail of each chain so matches stay in build order
    if (ranged) last.assign(first.size(), -1);
    for (int64_t r = 0; r < rows.n; ++r) {
      bool has_null = false;
      for (size_t k = 0; k < nk; ++k) { row_ids[k] = ids[k][r]; has_null |= ids[k][r] == 0; }
      if (has_null) continue;
      int32_t e;
      bool is_new = false;
      if (ranged) {  // ids index the chains directly; -1 = no row with this key yet
        e = row_ids[0] - 1;
        if (first[e] < 0) { first[e] = static_cast<int32_t>(r); last[e] = static_cast<int32_t>(r); }
        else { next_row[last[e]] = static_cast<int32_t>(r); last[e] = static_cast<int32_t>(r); }
        continue;
      }
      if (nk == 1) {  // value ids are dense (1..N): they index the chains directly
        e = row_ids[0] - 1;
        is_new = static_cast<size_t>(e) >= first.size();
      } else if (nk <= 2) e = combos.find_or_insert(pack(row_ids), static_cast<int32_t>(first.size()), is_new);
      else {
        auto it = wide.emplace(std::string(reinterpret_cast<const char*>(row_ids.data()), nk * 4), static_cast<int32_t>(first.size()));
        is_new = it.second; e = it.first->second;
      }
      if (is_new) { first.push_back(static_cast<int32_t>(r)); last.push_back(static_cast<int32_t>(r)); }
      else { next_row[last[e]] = static_cast<int32_t>(r); last[e] = static_cast<int32_t>(r); }
    }
  }
  // first build row matching probe row ids, or -1
  int32_t lookup(const std::vector<int32_t>& row_ids) const {
    for (auto id : row_ids) if (id <= 0) return -1;  // null (0) or unseen value (-1)
    int32_t e;
    if (nk == 1) e = row_ids[0] - 1 < static_cast<int32_t>(first.size()) ? row_ids[0] - 1 : -1;
    else if (nk <= 2) e = combos.find(pack(row_ids));
    else {
      auto it = wide.find(std::string(reinterpret_cast<const char*>(row_ids.data()), nk * 4));
      e = it == wide.end() ? -1 : it->second;
    }
    return e < 0 ? -1 : first[e];
  }
};

static Batch probe_join(const Node& n, JoinTable& jt, const Batch& b) {
  const size_t nk = jt.nk;
  std::vector<std::vector<int32_t>> ids(nk);
  for (size_t k = 0; k < nk; ++k) {
    ValueIds& h = jt.hashers[k];
    h.ids(b.cols[n.probe_keys[k]], b.n, false, ids[k]);
  }
  std::vector<int32_t> pr, br;  // candidate pairs in probe order
  std::vector<int32_t> row_ids(nk);
  const std::string& jt_type = n.join_type;
  bool inner = jt_type == "inner", left = jt_type == "left", semi = jt_type == "semi", anti = jt_type == "anti";
  if (!(inner || left || semi || anti)) throw std::runtime_error("unsupported join type " + jt_type);
  for (int64_t r = 0; r < b.n; ++r) {
    for (size_t k = 0; k < nk; ++k) row_ids[k] = ids[k][r];
    for (int32_t m = jt.lookup(row_ids); m >= 0; m = jt.next_row[m]) { pr.push_back(static_cast<int32_t>(r)); br.push_back(m); }
  }
  // optional filter over (probe cols ++ build cols)
  std::vector<uint8_t> pass(pr.size(), 1);
  if (n.filter && !pr.empty()) {
    Batch fb;
    fb.n = static_cast<int64_t>(pr.size());
    for (auto& c : b.cols) fb.cols.push_back(gather(c, pr.data(), fb.n));
    for (auto& c : jt.flat_cols) fb.cols.push_back(gather(c, br.data(), fb.n));
    EvalCtx ctx{&fb, {}};
    Rows rows; rows.n = fb.n;
    VecPtr v = eval(*n.filter, ctx, rows);
    Acc<uint8_t> x(*v);
    for (int64_t i = 0; i < fb.n; ++i) pass[i] = !v->null_at(i) && x[i];
  }
  std::vector<int32_t> op, ob;
  if (inner) {
    for (size_t i = 0; i < pr.size(); ++i) if (pass[i]) { op.push_back(pr[i]); ob.push_back(br[i]); }
  } else {
    std::vector<uint8_t> matched(b.n, 0);
    for (size_t i = 0; i < pr.size(); ++i) if (pass[i]) matched[pr[i]] = 1;
    if (left) {
      size_t i = 0;
      for (int64_t r = 0; r < b.n; ++r) {
        bool any = false;
        for (; i < pr.size() && pr[i] == r; ++i) if (pass[i]) { op.push_back(pr[i]); ob.push_back(br[i]); any = true; }
        if (!any) { op.push_back(static_cast<int32_t>(r)); ob.push_back(-1); }
      }
    } else if (semi) {
      for (int64_t r = 0; r < b.n; ++r) if (matched[r]) { op.push_back(static_cast<int32_t>(r)); ob.push_back(-1); }
    } else {  // anti (not null-aware): probe rows with no match
      for (int64_t r = 0; r < b.n; ++r) if (!matched[r]) { op.push_back(static_cast<int32_t>(r)); ob.push_back(-1); }
    }
  }
  Batch out;
  out.n = static_cast<int64_t>(op.size());
  for (auto& o : n.join_out) {
    if (o.first == 'p') out.cols.push_back(gather(b.cols[o.second], op.data(), out.n));
    else out.cols.push_back(gather(jt.flat_cols[o.second], ob.data(), out.n));
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// FilterProject
// ----------------------------------------------------------------------------------------------
static Batch apply_filter(const Node& n, const Batch& b) {
  EvalCtx ctx{&b, {}};
  Rows rows; rows.n = b.n;
  VecPtr v = eval(*n.filter, ctx, rows);
  Acc<uint8_t> x(*v);
  std::vector<int32_t> sel;
  sel.reserve(b.n);
  // keep rows whose predicate is true and not null (exec/OperatorUtils.cpp:238-248)
  for (int64_t r = 0; r < b.n; ++r) if (!v->null_at(r) && x[r]) sel.push_back(static_cast<int32_t>(r));
  Batch out;
  out.n = static_cast<int64_t>(sel.size());
  if (out.n == b.n) { out.cols = b.cols; return out; }
  if (out.n == 0) return out;
  // The reference wraps survivors in a dictionary (exec/Operator.cpp:270-303); flat gather here,
  // except that already-wrapped columns stay wrapped (indices composed).
  for (auto& c : b.cols) {
    if (c->idx && !c->nulls) {
      auto w = std::make_shared<Vec>();
      w->type = c->type; w->n = out.n; w->base = c->base;
      w->own_idx = std::make_shared<std::vector<int32_t>>(out.n);
      for (int64_t i = 0; i < out.n; ++i) (*w->own_idx)[i] = c->idx[sel[i]];
      w->idx = w->own_idx->data();
      out.cols.push_back(w);
    } else {
      out.cols.push_back(gather(c, sel.data(), out.n));
    }
  }
  return out;
}

static Batch apply_project(const Node& n, const Batch& b) {
  EvalCtx ctx{&b, {}};
  Rows rows; rows.n = b.n;
  Batch out;
  out.n = b.n;
  for (auto& e : n.projections) {
    if (e->kind == Expr::FIELD) { out.cols.push_back(b.cols[e->field]); continue; }  // identity: zero copy
    VecPtr v = eval(*e, ctx, rows);
    if (v->is_const) {
      std::vector<int32_t> z(b.n, 0);
      v->is_const = false; v->n = 1;
      auto one = v;
      v = gather(one, z.data(), b.n);
    }
    out.cols.push_back(v);
  }
  return out;
}

// ----------------------------------------------------------------------------------------------
// Executor: N driver threads over contiguous shards of the driving source; blocking operators
// (aggregation, join build) merge per-driver state, mirroring partial -> localPartition -> final
// (exec/tests/utils/TpchQueryBuilder.cpp:235-246).
// ----------------------------------------------------------------------------------------------
struct Executor {
  const orc_table* sources;
  int nsources;
  int threads;
  int64_t batch_rows;

  // OrderBy (exec/OrderBy.cpp + exec/SortBuffer.cpp): materialise, sort row numbers with a stable
  // sort, gather. NULLs go first or last as the key says, independent of the direction; doubles
  // compare NaN-largest (type/FloatingPointUtil.h), strings bytewise.
  Table run_orderby(const NodePtr& node) {
    Table in = materialise(node->child);
    const size_t nc = node->schema.size();
    std::vector<ColBuilder> cb;
    for (int ty : node->schema) cb.emplace_back(ty);
    for (auto& b : in)
      for (size_t c = 0; c < nc; ++c) {
        VecPtr f = flatten(b.cols[c]);
        for (int64_t r = 0; r < b.n; ++r) cb[c].push_from(*f, r);
      }
    std::vector<VecPtr> flat;
    for (auto& c : cb) flat.push_back(c.finish());
    const int64_t n = flat.empty() ? 0 : flat[0]->n;
    std::vector<int64_t> idx(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) idx[static_cast<size_t>(i)] = i;
    auto cmp3 = [&](const Vec& v, int64_t a, int64_t b) -> int {
      switch (v.type) {
        case ORC_BIGINT: { auto x = v.as<int64_t>()[a], y = v.as<int64_t>()[b]; return x < y ? -1 : (x > y ? 1 : 0); }
        case ORC_INTEGER: { auto x = v.as<int32_t>()[a], y = v.as<int32_t>()[b]; return x < y ? -1 : (x > y ? 1 : 0); }
        case ORC_BOOLEAN: { auto x = v.as<uint8_t>()[a], y = v.as<uint8_t>()[b]; return x < y ? -1 : (x > y ? 1 : 0); }
        case ORC_DOUBLE: {
          const double x = v.as<double>()[a], y = v.as<double>()[b];
          const bool xn = std::isnan(x), yn = std::isnan(y);
          if (xn || yn) return xn == yn ? 0 : (xn ? 1 : -1);  // NaN is the largest value
          return x < y ? -1 : (x > y ? 1 : 0);
        }
        default: {
          const int32_t la = v.off[a + 1] - v.off[a], lb = v.off[b + 1] - v.off[b];
          const int c = std::memcmp(v.chars + v.off[a], v.chars + v.off[b], static_cast<size_t>(std::min(la, lb)));
          return c != 0 ? (c < 0 ? -1 : 1) : (la < lb ? -1 : (la > lb ? 1 : 0));
        }
      }
    };
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
      for (auto& k : node->sort_keys) {
        const Vec& v = *flat[static_cast<size_t>(k.col)];
        const bool an = v.null_at(a), bn = v.null_at(b);
        if (an || bn) {
          if (an == bn) continue;
          return an ? k.nulls_first : !k.nulls_first;
        }
        const int c = cmp3(v, a, b);
        if (c != 0) return k.asc ? c < 0 : c > 0;
      }
      return false;
    });
    std::vector<ColBuilder> ob;
    for (int ty : node->schema) ob.emplace_back(ty);
    const int64_t kept = node->limit >= 0 ? std::min<int64_t>(n, node->limit) : n;
    for (int64_t i = 0; i < kept; ++i)
      for (size_t c = 0; c < nc; ++c) ob[c].push_from(*flat[c], idx[static_cast<size_t>(i)]);
    Batch out;
    out.n = kept;
    for (auto& c : ob) out.cols.push_back(c.finish());
    Table t;
    if (kept > 0) t.push_back(std::move(out));
    return t;
  }

  Table materialise(const NodePtr& node) {
    if (node->kind == Node::AGG) return run_aggregation(node);
    if (node->kind == Node::ORDERBY) return run_orderby(node);
    std::vector<Table> per_thread(threads);
    stream(node, [&](int t, Batch&& b) { per_thread[t].push_back(std::move(b)); });
    Table out;
    for (auto& t : per_thread) for (auto& b : t) out.push_back(std::move(b));
    return out;
  }

  Table run_aggregation(const NodePtr& node) {
    bool single_driver = !raw_input_step(node->step);  // final / intermediate run on one driver
    for (auto& a : node->aggs) single_driver = single_driver || a.distinct;  // distinct sets do not merge across drivers
    int T = single_driver ? 1 : threads;
    std::vector<std::unique_ptr<GroupBy>> gbs;
    Node partial = *node;
    if (node->step == "single" && T > 1) partial.step = "partial";
    // per-driver nodes must outlive the GroupBy objects
    std::vector<std::unique_ptr<Node>> keep;
    for (int t = 0; t < T; ++t) {
      keep.push_back(std::make_unique<Node>(partial));
      gbs.push_back(std::make_unique<GroupBy>(*keep.back()));
    }
    stream(node->child, [&](int t, Batch&& b) { gbs[t]->add(b); }, T);
    Table out;
    if (node->step == "single" && T > 1) {
      // merge: final aggregation over the drivers' intermediate outputs
      auto fin = std::make_unique<Node>(*node);
      fin->step = "final";
      auto mid = std::make_shared<Node>();
      mid->schema = keep[0]->schema;  // partial output schema
      // recompute partial schema
      {
        mid->schema.clear();
        for (int k : node->keys) mid->schema.push_back(node->child->schema[k]);
        for (auto& a : node->aggs) {
          if (a.fn == "sum") mid->schema.push_back(sum_type(a.in_type));
          else if (a.fn == "count") mid->schema.push_back(ORC_BIGINT);
          else if (a.fn == "avg") { mid->schema.push_back(ORC_DOUBLE); mid->schema.push_back(ORC_BIGINT); }
          else mid->schema.push_back(a.in_type);
        }
      }
      fin->child = mid;
      fin->keys.clear();
      int c = 0;
      for (size_t k = 0; k < node->keys.size(); ++k) fin->keys.push_back(c++);
      for (auto& a : fin->aggs) {
        a.input = c;
        a.mask = -1;
        a.in_type = mid->schema[c];
        c += a.fn == "avg" ? 2 : 1;
      }
      GroupBy merged(*fin);
      for (auto& g : gbs) {
        Batch b = g->output();
        if (b.n > 0) merged.add(b);
      }
      out.push_back(merged.output());
    } else {
      for (auto& g : gbs) {
        Batch b = g->output();
        if (b.n > 0 || node->keys.empty()) out.push_back(std::move(b));
      }
    }
    return out;
  }

  // Runs the streaming chain that ends at `node`, handing each output batch to sink(thread, batch).
  // drivers: number of driver threads of THIS chain (blocking children below it are materialised
  // with the executor's full thread count).
  void stream(const NodePtr& node, const std::function<void(int, Batch&&)>& sink, int drivers = 0) {
    const int threads = drivers > 0 ? drivers : this->threads;
    std::vector<const Node*> ops;
    NodePtr cur = node;
    while (cur->kind == Node::FILTER || cur->kind == Node::PROJECT || cur->kind == Node::JOIN) {
      ops.push_back(cur.get());
      cur = cur->child;
    }
    std::reverse(ops.begin(), ops.end());
    std::unordered_map<const Node*, std::shared_ptr<JoinTable>> tables;
    for (auto* op : ops)
      if (op->kind == Node::JOIN) {
        auto jt = std::make_shared<JoinTable>();
        jt->build(*op, materialise(op->build), this->threads);
        tables[op] = jt;
      }
    auto run_ops = [&](int t, Batch b) {
      for (auto* op : ops) {
        if (b.n == 0) return;
        switch (op->kind) {
          case Node::FILTER: b = apply_filter(*op, b); break;
          case Node::PROJECT: b = apply_project(*op, b); break;
          default: b = probe_join(*op, *tables.at(op), b);
        }
      }
      if (b.n > 0) sink(t, std::move(b));
    };
    std::vector<std::string> errors(threads);
    std::vector<int> user_err(threads, 0);
    auto guarded = [&](int t, const std::function<void()>& f) {
      try { f(); } catch (const UserError& e) { errors[t] = e.what(); user_err[t] = 1; } catch (const std::exception& e) { errors[t] = e.what(); }
    };
    if (cur->kind == Node::VALUES) {
      if (cur->source < 0 || cur->source >= nsources) throw std::runtime_error("values: no such source");
      const orc_table& src = sources[cur->source];
      if (static_cast<size_t>(src.ncols) != cur->schema.size()) throw std::runtime_error("values: column count mismatch");
      for (int c = 0; c < src.ncols; ++c)
        if (src.cols[c].type != cur->schema[c]) throw std::runtime_error("values: column type mismatch");
      int64_t nb = (src.rows + batch_rows - 1) / batch_rows;
      auto work = [&](int t) {
        guarded(t, [&] {
          std::unordered_map<const void*, VecPtr> base_cache;
          int64_t b0 = nb * t / threads, b1 = nb * (t + 1) / threads;
          for (int64_t bi = b0; bi < b1; ++bi) {
            int64_t r0 = bi * batch_rows, n = std::min<int64_t>(batch_rows, src.rows - r0);
            Batch b;
            b.n = n;
            for (int c = 0; c < src.ncols; ++c) b.cols.push_back(slice_column(src.cols[c], r0, n, base_cache));
            run_ops(t, std::move(b));
          }
        });
      };
      run_threads(work, threads);
    } else {  // blocking child (aggregation): its output batches are distributed over drivers
      Table in = materialise(cur);
      auto work = [&](int t) {
        guarded(t, [&] {
          for (size_t i = t; i < in.size(); i += threads) run_ops(t, in[i]);
        });
      };
      run_threads(work, threads);
    }
    for (int t = 0; t < threads; ++t)
      if (!errors[t].empty()) {
        if (user_err[t]) throw UserError(errors[t]);
        throw std::runtime_error(errors[t]);
      }
  }

  void run_threads(const std::function<void(int)>& work, int threads) {
    if (threads == 1) { work(0); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t) ts.emplace_back(work, t);
    for (auto& t : ts) t.join();
  }
};

struct Result {
  std::vector<int> types;
  std::vector<VecPtr> cols;  // flat, concatenated
  int64_t rows = 0;
};

static Result* collect(const NodePtr& root, Table&& t) {
  auto res = new Result();
  res->types = root->schema;
  std::vector<ColBuilder> cb;
  for (int ty : root->schema) cb.emplace_back(ty);
  for (auto& b : t) {
    for (size_t c = 0; c < root->schema.size(); ++c) {
      VecPtr f = flatten(b.cols[c]);
      for (int64_t r = 0; r < b.n; ++r) cb[c].push_from(*f, r);
    }
    res->rows += b.n;
  }
  for (auto& c : cb) res->cols.push_back(c.finish());
  return res;
}

// VectorHasher::hash over columns (exec/VectorHasher.cpp:87-126,567-594).
static void hash_columns(const orc_column* cols, int ncols, int64_t rows, uint64_t* out) {
  std::unordered_map<const void*, VecPtr> cache;
  for (int c = 0; c < ncols; ++c) {
    VecPtr v = flatten(slice_column(cols[c], 0, rows, cache));
    bool mix = c > 0;
    for (int64_t r = 0; r < rows; ++r) {
      uint64_t h;
      if (v->null_at(r)) h = kNullHash;
      else switch (v->type) {
        case ORC_BIGINT: h = hash_i64(v->as<int64_t>()[r]); break;
        case ORC_INTEGER: h = hash_i32(v->as<int32_t>()[r]); break;
        case ORC_BOOLEAN: h = hash_bool(v->as<uint8_t>()[r]); break;
        case ORC_DOUBLE: h = hash_f64(v->as<double>()[r]); break;
        default: h = hash_string(v->chars + v->off[r], v->off[r + 1] - v->off[r]);
      }
      out[r] = mix ? hash_mix(out[r], h) : h;
    }
  }
}

}  // namespace orc

using namespace orc;

static void set_err(char* err, int32_t errlen, const std::string& prefix, const char* what) {
  if (!err || errlen <= 0) return;
  std::string m = prefix + what;
  std::strncpy(err, m.c_str(), errlen - 1);
  err[errlen - 1] = 0;
}

extern "C" {

void* orc_run_plan(const char* plan, int32_t n_sources, const orc_table* sources, int32_t threads,
                   int32_t batch_rows, char* err, int32_t errlen) {
  try {
    std::string text(plan);
    SNode s = SParser(text).parse();
    NodePtr root = parse_plan(s);
    Executor ex{sources, n_sources, std::max(1, threads), std::max<int64_t>(1, batch_rows)};
    return collect(root, ex.materialise(root));
  } catch (const UserError& e) {
    set_err(err, errlen, "VeloxUserError: ", e.what());
  } catch (const std::exception& e) {
    set_err(err, errlen, "VeloxRuntimeError: ", e.what());
  }
  return nullptr;
}

int64_t orc_result_rows(void* r) { return static_cast<Result*>(r)->rows; }
int32_t orc_result_cols(void* r) { return static_cast<int32_t>(static_cast<Result*>(r)->cols.size()); }
int32_t orc_result_type(void* r, int32_t col) { return static_cast<Result*>(r)->types[col]; }
void orc_result_copy(void* r, int32_t col, void* values, uint8_t* nulls) {
  auto* res = static_cast<Result*>(r);
  const Vec& v = *res->cols[col];
  if (values && v.data) std::memcpy(values, v.data, res->rows * width_of(v.type));
  if (nulls) {
    if (v.nulls) std::memcpy(nulls, v.nulls, res->rows);
    else std::memset(nulls, 0, res->rows);
  }
}
int64_t orc_result_str_bytes(void* r, int32_t col) {
  auto* res = static_cast<Result*>(r);
  const Vec& v = *res->cols[col];
  return v.off ? v.off[res->rows] : 0;
}
void orc_result_copy_str(void* r, int32_t col, int32_t* offsets, char* chars, uint8_t* nulls) {
  auto* res = static_cast<Result*>(r);
  const Vec& v = *res->cols[col];
  std::memcpy(offsets, v.off, (res->rows + 1) * 4);
  std::memcpy(chars, v.chars, v.off[res->rows]);
  if (nulls) {
    if (v.nulls) std::memcpy(nulls, v.nulls, res->rows);
    else std::memset(nulls, 0, res->rows);
  }
}
void orc_result_free(void* r) { delete static_cast<Result*>(r); }

int32_t orc_hash_columns(const orc_column* cols, int32_t ncols, int64_t rows, uint64_t* out) {
  try { hash_columns(cols, ncols, rows, out); return 0; } catch (...) { return 1; }
}
int32_t orc_partition(const orc_column* cols, int32_t ncols, int64_t rows, int32_t num_partitions, uint32_t* out) {
  try {
    std::vector<uint64_t> h(rows);
    hash_columns(cols, ncols, rows, h.data());
    for (int64_t r = 0; r < rows; ++r) out[r] = static_cast<uint32_t>(h[r] % static_cast<uint64_t>(num_partitions));
    return 0;
  } catch (...) { return 1; }
}

uint64_t orc_twang_mix64(uint64_t v) { return twang_mix64(v); }
uint32_t orc_jenkins_rev_mix32(uint32_t v) { return jenkins_rev_mix32(v); }
uint64_t orc_hash_mix(uint64_t upper, uint64_t lower) { return hash_mix(upper, lower); }
uint64_t orc_hash_bytes(uint64_t seed, const char* data, int64_t size) { return hash_bytes(seed, data, size); }
uint64_t orc_hash_f64(double v) { return hash_f64(v); }
int32_t orc_compare_f64(int32_t op, double a, double b) { return cmp_f64(op, a, b); }
int32_t orc_checked_i64(int32_t op, int64_t a, int64_t b, int64_t* out) { return checked_arith<int64_t>(op, a, b, out) ? 0 : 1; }

}  // extern "C"
