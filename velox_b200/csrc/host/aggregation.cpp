// B200HashAggregation: GROUP BY on the device (array mode / normalized-key hash mode), with the
// fused scan->filter->[probe]->project->aggregate fast path for null-free flat batches.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <map>
#include <set>

#include "join.h"
#include "operators.h"
#include "plan_resolve.h"

namespace velox_b200 {

namespace {

using Step = core::AggregationNode::Step;

constexpr uint64_t kArrayModeMax = 1ull << 22;  // slots; accumulators are SoA arrays of this length
constexpr int kFusedMaxGroups = 48;

uint64_t nextPow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

enum class Mode { kGlobal, kArray, kHash, kKeyed };  // kKeyed: the reference's kHash mode (rows store their key columns)

struct KeyState {
  TypeKind kind;
  bool isVarchar = false;
  bool isDouble = false;      // DOUBLE keys live in keyed tables only (no value-id range)
  // VARCHAR keys arrive dictionary-encoded; every distinct value gets a global id (1-based)
  std::unordered_map<std::string, int32_t> valueIds;
  std::vector<std::string> valuesById;
  // cache: alphabet -> device LUT (dictionary entry -> global id, 0 = NULL entry)
  std::shared_ptr<const HostAlphabet> cachedAlphabet;  // held, so its address cannot be recycled while cached
  DeviceBufferPtr lut;
  DeviceBufferPtr fusedLut;   // dictionary entry -> id under the current layout (fused kernel)
  bool hasRange = false;
  int64_t lo = 0, hi = 0;
  bool nullable = false;      // a NULL key was (or may have been) seen: the layout reserves id 0
  // device copy of the global alphabet for output columns (dictionary base), refreshed when it grew
  DeviceBufferPtr devOffsets, devChars;
  size_t devAlphabetCount = 0;
  std::vector<int32_t> hostOffsets;
  std::string hostChars;
  std::shared_ptr<const HostAlphabet> outAlphabet;
};

struct AccState {
  std::string fn;
  int32_t kind = 0;          // vb2_agg_kind of the main accumulator
  bool isDouble = false;     // accumulator holds doubles
  TypePtr inputType;
  int32_t accWord = -1;      // word of the accumulator inside the group row
  int32_t nnWord = -1;       // word of the non-null counter (-1: count(), which is its own counter)
  // The non-null counter only decides NULL-ness for sum / min / max; it is not maintained while
  // every input seen so far was non-null and unmasked (all groups trivially non-null). AVG always
  // tracks (the counter is its divisor).
  bool nnTracked = false;
};

}  // namespace

void registerB200Aggregate(const std::string& name, const std::string& family, const std::string& inputFunction, const std::string& finalFunction) {
  static const std::set<std::string> kFamilies = {"sum", "avg", "count", "min", "max"};
  VELOX_CHECK(kFamilies.count(family) == 1, "aggregate family must be one of sum avg count min max, not '" + family + "'");
  auto sig = std::make_shared<exec::AggregateFunctionSignature>();
  sig->argTypes = {"T"};
  sig->returnType = family == "count" ? "bigint" : (family == "avg" ? "double" : "T");
  sig->intermediateType = family == "avg" ? "row(double,bigint)" : sig->returnType;
  exec::registerAggregateFunction(
      name, {sig},
      [family, inputFunction, finalFunction](core::AggregationNode::Step, const std::vector<TypePtr>&, const TypePtr& resultType, const core::QueryConfig&) {
        return std::make_unique<B200Aggregate>(resultType, family, inputFunction, finalFunction);
      },
      /*registerCompanionFunctions=*/false, /*overwrite=*/true);
}

void registerB200Aggregates() {
  static std::once_flag once;
  std::call_once(once, [] {
    // prestosql/aggregates/{Sum,Average,Count,MinMax}Aggregate.cpp register these names on the CPU
    for (const char* fn : {"sum", "avg", "count", "min", "max"}) registerB200Aggregate(fn, fn, "", "");
  });
}

TypePtr scalarFunctionReturnType(const std::string& name, const TypePtr& argType) {
  registerB200Functions();
  auto fn = exec::getVectorFunction(name);
  VELOX_CHECK(fn != nullptr, "scalar function '" + name + "' is not registered");
  if (auto* user = dynamic_cast<const B200DeviceFunction*>(fn.get())) return user->returnType();
  static const std::set<std::string> kBool = {"lt", "lte", "gt", "gte", "eq", "neq", "between", "like", "not", "is_null"};
  return kBool.count(name) ? BOOLEAN() : argType;
}

struct B200HashAggregation::Impl {
  B200HashAggregation* self;
  std::shared_ptr<const core::AggregationNode> node;
  ResolvedAggregation resolved;  // the node's column names resolved to input channels (HashAggregation::initialize does the same)
  std::vector<std::unique_ptr<exec::Operator>> absorbed;
  std::shared_ptr<DeviceContext> dev;
  bool raw, fin;
  Mode mode = Mode::kGlobal;
  std::vector<KeyState> keys;
  KeyLayout layout;             // mins are adjusted: id = v - mins[k] + 1 (mins[k] = lo + 1 - nullReserved[k])
  std::vector<int64_t> covLo, covHi;  // value interval each key's layout covers
  std::vector<int32_t> nullReserved;  // 1 = id 0 means NULL for that key
  int64_t capacity = 1;
  // Group rows [capacity][rowWords]: word 0 = normalized key (hash mode) / rows seen (array, global),
  // then the accumulator and non-null-counter words of every aggregate (vb2_group_table).
  DeviceBufferPtr rowsBuf;
  int32_t rowWords = 1;        // words of an array / normalized-key row; keyed rows add their key words (rowWordsFor)
  int32_t baseWords = 1;       // 1 + accumulator / counter words, before padding
  bool forceKeyed = false;
  bool keyedReady = false;     // the keyed table has its hash-mode storage (the initial 1-row storage does not count)     // a key type without a value-id range (DOUBLE): keyed table from the start
  std::vector<uint64_t> rowInit;
  bool sawGeneric = false;     // a batch went through vb2k_group_update (untracked counters need a fix-up to start tracking)
  DeviceBufferPtr numGroupsDev;
  int64_t numGroupsUpper = 0;  // upper bound of distinct groups seen (hash mode sizing)
  std::vector<AccState> accs;
  // registered aggregates with an input / final transform (B200Aggregate): evaluated by the expression engine
  std::unique_ptr<CompiledProgram> inputProgram, finalProgram;
  RowTypePtr inputProgramType, extractType;
  DeviceBufferPtr errorFlag;
  bool sawInput = false;
  bool outputDone = false;

  // ---- fused fast path --------------------------------------------------------------------------
  bool fusedPlanned = false;      // plan shape allows it
  int fusedId = -1;
  FusedBinding binding;
  std::vector<int> aggToProj;     // aggregate -> fused projection (-1 for count(*))
  std::vector<int> fusedKeySourceCols;  // source-batch column of each group key
  int fusedJoinKeySourceCol = -1;
  core::TypedExprPtr joinFlagExpr;     // over build columns (field index = build column)
  B200HashProbe* fusedProbe = nullptr;
  DeviceBufferPtr joinSlotFlags;
  DeviceBufferPtr fusedSums, fusedCounts, fusedWs;
  size_t fusedWsBytes = 0;
  int fusedGroups = 0;
  int64_t fusedBatches = 0, genericBatches = 0, selectiveBatches = 0, partitionedBatches = 0;
  DeviceBufferPtr partStartDev, barrierWord;  // set while a radix-partitioned batch is being applied
  // ---- shared-memory slice aggregation (slice_agg.cu): high-cardinality input is buffered and aggregated at the end
  enum class SliceState { kUndecided, kBuffering, kOff };
  SliceState sliceState = SliceState::kUndecided;
  std::vector<B200VectorPtr> sliceChunks;  // buffered input batches (zero copy: the kernels read their columns in place)
  std::vector<int32_t> slicePayload;       // input columns the aggregates read
  int64_t sliceRows = 0, sliceBatches = 0;
  bool selectiveDecided = false, selectiveUsable = true;
  double selectivity = 1.0;
  struct FusedTiming { cudaEvent_t begin = nullptr, end = nullptr; int64_t rows = 0; };
  std::vector<FusedTiming> fusedTimings;
  ~Impl() {
    for (auto& t : fusedTimings) { cudaEventDestroy(t.begin); cudaEventDestroy(t.end); }
  }
  // Device time of the fused launches so far (call after the stream was synchronised).
  void reportFusedTimings() {
    double ns = 0;
    int64_t rows = 0;
    for (auto& t : fusedTimings) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, t.begin, t.end) == cudaSuccess) { ns += static_cast<double>(ms) * 1e6; rows += t.rows; }
      cudaEventDestroy(t.begin);
      cudaEventDestroy(t.end);
    }
    if (!fusedTimings.empty()) {
      self->addRuntimeStat("b200.fusedScanNanos", exec::RuntimeCounter{static_cast<int64_t>(ns)});
      self->addRuntimeStat("b200.fusedScanRows", exec::RuntimeCounter{rows});
    }
    fusedTimings.clear();
  }

  cudaStream_t st() const { return dev->stream; }

  // =================================================================================================
  void init() {
    raw = node->isRawInput();
    fin = node->isFinalOutput();
    const auto& inType = node->sources()[0]->outputType();
    resolved = resolveAggregation(*node);
    resolveRegisteredAggregates(inType);
    for (int32_t k : resolved.keys) {
      KeyState ks;
      ks.kind = inType->childAt(k)->kind();
      ks.isVarchar = ks.kind == TypeKind::VARCHAR;
      ks.isDouble = ks.kind == TypeKind::DOUBLE;
      if (ks.isDouble) forceKeyed = true;
      keys.push_back(std::move(ks));
    }
    VELOX_CHECK(keys.size() <= 4, "at most 4 grouping keys");
    mode = keys.empty() ? Mode::kGlobal : (forceKeyed ? Mode::kKeyed : Mode::kArray);
    for (auto& a : resolved.aggregates) {
      AccState s;
      s.fn = a.function;
      s.inputType = a.rawInputType;
      const bool dbl = a.rawInputType && a.rawInputType->kind() == TypeKind::DOUBLE;
      if (a.function == "sum") { s.isDouble = dbl; s.kind = dbl ? VB2_AGG_SUM_F64 : VB2_AGG_SUM_I64; }
      else if (a.function == "avg") { s.isDouble = true; s.kind = VB2_AGG_SUM_F64; }
      else if (a.function == "count") { s.isDouble = false; s.kind = raw ? VB2_AGG_COUNT : VB2_AGG_COUNT_MERGE; }
      else if (a.function == "min") { s.isDouble = dbl; s.kind = dbl ? VB2_AGG_MIN_F64 : VB2_AGG_MIN_I64; }
      else if (a.function == "max") { s.isDouble = dbl; s.kind = dbl ? VB2_AGG_MAX_F64 : VB2_AGG_MAX_I64; }
      else VELOX_UNSUPPORTED("aggregate function " + a.function);
      if (a.rawInputType && a.rawInputType->kind() == TypeKind::VARCHAR) VELOX_UNSUPPORTED("aggregates over VARCHAR");
      accs.push_back(std::move(s));
    }
    int32_t w = 1;
    for (auto& a : accs) {
      a.accWord = w++;
      if (a.fn != "count") a.nnWord = w++;
      // a global aggregation emits its row even when no input arrived, so its counters always run
      a.nnTracked = a.fn == "avg" || keys.empty();
    }
    baseWords = w;
    rowWords = w <= 2 ? w : (w + 3) / 4 * 4;  // whole 32-byte sectors
    VELOX_CHECK(rowWordsFor(Mode::kKeyed) <= VB2_MAX_ROW_WORDS, "too many aggregates for one group row");
    errorFlag = allocDeviceZeroed(8, st());
    numGroupsDev = allocDeviceZeroed(8, st());
    layout.mins.assign(keys.size(), 1);
    layout.ranges.assign(keys.size(), 0);
    layout.mults.assign(keys.size(), 1);
    layout.product = 1;
    covLo.assign(keys.size(), 0);
    covHi.assign(keys.size(), -1);
    nullReserved.assign(keys.size(), 0);
    capacity = 1;
    allocateStorage(capacity, mode);
    planFused();
  }

  // Every aggregate name goes through the registry (Aggregate::create, exec/Aggregate.h:361): the
  // B200Aggregate it yields names the device accumulator family and the optional transforms.
  void resolveRegisteredAggregates(const RowTypePtr& inType) {
    registerB200Aggregates();
    const auto& cfg = self->driverCtx()->queryConfig();
    std::vector<core::TypedExprPtr> inExprs, finExprs;
    std::vector<std::string> inNames = inType->names();
    std::vector<TypePtr> inTypes = inType->children();
    for (uint32_t i = 0; i < inType->size(); ++i) inExprs.push_back(std::make_shared<core::FieldAccessTypedExpr>(inType->childAt(i), inType->nameOf(i)));
    const RowTypePtr& outType = node->outputType();
    const size_t nk = resolved.keys.size();
    for (size_t k = 0; k < nk; ++k) finExprs.push_back(std::make_shared<core::FieldAccessTypedExpr>(outType->childAt(k), outType->nameOf(k)));
    bool anyInput = false, anyFinal = false;
    std::vector<TypePtr> preTypes(outType->children().begin(), outType->children().begin() + nk);  // output of the extraction, before final transforms
    uint32_t outCol = static_cast<uint32_t>(nk);
    for (size_t i = 0; i < resolved.aggregates.size(); ++i) {
      auto& a = resolved.aggregates[i];
      const auto& call = node->aggregates()[i].call;
      std::vector<TypePtr> argTypes;
      for (auto& in : call->inputs()) argTypes.push_back(in->type());
      auto created = exec::Aggregate::create(a.function, node->step(), argTypes, call->type(), cfg);
      auto* agg = dynamic_cast<B200Aggregate*>(created.get());
      if (!agg) VELOX_UNSUPPORTED("aggregate function " + a.function + " has no B200 implementation (register a B200Aggregate)");
      a.function = agg->family();
      if (!agg->inputFunction().empty() && raw && !a.inputs.empty()) {
        // aggregate the transformed column: it is appended to the input batch by one projection kernel
        const int32_t src = a.inputs[0];
        const TypePtr t = scalarFunctionReturnType(agg->inputFunction(), inType->childAt(src));
        inExprs.push_back(std::make_shared<core::CallTypedExpr>(t, std::vector<core::TypedExprPtr>{inExprs[src]}, agg->inputFunction()));
        a.inputs[0] = static_cast<int32_t>(inTypes.size());
        inNames.push_back("\x01" "b200.agg_in#" + std::to_string(i));
        inTypes.push_back(t);
        a.rawInputType = t;
        anyInput = true;
      } else if (!agg->inputFunction().empty() && a.rawInputType) {
        a.rawInputType = scalarFunctionReturnType(agg->inputFunction(), a.rawInputType);  // later steps: the type the accumulator holds
      }
      const bool twoCols = a.function == "avg" && !fin;  // intermediate avg: (sum, count)
      TypePtr extracted = outType->childAt(outCol);
      if (fin && !agg->finalFunction().empty()) {
        // type the family itself produces, before the final transform
        if (a.function == "avg") extracted = DOUBLE();
        else if (a.function == "count") extracted = BIGINT();
        else if (a.function == "sum") extracted = (a.rawInputType && a.rawInputType->kind() == TypeKind::DOUBLE) ? DOUBLE() : BIGINT();
        else extracted = a.rawInputType;
        auto f = std::make_shared<core::FieldAccessTypedExpr>(extracted, outType->nameOf(outCol));
        const TypePtr t = scalarFunctionReturnType(agg->finalFunction(), extracted);
        VELOX_CHECK(t->kind() == outType->childAt(outCol)->kind(), "aggregate " + call->name() + ": the plan's result type differs from its final function's");
        finExprs.push_back(std::make_shared<core::CallTypedExpr>(t, std::vector<core::TypedExprPtr>{f}, agg->finalFunction()));
        anyFinal = true;
      } else {
        finExprs.push_back(std::make_shared<core::FieldAccessTypedExpr>(outType->childAt(outCol), outType->nameOf(outCol)));
        if (twoCols) finExprs.push_back(std::make_shared<core::FieldAccessTypedExpr>(outType->childAt(outCol + 1), outType->nameOf(outCol + 1)));
      }
      preTypes.push_back(extracted);
      if (twoCols) preTypes.push_back(outType->childAt(outCol + 1));
      outCol += twoCols ? 2 : 1;
    }
    if (anyInput) {
      inputProgramType = ROW(inNames, inTypes);
      inputProgram = std::make_unique<CompiledProgram>(compileExprs(inExprs, false, inType));
      inputProgram->uploadConstants(st());
    }
    if (anyFinal) {
      extractType = ROW(outType->names(), preTypes);
      finalProgram = std::make_unique<CompiledProgram>(compileExprs(finExprs, false, extractType));
      finalProgram->uploadConstants(st());
    }
  }

  uint64_t identityBits(const AccState& s) const {
    switch (s.kind) {
      case VB2_AGG_MIN_F64: return 0x7ff8000000000000ull;                      // NaN: the largest value
      case VB2_AGG_MAX_F64: return 0xfff0000000000000ull;                      // -inf
      case VB2_AGG_MIN_I64: return static_cast<uint64_t>(INT64_MAX);
      case VB2_AGG_MAX_I64: return static_cast<uint64_t>(INT64_MIN);
      default: return 0;
    }
  }

  // Keyed rows: [state | key words | NULL mask | accumulators ...]: accumulator word w of the other
  // layouts sits at w + shiftFor(kKeyed).
  int32_t shiftFor(Mode m) const { return m == Mode::kKeyed ? static_cast<int32_t>(keys.size()) + 1 : 0; }
  int32_t rowWordsFor(Mode m) const {
    if (m != Mode::kKeyed) return rowWords;
    const int32_t w = baseWords + shiftFor(m);
    return (w + 3) / 4 * 4;
  }
  int32_t shift() const { return shiftFor(mode); }
  vb2_group_table tableOf(const DeviceBufferPtr& buf, int64_t cap, Mode m) const {
    vb2_group_table t{};
    t.rows = buf->as<uint64_t>();
    t.capacity = cap;
    t.row_words = rowWordsFor(m);
    t.hash_mode = m == Mode::kHash ? 1 : (m == Mode::kKeyed ? VB2_GROUP_KEYED : 0);
    return t;
  }
  vb2_group_table table() const { return tableOf(rowsBuf, capacity, mode); }

  DeviceBufferPtr makeStorage(int64_t cap, Mode m) {
    const int32_t rw = rowWordsFor(m), sh = shiftFor(m);
    std::vector<uint64_t> init(rw, 0);
    init[0] = (m == Mode::kHash || m == Mode::kKeyed) ? VB2_EMPTY_KEY : 0;
    for (auto& a : accs) init[a.accWord + sh] = identityBits(a);
    auto buf = allocDevice(static_cast<size_t>(cap) * rw * 8, st());
    const vb2_group_table t = tableOf(buf, cap, m);
    kernelCheck(vb2k_group_table_init(&t, init.data(), st()));
    return buf;
  }
  void allocateStorage(int64_t cap, Mode m) { rowsBuf = makeStorage(cap, m); }

  // Occupied slots of the current table (synchronises for the count).
  DeviceBufferPtr occupiedSlots(int64_t& count) {
    auto slots = allocDevice(static_cast<size_t>(capacity) * 4, st());
    auto cnt = allocDevice(8, st());
    const size_t wsb = vb2k_group_occupied_workspace(capacity);
    auto ws = allocDevice(wsb, st());
    const vb2_group_table t = table();
    kernelCheck(vb2k_group_occupied(&t, slots->as<int32_t>(), cnt->as<int64_t>(), ws->data(), wsb, st()));
    VB2_CU(cudaMemcpyAsync(&count, cnt->data(), 8, cudaMemcpyDeviceToHost, st()));
    VB2_CU(cudaStreamSynchronize(st()));
    return slots;
  }

  // Starts maintaining aggregate i's non-null counter: groups that exist already have only seen
  // non-null, unmasked inputs, so their counters become 1 ("some non-null input").
  void trackNonNull(size_t i) {
    AccState& s = accs[i];
    if (s.nnTracked || s.nnWord < 0) return;
    if (sawGeneric) {
      const vb2_group_table t = table();
      kernelCheck(vb2k_group_set_word(&t, s.nnWord + shift(), 1, st()));
    }
    s.nnTracked = true;
  }

  // Moves every group to a new layout / mode / capacity (ranges grew or the table filled up).
  void relayout(const KeyLayout& nl, const std::vector<int32_t>& newNullReserved, Mode nm, int64_t ncap) {
    flushFused();  // fused partial sums are laid out by the old group ids
    int64_t m = 0;
    DeviceBufferPtr slots = sawInput ? occupiedSlots(m) : nullptr;
    DeviceBufferPtr ns = makeStorage(ncap, nm);
    if (m > 0 && nm == Mode::kKeyed) {
      // into a keyed table: from a smaller keyed table (keys are in the rows) or from an array /
      // normalized-key table (keys decoded from the value ids of the old layout)
      const vb2_group_table from = table(), to = tableOf(ns, ncap, nm);
      const int32_t nk = static_cast<int32_t>(keys.size());
      VB2_CU(cudaMemsetAsync(numGroupsDev->data(), 0, 8, st()));
      if (mode == Mode::kKeyed)
        kernelCheck(vb2k_group_move_keyed(&from, slots->as<int32_t>(), m, nk, &to, numGroupsDev->as<int64_t>(), errorFlag->as<int32_t>(), st()));
      else
        kernelCheck(vb2k_group_move_to_keyed(&from, slots->as<int32_t>(), m, nk, layout.mins.data(), layout.mults.data(), layout.ranges.data(),
                                             nullReserved.data(), shiftFor(Mode::kKeyed), &to, numGroupsDev->as<int64_t>(), errorFlag->as<int32_t>(), st()));
      checkDeviceError(errorFlag, st(), "aggregation rehash");
    } else if (m > 0) {
      const vb2_group_table from = table(), to = tableOf(ns, ncap, nm);
      auto newKeys = allocDevice(static_cast<size_t>(m) * 8, st());
      kernelCheck(vb2k_group_rekey(&from, slots->as<int32_t>(), m, static_cast<int32_t>(keys.size()), layout.mins.data(), layout.mults.data(),
                                   layout.ranges.data(), nullReserved.data(), nl.mins.data(), nl.mults.data(), newKeys->as<uint64_t>(), st()));
      VB2_CU(cudaMemsetAsync(numGroupsDev->data(), 0, 8, st()));  // the new table counts its groups afresh
      kernelCheck(vb2k_group_move(&from, slots->as<int32_t>(), newKeys->as<uint64_t>(), m, &to, numGroupsDev->as<int64_t>(),
                                  errorFlag->as<int32_t>(), st()));
      checkDeviceError(errorFlag, st(), "aggregation rehash");
    }
    layout = nl;
    nullReserved = newNullReserved;
    for (auto& ks : keys) ks.fusedLut = nullptr;
    mode = nm;
    capacity = ncap;
    rowsBuf = std::move(ns);
    self->addRuntimeStat("b200.aggRelayouts", exec::RuntimeCounter{1});
  }

  // ---- key handling -----------------------------------------------------------------------------
  // Returns the column the normalize kernel should read for key k of this batch and updates the
  // key's observed range. VARCHAR dictionary keys become INTEGER dictionary columns over a LUT of
  // global value ids.
  vb2_column keyColumn(size_t k, const DeviceColumn& col, int64_t rows, std::vector<DeviceBufferPtr>& keep) {
    KeyState& ks = keys[k];
    vb2_column d = col.desc;
    if (col.mayHaveNulls()) ks.nullable = true;
    if (ks.isVarchar) {
      if (d.encoding == VB2_FLAT) VELOX_UNSUPPORTED("GROUP BY on flat VARCHAR keys with more than 65536 distinct values per batch (flat strings are dictionary-encoded on upload up to that size)");
      VELOX_CHECK(col.alphabet != nullptr, "VARCHAR key without a host alphabet (dictionary above 65536 entries)");
      if (ks.cachedAlphabet != col.alphabet) {
        std::vector<int32_t> lut(col.alphabet->values.size());
        for (size_t i = 0; i < lut.size(); ++i) {
          if (col.alphabet->nulls[i]) { lut[i] = 0; ks.nullable = true; continue; }
          auto it = ks.valueIds.find(col.alphabet->values[i]);
          if (it == ks.valueIds.end()) {
            ks.valuesById.push_back(col.alphabet->values[i]);
            it = ks.valueIds.emplace(col.alphabet->values[i], static_cast<int32_t>(ks.valuesById.size())).first;
          }
          lut[i] = it->second;
        }
        ks.lut = allocDevice(lut.size() * 4 + 4, st());
        // pageable source: the call returns once the bytes are staged, the vector may go away
        VB2_CU(cudaMemcpyAsync(ks.lut->data(), lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, st()));
        ks.cachedAlphabet = col.alphabet;
        ks.fusedLut = nullptr;
      }
      keep.push_back(ks.lut);
      d.type = VB2_INTEGER;
      d.values = ks.lut->data();
      d.aux = nullptr;
      if (d.encoding == VB2_CONSTANT) d.dict_size = 1;
      ks.hasRange = true;
      ks.lo = 1;
      ks.hi = std::max<int64_t>(1, static_cast<int64_t>(ks.valuesById.size()));
      return d;
    }
    if (ks.isDouble || mode == Mode::kKeyed) return d;  // keyed tables read the key columns as they are: no value-id range needed
    int64_t lo, hi, nn;
    columnMinMax(col, rows, st(), lo, hi, nn);
    if (nn > 0) {
      if (!ks.hasRange) { ks.lo = lo; ks.hi = hi; ks.hasRange = true; }
      else { ks.lo = std::min(ks.lo, lo); ks.hi = std::max(ks.hi, hi); }
    }
    return d;
  }

  // Makes sure the layout covers the observed key ranges and the table can take `incoming` more
  // groups; relayouts otherwise.
  void ensureLayout(int64_t incoming) {
    if (keys.empty()) return;
    if (mode == Mode::kKeyed) {
      // keyed tables only ever grow (sized by groups so far + one pass, as the normalized-key tables)
      const uint64_t want = nextPow2(static_cast<uint64_t>(numGroupsUpper + incoming) * 2 + 16);
      if (!rowsBuf || static_cast<uint64_t>(capacity) < want || !keyedReady) {
        VELOX_CHECK(want <= (1ull << 31), "hash aggregation above 2^31 slots");
        relayout(layout, nullReserved, Mode::kKeyed, static_cast<int64_t>(std::max<uint64_t>(want, static_cast<uint64_t>(capacity))));
        keyedReady = true;
      }
      return;
    }
    bool covers = true;
    for (size_t k = 0; k < keys.size(); ++k) {
      const KeyState& ks = keys[k];
      if (ks.nullable && !nullReserved[k]) covers = false;
      if (ks.hasRange && (covHi[k] < covLo[k] || ks.lo < covLo[k] || ks.hi > covHi[k])) covers = false;
    }
    KeyLayout nl = layout;
    std::vector<int64_t> nLo = covLo, nHi = covHi;
    std::vector<int32_t> nRes = nullReserved;
    if (!covers) {
      // exact ranges first; if that already needs hash mode, widen integer ranges (free there)
      auto build = [&](bool widen, unsigned __int128& product) {
        product = 1;
        bool overflow = false;
        for (size_t k = 0; k < keys.size(); ++k) {
          const KeyState& ks = keys[k];
          __int128 lo = ks.hasRange ? ks.lo : 0, hi = ks.hasRange ? ks.hi : 0;
          if (covHi[k] >= covLo[k]) { lo = std::min<__int128>(lo, covLo[k]); hi = std::max<__int128>(hi, covHi[k]); }
          if (widen && !ks.isVarchar) {
            const __int128 span = hi - lo + 1;
            lo = std::max<__int128>(lo - span, INT64_MIN / 2);
            hi = std::min<__int128>(hi + span, INT64_MAX / 2);
          }
          nRes[k] = (ks.nullable || nullReserved[k]) ? 1 : 0;
          const unsigned __int128 range = static_cast<unsigned __int128>(hi - lo) + 1 + nRes[k];
          if (range > (static_cast<unsigned __int128>(1) << 62)) overflow = true;
          nLo[k] = static_cast<int64_t>(lo);
          nHi[k] = static_cast<int64_t>(hi);
          nl.mins[k] = static_cast<int64_t>(lo) + 1 - nRes[k];
          nl.ranges[k] = static_cast<uint64_t>(range);
          product *= range;
          if (product > (static_cast<unsigned __int128>(1) << 62)) overflow = true;
        }
        return !overflow;
      };
      unsigned __int128 product;
      bool ok = build(false, product);
      if (ok && product > kArrayModeMax) {
        unsigned __int128 wide;
        if (build(true, wide)) product = wide;
        else ok = build(false, product);
      }
      if (!ok) {
        // the packed value ids need more than one 64-bit word: keyed table (the reference's kHash mode,
        // exec/HashTable.cpp:1751-1838); the groups collected so far are decoded and moved over
        const uint64_t want = nextPow2(static_cast<uint64_t>(numGroupsUpper + incoming) * 2 + 16);
        VELOX_CHECK(want <= (1ull << 31), "hash aggregation above 2^31 slots");
        relayout(layout, nullReserved, Mode::kKeyed, static_cast<int64_t>(want));
        keyedReady = true;
        return;
      }
      nl.product = static_cast<uint64_t>(product);
      nl.mults.assign(keys.size(), 1);
      for (int i = static_cast<int>(keys.size()) - 2; i >= 0; --i) nl.mults[i] = nl.mults[i + 1] * nl.ranges[i + 1];
    }
    Mode nm = nl.product <= kArrayModeMax ? Mode::kArray : Mode::kHash;
    int64_t ncap = capacity;
    if (nm == Mode::kArray) {
      ncap = static_cast<int64_t>(nl.product);
    } else {
      const uint64_t bound = std::min<uint64_t>(nl.product, static_cast<uint64_t>(numGroupsUpper + incoming));
      const uint64_t want = nextPow2(bound * 2 + 16);
      if (mode != Mode::kHash || static_cast<uint64_t>(capacity) < want) ncap = static_cast<int64_t>(want);
      VELOX_CHECK(ncap <= (1ll << 31), "hash aggregation above 2^31 slots");
    }
    if (!covers || nm != mode || ncap != capacity) {
      relayout(nl, nRes, nm, ncap);
      covLo = nLo;
      covHi = nHi;
    }
  }

  // ---- generic path -------------------------------------------------------------------------------
  // Distinct-key estimate of a batch from the HyperLogLog sketch the radix histogram kernel fills
  // (registers over the keys whose hash ends in 000, hence the factor 8).
  static double hllEstimate(const std::vector<int32_t>& regs) {
    const double m = static_cast<double>(regs.size());
    double sum = 0;
    int zeros = 0;
    for (int32_t r : regs) {
      sum += std::ldexp(1.0, -r);
      zeros += r == 0;
    }
    const double alpha = 0.7213 / (1.0 + 1.079 / m);
    double e = alpha * m * m / sum;
    if (e <= 2.5 * m && zeros > 0) e = m * std::log(m / zeros);  // small-range correction (linear counting)
    return 8.0 * e;
  }

  // ---- shared-memory slice aggregation ---------------------------------------------------------------
  // A GROUP BY with millions of groups is bound by random DRAM sectors and page walks when every row
  // probes a global table. When the first large batch shows that many distinct keys, the operator
  // buffers its input instead (zero copy) and aggregates everything at the end with two partition
  // passes and per-slice shared-memory tables (slice_agg.cu). The reference makes the same kind of
  // call at run time when it abandons partial aggregation for nearly-unique keys
  // (exec/HashAggregation.cpp:185-189,320-336); results are identical either way.
  // Shape: one flat NULL-free BIGINT key, unmasked aggregates over flat NULL-free BIGINT / DOUBLE columns.
  bool sliceEligible(const B200VectorPtr& in, std::vector<int32_t>& payload) const {
    if (keys.size() != 1 || mode == Mode::kKeyed) return false;
    const vb2_column& k = in->column(resolved.keys[0])->desc;
    if (k.encoding != VB2_FLAT || k.nulls || k.type != VB2_BIGINT) return false;
    for (auto& a : resolved.aggregates) {
      if (a.mask >= 0) return false;
      for (int32_t c : a.inputs) {
        const vb2_column& d = in->column(c)->desc;
        if (d.encoding != VB2_FLAT || d.nulls || (d.type != VB2_BIGINT && d.type != VB2_DOUBLE)) return false;
        if (std::find(payload.begin(), payload.end(), c) == payload.end()) payload.push_back(c);
      }
    }
    return payload.size() <= VB2_SLICE_MAX_COLS;  // the accumulator-word limit is checked when the ops are listed
  }

  bool addSliceBuffered(const B200VectorPtr& in) {
    if (sliceState == SliceState::kOff) return false;
    const auto& cfg = self->driverCtx()->queryConfig();
    std::vector<int32_t> payload;
    const bool eligible = cfg.get<bool>("b200.agg_slice_aggregation", true) && sliceEligible(in, payload);
    if (sliceState == SliceState::kUndecided) {
      // decided once, on the first batch that reaches the generic path: the table must still be empty
      sliceState = SliceState::kOff;
      const int64_t n = in->size();
      if (!eligible || sawGeneric || n < cfg.get<int64_t>("b200.agg_slice_min_rows", 1 << 23)) return false;
      // distinct keys of this batch (HyperLogLog over the raw keys: the hash of v - min + 1 with any min sketches the same count)
      // over the first 64 M rows at most: a prefix with that many distinct keys settles the question, and the sketch pass
      // stays a fraction of a millisecond
      const vb2_column& k = in->column(resolved.keys[0])->desc;
      const int64_t probe = std::min<int64_t>(n, 64ll << 20);
      const size_t wsBytes = vb2k_radix_workspace_bytes(probe);
      auto ws = allocDevice(wsBytes, st());
      const int32_t nregs = vb2k_radix_hll_registers();
      auto hll = allocDevice(static_cast<size_t>(nregs) * 4, st());
      kernelCheck(vb2k_radix_histogram(nullptr, k.values, 1, 0, probe, ws->data(), wsBytes, hll->as<int32_t>(), st()));
      std::vector<int32_t> regs(nregs);
      VB2_CU(cudaMemcpyAsync(regs.data(), hll->data(), static_cast<size_t>(nregs) * 4, cudaMemcpyDeviceToHost, st()));
      VB2_CU(cudaStreamSynchronize(st()));
      // worth it once the table (32-byte rows at load 0.5) outgrows the L2 by a wide margin
      if (hllEstimate(regs) < static_cast<double>(cfg.get<int64_t>("b200.agg_slice_min_groups", 4 << 20))) return false;
      sliceState = SliceState::kBuffering;
      slicePayload = payload;
    } else if (!eligible || payload != slicePayload) {
      // a batch the slice kernels cannot read: everything buffered so far goes through the table path
      drainSliceBuffer();
      return false;
    }
    std::vector<DeviceBufferPtr> keep;
    keyColumn(0, *in->column(resolved.keys[0]), in->size(), keep);  // tracks the key range (one min/max pass)
    sliceChunks.push_back(in);
    sliceRows += in->size();
    ++sliceBatches;
    if (sliceRows >= (1ll << 32) - (1ll << 28)) drainSliceBuffer();  // the slice kernels index rows with 32 bits
    return true;
  }

  // Buffered batches through the table path (fallback; also ends slice mode).
  void drainSliceBuffer() {
    sliceState = SliceState::kOff;
    std::vector<B200VectorPtr> chunks = std::move(sliceChunks);
    sliceChunks.clear();
    sliceRows = 0;
    const int64_t chunk = std::max<int64_t>(1 << 16, self->driverCtx()->queryConfig().b200AggProbeChunkRows()) / 64 * 64;
    for (auto& b : chunks)
      for (int64_t off = 0; off < b->size(); off += chunk) addGeneric(sliceVector(b, off, std::min<int64_t>(chunk, b->size() - off)));
  }

  // Aggregates the buffered input; the compact group rows become the operator's table (never probed again).
  void finishSliceAggregation() {
    if (sliceState != SliceState::kBuffering || sliceChunks.empty()) return;
    ensureLayout(0);  // the layout covers every buffered key range
    if (mode != Mode::kHash || layout.mults[0] != 1 || nullReserved[0] != 0) { drainSliceBuffer(); return; }
    const int32_t ncols = static_cast<int32_t>(slicePayload.size());
    std::vector<vb2_slice_chunk> chunks;
    for (auto& b : sliceChunks) {
      vb2_slice_chunk c{};
      c.raw_keys = static_cast<const int64_t*>(b->column(resolved.keys[0])->desc.values);
      c.key_min = layout.mins[0];
      for (int32_t i = 0; i < ncols; ++i) c.cols[i] = b->column(slicePayload[i])->desc.values;
      c.rows = b->size();
      chunks.push_back(c);
    }
    auto colOf = [&](int32_t inputColumn) { return static_cast<int32_t>(std::find(slicePayload.begin(), slicePayload.end(), inputColumn) - slicePayload.begin()); };
    std::vector<vb2_slice_op> ops;
    for (size_t i = 0; i < accs.size(); ++i) {
      const auto& a = resolved.aggregates[i];
      const AccState& s = accs[i];
      if (a.function == "count") {
        if (raw) ops.push_back(vb2_slice_op{VB2_AGG_COUNT, -1, s.accWord});
        else ops.push_back(vb2_slice_op{VB2_AGG_COUNT_MERGE, colOf(a.inputs[0]), s.accWord});
        continue;
      }
      ops.push_back(vb2_slice_op{s.kind, colOf(a.inputs[0]), s.accWord});
      if (a.function == "avg" && !raw) ops.push_back(vb2_slice_op{VB2_AGG_COUNT_MERGE, colOf(a.inputs[1]), s.nnWord});
      else if (s.nnTracked) ops.push_back(vb2_slice_op{VB2_AGG_COUNT, -1, s.nnWord});  // inputs are NULL-free: every row counts
    }
    const auto& cfg = self->driverCtx()->queryConfig();
    const size_t wsBytes = vb2k_slice_agg_workspace(sliceRows, ncols);
    bool done = false;
    if (ops.size() <= VB2_SLICE_MAX_OPS) {
      auto ws = allocDevice(wsBytes, st());
      std::vector<int32_t> regs(vb2k_slice_agg_hll_registers());
      kernelCheck(vb2k_slice_agg_partition(chunks.data(), static_cast<int32_t>(chunks.size()), ncols, sliceRows, ws->data(), wsBytes, regs.data(), st()));
      int64_t distinct = std::min<int64_t>(sliceRows, static_cast<int64_t>(hllEstimate(regs) * 1.15) + 1024);  // + 9 sigma of the sketch's error
      distinct = cfg.get<int64_t>("b200.agg_slice_distinct_hint", distinct);
      const int32_t rw = rowWordsFor(Mode::kHash);
      std::vector<uint64_t> init(rw, 0);
      for (auto& a : accs) init[a.accWord] = identityBits(a);
      // the group rows land in a table-shaped buffer (power-of-two rows, unused rows keep the EMPTY key) so that
      // the extraction reads it like any hash-mode table
      const int64_t cap = static_cast<int64_t>(nextPow2(static_cast<uint64_t>(std::max<int64_t>(16, vb2k_slice_agg_output_rows(std::min<int64_t>(sliceRows, distinct))))));
      auto rowsOut = makeStorage(cap, Mode::kHash);
      auto words = allocDeviceZeroed(32, st());  // [0] groups, [1] rows reserved (int64), error (int32 at byte 16), overflow (int32 at byte 24)
      int64_t* groupsDev = words->as<int64_t>();
      int32_t* errDev = reinterpret_cast<int32_t*>(words->as<uint8_t>() + 16);
      int32_t* ovfDev = reinterpret_cast<int32_t*>(words->as<uint8_t>() + 24);
      const int rc = vb2k_slice_agg_finish(sliceRows, ncols, distinct, ops.data(), static_cast<int32_t>(ops.size()), rw, init.data(), rowsOut->as<uint64_t>(), cap,
                                           groupsDev, groupsDev + 1, errDev, ovfDev, ws->data(), wsBytes, st());
      if (rc == VB2_OK) {
        uint8_t host[32];
        VB2_CU(cudaMemcpyAsync(host, words->data(), 32, cudaMemcpyDeviceToHost, st()));
        VB2_CU(cudaStreamSynchronize(st()));
        int64_t m;
        int32_t err, ovf;
        std::memcpy(&m, host, 8);
        std::memcpy(&err, host + 16, 4);
        std::memcpy(&ovf, host + 24, 4);
        if (err == 1) throw VeloxUserError("integer overflow in sum");  // SumAggregate.cpp:24
        if (err == 0 && ovf == 0 && m > 0) {
          rowsBuf = rowsOut;
          capacity = cap;
          mode = Mode::kHash;
          numGroupsUpper = m;
          done = true;
          self->addRuntimeStat("b200.sliceAggRows", exec::RuntimeCounter{sliceRows});
          self->addRuntimeStat("b200.sliceAggGroups", exec::RuntimeCounter{m});
        }
      } else if (rc != VB2_ERR_UNSUPPORTED) {
        kernelCheck(rc);
      }
    }
    if (!done) { drainSliceBuffer(); return; }
    sliceChunks.clear();
    sliceState = SliceState::kOff;
  }

  // Large batches over a large hash-mode table: the rows are radix-partitioned by the top bits of
  // their table hash first, so that the find-or-insert + update pass walks the table slice by slice
  // (L2 resident) instead of paying a DRAM round trip per row. Returns false when the batch does not
  // qualify (then the plain path runs).
  bool addPartitioned(const B200VectorPtr& in) {
    const int64_t n = in->size();
    const auto& cfg = self->driverCtx()->queryConfig();
    if (keys.empty() || mode == Mode::kKeyed || n < cfg.get<int64_t>("b200.agg_partition_min_rows", 1 << 23)) return false;
    // Off by default: measured on B200 (1 B rows / 100 M keys, profiles/r02_config5_*.json) the partition
    // passes cost what the L2-resident update saves (31.7 ms vs 31 ms per 250 M-row batch) — the update
    // stays latency-bound behind its grid barriers. Kept, tested, as the basis for the shared-memory-slice variant.
    if (!cfg.get<bool>("b200.agg_radix_partition", false)) return false;
    // aggregate inputs: flat, NULL-free, unmasked, 4 or 8 bytes wide, at most four distinct columns
    const auto& aggs = resolved.aggregates;
    std::vector<int32_t> payload;
    for (auto& a : aggs) {
      if (a.mask >= 0) return false;
      for (int32_t c : a.inputs) {
        const vb2_column& d = in->column(c)->desc;
        if (d.encoding != VB2_FLAT || d.nulls || (d.type != VB2_BIGINT && d.type != VB2_DOUBLE && d.type != VB2_INTEGER)) return false;
        if (std::find(payload.begin(), payload.end(), c) == payload.end()) payload.push_back(c);
      }
    }
    if (payload.size() > 4) return false;
    std::vector<DeviceBufferPtr> keep;
    std::vector<vb2_column> keyCols;
    for (size_t k = 0; k < keys.size(); ++k) keyCols.push_back(keyColumn(k, *in->column(resolved.keys[k]), n, keep));
    ensureLayout(0);  // the layout must cover this batch's key ranges before its keys can be normalized
    if (mode != Mode::kHash) return false;  // small key space: array mode needs no partitioning
    // keys: one flat NULL-free integer column is normalized on the fly, anything else through vb2k_normalize_keys
    const uint64_t* norm = nullptr;
    const void* keyValues = nullptr;
    int32_t keyIs64 = 0;
    int64_t keyMin = 0;
    DeviceBufferPtr normBuf;
    auto prepareKeys = [&]() {
      const vb2_column& k0 = keyCols[0];
      if (keyCols.size() == 1 && k0.encoding == VB2_FLAT && !k0.nulls && (k0.type == VB2_BIGINT || k0.type == VB2_INTEGER) && layout.mults[0] == 1) {
        keyValues = k0.values;
        keyIs64 = k0.type == VB2_BIGINT;
        keyMin = layout.mins[0];
        norm = nullptr;
        return;
      }
      normBuf = allocDevice(static_cast<size_t>(n) * 8, st());
      kernelCheck(vb2k_normalize_keys(keyCols.data(), static_cast<int32_t>(keyCols.size()), layout.mins.data(), layout.mults.data(), nullptr, 0, nullptr, n,
                                      normBuf->as<uint64_t>(), nullptr, st()));
      norm = normBuf->as<uint64_t>();
    };
    prepareKeys();
    const size_t wsBytes = vb2k_radix_workspace_bytes(n);
    auto ws = allocDevice(wsBytes, st());
    const int32_t nregs = vb2k_radix_hll_registers();
    auto hll = allocDevice(static_cast<size_t>(nregs) * 4, st());
    kernelCheck(vb2k_radix_histogram(norm, keyValues, keyIs64, keyMin, n, ws->data(), wsBytes, hll->as<int32_t>(), st()));
    std::vector<int32_t> regs(nregs);
    VB2_CU(cudaMemcpyAsync(regs.data(), hll->data(), static_cast<size_t>(nregs) * 4, cudaMemcpyDeviceToHost, st()));
    VB2_CU(cudaStreamSynchronize(st()));
    const int64_t distinct = std::min<int64_t>(n, static_cast<int64_t>(hllEstimate(regs) * 1.15) + 1024);  // + 9 sigma of the sketch's error
    const Mode before = mode;
    const KeyLayout layoutBefore = layout;
    ensureLayout(distinct);  // grows the table ONCE for the whole batch (a relayout keeps the layout: the ranges are already covered)
    VELOX_CHECK(mode == before && layout.mins == layoutBefore.mins && layout.mults == layoutBefore.mults, "layout changed after the keys were read");
    // keys + payload columns into partition order
    auto partKeys = allocDevice(static_cast<size_t>(n) * 8, st());
    auto partStart = allocDevice(257 * 8, st());
    std::vector<const void*> colsIn;
    std::vector<void*> colsOut;
    std::vector<int32_t> colBytes;
    std::map<int32_t, const void*> permuted;
    for (int32_t c : payload) {
      const vb2_column& d = in->column(c)->desc;
      const int w = d.type == VB2_INTEGER ? 4 : 8;
      auto out = allocDevice(static_cast<size_t>(n) * w, st());
      keep.push_back(out);
      colsIn.push_back(d.values);
      colsOut.push_back(out->data());
      colBytes.push_back(w);
      permuted[c] = out->data();
    }
    kernelCheck(vb2k_radix_scatter(norm, keyValues, keyIs64, keyMin, n, ws->data(), wsBytes, colsIn.data(), colsOut.data(), colBytes.data(),
                                   static_cast<int32_t>(colsIn.size()), partKeys->as<uint64_t>(), partStart->as<int64_t>(), st()));
    ++genericBatches;
    ++partitionedBatches;
    partStartDev = partStart;
    applyUpdates(in, n, partKeys, keyCols, keep, &permuted);
    partStartDev = nullptr;
    return true;
  }

  void addGeneric(const B200VectorPtr& in) {
    const int64_t n = in->size();
    ++genericBatches;
    std::vector<DeviceBufferPtr> keep;
    DeviceBufferPtr rowKeys;
    std::vector<vb2_column> keyCols;
    if (!keys.empty()) {
      for (size_t k = 0; k < keys.size(); ++k) keyCols.push_back(keyColumn(k, *in->column(resolved.keys[k]), n, keep));
      ensureLayout(n);
    }
    if (!keys.empty() && mode != Mode::kKeyed) {
      auto nk = allocDevice(static_cast<size_t>(n) * 8, st());
      kernelCheck(vb2k_normalize_keys(keyCols.data(), static_cast<int32_t>(keyCols.size()), layout.mins.data(), layout.mults.data(), nullptr, 0,
                                      nullptr, n, nk->as<uint64_t>(), nullptr, st()));
      rowKeys = nk;
    }
    applyUpdates(in, n, rowKeys, keyCols, keep, nullptr);
  }

  // Find-or-insert + every aggregate update of one batch. rowKeys: normalized keys in the order the
  // rows are visited (input order, or partition order with `permuted` giving the reordered input columns).
  void applyUpdates(const B200VectorPtr& in, int64_t n, const DeviceBufferPtr& rowKeys, std::vector<vb2_column>& keyCols, std::vector<DeviceBufferPtr>& keep,
                    const std::map<int32_t, const void*>* permuted) {
    std::vector<vb2_agg_update> ups;
    // Points an update at its input column. Flat columns are read in place; a dictionary wrap over
    // fixed-width values (what a filter leaves behind, exec/Operator.cpp:270-303 wrapChild) is read
    // through its indices by the update kernel — no flattening pass; anything else (constant,
    // BOOLEAN bits) is decoded once per column and batch.
    std::map<int32_t, FlatColumn> flattened;
    auto flatInput = [&](int32_t colIdx, vb2_agg_update& u) {
      const DeviceColumnPtr& c = in->column(colIdx);
      u.input_type = c->desc.type;
      if (permuted) {  // partition order: the flat NULL-free column was reordered together with the keys
        u.input = permuted->at(colIdx);
        return;
      }
      if (c->desc.encoding == VB2_FLAT && u.input_type != VB2_BOOLEAN) {
        u.input = c->desc.values;
        u.nulls = c->desc.nulls;
        return;
      }
      if (c->desc.encoding == VB2_DICTIONARY && u.input_type != VB2_BOOLEAN && u.input_type != VB2_VARCHAR) {
        u.input = c->desc.values;
        u.indices = c->desc.indices;
        u.nulls = c->desc.nulls;
        u.base_nulls = c->desc.dict_nulls;
        return;
      }
      auto it = flattened.find(colIdx);
      if (it == flattened.end()) {
        it = flattened.emplace(colIdx, flattenColumn(c, nullptr, n, st())).first;
        keep.push_back(it->second.values);
        if (it->second.nulls) keep.push_back(it->second.nulls);
      }
      u.input = it->second.values->data();
      u.nulls = it->second.nulls ? it->second.nulls->as<uint64_t>() : nullptr;
    };
    const auto& aggs = resolved.aggregates;
    for (size_t i = 0; i < aggs.size(); ++i) {
      const auto& a = aggs[i];
      AccState& s = accs[i];
      const uint64_t* mask = nullptr;
      if (a.mask >= 0) {
        const DeviceColumnPtr& mc = in->column(a.mask);
        VELOX_CHECK(mc->desc.type == VB2_BOOLEAN && mc->desc.encoding == VB2_FLAT, "aggregate masks must be flat BOOLEAN columns");
        if (mc->desc.nulls) {
          auto mb = allocDevice(bits::nbytes(n), st());
          kernelCheck(vb2k_and_bits(reinterpret_cast<const uint64_t*>(mc->desc.values), mc->desc.nulls, n, mb->as<uint64_t>(), st()));
          keep.push_back(mb);
          mask = mb->as<uint64_t>();
        } else {
          mask = reinterpret_cast<const uint64_t*>(mc->desc.values);
        }
      }
      vb2_agg_update u{};
      u.mask = mask;
      u.acc_word = s.accWord + shift();
      u.nonnull_word = -1;
      if (a.function == "count") {
        u.kind = raw ? VB2_AGG_COUNT : VB2_AGG_COUNT_MERGE;
        if (!a.inputs.empty()) flatInput(a.inputs[0], u);
        ups.push_back(u);
        continue;
      }
      flatInput(a.inputs[0], u);
      u.kind = s.kind;
      if (a.function == "avg" && !raw) {
        // intermediate (sum, count): add the sums, merge the counts into the non-null counter
        ups.push_back(u);
        vb2_agg_update c{};
        c.kind = VB2_AGG_COUNT_MERGE;
        c.mask = mask;
        c.nonnull_word = -1;
        flatInput(a.inputs[1], c);
        c.acc_word = s.nnWord + shift();
        ups.push_back(c);
        continue;
      }
      if (mask || u.nulls || u.base_nulls) trackNonNull(i);
      if (s.nnTracked) u.nonnull_word = s.nnWord + shift();
      ups.push_back(u);
    }
    const vb2_group_table t = table();
    const uint64_t* rk = rowKeys ? rowKeys->as<uint64_t>() : nullptr;
    for (size_t i = 0; i == 0 || i < ups.size(); i += 16) {
      // the first call inserts the groups; later slices of a long aggregate list find them again
      const int32_t cnt = static_cast<int32_t>(std::min<size_t>(16, ups.size() - i));
      if (partStartDev && mode == Mode::kHash && capacity >= 65536) {
        if (!barrierWord) barrierWord = allocDeviceZeroed(8, st());
        kernelCheck(vb2k_group_update_partitioned(&t, rk, partStartDev->as<int64_t>(), 256, n, ups.data() + i, cnt, i == 0 ? numGroupsDev->as<int64_t>() : nullptr,
                                                  errorFlag->as<int32_t>(), barrierWord->as<uint32_t>(), st()));
      } else if (mode == Mode::kKeyed)
        kernelCheck(vb2k_group_update_keyed(&t, keyCols.data(), static_cast<int32_t>(keyCols.size()), n, ups.data() + i, cnt,
                                            i == 0 ? numGroupsDev->as<int64_t>() : nullptr, errorFlag->as<int32_t>(), st()));
      else
        kernelCheck(vb2k_group_update(&t, rk, nullptr, n, ups.data() + i, cnt, i == 0 ? numGroupsDev->as<int64_t>() : nullptr,
                                      errorFlag->as<int32_t>(), st()));
    }
    sawGeneric = true;
    numGroupsUpper += n;
    // an array-mode table cannot hold more groups than it has slots, however many rows went in
    if (mode == Mode::kArray) numGroupsUpper = std::min<int64_t>(numGroupsUpper, capacity);
    if (mode == Mode::kHash || mode == Mode::kKeyed) {
      // tighten the bound with the real group count (one 8-byte read per batch)
      int64_t g = 0;
      VB2_CU(cudaMemcpyAsync(&g, numGroupsDev->data(), 8, cudaMemcpyDeviceToHost, st()));
      VB2_CU(cudaStreamSynchronize(st()));
      numGroupsUpper = g;
    }
    // buffers in `keep` are freed stream-ordered after the kernels above
  }

  // ---- fused path -----------------------------------------------------------------------------------
  // Decides at plan time whether the absorbed chain can be expressed as one fused pipeline.
  void planFused() {
    if (!raw || !self->driverCtx()->queryConfig().b200FusedPipelines()) return;
    if (inputProgram) return;  // transformed inputs are columns the fused scan does not produce
    if (keys.size() > VB2_FUSED_MAX_KEYS) return;
    // shape of the absorbed chain
    B200FilterProject *fp1 = nullptr, *fp2 = nullptr;
    B200HashProbe* probe = nullptr;
    if (absorbed.size() == 1) fp1 = dynamic_cast<B200FilterProject*>(absorbed[0].get());
    else if (absorbed.size() == 2) {
      probe = dynamic_cast<B200HashProbe*>(absorbed[0].get());
      fp2 = dynamic_cast<B200FilterProject*>(absorbed[1].get());
      if (!probe || !fp2 || fp2->hasFilter()) return;
      if (probe->node()->joinType() != core::JoinType::kInner || probe->node()->filter() || probe->node()->leftKeys().size() != 1) return;
    } else if (absorbed.size() == 3) {
      fp1 = dynamic_cast<B200FilterProject*>(absorbed[0].get());
      probe = dynamic_cast<B200HashProbe*>(absorbed[1].get());
      fp2 = dynamic_cast<B200FilterProject*>(absorbed[2].get());
      if (!probe || !fp2 || fp2->hasFilter()) return;
      if (probe->node()->joinType() != core::JoinType::kInner || probe->node()->filter() || probe->node()->leftKeys().size() != 1) return;
    }
    if (!fp1 && !probe) return;
    const RowTypePtr srcType = fp1 ? fp1->inputType() : probe->node()->sources()[0]->outputType();
    // expressions of the stage feeding the aggregation, in terms of source columns; build-side
    // columns are marked with index = -(buildColumn + 1)
    std::vector<core::TypedExprPtr> stage;
    core::TypedExprPtr filter;
    if (fp1) {
      const auto& e = fp1->exprs();
      size_t first = 0;
      if (fp1->hasFilter()) { filter = e[0]; first = 1; }
      for (size_t i = first; i < e.size(); ++i) stage.push_back(e[i]);
    } else {
      // the probe reads the source batch directly: the stage is the identity over its columns
      for (uint32_t i = 0; i < srcType->size(); ++i)
        stage.push_back(std::make_shared<core::FieldAccessTypedExpr>(srcType->childAt(i), srcType->nameOf(i)));
    }
    int joinKeySource = -1;
    if (probe) {
      const ResolvedJoin rj = resolveJoin(*probe->node());
      auto keyExpr = stage.at(rj.leftKeys[0]);
      auto kf = dynamic_cast<const core::FieldAccessTypedExpr*>(keyExpr.get());
      if (!kf) return;
      joinKeySource = channelOf(srcType, *kf);
      std::vector<core::TypedExprPtr> joined;
      const auto& bt = probe->node()->sources()[1]->outputType();
      for (auto& o : rj.outputs) {
        if (o.fromProbe) joined.push_back(stage.at(o.column));
        else joined.push_back(std::make_shared<core::FieldAccessTypedExpr>(bt->childAt(o.column), buildFieldName(o.column)));
      }
      std::vector<core::TypedExprPtr> after;
      for (auto& e : fp2->exprs()) after.push_back(substituteFields(e, joined, probe->node()->outputType()));
      stage = after;
    }
    // aggregate inputs -> deduplicated projections
    std::vector<core::TypedExprPtr> projs;
    std::vector<std::string> projKeys;
    aggToProj.clear();
    for (auto& a : resolved.aggregates) {
      if (a.mask >= 0) return;
      if (a.function == "count") {
        if (!a.inputs.empty()) return;  // count(x) needs x's nulls; count(*) only
        aggToProj.push_back(-1);
        continue;
      }
      if (a.function != "sum" && a.function != "avg") return;
      auto e = stage.at(a.inputs[0]);
      if (e->type()->kind() != TypeKind::DOUBLE) return;
      const std::string key = e->toString() + "#" + std::to_string(reinterpret_cast<uintptr_t>(e.get()));
      int found = -1;
      for (size_t i = 0; i < projs.size(); ++i)
        if (projs[i].get() == e.get()) found = static_cast<int>(i);
      if (found < 0) { projs.push_back(e); found = static_cast<int>(projs.size()) - 1; }
      aggToProj.push_back(found);
    }
    if (projs.empty()) return;
    // group keys must be plain source columns
    fusedKeySourceCols.clear();
    for (int32_t k : resolved.keys) {
      auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(stage.at(k).get());
      if (!f || buildFieldColumn(*f) >= 0) return;
      fusedKeySourceCols.push_back(channelOf(srcType, *f));
    }
    // the build-side predicate: the one BOOLEAN sub-expression that touches build columns
    const core::ITypedExpr* flag = nullptr;
    if (probe) {
      std::function<bool(const core::TypedExprPtr&)> touchesBuild = [&](const core::TypedExprPtr& e) {
        if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) return buildFieldColumn(*f) >= 0;
        for (auto& in : e->inputs()) if (touchesBuild(in)) return true;
        return false;
      };
      std::function<bool(const core::TypedExprPtr&)> onlyBuild = [&](const core::TypedExprPtr& e) {
        if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) return buildFieldColumn(*f) >= 0;
        if (dynamic_cast<const core::ConstantTypedExpr*>(e.get())) return true;
        for (auto& in : e->inputs()) if (!onlyBuild(in)) return false;
        return true;
      };
      bool bad = false;
      std::function<void(const core::TypedExprPtr&)> find = [&](const core::TypedExprPtr& e) {
        if (!touchesBuild(e)) return;
        if (onlyBuild(e) && e->type()->kind() == TypeKind::BOOLEAN) {
          if (flag && flag != e.get() && flag->toString() != e->toString()) bad = true;
          if (!flag) { flag = e.get(); joinFlagExpr = e; }
          return;
        }
        if (dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) { bad = true; return; }
        for (auto& in : e->inputs()) find(in);
      };
      for (auto& p : projs) find(p);
      if (bad) return;
    }
    binding = fusedSignature(filter, projs, srcType, joinKeySource, flag);
    if (!binding.ok) return;
    fusedId = vb2k_fused_find(binding.signature.c_str());
    self->addRuntimeStat("b200.fusedSignatureMatched", exec::RuntimeCounter{fusedId >= 0 ? 1 : 0});
    if (fusedId < 0) return;
    fusedProbe = probe;
    fusedJoinKeySourceCol = joinKeySource;
    fusedPlanned = true;
  }

  // Build-side preparation of the fused probe: dense key -> flag table. Returns false when the
  // build side does not qualify (duplicate keys, hash mode ...).
  bool prepareFusedJoin() {
    if (!fusedProbe) return true;
    if (joinSlotFlags) return true;
    auto jt = fusedProbe->table();
    if (!jt || jt->table.mode != 0 || jt->hasDuplicateKeys || !jt->rows) return false;
    DeviceBufferPtr flagBuf;
    const int32_t* codes = nullptr;
    if (joinFlagExpr) {
      // the predicate's single build column
      int buildCol = -1;
      std::function<void(const core::TypedExprPtr&)> scan = [&](const core::TypedExprPtr& e) {
        if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) {
          const int c = buildFieldColumn(*f);
          if (buildCol >= 0 && buildCol != c) buildCol = -2;
          else if (buildCol != -2) buildCol = c;
        }
        for (auto& in : e->inputs()) scan(in);
      };
      scan(joinFlagExpr);
      if (buildCol < 0) return false;
      const DeviceColumnPtr& bc = jt->rows->column(buildCol);
      // evaluate the predicate once per dictionary entry (or per build row when flat)
      vb2_column view = bc->desc;
      int64_t entries = bc->desc.size;
      if (bc->desc.encoding == VB2_DICTIONARY) {
        if (bc->desc.nulls) return false;
        view.encoding = VB2_FLAT;
        view.size = bc->desc.dict_size;
        view.nulls = bc->desc.dict_nulls;
        view.indices = nullptr;
        entries = bc->desc.dict_size;
        codes = bc->desc.indices;
      } else if (bc->desc.encoding != VB2_FLAT) {
        return false;
      }
      std::vector<core::TypedExprPtr> one{std::make_shared<core::FieldAccessTypedExpr>(bc->type, "b")};
      auto pred = substituteBuild(joinFlagExpr, one[0]);
      CompiledProgram prog = compileExprs({pred}, false, ROW({"b"}, {bc->type}));
      prog.uploadConstants(st());
      const vb2_program pv = prog.view();
      flagBuf = allocDevice(static_cast<size_t>(entries) + 8, st());
      auto nulls = allocDevice(bits::nbytes(entries), st());
      vb2_output o{prog.outputs[0].reg, VB2_BOOLEAN, flagBuf->data(), nulls->as<uint64_t>()};
      kernelCheck(vb2k_eval_project(&pv, &view, 1, nullptr, entries, &o, 1, errorFlag->as<int32_t>(), st()));
      // NULL predicate results were written as 0 = false, matching CASE semantics
    }
    joinSlotFlags = allocDevice(static_cast<size_t>(jt->table.capacity) + 8, st());
    kernelCheck(vb2k_join_slot_flags(jt->table.head, codes, flagBuf ? flagBuf->as<uint8_t>() : nullptr, jt->table.capacity,
                                     joinSlotFlags->as<uint8_t>(), st()));
    return true;
  }
  static core::TypedExprPtr substituteBuild(const core::TypedExprPtr& e, const core::TypedExprPtr& field) {
    if (dynamic_cast<const core::FieldAccessTypedExpr*>(e.get())) return field;
    if (dynamic_cast<const core::ConstantTypedExpr*>(e.get())) return e;
    std::vector<core::TypedExprPtr> in;
    for (auto& i : e->inputs()) in.push_back(substituteBuild(i, field));
    if (auto c = dynamic_cast<const core::CallTypedExpr*>(e.get())) return std::make_shared<core::CallTypedExpr>(e->type(), std::move(in), c->name());
    if (auto c = dynamic_cast<const core::CastTypedExpr*>(e.get())) return std::make_shared<core::CastTypedExpr>(e->type(), in[0], c->nullOnFailure());
    VELOX_UNSUPPORTED("unknown expression node");
  }

  // Tries the fused kernel for this source batch. Returns false when the batch does not qualify.
  bool tryFused(const B200VectorPtr& src) {
    if (!fusedPlanned || fusedId < 0) return false;
    const int64_t n = src->size();
    vb2_fused_args a{};
    for (size_t i = 0; i < binding.columns.size(); ++i) {
      const vb2_column& d = src->column(binding.columns[i])->desc;
      if (d.encoding != VB2_FLAT || d.nulls) return false;
      a.cols[i] = d.values;
    }
    for (size_t i = 0; i < binding.pf.size(); ++i) a.pf[i] = binding.pf[i];
    for (size_t i = 0; i < binding.pl.size(); ++i) a.pl[i] = binding.pl[i];
    for (size_t i = 0; i < binding.pi.size(); ++i) a.pi[i] = binding.pi[i];
    a.rows = n;
    std::vector<DeviceBufferPtr> keep;
    // group keys: dictionary VARCHAR (indices + LUT of global ids) or flat INTEGER/BIGINT
    a.nkeys = static_cast<int32_t>(keys.size());
    for (size_t k = 0; k < keys.size(); ++k) {
      const DeviceColumn& col = *src->column(fusedKeySourceCols[k]);
      const vb2_column& d = col.desc;
      if (d.nulls || d.dict_nulls) return false;
      if (keys[k].isVarchar) {
        if (d.encoding != VB2_DICTIONARY || !col.alphabet) return false;
        for (bool isNull : col.alphabet->nulls) if (isNull) return false;
        (void)keyColumn(k, col, n, keep);  // assigns global ids, refreshes the LUT and the range
        a.key[k] = d.indices;
        a.key_is64[k] = 0;
      } else {
        if (d.encoding != VB2_FLAT || (d.type != VB2_INTEGER && d.type != VB2_BIGINT)) return false;
        (void)keyColumn(k, col, n, keep);
        a.key[k] = d.values;
        a.key_is64[k] = d.type == VB2_BIGINT;
      }
    }
    if (!keys.empty()) {
      ensureLayout(n);
      if (mode != Mode::kArray || layout.product > static_cast<uint64_t>(kFusedMaxGroups)) return false;
      for (size_t k = 0; k < keys.size(); ++k) {
        // kernel: id = lut ? lut[v - key_min] : v - key_min ; layout: id = v - mins[k] + 1
        a.key_mult[k] = static_cast<int32_t>(layout.mults[k]);
        if (!keys[k].isVarchar) { a.key_min[k] = layout.mins[k] - 1; continue; }
        // VARCHAR: per-dictionary LUT of layout ids (global id - mins + 1), rebuilt when either changes
        KeyState& ks = keys[k];
        const DeviceColumn& col = *src->column(fusedKeySourceCols[k]);
        if (!ks.fusedLut) {
          std::vector<int32_t> lut(col.alphabet->values.size());
          for (size_t i = 0; i < lut.size(); ++i) lut[i] = static_cast<int32_t>(ks.valueIds.at(col.alphabet->values[i]) - layout.mins[k] + 1);
          ks.fusedLut = allocDevice(lut.size() * 4 + 4, st());
          VB2_CU(cudaMemcpyAsync(ks.fusedLut->data(), lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, st()));
        }
        a.key_min[k] = 0;
        a.key_lut[k] = ks.fusedLut->as<int32_t>();
      }
    }
    const int groups = keys.empty() ? 1 : static_cast<int>(layout.product);
    if (!prepareFusedJoin()) return false;
    if (fusedProbe) {
      auto jt = fusedProbe->table();
      a.join_slot_flags = joinSlotFlags->as<uint8_t>();
      // slot = normalized key = v - min + 1  =>  v - (min - 1)
      a.join_min = jt->layout.mins[0] - 1;
      a.join_range = jt->table.capacity;
    }
    a.ngroups = groups;
    const int np = vb2k_fused_nproj(fusedId);
    if (!fusedSums || fusedGroups != groups) {
      flushFused();
      fusedGroups = groups;
      fusedSums = allocDeviceZeroed(static_cast<size_t>(groups) * np * 8, st());
      fusedCounts = allocDeviceZeroed(static_cast<size_t>(groups) * 8, st());
      fusedWsBytes = vb2k_fused_workspace_bytes(fusedId, groups);
      fusedWs = allocDevice(fusedWsBytes, st());
    }
    // The fused launch is bracketed by CUDA events on the launching stream; the elapsed device
    // time is reported as b200.fusedScanNanos (what bench.py's roofline divides the bytes by).
    FusedTiming ft;
    VB2_CU(cudaEventCreate(&ft.begin));
    VB2_CU(cudaEventCreate(&ft.end));
    VB2_CU(cudaEventRecord(ft.begin, st()));
    int rc = VB2_ERR_UNSUPPORTED;
    if (chooseSelective(a, n, groups)) {
      // Late materialisation: filter columns only -> selection bitmap -> row numbers -> probe /
      // project / aggregate over the surviving rows. Everything stays on the stream; the row count
      // is read by the gather kernel on the device.
      auto bitsBuf = allocDevice(bits::nbytes(n), st());
      auto kept = allocDeviceZeroed(16, st());
      rc = vb2k_fused_filter_bits(fusedId, &a, 1, bitsBuf->as<uint64_t>(), kept->as<int64_t>(), st());
      if (rc == VB2_OK) {
        // the filter kernel counted the survivors: one 16-byte read sizes the row-number buffer exactly
        int64_t h[2] = {0, 0};
        VB2_CU(cudaMemcpyAsync(h, kept->data(), 16, cudaMemcpyDeviceToHost, st()));
        VB2_CU(cudaStreamSynchronize(st()));
        auto sel = allocDevice(static_cast<size_t>(h[0] ? h[0] : 1) * 4, st());
        auto cnt = allocDevice(8, st());
        const size_t wsb = vb2k_bits_to_indices_workspace(n);
        auto ws = allocDevice(wsb, st());
        kernelCheck(vb2k_bits_to_indices(bitsBuf->as<uint64_t>(), n, sel->as<int32_t>(), cnt->as<int64_t>(), ws->data(), wsb, st()));
        const int64_t hint = std::max<int64_t>(1024, h[0]);
        rc = vb2k_fused_gather_agg(fusedId, &a, sel->as<int32_t>(), cnt->as<int64_t>(), hint, fusedSums->as<double>(), fusedCounts->as<int64_t>(), fusedWs->data(),
                                   fusedWsBytes, st());
        if (rc == VB2_OK) ++selectiveBatches;
      }
      if (rc != VB2_OK) selectiveUsable = false;  // e.g. unaligned filter columns: scan everything instead
    }
    if (rc == VB2_ERR_UNSUPPORTED)
      rc = vb2k_fused_scan_agg(fusedId, &a, fusedSums->as<double>(), fusedCounts->as<int64_t>(), fusedWs->data(), fusedWsBytes, st());
    if (rc == VB2_OK) VB2_CU(cudaEventRecord(ft.end, st()));
    if (rc != VB2_OK) { cudaEventDestroy(ft.begin); cudaEventDestroy(ft.end); }
    if (rc == VB2_ERR_UNSUPPORTED) return false;
    kernelCheck(rc);
    ft.rows = n;
    fusedTimings.push_back(ft);
    ++fusedBatches;
    return true;
  }

  // Bytes per row the filter alone reads / the rest of the pipeline reads, from the signature's
  // column tokens (f = 8, l = 8, i = 4 bytes).
  void signatureBytes(double& filterBytes, double& restBytes) const {
    const std::string& sg = binding.signature;
    const size_t pp = sg.find(";P:");
    std::set<std::string> filterCols, allCols;
    auto scan = [](const std::string& part, std::set<std::string>& out) {
      size_t i = 0;
      while (i < part.size()) {
        if (!std::isalnum(static_cast<unsigned char>(part[i]))) { ++i; continue; }
        size_t j = i;
        while (j < part.size() && std::isalnum(static_cast<unsigned char>(part[j]))) ++j;
        const std::string id = part.substr(i, j - i);
        if (id.size() >= 2 && (id[0] == 'f' || id[0] == 'i' || id[0] == 'l') && std::isdigit(static_cast<unsigned char>(id[1]))) out.insert(id);
        i = j;
      }
    };
    scan(sg.substr(0, pp), filterCols);
    scan(sg, allCols);
    auto width = [](const std::string& id) { return id[0] == 'i' ? 4.0 : 8.0; };
    filterBytes = restBytes = 0;
    for (auto& c : allCols) (filterCols.count(c) ? filterBytes : restBytes) += width(c);
  }

  // Decides, once per operator, whether this pipeline's filter is selective enough for late
  // materialisation: a strided sample of the first batch (every k-th 1024-row tile, about a million
  // rows) gives the selectivity; the gather touches whole 32-byte sectors, so a column of 8-byte
  // values costs about 1 - (1 - s)^4 of its full traffic.
  bool chooseSelective(const vb2_fused_args& a, int64_t n, int groups) {
    if (!selectiveUsable || groups > 4 || n < self->driverCtx()->queryConfig().get<int64_t>("b200.late_materialization_min_rows", 1 << 20)) return false;
    if (!selectiveDecided) {
      selectiveDecided = true;
      selectiveUsable = false;
      if (!self->driverCtx()->queryConfig().get<bool>("b200.late_materialization", true)) return false;
      if (!vb2k_fused_has_filter(fusedId)) return false;
      double fb = 0, rb = 0;
      signatureBytes(fb, rb);
      if (rb <= 0) return false;
      auto counters = allocDeviceZeroed(16, st());
      const int stride = static_cast<int>(std::max<int64_t>(1, (n / 1024) / 1024));
      const int rc = vb2k_fused_filter_bits(fusedId, &a, stride, nullptr, counters->as<int64_t>(), st());
      if (rc != VB2_OK) return false;
      int64_t h[2] = {0, 0};
      VB2_CU(cudaMemcpyAsync(h, counters->data(), 16, cudaMemcpyDeviceToHost, st()));
      VB2_CU(cudaStreamSynchronize(st()));
      if (h[1] <= 0) return false;
      selectivity = static_cast<double>(h[0]) / static_cast<double>(h[1]);
      const double sectors = 1.0 - std::pow(1.0 - selectivity, 4.0);
      const double lateCost = (fb + sectors * rb + selectivity * 12.0) * 1.15;  // + row numbers written and read back, + launch slack
      selectiveUsable = lateCost < fb + rb;
      self->addRuntimeStat("b200.sampledSelectivityPpm", exec::RuntimeCounter{static_cast<int64_t>(selectivity * 1e6)});
    }
    return selectiveUsable;
  }

  // Folds the fused partial sums into the group rows (small: <= kFusedMaxGroups groups) on the
  // device: one launch, no host round trip.
  void flushFused() {
    if (!fusedSums) return;
    const int np = vb2k_fused_nproj(fusedId);
    VELOX_CHECK(static_cast<int64_t>(fusedGroups) <= capacity && mode != Mode::kHash && mode != Mode::kKeyed, "fused group space does not match the table");
    std::vector<int32_t> words{0}, projs{-1};  // word 0: rows seen (occupancy)
    for (size_t i = 0; i < accs.size(); ++i) {
      trackNonNull(i);
      const int p = aggToProj[i];
      words.push_back(accs[i].accWord);
      projs.push_back(p);  // p < 0: count(*)
      if (p >= 0 && accs[i].nnWord >= 0) {
        words.push_back(accs[i].nnWord);
        projs.push_back(-1);
      }
    }
    const vb2_group_table t = table();
    kernelCheck(vb2k_group_merge_partials(&t, fusedSums->as<double>(), fusedCounts->as<int64_t>(), fusedGroups, np, words.data(), projs.data(),
                                          static_cast<int32_t>(words.size()), st()));
    fusedSums = nullptr;
    fusedCounts = nullptr;
  }

  // ---- input ------------------------------------------------------------------------------------
  void addInput(const B200VectorPtr& src) {
    sawInput = true;
    if (tryFused(src)) return;
    B200VectorPtr cur = src;
    for (auto& op : absorbed) {
      if (auto fp = dynamic_cast<B200FilterProject*>(op.get())) cur = fp->apply(cur);
      else if (auto pr = dynamic_cast<B200HashProbe*>(op.get())) cur = pr->apply(cur);
      if (!cur) return;  // every row filtered out
    }
    if (inputProgram) cur = evalProjections(*inputProgram, cur, nullptr, cur->size(), st(), errorFlag, inputProgramType, self->pool());
    // Hash-mode tables are sized by (groups so far + rows of one pass): large batches go through
    // find-or-insert in bounded passes so a high-cardinality table ends near 2x its group count
    // instead of 2x the batch.
    if (addSliceBuffered(cur)) return;
    if (addPartitioned(cur)) return;
    const int64_t chunk = std::max<int64_t>(1 << 16, self->driverCtx()->queryConfig().b200AggProbeChunkRows()) / 64 * 64;
    if (!keys.empty() && cur->size() > chunk) {
      // the first pass shows whether these keys need a hash table at all; array mode takes the rest at once
      for (int64_t off = 0; off < cur->size();) {
        const int64_t len = (off == 0 || mode == Mode::kHash || mode == Mode::kKeyed) ? std::min<int64_t>(chunk, cur->size() - off) : cur->size() - off;
        addGeneric(sliceVector(cur, off, len));
        off += len;
      }
      return;
    }
    addGeneric(cur);
  }

  // ---- output -----------------------------------------------------------------------------------
  DeviceColumnPtr flatOutput(const TypePtr& type, DeviceBufferPtr values, DeviceBufferPtr valid, int64_t n) {
    auto col = std::make_shared<DeviceColumn>();
    col->type = type;
    col->desc.type = veloxTypeToVb2(type);
    col->desc.encoding = VB2_FLAT;
    col->desc.size = n;
    col->desc.values = values->data();
    col->owners.push_back(values);
    if (valid) {
      col->desc.nulls = valid->as<uint64_t>();
      col->owners.push_back(valid);
    }
    return col;
  }

  // Device copy of a VARCHAR key's global alphabet (offsets + chars), refreshed when ids were added.
  void ensureDeviceAlphabet(KeyState& ks) {
    if (ks.devOffsets && ks.devAlphabetCount == ks.valuesById.size()) return;
    ks.hostOffsets.assign(ks.valuesById.size() + 1, 0);
    ks.hostChars.clear();
    for (size_t i = 0; i < ks.valuesById.size(); ++i) {
      ks.hostChars += ks.valuesById[i];
      ks.hostOffsets[i + 1] = static_cast<int32_t>(ks.hostChars.size());
    }
    ks.devOffsets = allocDevice(ks.hostOffsets.size() * 4, st());
    ks.devChars = allocDevice(ks.hostChars.size() + 1, st());
    VB2_CU(cudaMemcpyAsync(ks.devOffsets->data(), ks.hostOffsets.data(), ks.hostOffsets.size() * 4, cudaMemcpyHostToDevice, st()));
    if (!ks.hostChars.empty()) VB2_CU(cudaMemcpyAsync(ks.devChars->data(), ks.hostChars.data(), ks.hostChars.size(), cudaMemcpyHostToDevice, st()));
    ks.devAlphabetCount = ks.valuesById.size();
    auto alpha = std::make_shared<HostAlphabet>();
    alpha->values = ks.valuesById;
    alpha->nulls.assign(alpha->values.size(), false);
    ks.outAlphabet = alpha;
  }

  // All output columns are carved out of ONE device arena and written by ONE kernel
  // (vb2k_group_extract). Small tables (<= VB2_EXTRACT_SMALL_CAPACITY rows, i.e. every TPC-H shape)
  // are compacted inside that kernel and the whole arena comes back to pinned host memory with the
  // row count in a single copy + synchronisation; B200ToHost then reads the host mirror.
  B200VectorPtr output() {
    auto out = extractGroups();
    if (out && finalProgram) out = evalProjections(*finalProgram, out, nullptr, out->size(), st(), errorFlag, node->outputType(), self->pool());
    return out;
  }

  B200VectorPtr extractGroups() {
    flushFused();
    finishSliceAggregation();
    if (!sawInput && mode != Mode::kGlobal) return nullptr;  // a grouped aggregation over no input emits nothing
    const bool small = capacity <= VB2_EXTRACT_SMALL_CAPACITY;
    int64_t m = 0;
    DeviceBufferPtr slots;
    if (!small) {
      checkDeviceError(errorFlag, st(), "sum");  // SUM(BIGINT) overflow (functions/prestosql/aggregates/SumAggregate.cpp:24)
      slots = occupiedSlots(m);
      if (m == 0) return nullptr;
    }
    const int64_t rowsCap = small ? capacity : m;
    const RowTypePtr& outType = finalProgram ? extractType : node->outputType();  // final transforms run on the extracted batch
    const auto& inType = node->sources()[0]->outputType();

    struct ColPlan {
      vb2_extract_col ec{};
      TypePtr type;
      int width = 8;
      bool varcharKey = false;
      size_t key = 0;
      size_t valuesOff = 0, validOff = 0;
      bool hasValid = false;
    };
    std::vector<ColPlan> plan;
    uint32_t oc = 0;
    for (size_t k = 0; k < keys.size(); ++k, ++oc) {
      ColPlan c;
      c.type = inType->childAt(resolved.keys[k]);
      c.key = k;
      c.ec.mult = c.ec.range = 1;
      if (mode == Mode::kKeyed) {
        // the key column is word 1 + k of the row, its NULL bit is bit k of word 1 + K
        c.ec.kind = VB2_EXTRACT_KEYWORD;
        c.ec.word = 1 + static_cast<int32_t>(k);
        c.ec.count_word = 1 + static_cast<int32_t>(keys.size());
        c.ec.null_reserved = static_cast<int32_t>(k);
        c.hasValid = keys[k].nullable;
        c.ec.min = 0;
      } else {
        c.ec.kind = VB2_EXTRACT_KEY;
        c.ec.mult = layout.mults[k] ? layout.mults[k] : 1;
        c.ec.range = layout.ranges[k] ? layout.ranges[k] : 1;
        c.ec.null_reserved = nullReserved[k];
        c.ec.count_word = -1;
        c.hasValid = nullReserved[k] != 0;
        c.ec.min = layout.mins[k];
      }
      if (keys[k].isVarchar) {
        c.varcharKey = true;  // dictionary over the global alphabet: index = global id - 1
        c.ec.type = VB2_INTEGER;
        c.ec.min -= 1;
        c.width = 4;
      } else {
        c.ec.type = veloxTypeToVb2(c.type);
        c.width = c.ec.type == VB2_INTEGER ? 4 : (c.ec.type == VB2_BOOLEAN ? 0 : 8);
      }
      plan.push_back(c);
    }
    for (size_t i = 0; i < accs.size(); ++i) {
      const AccState& s = accs[i];
      auto add = [&](int32_t kind, int32_t word, int32_t countWord, bool valid, int width) {
        ColPlan c;
        c.type = outType->childAt(oc++);
        c.ec.kind = kind;
        c.ec.word = word + shift();
        c.ec.count_word = countWord >= 0 ? countWord + shift() : -1;
        c.ec.mult = c.ec.range = 1;
        c.hasValid = valid;
        c.width = width;
        plan.push_back(c);
      };
      const bool tracked = s.nnTracked;  // untracked: every input of every group was non-null
      if (s.fn == "count") add(VB2_EXTRACT_WORD, s.accWord, -1, false, 8);
      else if (s.fn == "avg") {
        if (fin) add(VB2_EXTRACT_AVG, s.accWord, s.nnWord, tracked, 8);
        else { add(VB2_EXTRACT_WORD, s.accWord, tracked ? s.nnWord : -1, tracked, 8); add(VB2_EXTRACT_WORD, s.nnWord, tracked ? s.nnWord : -1, tracked, 8); }
      } else {
        const bool narrow = outType->childAt(oc)->kind() == TypeKind::INTEGER;  // min/max over INTEGER keep their type
        add(narrow ? VB2_EXTRACT_WORD_I32 : VB2_EXTRACT_WORD, s.accWord, tracked ? s.nnWord : -1, tracked, narrow ? 4 : 8);
      }
    }
    VELOX_CHECK(plan.size() <= VB2_EXTRACT_MAX_COLS, "too many aggregation output columns");
    // arena: [header 64 B | slot scratch (small tables) | per column: values, validity]
    auto align = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    size_t off = 64;
    const size_t scratchOff = off;
    if (small) off = align(off + static_cast<size_t>(rowsCap) * 4, 256);
    for (auto& c : plan) {
      c.valuesOff = off;
      off = align(off + (c.width == 0 ? bits::nbytes(rowsCap) : static_cast<size_t>(rowsCap) * c.width), 256);
      if (c.hasValid) { c.validOff = off; off = align(off + bits::nbytes(rowsCap), 256); }
    }
    const size_t arenaBytes = off;
    auto arena = allocDevice(arenaBytes, st());
    uint8_t* base = arena->as<uint8_t>();
    VB2_CU(cudaMemsetAsync(base, 0, 64, st()));
    std::vector<vb2_extract_col> ecs;
    for (auto& c : plan) {
      c.ec.values = base + c.valuesOff;
      c.ec.valid = c.hasValid ? reinterpret_cast<uint64_t*>(base + c.validOff) : nullptr;
      ecs.push_back(c.ec);
    }
    const vb2_group_table tab = table();
    int64_t* header = reinterpret_cast<int64_t*>(base);
    if (mode == Mode::kGlobal) {
      // a global aggregation always emits one row (exec/GroupingSet.cpp:499-770): slot 0, listed explicitly
      kernelCheck(vb2k_group_extract(&tab, reinterpret_cast<const int32_t*>(base + 16), 1, nullptr, ecs.data(), static_cast<int32_t>(ecs.size()), header,
                                     errorFlag->as<int32_t>(), st()));
    } else if (small) {
      kernelCheck(vb2k_group_extract(&tab, nullptr, 0, reinterpret_cast<int32_t*>(base + scratchOff), ecs.data(), static_cast<int32_t>(ecs.size()), header,
                                     errorFlag->as<int32_t>(), st()));
    } else {
      kernelCheck(vb2k_group_extract(&tab, slots->as<int32_t>(), m, nullptr, ecs.data(), static_cast<int32_t>(ecs.size()), header, nullptr, st()));
    }
    std::shared_ptr<HostMirror> mirror;
    if (small) {
      mirror = std::make_shared<HostMirror>();
      mirror->host = acquirePinned(arenaBytes);
      mirror->devBase = base;
      mirror->bytes = arenaBytes;
      VB2_CU(cudaMemcpyAsync(mirror->host.get(), base, arenaBytes, cudaMemcpyDeviceToHost, st()));
      VB2_CU(cudaStreamSynchronize(st()));
      const int64_t* h = static_cast<const int64_t*>(mirror->host.get());
      if (h[1] != 0) checkDeviceError(errorFlag, st(), "sum");  // throws (SumAggregate.cpp:24 overflow)
      m = h[0];
      if (m == 0) return nullptr;
    }
    std::vector<DeviceColumnPtr> cols;
    for (auto& c : plan) {
      auto col = std::make_shared<DeviceColumn>();
      col->type = c.type;
      col->desc.size = m;
      col->owners = {arena};
      const uint64_t* valid = c.hasValid ? reinterpret_cast<const uint64_t*>(base + c.validOff) : nullptr;
      if (c.varcharKey) {
        KeyState& ks = keys[c.key];
        ensureDeviceAlphabet(ks);
        col->desc.type = VB2_VARCHAR;
        col->desc.encoding = VB2_DICTIONARY;
        col->desc.indices = reinterpret_cast<const int32_t*>(base + c.valuesOff);
        col->desc.values = ks.devOffsets->data();
        col->desc.aux = ks.devChars->data();
        col->desc.dict_size = static_cast<int64_t>(ks.valuesById.size());
        col->desc.nulls = valid;  // NULL key group: id 0
        col->owners.push_back(ks.devOffsets);
        col->owners.push_back(ks.devChars);
        col->alphabet = ks.outAlphabet;
      } else {
        col->desc.type = veloxTypeToVb2(c.type);
        col->desc.encoding = VB2_FLAT;
        col->desc.values = base + c.valuesOff;
        col->desc.nulls = valid;
      }
      cols.push_back(std::move(col));
    }
    reportFusedTimings();  // every fused launch precedes the synchronisation above
    self->addRuntimeStat("b200.fusedBatches", exec::RuntimeCounter{fusedBatches});
    self->addRuntimeStat("b200.selectiveBatches", exec::RuntimeCounter{selectiveBatches});
    self->addRuntimeStat("b200.partitionedBatches", exec::RuntimeCounter{partitionedBatches});
    self->addRuntimeStat("b200.genericBatches", exec::RuntimeCounter{genericBatches});
    self->addRuntimeStat("b200.aggMode", exec::RuntimeCounter{static_cast<int64_t>(mode)});
    auto out = std::make_shared<B200Vector>(self->pool(), outType, static_cast<vector_size_t>(m), std::move(cols), st());
    if (mirror) out->setMirror(mirror);
    return out;
  }
};

B200HashAggregation::B200HashAggregation(int32_t id, exec::DriverCtx* ctx, std::shared_ptr<const core::AggregationNode> node,
                                         std::vector<std::unique_ptr<exec::Operator>> absorbed)
    : Operator(ctx, node->outputType(), id, node->id(), "B200HashAggregation"), impl_(std::make_unique<Impl>()) {
  impl_->self = this;
  impl_->node = std::move(node);
  impl_->absorbed = std::move(absorbed);
}
B200HashAggregation::~B200HashAggregation() = default;

void B200HashAggregation::initialize() {
  Operator::initialize();
  impl_->dev = driverDeviceContext(driverCtx_);
  for (auto& op : impl_->absorbed) op->initialize();
  impl_->init();
}

exec::BlockingReason B200HashAggregation::isBlocked(exec::ContinueFuture* future) {
  for (auto& op : impl_->absorbed) {
    auto r = op->isBlocked(future);
    if (r != exec::BlockingReason::kNotBlocked) return r;
  }
  return exec::BlockingReason::kNotBlocked;
}

void B200HashAggregation::addInput(RowVectorPtr input) {
  B200_NVTX_OPERATOR_RANGE("addInput");
  auto in = std::dynamic_pointer_cast<B200Vector>(input);
  VELOX_CHECK(in != nullptr, "B200HashAggregation expects device-resident input");
  orderAfterProducer(*in, impl_->dev->stream);
  impl_->addInput(in);
}

void B200HashAggregation::noMoreInput() { Operator::noMoreInput(); }

RowVectorPtr B200HashAggregation::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!noMoreInput_ || finished_) return nullptr;
  finished_ = true;
  return impl_->output();
}

}  // namespace velox_b200
