// Contract shim — the plan / expression / operator boundary of the reference, restated without
// its dependencies. The B200 operators are written against exactly these names; with real Velox
// headers (VELOX_B200_WITH_REAL_VELOX) the same operator sources bind to the real classes.
//
//   core::PlanNode family ....... velox/core/PlanNode.h:175 (FilterNode :671, ProjectNode :876,
//                                 AggregationNode :1120, HashJoinNode :3437, ValuesNode)
//   core::ITypedExpr family ..... velox/core/Expressions.h (FieldAccess / Constant / Call / Cast)
//   core::QueryConfig ........... velox/core/QueryConfig.h (string map with typed accessors)
//   exec::ExprSet ............... velox/expression/Expr.h:837
//   exec::VectorFunction + registry  velox/expression/VectorFunction.h:38,241,293
//   exec::Operator .............. velox/exec/Operator.h:120-702 (virtuals :232-342)
//   exec::Driver / DriverFactory / DriverAdapter  velox/exec/Driver.h:364,789-847;
//                                 loop order of Driver::runInternal velox/exec/Driver.cpp:538-850
//   exec::LocalPlanner .......... velox/exec/LocalPlanner.cpp:374-810 (adapters run :762-766)
//   exec::HashJoinBridge ........ velox/exec/HashJoinBridge.h:42-130
//   exec::Task .................. velox/exec/Task.h (single-threaded serial execution mode)
#pragma once
#ifndef VELOX_B200_WITH_REAL_VELOX

#include <atomic>
#include <functional>
#include <chrono>
#include <map>
#include <mutex>
#include <optional>
#include <set>
#include <unordered_map>
#include <variant>

#include "vector_abi.h"

namespace facebook::velox {

// A constant value (velox/type/Variant.h, the subset of kinds on the hot path).
struct Variant {
  TypeKind kind = TypeKind::UNKNOWN;
  bool isNull = true;
  std::variant<bool, int32_t, int64_t, double, std::string> value;
  static Variant null(TypeKind k) { Variant v; v.kind = k; return v; }
  template <class T> static Variant of(TypeKind k, T x) { Variant v; v.kind = k; v.isNull = false; v.value = std::move(x); return v; }
};

namespace core {

// ---- expressions ------------------------------------------------------------------------------
class ITypedExpr {
 public:
  ITypedExpr(TypePtr type, std::vector<std::shared_ptr<const ITypedExpr>> inputs) : type_(std::move(type)), inputs_(std::move(inputs)) {}
  virtual ~ITypedExpr() = default;
  const TypePtr& type() const { return type_; }
  const std::vector<std::shared_ptr<const ITypedExpr>>& inputs() const { return inputs_; }
  virtual std::string toString() const = 0;

 private:
  TypePtr type_;
  std::vector<std::shared_ptr<const ITypedExpr>> inputs_;
};
using TypedExprPtr = std::shared_ptr<const ITypedExpr>;

// velox/core/Expressions.h FieldAccessTypedExpr: a column of the input row, by NAME (operators resolve
// names against their input type, exec/OperatorUtils.h exprToChannel).
class FieldAccessTypedExpr : public ITypedExpr {
 public:
  FieldAccessTypedExpr(TypePtr type, std::string name) : ITypedExpr(std::move(type), {}), name_(std::move(name)) {}
  const std::string& name() const { return name_; }
  bool isInputColumn() const { return true; }
  std::string toString() const override { return name_; }

 private:
  std::string name_;
};
using FieldAccessTypedExprPtr = std::shared_ptr<const FieldAccessTypedExpr>;
class ConstantTypedExpr : public ITypedExpr {
 public:
  ConstantTypedExpr(TypePtr type, Variant value) : ITypedExpr(std::move(type), {}), value_(std::move(value)) {}
  const Variant& value() const { return value_; }
  std::string toString() const override { return "const"; }

 private:
  Variant value_;
};
class CallTypedExpr : public ITypedExpr {
 public:
  CallTypedExpr(TypePtr type, std::vector<TypedExprPtr> inputs, std::string name) : ITypedExpr(std::move(type), std::move(inputs)), name_(std::move(name)) {}
  const std::string& name() const { return name_; }
  std::string toString() const override {
    std::string s = name_ + "(";
    for (size_t i = 0; i < inputs().size(); ++i) s += (i ? "," : "") + inputs()[i]->toString();
    return s + ")";
  }

 private:
  std::string name_;
};
using CallTypedExprPtr = std::shared_ptr<const CallTypedExpr>;
class CastTypedExpr : public ITypedExpr {
 public:
  CastTypedExpr(TypePtr type, TypedExprPtr input, bool nullOnFailure = false) : ITypedExpr(std::move(type), {std::move(input)}), nullOnFailure_(nullOnFailure) {}
  bool nullOnFailure() const { return nullOnFailure_; }
  std::string toString() const override { return "cast(" + inputs()[0]->toString() + " as " + type()->toString() + ")"; }

 private:
  bool nullOnFailure_;
};

// ---- plan nodes -------------------------------------------------------------------------------
using PlanNodeId = std::string;
class PlanNode {
 public:
  explicit PlanNode(PlanNodeId id) : id_(std::move(id)) {}
  virtual ~PlanNode() = default;
  const PlanNodeId& id() const { return id_; }
  virtual const RowTypePtr& outputType() const = 0;
  virtual const std::vector<std::shared_ptr<const PlanNode>>& sources() const = 0;
  virtual std::string_view name() const = 0;

 private:
  PlanNodeId id_;
};
using PlanNodePtr = std::shared_ptr<const PlanNode>;

class ValuesNode : public PlanNode {
 public:
  ValuesNode(PlanNodeId id, RowTypePtr type, int32_t sourceId) : PlanNode(std::move(id)), type_(std::move(type)), sourceId_(sourceId) {}
  const RowTypePtr& outputType() const override { return type_; }
  const std::vector<PlanNodePtr>& sources() const override { static const std::vector<PlanNodePtr> kEmpty; return kEmpty; }
  std::string_view name() const override { return "Values"; }
  int32_t sourceId() const { return sourceId_; }  // batches are supplied at run time (Task::addInput)

 private:
  RowTypePtr type_;
  int32_t sourceId_;
};
class FilterNode : public PlanNode {
 public:
  FilterNode(PlanNodeId id, TypedExprPtr filter, PlanNodePtr source) : PlanNode(std::move(id)), sources_{std::move(source)}, filter_(std::move(filter)) {}
  const RowTypePtr& outputType() const override { return sources_[0]->outputType(); }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  const TypedExprPtr& filter() const { return filter_; }
  std::string_view name() const override { return "Filter"; }

 private:
  std::vector<PlanNodePtr> sources_;
  TypedExprPtr filter_;
};
class ProjectNode : public PlanNode {
 public:
  ProjectNode(PlanNodeId id, std::vector<std::string> names, std::vector<TypedExprPtr> projections, PlanNodePtr source)
      : PlanNode(std::move(id)), sources_{std::move(source)}, names_(std::move(names)), projections_(std::move(projections)) {
    std::vector<TypePtr> types;
    for (auto& p : projections_) types.push_back(p->type());
    type_ = ROW(names_, types);
  }
  const RowTypePtr& outputType() const override { return type_; }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  const std::vector<std::string>& names() const { return names_; }
  const std::vector<TypedExprPtr>& projections() const { return projections_; }
  std::string_view name() const override { return "Project"; }

 private:
  std::vector<PlanNodePtr> sources_;
  std::vector<std::string> names_;
  std::vector<TypedExprPtr> projections_;
  RowTypePtr type_;
};
// velox/core/PlanNode.h:64-95 SortOrder
class SortOrder {
 public:
  SortOrder(bool ascending, bool nullsFirst) : ascending_(ascending), nullsFirst_(nullsFirst) {}
  bool isAscending() const { return ascending_; }
  bool isNullsFirst() const { return nullsFirst_; }
  bool operator==(const SortOrder& o) const { return ascending_ == o.ascending_ && nullsFirst_ == o.nullsFirst_; }
  std::string toString() const { return std::string(ascending_ ? "ASC" : "DESC") + (nullsFirst_ ? " NULLS FIRST" : " NULLS LAST"); }

 private:
  bool ascending_;
  bool nullsFirst_;
};
inline const SortOrder kAscNullsFirst{true, true};
inline const SortOrder kAscNullsLast{true, false};
inline const SortOrder kDescNullsFirst{false, true};
inline const SortOrder kDescNullsLast{false, false};

// velox/core/PlanNode.h:1120-1300
class AggregationNode : public PlanNode {
 public:
  enum class Step { kPartial, kFinal, kIntermediate, kSingle };
  // velox/core/PlanNode.h:1136-1158
  struct Aggregate {
    CallTypedExprPtr call;                 // function name and input columns (FieldAccessTypedExpr inputs)
    std::vector<TypePtr> rawInputTypes;    // raw argument types (differ from call's inputs for kIntermediate / kFinal)
    FieldAccessTypedExprPtr mask{};        // optional BOOLEAN mask column
    std::vector<FieldAccessTypedExprPtr> sortingKeys{};
    std::vector<SortOrder> sortingOrders{};
    bool distinct{false};
  };
  // velox/core/PlanNode.h:1165-1174: the output type is the grouping keys followed by aggregateNames typed by their calls
  AggregationNode(PlanNodeId id, Step step, std::vector<FieldAccessTypedExprPtr> groupingKeys, std::vector<FieldAccessTypedExprPtr> preGroupedKeys,
                  std::vector<std::string> aggregateNames, std::vector<Aggregate> aggregates, bool ignoreNullKeys, bool noGroupsSpanBatches,
                  PlanNodePtr source)
      : PlanNode(std::move(id)), step_(step), keys_(std::move(groupingKeys)), preGroupedKeys_(std::move(preGroupedKeys)),
        aggregateNames_(std::move(aggregateNames)), aggregates_(std::move(aggregates)), ignoreNullKeys_(ignoreNullKeys),
        noGroupsSpanBatches_(noGroupsSpanBatches), sources_{std::move(source)} {
    std::vector<std::string> names;
    std::vector<TypePtr> types;
    for (auto& k : keys_) { names.push_back(k->name()); types.push_back(k->type()); }
    for (size_t i = 0; i < aggregates_.size(); ++i) { names.push_back(aggregateNames_[i]); types.push_back(aggregates_[i].call->type()); }
    type_ = ROW(std::move(names), std::move(types));
  }
  Step step() const { return step_; }
  const std::vector<FieldAccessTypedExprPtr>& groupingKeys() const { return keys_; }
  const std::vector<FieldAccessTypedExprPtr>& preGroupedKeys() const { return preGroupedKeys_; }
  const std::vector<std::string>& aggregateNames() const { return aggregateNames_; }
  const std::vector<Aggregate>& aggregates() const { return aggregates_; }
  bool ignoreNullKeys() const { return ignoreNullKeys_; }
  bool noGroupsSpanBatches() const { return noGroupsSpanBatches_; }
  const RowTypePtr& outputType() const override { return type_; }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  std::string_view name() const override { return "Aggregation"; }
  bool isRawInput() const { return step_ == Step::kPartial || step_ == Step::kSingle; }
  bool isFinalOutput() const { return step_ == Step::kFinal || step_ == Step::kSingle; }
  // shim: replaces the derived output type (the flattened (sum, count) pair of an intermediate avg)
  void setOutputType(RowTypePtr t) { type_ = std::move(t); }

 private:
  Step step_;
  std::vector<FieldAccessTypedExprPtr> keys_, preGroupedKeys_;
  std::vector<std::string> aggregateNames_;
  std::vector<Aggregate> aggregates_;
  bool ignoreNullKeys_;
  bool noGroupsSpanBatches_;
  RowTypePtr type_;
  std::vector<PlanNodePtr> sources_;
};
enum class JoinType { kInner, kLeft, kLeftSemiFilter, kAnti };
// velox/core/PlanNode.h:3437-3470 (AbstractJoinNode :3330-3420): keys are columns by name, the output
// type names the columns taken from the left (probe) and the right (build) side.
class HashJoinNode : public PlanNode {
 public:
  HashJoinNode(PlanNodeId id, JoinType joinType, bool nullAware, std::vector<FieldAccessTypedExprPtr> leftKeys, std::vector<FieldAccessTypedExprPtr> rightKeys,
               TypedExprPtr filter, PlanNodePtr left, PlanNodePtr right, RowTypePtr outputType, bool useHashTableCache = false, bool nullAsValue = false,
               std::optional<std::string> cacheKey = std::nullopt)
      : PlanNode(std::move(id)), joinType_(joinType), nullAware_(nullAware), leftKeys_(std::move(leftKeys)), rightKeys_(std::move(rightKeys)),
        filter_(std::move(filter)), sources_{std::move(left), std::move(right)}, type_(std::move(outputType)), useHashTableCache_(useHashTableCache),
        nullAsValue_(nullAsValue), cacheKey_(std::move(cacheKey)) {}
  bool useHashTableCache() const { return useHashTableCache_; }
  bool nullAsValue() const { return nullAsValue_; }
  JoinType joinType() const { return joinType_; }
  bool isNullAware() const { return nullAware_; }
  bool isInnerJoin() const { return joinType_ == JoinType::kInner; }
  bool isLeftJoin() const { return joinType_ == JoinType::kLeft; }
  bool isLeftSemiFilterJoin() const { return joinType_ == JoinType::kLeftSemiFilter; }
  bool isAntiJoin() const { return joinType_ == JoinType::kAnti; }
  const std::vector<FieldAccessTypedExprPtr>& leftKeys() const { return leftKeys_; }    // probe side
  const std::vector<FieldAccessTypedExprPtr>& rightKeys() const { return rightKeys_; }  // build side
  const TypedExprPtr& filter() const { return filter_; }
  const RowTypePtr& outputType() const override { return type_; }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  std::string_view name() const override { return "HashJoin"; }

 private:
  JoinType joinType_;
  bool nullAware_;
  std::vector<FieldAccessTypedExprPtr> leftKeys_, rightKeys_;
  TypedExprPtr filter_;
  std::vector<PlanNodePtr> sources_;
  RowTypePtr type_;
  bool useHashTableCache_;
  bool nullAsValue_;
  std::optional<std::string> cacheKey_;
};

// velox/core/PlanNode.h:2600-2640 PartitionFunctionSpec: how rows map to partitions. The B200
// PartitionedOutput understands the hash spec (exec/HashPartitionFunction.h:75
// HashPartitionFunctionSpec: hash of the key channels modulo the partition count) and nullptr
// (gather / broadcast need no function).
class PartitionFunctionSpec {
 public:
  virtual ~PartitionFunctionSpec() = default;
  virtual std::string toString() const = 0;
};
using PartitionFunctionSpecPtr = std::shared_ptr<const PartitionFunctionSpec>;

// velox/core/PlanNode.h:2712 PartitionedOutputNode — root of a producing plan fragment: rows are
// partitioned by hash(keys) % numPartitions (kPartitioned; HashPartitionFunction,
// velox/exec/HashPartitionFunction.cpp:75-118), replicated to every consumer (kBroadcast) or sent
// anywhere (kArbitrary). numPartitions == 0 in the shim means "one partition per rank of the task's
// exchange transport", resolved when the operator is created.
class PartitionedOutputNode : public PlanNode {
 public:
  enum class Kind { kPartitioned, kBroadcast, kArbitrary };
  PartitionedOutputNode(PlanNodeId id, Kind kind, std::vector<TypedExprPtr> keys, int numPartitions, bool replicateNullsAndAny,
                        PartitionFunctionSpecPtr partitionFunctionSpec, RowTypePtr outputType, std::string serdeKind, PlanNodePtr source)
      : PlanNode(std::move(id)), kind_(kind), keys_(std::move(keys)), numPartitions_(numPartitions), replicateNullsAndAny_(replicateNullsAndAny),
        partitionFunctionSpec_(std::move(partitionFunctionSpec)), type_(std::move(outputType)), serdeKind_(std::move(serdeKind)), sources_{std::move(source)} {}
  const PartitionFunctionSpecPtr& partitionFunctionSpecPtr() const { return partitionFunctionSpec_; }
  Kind kind() const { return kind_; }
  bool isBroadcast() const { return kind_ == Kind::kBroadcast; }
  const std::vector<TypedExprPtr>& keys() const { return keys_; }
  int numPartitions() const { return numPartitions_; }
  bool isReplicateNullsAndAny() const { return replicateNullsAndAny_; }
  const std::string& serdeKind() const { return serdeKind_; }
  const RowTypePtr& outputType() const override { return type_; }
  const RowTypePtr& inputType() const { return sources_[0]->outputType(); }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  std::string_view name() const override { return "PartitionedOutput"; }

 private:
  Kind kind_;
  std::vector<TypedExprPtr> keys_;
  int numPartitions_;
  bool replicateNullsAndAny_;
  PartitionFunctionSpecPtr partitionFunctionSpec_;
  RowTypePtr type_;
  std::string serdeKind_;
  std::vector<PlanNodePtr> sources_;
};
// velox/core/PlanNode.h:2182 ExchangeNode — leaf of a consuming fragment. In the reference the
// producing fragment is another Task reached through remote splits; one process per GPU runs every
// fragment of the (SPMD) plan inside one Task, so the shim links the producer directly
// (`upstream()`, shim only) and the planner turns the pair into two pipelines joined by an
// ExchangeQueue.
class ExchangeNode : public PlanNode {
 public:
  ExchangeNode(PlanNodeId id, RowTypePtr type, std::string serdeKind) : PlanNode(std::move(id)), type_(std::move(type)), serdeKind_(std::move(serdeKind)) {}
  const RowTypePtr& outputType() const override { return type_; }
  const std::vector<PlanNodePtr>& sources() const override { static const std::vector<PlanNodePtr> kEmpty; return kEmpty; }
  std::string_view name() const override { return "Exchange"; }
  const std::string& serdeKind() const { return serdeKind_; }
  void setUpstream(std::shared_ptr<const PartitionedOutputNode> producer) { upstream_ = std::move(producer); }
  const std::shared_ptr<const PartitionedOutputNode>& upstream() const { return upstream_; }

 private:
  RowTypePtr type_;
  std::string serdeKind_;
  std::shared_ptr<const PartitionedOutputNode> upstream_;
};

// velox/core/PlanNode.h:4221-4330 OrderByNode; :4587-4680 TopNNode
class OrderByNode : public PlanNode {
 public:
  OrderByNode(const PlanNodeId& id, const std::vector<FieldAccessTypedExprPtr>& sortingKeys, const std::vector<SortOrder>& sortingOrders,
              bool isPartial, const PlanNodePtr& source)
      : PlanNode(id), sortingKeys_(sortingKeys), sortingOrders_(sortingOrders), isPartial_(isPartial), sources_{source} {
    if (sortingKeys.empty()) throw VeloxUserError("OrderBy must specify sorting keys");
    if (sortingKeys.size() != sortingOrders.size()) throw VeloxUserError("Number of sorting keys and sorting orders in OrderBy must be the same");
    std::set<std::string> unique;
    for (auto& k : sortingKeys)
      if (!k || !unique.insert(k->name()).second) throw VeloxUserError("Duplicate sorting keys are not allowed");
  }
  const std::vector<FieldAccessTypedExprPtr>& sortingKeys() const { return sortingKeys_; }
  const std::vector<SortOrder>& sortingOrders() const { return sortingOrders_; }
  bool isPartial() const { return isPartial_; }
  const RowTypePtr& outputType() const override { return sources_[0]->outputType(); }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  std::string_view name() const override { return "OrderBy"; }

 private:
  const std::vector<FieldAccessTypedExprPtr> sortingKeys_;
  const std::vector<SortOrder> sortingOrders_;
  const bool isPartial_;
  const std::vector<PlanNodePtr> sources_;
};
class TopNNode : public PlanNode {
 public:
  TopNNode(const PlanNodeId& id, const std::vector<FieldAccessTypedExprPtr>& sortingKeys, const std::vector<SortOrder>& sortingOrders, int32_t count,
           bool isPartial, const PlanNodePtr& source)
      : PlanNode(id), sortingKeys_(sortingKeys), sortingOrders_(sortingOrders), count_(count), isPartial_(isPartial), sources_{source} {
    if (sortingKeys.empty()) throw VeloxUserError("TopN must specify sorting keys");
    if (sortingKeys.size() != sortingOrders.size()) throw VeloxUserError("Number of sorting keys and sorting orders in TopN must be the same");
    if (count <= 0) throw VeloxUserError("TopN must specify greater than zero number of rows to keep");
  }
  const std::vector<FieldAccessTypedExprPtr>& sortingKeys() const { return sortingKeys_; }
  const std::vector<SortOrder>& sortingOrders() const { return sortingOrders_; }
  int32_t count() const { return count_; }
  bool isPartial() const { return isPartial_; }
  const RowTypePtr& outputType() const override { return sources_[0]->outputType(); }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  std::string_view name() const override { return "TopN"; }

 private:
  const std::vector<FieldAccessTypedExprPtr> sortingKeys_;
  const std::vector<SortOrder> sortingOrders_;
  const int32_t count_;
  const bool isPartial_;
  const std::vector<PlanNodePtr> sources_;
};

// velox/core/PlanNode.h LocalPartitionNode (gather form): the boundary between a pipeline that may run
// on several drivers and the single-driver pipeline consuming their outputs.
class LocalPartitionNode : public PlanNode {
 public:
  enum class Type { kGather, kRepartition };
  LocalPartitionNode(const PlanNodeId& id, Type type, bool scaleWriter, PartitionFunctionSpecPtr partitionFunctionSpec, std::vector<PlanNodePtr> sources)
      : PlanNode(id), type_(type), scaleWriter_(scaleWriter), spec_(std::move(partitionFunctionSpec)), sources_(std::move(sources)) {
    if (sources_.empty()) throw VeloxUserError("Local repartitioning node requires at least one source");
  }
  Type type() const { return type_; }
  const RowTypePtr& outputType() const override { return sources_[0]->outputType(); }
  const std::vector<PlanNodePtr>& sources() const override { return sources_; }
  std::string_view name() const override { return "LocalPartition"; }

 private:
  const Type type_;
  const bool scaleWriter_;
  const PartitionFunctionSpecPtr spec_;
  const std::vector<PlanNodePtr> sources_;
};

// velox/core/PlanFragment.h:43 (the grouped-execution fields are not part of the path)
struct PlanFragment {
  std::shared_ptr<const PlanNode> planNode;
};

class QueryConfig {
 public:
  explicit QueryConfig(std::unordered_map<std::string, std::string> values = {}) : values_(std::move(values)) {}
  template <class T>
  T get(const std::string& key, T defaultValue) const {
    auto it = values_.find(key);
    if (it == values_.end()) return defaultValue;
    if constexpr (std::is_same_v<T, bool>) return it->second == "true" || it->second == "1";
    else if constexpr (std::is_integral_v<T>) return static_cast<T>(std::stoll(it->second));
    else return it->second;
  }
  // velox/core/QueryConfig.h:479-494
  int32_t preferredOutputBatchRows() const { return get<int32_t>("preferred_output_batch_rows", 1024); }
  // B200 keys, following the CudfConfig pattern (velox/experimental/cudf/CudfConfig.h:29-127)
  bool b200Enabled() const { return get<bool>("b200.enabled", true); }
  bool b200FusedPipelines() const { return get<bool>("b200.fused_pipelines", true); }
  int32_t b200DeviceId() const { return get<int32_t>("b200.device_id", -1); }
  // rows per find-or-insert pass of a hash-mode aggregation: the table grows between passes, so
  // it is sized by distinct groups + one pass, not by the whole input batch
  int32_t b200AggProbeChunkRows() const { return get<int32_t>("b200.agg_probe_chunk_rows", 1 << 25); }

 private:
  std::unordered_map<std::string, std::string> values_;
};

}  // namespace core

namespace common {
struct SpillConfig {};  // velox/common/base/SpillConfig.h: carried for signature parity, never used on the device
}  // namespace common
namespace wave {
// velox/exec/HashJoinBridge.h:25,63-65: the opaque table an accelerator backend hands from its build to its probe operator
struct HashTableHolder {
  virtual ~HashTableHolder() = default;
};
}  // namespace wave

namespace exec {

// ---- expression set -----------------------------------------------------------------------------
// Compiled expressions of an operator. The shim keeps the typed trees; the CPU interpreter of the
// reference is not part of this repo — B200 operators compile `exprs()` to a device program.
class ExprSet {
 public:
  explicit ExprSet(std::vector<core::TypedExprPtr> exprs) : exprs_(std::move(exprs)) {}
  const std::vector<core::TypedExprPtr>& exprs() const { return exprs_; }
  size_t size() const { return exprs_.size(); }

 private:
  std::vector<core::TypedExprPtr> exprs_;
};

// velox/expression/EvalCtx.h: the evaluation context a VectorFunction::apply receives. The device
// functions need only the pool of the vectors they hand back.
class EvalCtx {
 public:
  explicit EvalCtx(memory::MemoryPool* pool = nullptr) : pool_(pool) {}
  memory::MemoryPool* pool() const { return pool_; }

 private:
  memory::MemoryPool* pool_;
};

// ---- VectorFunction registry --------------------------------------------------------------------
struct FunctionSignature {
  std::string returnType;             // "boolean" "double" "bigint" "integer" "T"
  std::vector<std::string> argTypes;  // same vocabulary; "T" = any one type, all Ts equal
};
using FunctionSignaturePtr = std::shared_ptr<FunctionSignature>;
struct VectorFunctionMetadata {
  bool defaultNullBehavior = true;  // NULL in -> NULL out, function not called on those rows
  bool deterministic = true;
};
class VectorFunction {
 public:
  virtual ~VectorFunction() = default;
  // velox/expression/VectorFunction.h:81-86. `rows` selects the rows to compute; `result` may be
  // pre-allocated. Device implementations receive device-resident argument vectors.
  virtual void apply(const SelectivityVector& rows, std::vector<VectorPtr>& args, const TypePtr& outputType, EvalCtx& context,
                     VectorPtr& result) const = 0;
};
struct VectorFunctionEntry {
  std::vector<FunctionSignaturePtr> signatures;
  std::shared_ptr<VectorFunction> function;
  VectorFunctionMetadata metadata;
};
using VectorFunctionMap = std::unordered_map<std::string, VectorFunctionEntry>;
inline VectorFunctionMap& vectorFunctionFactories() {
  static VectorFunctionMap m;
  return m;
}
inline std::mutex& vectorFunctionMutex() {
  static std::mutex m;
  return m;
}
inline bool registerVectorFunction(std::string_view nameView, std::vector<FunctionSignaturePtr> signatures,
                                   std::unique_ptr<VectorFunction> func, VectorFunctionMetadata metadata = {}, bool overwrite = true) {
  const std::string name(nameView);
  std::lock_guard<std::mutex> l(vectorFunctionMutex());
  auto& m = vectorFunctionFactories();
  if (!overwrite && m.count(name)) return false;
  m[name] = VectorFunctionEntry{std::move(signatures), std::shared_ptr<VectorFunction>(std::move(func)), metadata};
  return true;
}
inline std::shared_ptr<VectorFunction> getVectorFunction(const std::string& name) {
  std::lock_guard<std::mutex> l(vectorFunctionMutex());
  auto& m = vectorFunctionFactories();
  auto it = m.find(name);
  return it == m.end() ? nullptr : it->second.function;
}

// ---- Aggregate registry (velox/exec/Aggregate.h:42-575) ------------------------------------------
// The CPU accumulator interface of the reference (addRawInput(char** groups, ...), extractValues ...)
// addresses host group rows and has no device counterpart; the shim carries what an accelerator
// needs from an Aggregate — its result type — and the registry it is created through.
class Aggregate {
 public:
  explicit Aggregate(TypePtr resultType) : resultType_(std::move(resultType)) {}
  virtual ~Aggregate() = default;
  const TypePtr& resultType() const { return resultType_; }
  // velox/exec/Aggregate.h:361-366
  static std::unique_ptr<Aggregate> create(const std::string& name, core::AggregationNode::Step step, const std::vector<TypePtr>& argTypes,
                                           const TypePtr& resultType, const core::QueryConfig& config);

 protected:
  const TypePtr resultType_;
};
struct AggregateFunctionSignature {  // velox/expression/FunctionSignature.h AggregateFunctionSignature
  std::string returnType, intermediateType;
  std::vector<std::string> argTypes;
};
using AggregateFunctionSignaturePtr = std::shared_ptr<AggregateFunctionSignature>;
using AggregateFunctionFactory = std::function<std::unique_ptr<Aggregate>(core::AggregationNode::Step step, const std::vector<TypePtr>& argTypes,
                                                                          const TypePtr& resultType, const core::QueryConfig& config)>;
struct AggregateFunctionMetadata {
  bool ignoreDuplicates{false};
  bool orderSensitive{true};
  bool companionFunction{false};
};
struct AggregateRegistrationResult {  // velox/exec/AggregateUtil.h:21
  bool mainFunction{false};
  bool partialFunction{false};
  bool mergeFunction{false};
  bool extractFunction{false};
  bool mergeExtractFunction{false};
};
struct AggregateFunctionEntry {
  std::vector<AggregateFunctionSignaturePtr> signatures;
  AggregateFunctionFactory factory;
  AggregateFunctionMetadata metadata;
};
using AggregateFunctionMap = std::unordered_map<std::string, AggregateFunctionEntry>;
inline AggregateFunctionMap& aggregateFunctions() {
  static AggregateFunctionMap m;
  return m;
}
inline std::mutex& aggregateFunctionMutex() {
  static std::mutex m;
  return m;
}
// velox/exec/Aggregate.h:537-542. Companion functions (name_partial / _merge / _extract) are scalar
// and aggregate wrappers of the CPU engine and are not generated here.
inline AggregateRegistrationResult registerAggregateFunction(const std::string& name, const std::vector<std::shared_ptr<AggregateFunctionSignature>>& signatures,
                                                             const AggregateFunctionFactory& factory, bool registerCompanionFunctions, bool overwrite) {
  (void)registerCompanionFunctions;
  std::lock_guard<std::mutex> l(aggregateFunctionMutex());
  auto& m = aggregateFunctions();
  AggregateRegistrationResult r;
  if (!overwrite && m.count(name)) return r;
  m[name] = AggregateFunctionEntry{signatures, factory, {}};
  r.mainFunction = true;
  return r;
}
inline const AggregateFunctionEntry* getAggregateFunctionEntry(const std::string& name) {
  std::lock_guard<std::mutex> l(aggregateFunctionMutex());
  auto& m = aggregateFunctions();
  auto it = m.find(name);
  return it == m.end() ? nullptr : &it->second;
}
inline std::unique_ptr<Aggregate> Aggregate::create(const std::string& name, core::AggregationNode::Step step, const std::vector<TypePtr>& argTypes,
                                                    const TypePtr& resultType, const core::QueryConfig& config) {
  AggregateFunctionFactory factory;
  {
    std::lock_guard<std::mutex> l(aggregateFunctionMutex());
    auto it = aggregateFunctions().find(name);
    if (it == aggregateFunctions().end()) throw VeloxUserError("Aggregate function not registered: " + name);
    factory = it->second.factory;
  }
  return factory(step, argTypes, resultType, config);
}

// ---- operators --------------------------------------------------------------------------------
enum class BlockingReason { kNotBlocked, kWaitForProducer, kWaitForJoinBuild, kWaitForConsumer };
// A future is a flag the unblocking side sets (folly::SemiFuture<Unit> in the reference); drivers may
// run on their own threads (Task::run with task.max_drivers > 1), so the flag is atomic. A blocked
// driver is re-polled by its thread.
struct ContinueFuture {
  std::shared_ptr<std::atomic<bool>> ready;
  bool valid() const { return ready != nullptr; }
  bool isReady() const { return ready && ready->load(std::memory_order_acquire); }
};
struct ContinuePromise {  // velox/common/future/VeloxPromise.h
  std::shared_ptr<std::atomic<bool>> flag = std::make_shared<std::atomic<bool>>(false);
  ContinueFuture getSemiFuture() const { return ContinueFuture{flag}; }
  void setValue() { flag->store(true, std::memory_order_release); }
};

class Task;
class Driver;
struct DriverCtx {  // velox/exec/Driver.h:263-300
  int32_t driverId = 0;
  int32_t pipelineId = 0;
  Task* task = nullptr;
  Driver* driver = nullptr;
  const core::QueryConfig* config = nullptr;
  memory::MemoryPool* pool = nullptr;
  const core::QueryConfig& queryConfig() const { return *config; }
};

struct RuntimeCounter { int64_t value = 0; };
struct OperatorStats {  // velox/exec/OperatorStats.h:93
  int32_t operatorId = 0;
  std::string operatorType;
  int64_t inputPositions = 0, inputVectors = 0, outputPositions = 0, outputVectors = 0;
  int64_t addInputWallNanos = 0, getOutputWallNanos = 0, finishWallNanos = 0;  // CpuWallTiming::wallNanos of each phase
  std::map<std::string, int64_t> runtimeStats;
};

using column_index_t = uint32_t;
// velox/exec/Operator.h:33-41
// velox/exec/HashPartitionFunction.h:76-101: partition = hash(key channels) % numPartitions
// (HashPartitionFunction.cpp:75-118); the B200 PartitionedOutput reads the key channels from here.
class HashPartitionFunctionSpec : public core::PartitionFunctionSpec {
 public:
  HashPartitionFunctionSpec(RowTypePtr inputType, std::vector<column_index_t> keyChannels, std::vector<VectorPtr> constValues = {})
      : inputType_(std::move(inputType)), keyChannels_(std::move(keyChannels)), constValues_(std::move(constValues)) {}
  const RowTypePtr& inputType() const { return inputType_; }
  const std::vector<column_index_t>& keyChannels() const { return keyChannels_; }  // shim accessor (private in the reference, read by create())
  const std::vector<VectorPtr>& constValues() const { return constValues_; }
  std::string toString() const override {
    std::string s = "HASH(";
    for (size_t i = 0; i < keyChannels_.size(); ++i) s += (i ? ", " : "") + inputType_->nameOf(keyChannels_[i]);
    return s + ")";
  }

 private:
  const RowTypePtr inputType_;
  const std::vector<column_index_t> keyChannels_;
  const std::vector<VectorPtr> constValues_;
};

struct IdentityProjection {
  IdentityProjection(column_index_t _inputChannel, column_index_t _outputChannel) : inputChannel(_inputChannel), outputChannel(_outputChannel) {}
  column_index_t inputChannel;
  column_index_t outputChannel;
};

class Operator {
 public:
  // velox/exec/Operator.h:216-222. Device operators never spill: spillConfig stays empty.
  Operator(DriverCtx* driverCtx, RowTypePtr outputType, int32_t operatorId, std::string planNodeId, std::string_view operatorType,
           std::optional<common::SpillConfig> spillConfig = std::nullopt)
      : driverCtx_(driverCtx), outputType_(std::move(outputType)), planNodeId_(std::move(planNodeId)), spillConfig_(std::move(spillConfig)) {
    stats_.operatorId = operatorId;
    stats_.operatorType = std::string(operatorType);
  }
  virtual ~Operator() = default;
  virtual void initialize() { initialized_ = true; }
  virtual bool needsInput() const = 0;
  virtual void addInput(RowVectorPtr input) = 0;
  virtual void noMoreInput() { noMoreInput_ = true; }
  virtual RowVectorPtr getOutput() = 0;
  virtual BlockingReason isBlocked(ContinueFuture* future) = 0;
  virtual bool isFinished() = 0;
  virtual void close() { input_ = nullptr; }
  virtual bool canReclaim() const { return false; }  // no spill on device
  virtual bool isSourceOperator() const { return false; }

  const RowTypePtr& outputType() const { return outputType_; }
  const std::string& planNodeId() const { return planNodeId_; }
  int32_t operatorId() const { return stats_.operatorId; }
  void setOperatorIdFromAdapter(int32_t id) { stats_.operatorId = id; }
  const std::string& operatorType() const { return stats_.operatorType; }
  OperatorStats& stats() { return stats_; }
  void addRuntimeStat(const std::string& name, const RuntimeCounter& c) { stats_.runtimeStats[name] += c.value; }
  memory::MemoryPool* pool() const { return driverCtx_->pool; }
  DriverCtx* driverCtx() const { return driverCtx_; }

 protected:
  DriverCtx* driverCtx_;
  RowTypePtr outputType_;
  std::string planNodeId_;
  OperatorStats stats_;
  std::optional<common::SpillConfig> spillConfig_;
  RowVectorPtr input_;
  bool noMoreInput_ = false;
  bool initialized_ = false;
};

class SourceOperator : public Operator {
 public:
  using Operator::Operator;
  bool needsInput() const override { return false; }
  void addInput(RowVectorPtr) override { VELOX_CHECK(false, "SourceOperator does not take input"); }
  bool isSourceOperator() const override { return true; }
};

// ---- join bridge --------------------------------------------------------------------------------
// One-shot hand-off of the finished build side from the build pipeline to the probe pipeline
// (velox/exec/HashJoinBridge.h:57 setHashTable, :116 tableOrFuture). The payload is opaque here;
// the B200 operators store their device table holder in it (the reference does the same for its
// Wave backend: setHashTable(shared_ptr<wave::HashTableHolder>) :63-65).
class JoinBridge {
 public:
  virtual ~JoinBridge() = default;
};
class HashJoinBridge : public JoinBridge {
 public:
  // velox/exec/HashJoinBridge.h:63-65 (the overload accelerator backends use) and :86-100,116
  void setHashTable(std::shared_ptr<wave::HashTableHolder> table, bool hasNullKeys) {
    std::lock_guard<std::mutex> l(mutex_);
    result_ = HashBuildResult{hasNullKeys, std::move(table)};
    for (auto& f : waiters_) f->store(true, std::memory_order_release);
    waiters_.clear();
  }
  struct HashBuildResult {
    bool hasNullKeys;
    std::shared_ptr<wave::HashTableHolder> waveTable;
  };
  std::optional<HashBuildResult> tableOrFuture(ContinueFuture* future) {
    std::lock_guard<std::mutex> l(mutex_);
    if (result_) return result_;
    future->ready = std::make_shared<std::atomic<bool>>(false);
    waiters_.push_back(future->ready);
    return std::nullopt;
  }

 private:
  std::mutex mutex_;
  std::optional<HashBuildResult> result_;
  std::vector<std::shared_ptr<std::atomic<bool>>> waiters_;
};

// ---- exchange hand-off ----------------------------------------------------------------------------
// What sits between a PartitionedOutput and the Exchange reading it inside one process: the role of
// OutputBufferManager (velox/exec/OutputBufferManager.h) + ExchangeClient / ExchangeQueue
// (velox/exec/ExchangeClient.h, ExchangeQueue.h). The producer deposits the pages this process
// receives (after the transport's all-to-all) and closes the queue; the consumer waits on a future.
class ExchangeQueue {
 public:
  void enqueue(RowVectorPtr page) {
    std::lock_guard<std::mutex> l(mutex_);
    pages_.push_back(std::move(page));
  }
  void noMoreData() {
    std::lock_guard<std::mutex> l(mutex_);
    done_ = true;
    for (auto& f : waiters_) f->store(true, std::memory_order_release);
    waiters_.clear();
  }
  // next page, or nullptr with *atEnd set; blocks (future) while the producer is still running
  RowVectorPtr dequeue(bool* atEnd, ContinueFuture* future) {
    std::lock_guard<std::mutex> l(mutex_);
    *atEnd = false;
    if (next_ < pages_.size()) return std::move(pages_[next_++]);
    if (done_) { *atEnd = true; return nullptr; }
    future->ready = std::make_shared<std::atomic<bool>>(false);
    waiters_.push_back(future->ready);
    return nullptr;
  }
  bool drained() const {
    std::lock_guard<std::mutex> l(mutex_);
    return done_ && next_ >= pages_.size();
  }

 private:
  mutable std::mutex mutex_;
  std::vector<RowVectorPtr> pages_;
  size_t next_ = 0;
  bool done_ = false;
  std::vector<std::shared_ptr<std::atomic<bool>>> waiters_;
};

// ---- local exchange ---------------------------------------------------------------------------------
// velox/exec/LocalPartition.h: LocalExchangeQueue between the N drivers of a producing pipeline and the
// consuming pipeline (gather: one queue). The reference's LocalPartition / LocalExchange operators only
// move RowVectorPtrs — device-resident batches pass through untouched — so they are real here.
class LocalExchangeQueue {
 public:
  explicit LocalExchangeQueue(int32_t producers) : producers_(producers) {}
  void enqueue(RowVectorPtr batch) {
    std::lock_guard<std::mutex> l(mutex_);
    queue_.push_back(std::move(batch));
  }
  void noMoreProducer() {
    std::lock_guard<std::mutex> l(mutex_);
    --producers_;
  }
  RowVectorPtr dequeue(bool* atEnd) {
    std::lock_guard<std::mutex> l(mutex_);
    *atEnd = false;
    if (next_ < queue_.size()) return std::move(queue_[next_++]);
    *atEnd = producers_ <= 0;
    return nullptr;
  }

 private:
  std::mutex mutex_;
  std::vector<RowVectorPtr> queue_;
  size_t next_ = 0;
  int32_t producers_;
};

// ---- CPU operators of the reference, as the LocalPlanner instantiates them ----------------------
// The reference's CPU implementations are not part of this repo (no CPU fallback): these classes
// carry the plan information and the accelerator hooks the adapter reads, and refuse to run.
class CpuOperatorStub : public Operator {
 public:
  using Operator::Operator;
  bool needsInput() const override { unavailable(); }
  void addInput(RowVectorPtr) override { unavailable(); }
  RowVectorPtr getOutput() override { unavailable(); }
  BlockingReason isBlocked(ContinueFuture*) override { return BlockingReason::kNotBlocked; }
  bool isFinished() override { unavailable(); }

 private:
  [[noreturn]] void unavailable() const {
    throw VeloxRuntimeError("CPU operator '" + stats_.operatorType +
                            "' is not available in this build: register the B200 adapter (registerB200())");
  }
};

class FilterProject : public CpuOperatorStub {
 public:
  FilterProject(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::FilterNode> filter,
                std::shared_ptr<const core::ProjectNode> project)
      : CpuOperatorStub(ctx, project ? project->outputType() : filter->outputType(), operatorId,
                        project ? project->id() : filter->id(), "FilterProject"),
        filter_(std::move(filter)), project_(std::move(project)) {
    std::vector<core::TypedExprPtr> all;
    if (filter_) all.push_back(filter_->filter());
    if (project_) for (auto& p : project_->projections()) all.push_back(p);
    exprs_ = std::make_unique<ExprSet>(std::move(all));
    // projections that are plain input columns pass through unevaluated (FilterProject::initialize, exec/FilterProject.cpp:90-130)
    if (project_) {
      const RowTypePtr& in = project_->sources()[0]->outputType();
      for (size_t i = 0; i < project_->projections().size(); ++i)
        if (auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(project_->projections()[i].get()))
          if (auto ch = in->getChildIdxIfExists(f->name())) resultProjections_.emplace_back(*ch, static_cast<column_index_t>(i));
    }
  }
  // velox/exec/FilterProject.h:70-78
  struct Export {
    const ExprSet* exprs;
    bool hasFilter;
    const std::vector<IdentityProjection>* resultProjections;
  };
  Export exprsAndProjection() const { return Export{exprs_.get(), filter_ != nullptr, &resultProjections_}; }
  const std::shared_ptr<const core::FilterNode>& filterNode() const { return filter_; }
  const std::shared_ptr<const core::ProjectNode>& projectNode() const { return project_; }
  RowTypePtr inputType() const { return (filter_ ? filter_->sources()[0] : project_->sources()[0])->outputType(); }

 private:
  std::shared_ptr<const core::FilterNode> filter_;
  std::shared_ptr<const core::ProjectNode> project_;
  std::unique_ptr<ExprSet> exprs_;
  std::vector<IdentityProjection> resultProjections_;
};
class HashAggregation : public CpuOperatorStub {
 public:
  HashAggregation(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::AggregationNode> node)
      : CpuOperatorStub(ctx, node->outputType(), operatorId, node->id(), "Aggregation"), node_(std::move(node)) {}
  const std::shared_ptr<const core::AggregationNode>& node() const { return node_; }

 private:
  std::shared_ptr<const core::AggregationNode> node_;
};
// velox/exec/OrderBy.h:34-39, velox/exec/TopN.h:23-28
class OrderBy : public CpuOperatorStub {
 public:
  OrderBy(int32_t operatorId, DriverCtx* driverCtx, const std::shared_ptr<const core::OrderByNode>& orderByNode)
      : CpuOperatorStub(driverCtx, orderByNode->outputType(), operatorId, orderByNode->id(), "OrderBy"), node_(orderByNode) {}
  const std::shared_ptr<const core::OrderByNode>& node() const { return node_; }

 private:
  std::shared_ptr<const core::OrderByNode> node_;
};
class TopN : public CpuOperatorStub {
 public:
  TopN(int32_t operatorId, DriverCtx* driverCtx, const std::shared_ptr<const core::TopNNode>& topNNode)
      : CpuOperatorStub(driverCtx, topNNode->outputType(), operatorId, topNNode->id(), "TopN"), node_(topNNode) {}
  const std::shared_ptr<const core::TopNNode>& node() const { return node_; }

 private:
  std::shared_ptr<const core::TopNNode> node_;
};
class HashBuild : public CpuOperatorStub {
 public:
  HashBuild(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::HashJoinNode> node, std::shared_ptr<HashJoinBridge> bridge)
      : CpuOperatorStub(ctx, nullptr, operatorId, node->id(), "HashBuild"), node_(std::move(node)), bridge_(std::move(bridge)) {}
  const std::shared_ptr<const core::HashJoinNode>& node() const { return node_; }
  const std::shared_ptr<HashJoinBridge>& joinBridge() const { return bridge_; }  // velox/exec/HashBuild.h:101-107

 private:
  std::shared_ptr<const core::HashJoinNode> node_;
  std::shared_ptr<HashJoinBridge> bridge_;
};
class HashProbe : public CpuOperatorStub {
 public:
  HashProbe(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::HashJoinNode> node, std::shared_ptr<HashJoinBridge> bridge)
      : CpuOperatorStub(ctx, node->outputType(), operatorId, node->id(), "HashProbe"), node_(std::move(node)), bridge_(std::move(bridge)) {}
  const std::shared_ptr<const core::HashJoinNode>& node() const { return node_; }
  const std::shared_ptr<HashJoinBridge>& joinBridge() const { return bridge_; }

 private:
  std::shared_ptr<const core::HashJoinNode> node_;
  std::shared_ptr<HashJoinBridge> bridge_;
};

// velox/exec/PartitionedOutput.h:157 / velox/exec/Exchange.h:51 — carriers like the stubs above.
class PartitionedOutput : public CpuOperatorStub {
 public:
  PartitionedOutput(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::PartitionedOutputNode> node, std::shared_ptr<ExchangeQueue> queue)
      : CpuOperatorStub(ctx, node->outputType(), operatorId, node->id(), "PartitionedOutput"), node_(std::move(node)), queue_(std::move(queue)) {}
  const std::shared_ptr<const core::PartitionedOutputNode>& node() const { return node_; }
  const std::shared_ptr<ExchangeQueue>& queue() const { return queue_; }

 private:
  std::shared_ptr<const core::PartitionedOutputNode> node_;
  std::shared_ptr<ExchangeQueue> queue_;
};
class Exchange : public CpuOperatorStub {
 public:
  Exchange(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::ExchangeNode> node, std::shared_ptr<ExchangeQueue> queue)
      : CpuOperatorStub(ctx, node->outputType(), operatorId, node->id(), "Exchange"), node_(std::move(node)), queue_(std::move(queue)) {}
  const std::shared_ptr<const core::ExchangeNode>& node() const { return node_; }
  const std::shared_ptr<ExchangeQueue>& queue() const { return queue_; }
  bool isSourceOperator() const override { return true; }

 private:
  std::shared_ptr<const core::ExchangeNode> node_;
  std::shared_ptr<ExchangeQueue> queue_;
};

// Source fed with batches at run time (velox/exec/Values.h:21 holds its vectors in the plan node;
// here the Task owns the queue so multi-GB inputs are not copied into the plan).
class Values : public SourceOperator {
 public:
  // The Values operator stands for the scan: with N drivers on the pipeline its batches are dealt like
  // splits, batch i to driver i % N (a fixed deal keeps floating-point sums reproducible).
  Values(int32_t operatorId, DriverCtx* ctx, std::shared_ptr<const core::ValuesNode> node, std::shared_ptr<std::vector<RowVectorPtr>> batches,
         int32_t numDrivers = 1)
      : SourceOperator(ctx, node->outputType(), operatorId, node->id(), "Values"), batches_(std::move(batches)),
        next_(static_cast<size_t>(ctx->driverId)), stride_(static_cast<size_t>(numDrivers < 1 ? 1 : numDrivers)) {}
  RowVectorPtr getOutput() override {
    if (next_ >= batches_->size()) return nullptr;
    RowVectorPtr b = (*batches_)[next_];
    next_ += stride_;
    return b;
  }
  BlockingReason isBlocked(ContinueFuture*) override { return BlockingReason::kNotBlocked; }
  bool isFinished() override { return next_ >= batches_->size(); }

 private:
  std::shared_ptr<std::vector<RowVectorPtr>> batches_;
  size_t next_, stride_;
};

// velox/exec/LocalPartition.h: sink of the producing pipeline / source of the consuming one.
class LocalPartition : public Operator {
 public:
  LocalPartition(int32_t operatorId, DriverCtx* ctx, const std::shared_ptr<const core::LocalPartitionNode>& node, std::shared_ptr<LocalExchangeQueue> queue)
      : Operator(ctx, node->outputType(), operatorId, node->id(), "LocalPartition"), queue_(std::move(queue)) {}
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override { queue_->enqueue(std::move(input)); }
  void noMoreInput() override {
    if (!noMoreInput_) queue_->noMoreProducer();
    Operator::noMoreInput();
  }
  RowVectorPtr getOutput() override { return nullptr; }
  BlockingReason isBlocked(ContinueFuture*) override { return BlockingReason::kNotBlocked; }
  bool isFinished() override { return noMoreInput_; }

 private:
  std::shared_ptr<LocalExchangeQueue> queue_;
};
class LocalExchange : public SourceOperator {
 public:
  LocalExchange(int32_t operatorId, DriverCtx* ctx, const std::shared_ptr<const core::LocalPartitionNode>& node, std::shared_ptr<LocalExchangeQueue> queue)
      : SourceOperator(ctx, node->outputType(), operatorId, node->id(), "LocalExchange"), queue_(std::move(queue)) {}
  RowVectorPtr getOutput() override {
    if (atEnd_) return nullptr;
    if (pending_) return std::move(pending_);
    return queue_->dequeue(&atEnd_);
  }
  BlockingReason isBlocked(ContinueFuture*) override {
    if (atEnd_ || pending_) return BlockingReason::kNotBlocked;
    pending_ = queue_->dequeue(&atEnd_);
    return (pending_ || atEnd_) ? BlockingReason::kNotBlocked : BlockingReason::kWaitForProducer;
  }
  bool isFinished() override { return atEnd_ && !pending_; }

 private:
  std::shared_ptr<LocalExchangeQueue> queue_;
  RowVectorPtr pending_;
  bool atEnd_ = false;
};

// Sink collecting the task's results (velox/exec/CallbackSink.h).
class CallbackSink : public Operator {
 public:
  CallbackSink(int32_t operatorId, DriverCtx* ctx, std::function<void(RowVectorPtr)> consumer)
      : Operator(ctx, nullptr, operatorId, "sink", "CallbackSink"), consumer_(std::move(consumer)) {}
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override { consumer_(std::move(input)); }
  RowVectorPtr getOutput() override { return nullptr; }
  BlockingReason isBlocked(ContinueFuture*) override { return BlockingReason::kNotBlocked; }
  bool isFinished() override { return noMoreInput_; }

 private:
  std::function<void(RowVectorPtr)> consumer_;
};

// ---- driver -------------------------------------------------------------------------------------
class Driver;
struct DriverFactory;
using AdaptDriverFunction = std::function<bool(const DriverFactory& factory, Driver& driver)>;  // velox/exec/Driver.h:786
struct DriverAdapter {  // velox/exec/Driver.h:789-793
  std::string label;
  std::function<void(const core::PlanFragment&)> inspect;  // sees the whole plan once, before the drivers exist (LocalPlanner.cpp:381)
  AdaptDriverFunction adapt;
};

struct DriverFactory {  // velox/exec/Driver.h:795-847
  std::vector<core::PlanNodePtr> planNodes;
  std::function<std::unique_ptr<Operator>(int32_t operatorId, DriverCtx* ctx)> consumerSupplier;
  bool inputDriver = false, outputDriver = false;
  int32_t pipelineId = 0;
  // Replaces operators [begin, end) of `driver` (velox/exec/LocalPlanner.cpp:772-810).
  std::vector<std::unique_ptr<Operator>> replaceOperators(Driver& driver, int32_t begin, int32_t end,
                                                          std::vector<std::unique_ptr<Operator>> replaceWith) const;
  static std::vector<DriverAdapter>& adapters() {
    static std::vector<DriverAdapter> a;
    return a;
  }
  static void registerAdapter(DriverAdapter adapter) { adapters().push_back(std::move(adapter)); }
};

class Driver {
 public:
  explicit Driver(std::unique_ptr<DriverCtx> ctx) : ctx_(std::move(ctx)) {}
  DriverCtx* driverCtx() const { return ctx_.get(); }
  std::vector<std::unique_ptr<Operator>>& operators() { return operators_; }
  const std::vector<std::unique_ptr<Operator>>& operators() const { return operators_; }
  void init(std::vector<std::unique_ptr<Operator>> ops) { operators_ = std::move(ops); }
  void initializeOperators() {
    if (initialized_) return;
    initialized_ = true;
    for (auto& op : operators_) op->initialize();
  }
  // One pass of the loop of Driver::runInternal (velox/exec/Driver.cpp:598-835): walks from the
  // sink toward the source, moving at most one batch per operator pair. Returns kNotBlocked with
  // *finished set when the pipeline has drained, or the reason it cannot progress.
  BlockingReason runOnce(bool* finished, bool* progressed);
  void close() {
    for (auto& op : operators_) op->close();
  }
  // velox/exec/Driver.h findOperator(planNodeId): how the last HashBuild peer reaches its siblings
  Operator* findOperator(std::string_view planNodeId) const {
    for (auto& op : operators_)
      if (op->planNodeId() == planNodeId) return op.get();
    return nullptr;
  }

 private:
  std::unique_ptr<DriverCtx> ctx_;
  std::vector<std::unique_ptr<Operator>> operators_;
  std::set<Operator*> signalled_;  // consumers already told noMoreInput
  bool initialized_ = false;
};

inline std::vector<std::unique_ptr<Operator>> DriverFactory::replaceOperators(Driver& driver, int32_t begin, int32_t end,
                                                                              std::vector<std::unique_ptr<Operator>> replaceWith) const {
  auto& ops = driver.operators();
  VELOX_CHECK(begin >= 0 && end <= static_cast<int32_t>(ops.size()) && begin <= end, "replaceOperators: bad range");
  std::vector<std::unique_ptr<Operator>> replaced;
  for (int32_t i = begin; i < end; ++i) replaced.push_back(std::move(ops[i]));
  ops.erase(ops.begin() + begin, ops.begin() + end);
  ops.insert(ops.begin() + begin, std::make_move_iterator(replaceWith.begin()), std::make_move_iterator(replaceWith.end()));
  for (size_t i = 0; i < ops.size(); ++i) ops[i]->setOperatorIdFromAdapter(static_cast<int32_t>(i));
  return replaced;
}

// Optional hook run before each timing lap (a device synchronize when VB2_SYNC_TIMING=1, so that
// the wall-time stats attribute asynchronous kernel time to the operator that launched it).
inline void (*g_timingSync)() = nullptr;

inline BlockingReason Driver::runOnce(bool* finished, bool* progressed) {
  initializeOperators();
  *finished = false;
  *progressed = false;
  const int32_t n = static_cast<int32_t>(operators_.size());
  ContinueFuture future;
  for (int32_t i = n - 1; i >= 0; --i) {
    Operator* op = operators_[i].get();
    BlockingReason r = op->isBlocked(&future);
    if (r != BlockingReason::kNotBlocked) return r;
    if (i == n - 1) {
      if (op->isFinished()) {
        *finished = true;
        return BlockingReason::kNotBlocked;
      }
      continue;
    }
    Operator* next = operators_[i + 1].get();
    if (!next->needsInput()) continue;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&t0]() {
      if (g_timingSync) g_timingSync();
      auto t1 = std::chrono::steady_clock::now();
      int64_t ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
      t0 = t1;
      return ns;
    };
    RowVectorPtr out = op->getOutput();
    op->stats().getOutputWallNanos += lap();
    if (out) {
      VELOX_CHECK(out->size() > 0, "operators must not emit empty vectors");
      op->stats().outputPositions += out->size();
      op->stats().outputVectors += 1;
      next->stats().inputPositions += out->size();
      next->stats().inputVectors += 1;
      next->addInput(std::move(out));
      next->stats().addInputWallNanos += lap();
      *progressed = true;
      return BlockingReason::kNotBlocked;  // restart from the sink, as runInternal does
    }
    if (op->isFinished() && !signalled_.count(next)) {
      next->noMoreInput();
      next->stats().finishWallNanos += lap();
      signalled_.insert(next);
      *progressed = true;
      return BlockingReason::kNotBlocked;
    }
  }
  return BlockingReason::kNotBlocked;
}

}  // namespace exec
}  // namespace facebook::velox

#endif  // VELOX_B200_WITH_REAL_VELOX
