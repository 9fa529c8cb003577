// B200 operators behind the exec::Operator contract (velox/exec/Operator.h:232-342). Each class
// states the reference operator it replaces. Installation: registerB200() (adapter.cpp).
#pragma once
#include <unordered_map>

#include "../../../include/velox_b200.h"
#include "device.h"
#include "plan_resolve.h"
#include "expr_compiler.h"
#include "task.h"

namespace velox_b200 {

// Throws VeloxUserError for the first data error a kernel recorded (integer overflow, division
// by zero, failed cast). Synchronises the stream.
void checkDeviceError(const DeviceBufferPtr& flag, cudaStream_t stream, const char* where);
std::vector<vb2_column> describe(const B200Vector& v);
// Zero-copy selection of rows sel[0..n) of `col` (dictionary wrap; exec/OperatorUtils.cpp:380 wrapChild).
DeviceColumnPtr wrapColumn(const DeviceColumnPtr& col, const DeviceBufferPtr& sel, int64_t n, cudaStream_t stream);
// Dense flat copy of (possibly wrapped) rows: values + validity bitmap (null when no nulls possible).
struct FlatColumn {
  DeviceBufferPtr values, nulls;
  int32_t type;
};
FlatColumn flattenColumn(const DeviceColumnPtr& col, const int32_t* sel, int64_t n, cudaStream_t stream);

// Orders `stream` after the work that produced `batch` when that happened on another stream (a batch
// that crossed a LocalExchange, an exchange page, a sibling driver's build rows).
void orderAfterProducer(const B200Vector& batch, cudaStream_t stream);

// Host RowVector -> device batch (velox/experimental/cudf/exec/CudfConversion.h:32 CudfFromVelox).
class B200FromHost : public exec::Operator {
 public:
  B200FromHost(int32_t id, exec::DriverCtx* ctx, RowTypePtr type) : Operator(ctx, std::move(type), id, "b200-from-host", "B200FromHost") {}
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_ && !input_; }
  void addInput(RowVectorPtr input) override { input_ = std::move(input); }
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture*) override { return exec::BlockingReason::kNotBlocked; }
  bool isFinished() override { return noMoreInput_ && !input_; }

 private:
  std::shared_ptr<DeviceContext> dev_;
  std::vector<RowVectorPtr> inFlight_;  // host batches whose copies may still be running
};

// Device batch -> host RowVector (CudfConversion.h:67 CudfToVelox).
class B200ToHost : public exec::Operator {
 public:
  B200ToHost(int32_t id, exec::DriverCtx* ctx, RowTypePtr type) : Operator(ctx, std::move(type), id, "b200-to-host", "B200ToHost") {}
  bool needsInput() const override { return !noMoreInput_ && !input_; }
  void addInput(RowVectorPtr input) override { input_ = std::move(input); }
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture*) override { return exec::BlockingReason::kNotBlocked; }
  bool isFinished() override { return noMoreInput_ && !input_; }
};

// Replaces exec::FilterProject (velox/exec/FilterProject.cpp:200-259): filter, then project the
// surviving rows; identity columns are dictionary-wrapped over the selection, computed columns
// are written densely. One VM kernel for the filter, one for all projections.
class B200FilterProject : public exec::Operator {
 public:
  B200FilterProject(int32_t id, exec::DriverCtx* ctx, const exec::FilterProject& cpu);
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_ && !input_; }
  void addInput(RowVectorPtr input) override { input_ = std::move(input); }
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture*) override { return exec::BlockingReason::kNotBlocked; }
  bool isFinished() override { return noMoreInput_ && !input_; }
  // used by fused operators that absorb this one
  B200VectorPtr apply(const B200VectorPtr& in);
  const std::vector<core::TypedExprPtr>& exprs() const { return exprs_; }
  bool hasFilter() const { return hasFilter_; }
  const RowTypePtr& inputType() const { return inputType_; }

 private:
  std::vector<core::TypedExprPtr> exprs_;
  bool hasFilter_;
  RowTypePtr inputType_;
  CompiledProgram program_;
  std::shared_ptr<DeviceContext> dev_;
  DeviceBufferPtr errorFlag_;
  // The filter as a TMA-staged bitmap kernel of the fused-pipeline family (registered or NVRTC-
  // instantiated from the expression templates), taken for flat NULL-free batches; -1 = not expressible.
  FusedBinding fastFilter_;
  int fastFilterId_ = -1;
};

// ---- aggregation ----------------------------------------------------------------------------------
struct JoinTableHolder;  // join.cpp
// keyed join mode: a probe-side VARCHAR dictionary and its device LUT of build-side string ids (join.cpp)
struct KeyedLutCache {
  std::shared_ptr<const HostAlphabet> alphabet;
  DeviceBufferPtr lut;
};

// An aggregate function of the B200 engine, created through the reference's registry
// (exec::registerAggregateFunction / Aggregate::create, velox/exec/Aggregate.h:361-575). The device
// keeps one 8-byte accumulator word (+ a non-null counter) per aggregate inside the group row
// (vb2_group_table); a function describes itself as
//   result = finalFunction( FAMILY( inputFunction(x) ) )
// where FAMILY is one of the device accumulator families "sum" "avg" "count" "min" "max"
// (functions/lib/aggregates/{SumAggregateBase,AverageAggregateBase,SimpleNumericAggregate}.h,
// prestosql/aggregates/CountAggregate.cpp) and the two optional transforms are names of registered
// scalar functions (built-in or B200DeviceFunction) that B200HashAggregation evaluates with the
// expression engine: inputFunction on the raw input of the partial / single step, finalFunction on
// the value the final / single step extracts. SUM / AVG / COUNT / MIN / MAX themselves are
// registered this way (registerB200Aggregates).
class B200Aggregate : public exec::Aggregate {
 public:
  B200Aggregate(TypePtr resultType, std::string family, std::string inputFunction = "", std::string finalFunction = "")
      : Aggregate(std::move(resultType)), family_(std::move(family)), inputFunction_(std::move(inputFunction)), finalFunction_(std::move(finalFunction)) {}
  const std::string& family() const { return family_; }
  const std::string& inputFunction() const { return inputFunction_; }
  const std::string& finalFunction() const { return finalFunction_; }

 private:
  std::string family_, inputFunction_, finalFunction_;
};
void registerB200Aggregates();
// Registers `name` as finalFunction(family(inputFunction(x))); the C ABI's vb2_register_aggregate_function.
void registerB200Aggregate(const std::string& name, const std::string& family, const std::string& inputFunction, const std::string& finalFunction);
// Return type of a registered scalar function applied to `argType` (built-ins: comparison -> BOOLEAN, else the argument's type).
TypePtr scalarFunctionReturnType(const std::string& name, const TypePtr& argType);

// Replaces exec::HashAggregation / GroupingSet / HashTable(group by) / RowContainer / Aggregate
// (velox/exec/HashAggregation.cpp:191-430, GroupingSet.cpp:190-884). Optionally absorbs the
// operators feeding it (FilterProject [-> HashProbe -> FilterProject]); when the absorbed
// expressions match an ahead-of-time fused pipeline and the batch is null-free flat data, the
// whole chain runs as one fused scan kernel, otherwise batch-by-batch through the same operators'
// generic kernels. Both paths update the same group layout.
class B200HashAggregation : public exec::Operator {
 public:
  B200HashAggregation(int32_t id, exec::DriverCtx* ctx, std::shared_ptr<const core::AggregationNode> node,
                      std::vector<std::unique_ptr<exec::Operator>> absorbed);
  ~B200HashAggregation() override;
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override;
  void noMoreInput() override;
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture* future) override;
  bool isFinished() override { return finished_; }

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
  bool finished_ = false;
};

// Replaces exec::OrderBy (velox/exec/OrderBy.cpp:60-110, SortBuffer.cpp): collects the input on the
// device, sorts once at noMoreInput (vb2k_sort_order), emits the input columns wrapped over the order.
class B200OrderBy : public exec::Operator {
 public:
  B200OrderBy(int32_t id, exec::DriverCtx* ctx, std::shared_ptr<const core::OrderByNode> node);
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override;
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture*) override { return exec::BlockingReason::kNotBlocked; }
  bool isFinished() override { return finished_; }

 private:
  std::shared_ptr<const core::OrderByNode> node_;
  std::shared_ptr<DeviceContext> dev_;
  std::vector<int32_t> channels_;
  std::vector<B200VectorPtr> batches_;
  bool finished_ = false;
};

// Replaces exec::TopN (velox/exec/TopN.cpp:60-150): ORDER BY ... LIMIT count. Keeps at most `count`
// rows between folds (sort of kept + pending rows, first `count` survive) instead of the CPU's heap.
class B200TopN : public exec::Operator {
 public:
  B200TopN(int32_t id, exec::DriverCtx* ctx, std::shared_ptr<const core::TopNNode> node);
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override;
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture*) override { return exec::BlockingReason::kNotBlocked; }
  bool isFinished() override { return finished_; }

 private:
  void fold();
  std::shared_ptr<const core::TopNNode> node_;
  std::shared_ptr<DeviceContext> dev_;
  std::vector<int32_t> channels_;
  std::vector<B200VectorPtr> pending_;
  int64_t pendingRows_ = 0;
  B200VectorPtr top_;
  bool finished_ = false;
};

// Replaces exec::HashBuild (velox/exec/HashBuild.cpp:442-598,819-993): collects the build side on
// the device, builds the join table at noMoreInput and publishes it on the HashJoinBridge.
class B200HashBuild : public exec::Operator {
 public:
  B200HashBuild(int32_t id, exec::DriverCtx* ctx, const exec::HashBuild& cpu);
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override;
  void noMoreInput() override;
  RowVectorPtr getOutput() override { return nullptr; }
  // Sibling builds (N drivers on the build pipeline) wait here until the last one has taken their rows
  // (exec/HashBuild.cpp:819-870: State::kWaitForBuild until the last driver fulfils the promise).
  exec::BlockingReason isBlocked(exec::ContinueFuture* future) override {
    if (!peerFuture_.valid() || peerFuture_.isReady()) return exec::BlockingReason::kNotBlocked;
    if (future) *future = peerFuture_;
    return exec::BlockingReason::kWaitForJoinBuild;
  }
  bool isFinished() override { return noMoreInput_ && (!peerFuture_.valid() || peerFuture_.isReady()); }
  // the batches collected so far, handed to the last peer (called under the barrier: this driver is parked)
  std::vector<B200VectorPtr> takeBatches() { return std::move(batches_); }

 private:
  void buildTable();
  void buildKeyedTable(const std::shared_ptr<JoinTableHolder>& holder, const std::vector<int32_t>& keys, int64_t n);
  std::shared_ptr<const core::HashJoinNode> node_;
  ResolvedJoin plan_;  // key / output column names resolved to channels
  std::shared_ptr<exec::HashJoinBridge> bridge_;
  std::shared_ptr<DeviceContext> dev_;
  std::vector<B200VectorPtr> batches_;
  exec::ContinueFuture peerFuture_;
};

// Replaces exec::HashProbe (velox/exec/HashProbe.cpp:796-900,1189-1437).
class B200HashProbe : public exec::Operator {
 public:
  B200HashProbe(int32_t id, exec::DriverCtx* ctx, const exec::HashProbe& cpu);
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_ && !input_; }
  void addInput(RowVectorPtr input) override { input_ = std::move(input); }
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture* future) override;
  bool isFinished() override { return noMoreInput_ && !input_; }
  B200VectorPtr apply(const B200VectorPtr& in);
  const std::shared_ptr<const core::HashJoinNode>& node() const { return node_; }
  std::shared_ptr<JoinTableHolder> table() const { return table_; }

 private:
  std::shared_ptr<const core::HashJoinNode> node_;
  ResolvedJoin plan_;  // key / output column names resolved to channels
  std::shared_ptr<exec::HashJoinBridge> bridge_;
  std::shared_ptr<JoinTableHolder> table_;
  std::shared_ptr<DeviceContext> dev_;
  std::unique_ptr<CompiledProgram> filterProgram_;
  DeviceBufferPtr errorFlag_;
  std::vector<KeyedLutCache> keyedLuts_;  // keyed mode: probe-side VARCHAR key dictionaries -> build-side string ids
};

// ---- exchange --------------------------------------------------------------------------------------
// The task's exchange transport over NCCL (one process per GPU; include/velox_b200.h vb2_comm_*).
class NcclTransport : public exec::ExchangeTransport {
 public:
  explicit NcclTransport(vb2_comm* comm) : comm_(comm) {}
  int world() const override { return vb2_comm_world(comm_); }
  int rank() const override { return vb2_comm_rank(comm_); }
  vb2_comm* comm() const { return comm_; }

 private:
  vb2_comm* comm_;
};

// Replaces exec::PartitionedOutput (velox/exec/PartitionedOutput.cpp; GPU pattern
// velox/experimental/ucx-exchange/UcxPartitionedOutput.h:29-110): collects the producing
// fragment's batches on the device, partitions the rows by VectorHasher-hash(keys) % partitions
// (bit exact with exec::HashPartitionFunction, exec/HashPartitionFunction.cpp:75-118), and moves
// every column of every partition to its rank in ONE grouped NCCL all-to-all at noMoreInput. The
// rows this rank receives are handed to the B200Exchange of the consuming fragment through the
// ExchangeQueue. Pages are columnar ("B200Columnar" serde: flat fixed-width values, validity as
// bytes, VARCHAR as dictionary codes over an alphabet merged across ranks) — no row serialisation.
class B200PartitionedOutput : public exec::Operator {
 public:
  B200PartitionedOutput(int32_t id, exec::DriverCtx* ctx, const exec::PartitionedOutput& cpu);
  void initialize() override;
  bool needsInput() const override { return !noMoreInput_; }
  void addInput(RowVectorPtr input) override;
  void noMoreInput() override;
  RowVectorPtr getOutput() override { return nullptr; }
  exec::BlockingReason isBlocked(exec::ContinueFuture*) override { return exec::BlockingReason::kNotBlocked; }
  bool isFinished() override { return noMoreInput_; }

 private:
  std::shared_ptr<const core::PartitionedOutputNode> node_;
  std::vector<int32_t> keyChannels_;  // of the node's HashPartitionFunctionSpec
  std::shared_ptr<exec::ExchangeQueue> queue_;
  std::shared_ptr<DeviceContext> dev_;
  std::vector<B200VectorPtr> batches_;
};

// Replaces exec::Exchange (velox/exec/Exchange.h:51): source operator of the consuming fragment.
class B200Exchange : public exec::SourceOperator {
 public:
  B200Exchange(int32_t id, exec::DriverCtx* ctx, const exec::Exchange& cpu);
  void initialize() override;
  RowVectorPtr getOutput() override;
  exec::BlockingReason isBlocked(exec::ContinueFuture* future) override;
  bool isFinished() override { return atEnd_; }

 private:
  std::shared_ptr<exec::ExchangeQueue> queue_;
  std::shared_ptr<DeviceContext> dev_;
  RowVectorPtr next_;
  bool atEnd_ = false;
};

// Installs the DriverAdapter that swaps the CPU operators for the classes above and inserts
// B200FromHost / B200ToHost at the seams (pattern: velox/experimental/cudf/exec/ToCudf.cpp:71-338).
void registerB200();

}  // namespace velox_b200
