#include "task.h"

namespace facebook::velox::exec {

namespace {

struct Pipeline {
  DriverFactory factory;
  std::unique_ptr<Driver> driver;
};

struct Planner {
  Task& task;
  std::map<int32_t, std::shared_ptr<std::vector<RowVectorPtr>>>& inputs;
  std::vector<std::unique_ptr<Pipeline>> pipelines;
  std::map<const core::PlanNode*, std::shared_ptr<HashJoinBridge>> bridges;
  std::map<const core::PlanNode*, std::shared_ptr<ExchangeQueue>> queues;

  // Collects the nodes of the pipeline ending at `node` (source first); every HashJoinNode's
  // build side becomes its own pipeline whose consumer is a HashBuild.
  void collect(const core::PlanNodePtr& node, std::vector<core::PlanNodePtr>& out) {
    if (auto ex = std::dynamic_pointer_cast<const core::ExchangeNode>(node)) {
      // fragment boundary: the producing fragment (… -> PartitionedOutput) becomes its own pipeline,
      // the Exchange is the source of the pipeline being collected
      auto queue = std::make_shared<ExchangeQueue>();
      queues[ex.get()] = queue;
      if (auto po = ex->upstream()) {
        auto prod = std::make_unique<Pipeline>();
        collect(po->sources()[0], prod->factory.planNodes);
        prod->factory.consumerSupplier = [po, queue](int32_t id, DriverCtx* ctx) -> std::unique_ptr<Operator> {
          return std::make_unique<PartitionedOutput>(id, ctx, po, queue);
        };
        prod->factory.pipelineId = static_cast<int32_t>(pipelines.size());
        pipelines.push_back(std::move(prod));
      }
      out.push_back(node);
      return;
    }
    if (auto join = std::dynamic_pointer_cast<const core::HashJoinNode>(node)) {
      auto bridge = std::make_shared<HashJoinBridge>();
      bridges[join.get()] = bridge;
      auto build = std::make_unique<Pipeline>();
      collect(join->sources()[1], build->factory.planNodes);
      build->factory.consumerSupplier = [join, bridge](int32_t id, DriverCtx* ctx) -> std::unique_ptr<Operator> {
        return std::make_unique<HashBuild>(id, ctx, join, bridge);
      };
      build->factory.pipelineId = static_cast<int32_t>(pipelines.size());
      pipelines.push_back(std::move(build));
      collect(join->sources()[0], out);
    } else if (!node->sources().empty()) {
      collect(node->sources()[0], out);
    }
    out.push_back(node);
  }

  std::unique_ptr<Driver> createDriver(const DriverFactory& f, std::function<void(RowVectorPtr)> sink,
                                       std::vector<std::unique_ptr<core::QueryConfig>>&) {
    auto ctx = std::make_unique<DriverCtx>();
    ctx->pipelineId = f.pipelineId;
    ctx->task = &task;
    ctx->config = &task.queryConfig();
    ctx->pool = task.pool();
    DriverCtx* c = ctx.get();
    auto driver = std::make_unique<Driver>(std::move(ctx));
    std::vector<std::unique_ptr<Operator>> ops;
    const auto& nodes = f.planNodes;
    for (size_t i = 0; i < nodes.size(); ++i) {
      const int32_t id = static_cast<int32_t>(ops.size());
      if (auto v = std::dynamic_pointer_cast<const core::ValuesNode>(nodes[i])) {
        auto it = inputs.find(v->sourceId());
        auto batches = it != inputs.end() ? it->second : std::make_shared<std::vector<RowVectorPtr>>();
        ops.push_back(std::make_unique<Values>(id, c, v, batches));
      } else if (auto fl = std::dynamic_pointer_cast<const core::FilterNode>(nodes[i])) {
        std::shared_ptr<const core::ProjectNode> pr;
        if (i + 1 < nodes.size()) pr = std::dynamic_pointer_cast<const core::ProjectNode>(nodes[i + 1]);
        if (pr) ++i;
        ops.push_back(std::make_unique<FilterProject>(id, c, fl, pr));
      } else if (auto pr = std::dynamic_pointer_cast<const core::ProjectNode>(nodes[i])) {
        ops.push_back(std::make_unique<FilterProject>(id, c, nullptr, pr));
      } else if (auto ag = std::dynamic_pointer_cast<const core::AggregationNode>(nodes[i])) {
        ops.push_back(std::make_unique<HashAggregation>(id, c, ag));
      } else if (auto jn = std::dynamic_pointer_cast<const core::HashJoinNode>(nodes[i])) {
        ops.push_back(std::make_unique<HashProbe>(id, c, jn, bridges.at(jn.get())));
      } else if (auto ex = std::dynamic_pointer_cast<const core::ExchangeNode>(nodes[i])) {
        ops.push_back(std::make_unique<Exchange>(id, c, ex, queues.at(ex.get())));
      } else {
        VELOX_UNSUPPORTED("plan node " + std::string(nodes[i]->name()));
      }
    }
    const int32_t sinkId = static_cast<int32_t>(ops.size());
    if (f.consumerSupplier) ops.push_back(f.consumerSupplier(sinkId, c));
    else ops.push_back(std::make_unique<CallbackSink>(sinkId, c, std::move(sink)));
    driver->init(std::move(ops));
    for (auto& adapter : DriverFactory::adapters())
      if (adapter.adapt) adapter.adapt(f, *driver);
    return driver;
  }
};

}  // namespace

Task::Task(core::PlanNodePtr plan, core::QueryConfig config) : plan_(std::move(plan)), config_(std::move(config)) {}
Task::~Task() = default;

void Task::addInput(int32_t sourceId, RowVectorPtr batch) {
  auto& q = inputs_[sourceId];
  if (!q) q = std::make_shared<std::vector<RowVectorPtr>>();
  if (batch && batch->size() > 0) q->push_back(std::move(batch));
}

std::vector<RowVectorPtr> Task::run() {
  Planner planner{*this, inputs_, {}, {}, {}};
  auto out = std::make_unique<Pipeline>();
  planner.collect(plan_, out->factory.planNodes);
  out->factory.outputDriver = true;
  out->factory.pipelineId = static_cast<int32_t>(planner.pipelines.size());
  planner.pipelines.push_back(std::move(out));
  const core::PlanFragment fragment{plan_};
  for (auto& adapter : DriverFactory::adapters())
    if (adapter.inspect) adapter.inspect(fragment);

  std::vector<RowVectorPtr> results;
  std::vector<std::unique_ptr<core::QueryConfig>> keep;
  for (auto& p : planner.pipelines)
    p->driver = planner.createDriver(p->factory, [&results](RowVectorPtr v) { results.push_back(std::move(v)); }, keep);

  // Serial scheduler: keep giving every unfinished driver a turn; a blocked driver is skipped.
  std::vector<bool> done(planner.pipelines.size(), false);
  size_t remaining = planner.pipelines.size();
  try {
    while (remaining > 0) {
      bool any = false;
      for (size_t i = 0; i < planner.pipelines.size(); ++i) {
        if (done[i]) continue;
        for (;;) {
          bool finished = false, progressed = false;
          const BlockingReason r = planner.pipelines[i]->driver->runOnce(&finished, &progressed);
          if (finished) { done[i] = true; --remaining; any = true; break; }
          if (r != BlockingReason::kNotBlocked) break;
          if (!progressed) break;
          any = true;
        }
      }
      VELOX_CHECK(any || remaining == 0, "task made no progress (deadlock between pipelines)");
    }
  } catch (...) {
    for (auto& p : planner.pipelines) p->driver->close();
    throw;
  }
  for (auto& p : planner.pipelines) {
    for (auto& op : p->driver->operators()) {
      const std::string prefix = std::to_string(p->factory.pipelineId) + "." + std::to_string(op->operatorId()) + "." + op->operatorType() + ".";
      stats_[prefix + "inputPositions"] = op->stats().inputPositions;
      stats_[prefix + "outputPositions"] = op->stats().outputPositions;
      stats_[prefix + "addInputWallNanos"] = op->stats().addInputWallNanos;
      stats_[prefix + "getOutputWallNanos"] = op->stats().getOutputWallNanos;
      stats_[prefix + "finishWallNanos"] = op->stats().finishWallNanos;
      for (auto& kv : op->stats().runtimeStats) stats_[prefix + kv.first] = kv.second;
    }
    p->driver->close();
  }
  return results;
}

}  // namespace facebook::velox::exec
