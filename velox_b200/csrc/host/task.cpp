#include "task.h"

#include <atomic>
#include <chrono>
#include <exception>
#include <thread>

namespace facebook::velox::exec {

namespace {

struct Pipeline {
  DriverFactory factory;
  std::vector<std::shared_ptr<Driver>> drivers;
  int32_t numDrivers = 1;
  std::shared_ptr<LocalExchangeQueue> localSink;  // set when the pipeline ends in a LocalPartition
  std::shared_ptr<const core::LocalPartitionNode> localSinkNode;
};

// How many drivers a plan node tolerates on its pipeline (PlanNodeTranslator::maxDrivers and the
// per-node limits of velox/exec/LocalPlanner.cpp:283-340): final / single aggregations and ORDER BY
// gather everything in one operator, exchanges talk to a communicator that is not shared between threads.
bool allowsManyDrivers(const core::PlanNode& node) {
  if (auto agg = dynamic_cast<const core::AggregationNode*>(&node)) return agg->step() == core::AggregationNode::Step::kPartial;
  if (auto ob = dynamic_cast<const core::OrderByNode*>(&node)) return ob->isPartial();
  if (auto tn = dynamic_cast<const core::TopNNode*>(&node)) return tn->isPartial();
  if (dynamic_cast<const core::ExchangeNode*>(&node) || dynamic_cast<const core::PartitionedOutputNode*>(&node)) return false;
  return true;
}

struct Planner {
  Task& task;
  std::map<int32_t, std::shared_ptr<std::vector<RowVectorPtr>>>& inputs;
  int32_t maxDrivers;
  std::vector<std::unique_ptr<Pipeline>> pipelines;
  std::map<const core::PlanNode*, std::shared_ptr<HashJoinBridge>> bridges;
  std::map<const core::PlanNode*, std::shared_ptr<ExchangeQueue>> queues;
  std::map<const core::PlanNode*, std::shared_ptr<LocalExchangeQueue>> localQueues;

  void addPipeline(std::unique_ptr<Pipeline> p) {
    p->factory.pipelineId = static_cast<int32_t>(pipelines.size());
    pipelines.push_back(std::move(p));
  }

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
        prod->factory.planNodes.push_back(po);  // counted for the driver limit, not instantiated (the consumer is)
        addPipeline(std::move(prod));
      }
      out.push_back(node);
      return;
    }
    if (auto lp = std::dynamic_pointer_cast<const core::LocalPartitionNode>(node)) {
      if (maxDrivers <= 1) {
        // serial execution mode: one driver runs producer and consumer back to back, the gather is the identity
        collect(lp->sources()[0], out);
        return;
      }
      auto prod = std::make_unique<Pipeline>();
      collect(lp->sources()[0], prod->factory.planNodes);
      prod->localSinkNode = lp;
      addPipeline(std::move(prod));
      out.push_back(node);  // LocalExchange: source of the consuming pipeline
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
      addPipeline(std::move(build));
      collect(join->sources()[0], out);
    } else if (!node->sources().empty()) {
      collect(node->sources()[0], out);
    }
    out.push_back(node);
  }

  // Driver counts: the task's maxDrivers where every node of the pipeline allows it. The output
  // pipeline (CallbackSink) keeps one driver: results are delivered in one order.
  void decideDriverCounts() {
    for (auto& p : pipelines) {
      bool many = maxDrivers > 1 && !p->factory.outputDriver;
      for (auto& n : p->factory.planNodes) many = many && allowsManyDrivers(*n);
      // a pipeline reading a LocalExchange has a single queue: one consumer
      for (auto& n : p->factory.planNodes)
        if (dynamic_cast<const core::LocalPartitionNode*>(n.get())) many = false;
      p->numDrivers = many ? maxDrivers : 1;
    }
    for (auto& p : pipelines)
      if (p->localSinkNode) {
        p->localSink = std::make_shared<LocalExchangeQueue>(p->numDrivers);
        localQueues[p->localSinkNode.get()] = p->localSink;
        auto node = p->localSinkNode;
        auto queue = p->localSink;
        p->factory.consumerSupplier = [node, queue](int32_t id, DriverCtx* ctx) -> std::unique_ptr<Operator> {
          return std::make_unique<LocalPartition>(id, ctx, node, queue);
        };
      }
  }

  std::shared_ptr<Driver> createDriver(const Pipeline& p, int32_t driverId, std::function<void(RowVectorPtr)> sink) {
    const DriverFactory& f = p.factory;
    auto ctx = std::make_unique<DriverCtx>();
    ctx->pipelineId = f.pipelineId;
    ctx->driverId = driverId;
    ctx->task = &task;
    ctx->config = &task.queryConfig();
    ctx->pool = task.pool();
    DriverCtx* c = ctx.get();
    auto driver = std::make_shared<Driver>(std::move(ctx));
    c->driver = driver.get();
    std::vector<std::unique_ptr<Operator>> ops;
    const auto& nodes = f.planNodes;
    for (size_t i = 0; i < nodes.size(); ++i) {
      const int32_t id = static_cast<int32_t>(ops.size());
      if (auto v = std::dynamic_pointer_cast<const core::ValuesNode>(nodes[i])) {
        auto it = inputs.find(v->sourceId());
        auto batches = it != inputs.end() ? it->second : std::make_shared<std::vector<RowVectorPtr>>();
        ops.push_back(std::make_unique<Values>(id, c, v, batches, p.numDrivers));
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
      } else if (auto lp = std::dynamic_pointer_cast<const core::LocalPartitionNode>(nodes[i])) {
        ops.push_back(std::make_unique<LocalExchange>(id, c, lp, localQueues.at(lp.get())));
      } else if (auto ob = std::dynamic_pointer_cast<const core::OrderByNode>(nodes[i])) {
        ops.push_back(std::make_unique<OrderBy>(id, c, ob));
      } else if (auto tn = std::dynamic_pointer_cast<const core::TopNNode>(nodes[i])) {
        ops.push_back(std::make_unique<TopN>(id, c, tn));
      } else if (std::dynamic_pointer_cast<const core::PartitionedOutputNode>(nodes[i])) {
        continue;  // instantiated by the consumer supplier below
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

bool Task::allPeersFinished(const core::PlanNodeId& planNodeId, Driver* caller, ContinueFuture* future, std::vector<ContinuePromise>& promises,
                            std::vector<std::shared_ptr<Driver>>& peers) {
  std::lock_guard<std::mutex> l(mutex_);
  auto& state = barriers_[planNodeId];
  const int32_t numPeers = numDrivers(caller->driverCtx()->pipelineId);
  if (++state.numRequested == numPeers) {
    peers = std::move(state.drivers);
    promises = std::move(state.allPeersFinishedPromises);
    barriers_.erase(planNodeId);
    return true;
  }
  std::shared_ptr<Driver> callerShared;
  for (auto& d : drivers_)
    if (d.get() == caller) { callerShared = d; break; }
  VELOX_CHECK(callerShared != nullptr, "Caller of Task::allPeersFinished is not a valid Driver");
  if (future != nullptr) {
    state.drivers.push_back(callerShared);
    state.allPeersFinishedPromises.emplace_back();
    *future = state.allPeersFinishedPromises.back().getSemiFuture();
  }
  return false;
}

std::vector<RowVectorPtr> Task::run() {
  const auto t0 = std::chrono::steady_clock::now();
  auto sinceStart = [&t0]() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); };
  const int32_t maxDrivers = std::max<int32_t>(1, config_.get<int32_t>("task.max_drivers", 1));
  Planner planner{*this, inputs_, maxDrivers, {}, {}, {}, {}};
  auto out = std::make_unique<Pipeline>();
  planner.collect(plan_, out->factory.planNodes);
  out->factory.outputDriver = true;
  planner.addPipeline(std::move(out));
  planner.decideDriverCounts();
  const core::PlanFragment fragment{plan_};
  for (auto& adapter : DriverFactory::adapters())
    if (adapter.inspect) adapter.inspect(fragment);

  std::vector<RowVectorPtr> results;
  stats_.clear();
  driversPerPipeline_.clear();
  drivers_.clear();
  barriers_.clear();
  for (auto& p : planner.pipelines) driversPerPipeline_.push_back(p->numDrivers);
  for (auto& p : planner.pipelines)
    for (int32_t d = 0; d < p->numDrivers; ++d) {
      p->drivers.push_back(planner.createDriver(*p, d, [&results](RowVectorPtr v) { results.push_back(std::move(v)); }));
      drivers_.push_back(p->drivers.back());
    }

  const int64_t plannedNanos = sinceStart();  // LocalPlanner + driver / operator construction + adapters
  struct Slot {
    Driver* driver;
    bool done = false;
  };
  std::vector<Slot> slots;
  for (auto& d : drivers_) slots.push_back(Slot{d.get()});

  auto closeAll = [&] {
    for (auto& d : drivers_) d->close();
  };
  if (maxDrivers <= 1) {
    // Serial execution mode: keep giving every unfinished driver a turn; a blocked driver is skipped.
    size_t remaining = slots.size();
    try {
      while (remaining > 0) {
        bool any = false;
        for (auto& s : slots) {
          if (s.done) continue;
          for (;;) {
            bool finished = false, progressed = false;
            const BlockingReason r = s.driver->runOnce(&finished, &progressed);
            if (finished) { s.done = true; --remaining; any = true; break; }
            if (r != BlockingReason::kNotBlocked) break;
            if (!progressed) break;
            any = true;
          }
        }
        VELOX_CHECK(any || remaining == 0, "task made no progress (deadlock between pipelines)");
      }
    } catch (...) {
      closeAll();
      drivers_.clear();
      throw;
    }
  } else {
    // One thread per driver (the reference schedules drivers on an executor; blocked drivers come back
    // when their future fires — here the thread re-polls).
    std::atomic<bool> failed{false};
    std::exception_ptr error;
    std::mutex errorMutex;
    std::vector<std::thread> threads;
    for (auto& s : slots)
      threads.emplace_back([&, driver = s.driver] {
        if (threadBegin_) threadBegin_();
        try {
          int idle = 0;
          for (;;) {
            if (failed.load(std::memory_order_acquire)) break;
            bool finished = false, progressed = false;
            const BlockingReason r = driver->runOnce(&finished, &progressed);
            if (finished) break;
            if (r != BlockingReason::kNotBlocked || !progressed) {
              // blocked on a peer / producer / join build: yield, then back off to 20 us naps
              if (++idle < 64) std::this_thread::yield();
              else std::this_thread::sleep_for(std::chrono::microseconds(20));
            } else {
              idle = 0;
            }
          }
        } catch (...) {
          std::lock_guard<std::mutex> l(errorMutex);
          if (!error) error = std::current_exception();
          failed.store(true, std::memory_order_release);
        }
        if (threadEnd_) threadEnd_();
      });
    for (auto& t : threads) t.join();
    if (error) {
      closeAll();
      drivers_.clear();
      std::rethrow_exception(error);
    }
  }
  const int64_t ranNanos = sinceStart();
  for (auto& p : planner.pipelines)
    for (auto& d : p->drivers)
      for (auto& op : d->operators()) {
        // sibling operators of the drivers of a pipeline add up under one key
        const std::string prefix = std::to_string(p->factory.pipelineId) + "." + std::to_string(op->operatorId()) + "." + op->operatorType() + ".";
        stats_[prefix + "inputPositions"] += op->stats().inputPositions;
        stats_[prefix + "outputPositions"] += op->stats().outputPositions;
        stats_[prefix + "addInputWallNanos"] += op->stats().addInputWallNanos;
        stats_[prefix + "getOutputWallNanos"] += op->stats().getOutputWallNanos;
        stats_[prefix + "finishWallNanos"] += op->stats().finishWallNanos;
        for (auto& kv : op->stats().runtimeStats) stats_[prefix + kv.first] += kv.second;
      }
  stats_["task.numDrivers"] = static_cast<int64_t>(drivers_.size());
  closeAll();
  drivers_.clear();
  // where a task's wall time goes outside its operators (host-side latency is what limits small per-GPU shards)
  stats_["task.planWallNanos"] = plannedNanos;
  stats_["task.driversWallNanos"] = ranNanos - plannedNanos;
  stats_["task.closeWallNanos"] = sinceStart() - ranNanos;
  return results;
}

}  // namespace facebook::velox::exec
