// Contract shim, continued: LocalPlanner + Task (serial execution mode).
//   LocalPlanner::plan / DriverFactory::createDriver ... velox/exec/LocalPlanner.cpp:374-810
//     (Filter+Project fused into one FilterProject :517-535, HashProbe :591, HashAggregation :618,
//      HashBuild as the build pipeline's consumer :239, adapters run after driver->init :762-766)
//   Task ........................................... velox/exec/Task.h
#pragma once
#include "../../abi/exec_abi.h"

namespace facebook::velox::exec {

// The transport behind PartitionedOutput / Exchange (the reference plugs ExchangeSource factories,
// velox/exec/ExchangeSource.h:139-145, selected by task URI). One process per GPU: the transport
// is the communicator of the ranks running the same plan.
class ExchangeTransport {
 public:
  virtual ~ExchangeTransport() = default;
  virtual int world() const = 0;
  virtual int rank() const = 0;
};

class Task {
 public:
  void setExchangeTransport(std::shared_ptr<ExchangeTransport> t) { transport_ = std::move(t); }
  const std::shared_ptr<ExchangeTransport>& exchangeTransport() const { return transport_; }
  Task(core::PlanNodePtr plan, core::QueryConfig config);
  ~Task();
  // Batches for a ValuesNode source (before run()).
  void addInput(int32_t sourceId, RowVectorPtr batch);
  // Plans, creates the drivers of every pipeline and runs them to completion; returns the batches that
  // reached the sink. QueryConfig "task.max_drivers" (default 1) = Task::start's maxDrivers: with 1
  // every pipeline has one driver and they take turns on the calling thread (serial execution mode);
  // above 1, pipelines whose nodes allow it get that many drivers, one thread each.
  std::vector<RowVectorPtr> run();
  const core::PlanNodePtr& plan() const { return plan_; }
  const core::QueryConfig& queryConfig() const { return config_; }
  memory::MemoryPool* pool() { return &pool_; }
  // runtime stats of every operator of every driver, "pipeline.operator.type.name" -> value
  std::map<std::string, int64_t> stats() const { return stats_; }

  // velox/exec/Task.cpp:2451 — barrier of the sibling operators of one plan node across the drivers of
  // a pipeline (HashBuild::finishHashBuild, exec/HashBuild.cpp:819): returns true for the LAST caller,
  // handing it the earlier callers' drivers and the promises that release them; earlier callers get
  // `future` and stay blocked until the last one fulfils their promise.
  bool allPeersFinished(const core::PlanNodeId& planNodeId, Driver* caller, ContinueFuture* future, std::vector<ContinuePromise>& promises,
                        std::vector<std::shared_ptr<Driver>>& peers);
  // Drivers running the given pipeline (Task::numDrivers).
  int32_t numDrivers(int32_t pipelineId) const { return pipelineId < static_cast<int32_t>(driversPerPipeline_.size()) ? driversPerPipeline_[pipelineId] : 1; }
  // Called at the start / end of every driver thread (task.max_drivers > 1): lets the embedding layer
  // attach its thread-local state (CUDA device, upload cache) to the thread.
  void setDriverThreadHooks(std::function<void()> begin, std::function<void()> end) { threadBegin_ = std::move(begin); threadEnd_ = std::move(end); }

 private:
  struct BarrierState {
    int32_t numRequested = 0;
    std::vector<std::shared_ptr<Driver>> drivers;
    std::vector<ContinuePromise> allPeersFinishedPromises;
  };
  std::mutex mutex_;
  std::map<core::PlanNodeId, BarrierState> barriers_;
  std::vector<std::shared_ptr<Driver>> drivers_;
  std::vector<int32_t> driversPerPipeline_;
  std::function<void()> threadBegin_, threadEnd_;
  core::PlanNodePtr plan_;
  core::QueryConfig config_;
  memory::MemoryPool pool_{"task"};
  std::map<int32_t, std::shared_ptr<std::vector<RowVectorPtr>>> inputs_;
  std::map<std::string, int64_t> stats_;
  std::shared_ptr<ExchangeTransport> transport_;
};

}  // namespace facebook::velox::exec
