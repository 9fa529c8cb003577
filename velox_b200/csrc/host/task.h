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
  // Plans, creates one driver per pipeline, runs them to completion (serially, honouring
  // isBlocked futures) and returns the batches that reached the sink.
  std::vector<RowVectorPtr> run();
  const core::PlanNodePtr& plan() const { return plan_; }
  const core::QueryConfig& queryConfig() const { return config_; }
  memory::MemoryPool* pool() { return &pool_; }
  // runtime stats of every operator of every driver, "pipeline.operator.type.name" -> value
  std::map<std::string, int64_t> stats() const { return stats_; }

 private:
  core::PlanNodePtr plan_;
  core::QueryConfig config_;
  memory::MemoryPool pool_{"task"};
  std::map<int32_t, std::shared_ptr<std::vector<RowVectorPtr>>> inputs_;
  std::map<std::string, int64_t> stats_;
  std::shared_ptr<ExchangeTransport> transport_;
};

}  // namespace facebook::velox::exec
