// registerB200(): installs the DriverAdapter that replaces the CPU operators of a pipeline with
// the B200 ones and inserts the host<->device conversion operators at the seams.
// Reference pattern: velox/experimental/cudf/exec/ToCudf.cpp:71-338 (per-operator adapters,
// automatic FromVelox / ToVelox insertion), Wave's fused runs velox/experimental/wave/exec/ToWave.cpp:221.
#include "operators.h"

namespace velox_b200 {

namespace {

// kSingle node equivalent to final(partial): same keys, same aggregates over the raw inputs, the
// final node's output type and id. nullptr when the pair does not line up.
std::shared_ptr<const core::AggregationNode> collapsePartialFinal(const core::AggregationNode& partial, const core::AggregationNode& fin) {
  using Step = core::AggregationNode::Step;
  if (partial.step() != Step::kPartial || fin.step() != Step::kFinal) return nullptr;
  // in serial execution mode the local gather between the two steps is the identity and the planner
  // drops it from the driver: look through it
  const core::PlanNode* below = fin.sources()[0].get();
  if (auto lp = dynamic_cast<const core::LocalPartitionNode*>(below)) below = lp->sources()[0].get();
  if (below != &partial) return nullptr;
  const size_t nk = partial.groupingKeys().size();
  if (fin.groupingKeys().size() != nk || fin.aggregates().size() != partial.aggregates().size()) return nullptr;
  const RowTypePtr& mid = partial.outputType();
  for (size_t k = 0; k < nk; ++k)
    if (fin.groupingKeys()[k]->name() != mid->nameOf(k)) return nullptr;
  uint32_t col = static_cast<uint32_t>(nk);
  std::vector<core::AggregationNode::Aggregate> aggregates;
  for (size_t i = 0; i < fin.aggregates().size(); ++i) {
    const auto& f = fin.aggregates()[i];
    const auto& p = partial.aggregates()[i];
    if (f.call->name() != p.call->name() || f.mask || f.distinct || p.distinct || f.call->inputs().empty()) return nullptr;
    auto in0 = dynamic_cast<const core::FieldAccessTypedExpr*>(f.call->inputs()[0].get());
    if (!in0 || col >= mid->size() || in0->name() != mid->nameOf(col)) return nullptr;
    col += static_cast<uint32_t>(f.call->inputs().size());  // avg's intermediate is a (sum, count) column pair here
    // the raw arguments and mask of the partial step, the result type of the final one
    aggregates.push_back({std::make_shared<core::CallTypedExpr>(f.call->type(), p.call->inputs(), p.call->name()), p.rawInputTypes, p.mask});
  }
  auto single = std::make_shared<core::AggregationNode>(fin.id(), Step::kSingle, partial.groupingKeys(), partial.preGroupedKeys(), fin.aggregateNames(),
                                                        std::move(aggregates), partial.ignoreNullKeys(), partial.noGroupsSpanBatches(), partial.sources()[0]);
  single->setOutputType(fin.outputType());
  return single;
}

// DISTINCT aggregates (exec/DistinctAggregations.cpp keeps a set of inputs per group and feeds every distinct value to
// the function once). On the device the set IS a group-by: a single-step aggregation whose aggregates are all DISTINCT
// over one common column x (unmasked) becomes
//     Aggregation[keys + x, no aggregates]  ->  Aggregation[keys, the same aggregates without DISTINCT]
// — the first operator emits every (keys, x) combination once (NULL x included: it keeps a group with only NULL inputs
// alive and the functions ignore it, as they do in the reference), the second one aggregates those rows.
// Returns false when the node has no DISTINCT aggregate; throws for DISTINCT shapes outside this rewrite.
bool splitDistinctAggregation(const core::AggregationNode& node, std::shared_ptr<const core::AggregationNode>& dedup,
                              std::shared_ptr<const core::AggregationNode>& aggregate) {
  bool any = false;
  for (auto& a : node.aggregates()) any = any || a.distinct;
  if (!any) return false;
  using Step = core::AggregationNode::Step;
  core::FieldAccessTypedExprPtr input;
  for (auto& a : node.aggregates()) {
    auto in0 = a.call->inputs().size() == 1 ? std::dynamic_pointer_cast<const core::FieldAccessTypedExpr>(a.call->inputs()[0]) : nullptr;
    if (!a.distinct || a.mask || !a.sortingKeys.empty() || !in0 || (input && input->name() != in0->name()) || node.step() != Step::kSingle)
      VELOX_UNSUPPORTED("DISTINCT aggregates: a single-step aggregation whose aggregates are all DISTINCT over one common input column, without masks, is supported; got " +
                        a.call->toString());
    input = in0;
  }
  if (input->type()->kind() == TypeKind::VARCHAR) VELOX_UNSUPPORTED("DISTINCT aggregates over VARCHAR");
  std::vector<core::FieldAccessTypedExprPtr> keys = node.groupingKeys();
  for (auto& k : keys)
    if (k->name() == input->name()) VELOX_UNSUPPORTED("DISTINCT aggregate over a grouping key");
  keys.push_back(input);
  dedup = std::make_shared<core::AggregationNode>(node.id() + ".distinct", Step::kSingle, keys, std::vector<core::FieldAccessTypedExprPtr>{},
                                                  std::vector<std::string>{}, std::vector<core::AggregationNode::Aggregate>{}, false, false,
                                                  node.sources()[0]);
  std::vector<core::AggregationNode::Aggregate> plain = node.aggregates();
  for (auto& a : plain) a.distinct = false;
  auto agg = std::make_shared<core::AggregationNode>(node.id(), Step::kSingle, node.groupingKeys(), node.preGroupedKeys(), node.aggregateNames(), std::move(plain),
                                                     node.ignoreNullKeys(), false, dedup);
  agg->setOutputType(node.outputType());
  aggregate = agg;
  return true;
}

bool adaptDriver(const exec::DriverFactory& factory, exec::Driver& driver) {
  const core::QueryConfig& config = driver.driverCtx()->queryConfig();
  if (!config.b200Enabled()) return false;
  auto& ops = driver.operators();
  exec::DriverCtx* ctx = driver.driverCtx();
  std::vector<std::unique_ptr<exec::Operator>> out;
  bool replacedAny = false;
  const bool fuse = config.b200FusedPipelines();
  for (size_t i = 0; i < ops.size(); ++i) {
    exec::Operator* op = ops[i].get();
    const int32_t id = static_cast<int32_t>(out.size());
    if (auto values = dynamic_cast<exec::Values*>(op)) {
      RowTypePtr type = values->outputType();
      out.push_back(std::move(ops[i]));
      out.push_back(std::make_unique<B200FromHost>(id + 1, ctx, type));
      replacedAny = true;
    } else if (auto fp = dynamic_cast<exec::FilterProject*>(op)) {
      out.push_back(std::make_unique<B200FilterProject>(id, ctx, *fp));
      replacedAny = true;
    } else if (auto agg = dynamic_cast<exec::HashAggregation*>(op)) {
      // Fuse the run of B200FilterProject / B200HashProbe operators feeding the aggregation into
      // it: the aggregation then sees the source batches and can run the whole chain as one
      // kernel; functionally the absorbed operators still run batch by batch when it cannot.
      std::vector<std::unique_ptr<exec::Operator>> absorbed;
      if (fuse && agg->node()->isRawInput()) {
        size_t first = out.size();
        while (first > 0 && (dynamic_cast<B200FilterProject*>(out[first - 1].get()) || dynamic_cast<B200HashProbe*>(out[first - 1].get()))) --first;
        // only the shapes the fused path understands; anything else stays unfused
        const size_t run = out.size() - first;
        // FilterProject | FilterProject -> HashProbe -> FilterProject | HashProbe -> FilterProject (the
        // probe side arrives ready-made, e.g. from an Exchange)
        const bool shapeOk = run == 1 ? dynamic_cast<B200FilterProject*>(out[first].get()) != nullptr
                           : run == 2 ? (dynamic_cast<B200HashProbe*>(out[first].get()) && dynamic_cast<B200FilterProject*>(out[first + 1].get()))
                                      : (run == 3 && dynamic_cast<B200FilterProject*>(out[first].get()) &&
                                         dynamic_cast<B200HashProbe*>(out[first + 1].get()) && dynamic_cast<B200FilterProject*>(out[first + 2].get()));
        if (shapeOk) {
          for (size_t j = first; j < out.size(); ++j) absorbed.push_back(std::move(out[j]));
          out.resize(first);
        }
      }
      std::shared_ptr<const core::AggregationNode> node = agg->node();
      {
        std::shared_ptr<const core::AggregationNode> dedup, plain;
        if (splitDistinctAggregation(*node, dedup, plain)) {
          out.push_back(std::make_unique<B200HashAggregation>(static_cast<int32_t>(out.size()), ctx, dedup, std::move(absorbed)));
          out.push_back(std::make_unique<B200HashAggregation>(static_cast<int32_t>(out.size()), ctx, plain, std::vector<std::unique_ptr<exec::Operator>>{}));
          replacedAny = true;
          continue;
        }
      }
      // partial -> final back to back in ONE driver (no exchange in between) is a single
      // aggregation: final(partial(x)) == single(x) group by group, so the pair becomes one
      // operator and the intermediate batch never exists.
      if (fuse && i + 1 < ops.size()) {
        if (auto next = dynamic_cast<exec::HashAggregation*>(ops[i + 1].get())) {
          if (auto merged = collapsePartialFinal(*node, *next->node())) {
            node = merged;
            ++i;
          }
        }
      }
      out.push_back(std::make_unique<B200HashAggregation>(static_cast<int32_t>(out.size()), ctx, node, std::move(absorbed)));
      replacedAny = true;
    } else if (auto build = dynamic_cast<exec::HashBuild*>(op)) {
      out.push_back(std::make_unique<B200HashBuild>(id, ctx, *build));
      replacedAny = true;
    } else if (auto probe = dynamic_cast<exec::HashProbe*>(op)) {
      out.push_back(std::make_unique<B200HashProbe>(id, ctx, *probe));
      replacedAny = true;
    } else if (auto ob = dynamic_cast<exec::OrderBy*>(op)) {
      out.push_back(std::make_unique<B200OrderBy>(id, ctx, ob->node()));
      replacedAny = true;
    } else if (auto tn = dynamic_cast<exec::TopN*>(op)) {
      out.push_back(std::make_unique<B200TopN>(id, ctx, tn->node()));
      replacedAny = true;
    } else if (auto po = dynamic_cast<exec::PartitionedOutput*>(op)) {
      out.push_back(std::make_unique<B200PartitionedOutput>(id, ctx, *po));
      replacedAny = true;
    } else if (auto ex = dynamic_cast<exec::Exchange*>(op)) {
      out.push_back(std::make_unique<B200Exchange>(id, ctx, *ex));
      replacedAny = true;
    } else if (dynamic_cast<exec::CallbackSink*>(op)) {
      // b200.result_on_device: the consumer takes device-resident batches (another GPU stage, or the C
      // ABI's direct copy-out): no B200ToHost in front of the sink
      if (!config.get<bool>("b200.result_on_device", false)) {
        RowTypePtr type = out.empty() ? nullptr : out.back()->outputType();
        out.push_back(std::make_unique<B200ToHost>(id, ctx, type));
      }
      out.push_back(std::move(ops[i]));
    } else {
      out.push_back(std::move(ops[i]));
    }
  }
  if (!replacedAny) return false;
  // every original operator was moved or superseded: swap the whole list
  std::vector<std::unique_ptr<exec::Operator>> replaced = factory.replaceOperators(driver, 0, static_cast<int32_t>(ops.size()), std::move(out));
  (void)replaced;
  return true;
}

}  // namespace

void registerB200() {
  static bool done = false;
  if (done) return;
  done = true;
  registerB200Functions();
  exec::DriverAdapter adapter;
  adapter.label = "b200";
  adapter.inspect = nullptr;
  adapter.adapt = adaptDriver;
  exec::DriverFactory::registerAdapter(std::move(adapter));
}

}  // namespace velox_b200
