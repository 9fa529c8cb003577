// Name -> channel resolution of plan nodes, the way the reference's operators do it in initialize()
// (exec/OperatorUtils.h exprToChannel; HashAggregation::initialize exec/HashAggregation.cpp:60-120;
// HashProbe / HashBuild constructors exec/HashProbe.cpp:120-190). The plan nodes carry column NAMES
// (core/PlanNode.h:1136-1158, :3400); these helpers turn them into input-column indices once.
#pragma once
#include <string>
#include <vector>

#include "device.h"

namespace velox_b200 {

inline int32_t channelOf(const RowTypePtr& type, const std::string& name) {
  auto idx = type->getChildIdxIfExists(name);
  if (!idx) throw VeloxRuntimeError("Field not found: " + name + " in " + type->toString());
  return static_cast<int32_t>(*idx);
}
inline int32_t channelOf(const RowTypePtr& type, const core::FieldAccessTypedExpr& f) { return channelOf(type, f.name()); }

// Synthesised references to columns of a join's build side inside a fused pipeline's expressions
// (they never appear in plans): the build column number rides in a reserved name.
inline std::string buildFieldName(int32_t column) { return "\x01" "b200.build#" + std::to_string(column); }
inline int32_t buildFieldColumn(const core::FieldAccessTypedExpr& f) {
  static const std::string kPrefix = "\x01" "b200.build#";
  if (f.name().compare(0, kPrefix.size(), kPrefix) != 0) return -1;
  return std::stoi(f.name().substr(kPrefix.size()));
}

struct ResolvedAggregate {
  std::string function;          // sum avg count min max
  std::vector<int32_t> inputs;   // input columns: raw argument, or the intermediate column(s) for kFinal / kIntermediate
  int32_t mask = -1;             // BOOLEAN mask column
  TypePtr rawInputType;          // type of the raw argument (decides sum's accumulator)
};
struct ResolvedAggregation {
  std::vector<int32_t> keys;
  std::vector<ResolvedAggregate> aggregates;
};
inline ResolvedAggregation resolveAggregation(const core::AggregationNode& node) {
  ResolvedAggregation r;
  const RowTypePtr& in = node.sources()[0]->outputType();
  for (auto& k : node.groupingKeys()) r.keys.push_back(channelOf(in, *k));
  for (auto& a : node.aggregates()) {
    ResolvedAggregate ra;
    ra.function = a.call->name();
    for (auto& arg : a.call->inputs()) {
      auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(arg.get());
      if (!f) {
        // count(0) and friends: a constant argument counts rows
        if (dynamic_cast<const core::ConstantTypedExpr*>(arg.get())) continue;
        throw VeloxRuntimeError("aggregate arguments must be input columns: " + a.call->toString());
      }
      ra.inputs.push_back(channelOf(in, *f));
    }
    if (a.mask) ra.mask = channelOf(in, *a.mask);
    if (a.distinct)
      VELOX_UNSUPPORTED("DISTINCT aggregate in this shape: " + a.call->toString() +
                        " (a single-step aggregation whose aggregates are all DISTINCT over one common column is split by the adapter)");
    if (!a.sortingKeys.empty()) VELOX_UNSUPPORTED("aggregates with ORDER BY: " + a.call->toString());
    if (!a.rawInputTypes.empty()) ra.rawInputType = a.rawInputTypes[0];
    r.aggregates.push_back(std::move(ra));
  }
  return r;
}

struct JoinOutput {
  bool fromProbe;
  int32_t column;
};
struct ResolvedJoin {
  std::vector<int32_t> leftKeys, rightKeys;
  std::vector<JoinOutput> outputs;  // the join's output columns: probe-side names first, then build-side names (exec/HashProbe.cpp:150-190)
};
inline ResolvedJoin resolveJoin(const core::HashJoinNode& node) {
  ResolvedJoin r;
  const RowTypePtr& probe = node.sources()[0]->outputType();
  const RowTypePtr& build = node.sources()[1]->outputType();
  for (auto& k : node.leftKeys()) r.leftKeys.push_back(channelOf(probe, *k));
  for (auto& k : node.rightKeys()) r.rightKeys.push_back(channelOf(build, *k));
  const RowTypePtr& out = node.outputType();
  for (uint32_t i = 0; i < out->size(); ++i) {
    if (auto p = probe->getChildIdxIfExists(out->nameOf(i))) r.outputs.push_back({true, static_cast<int32_t>(*p)});
    else r.outputs.push_back({false, channelOf(build, out->nameOf(i))});
  }
  return r;
}

}  // namespace velox_b200
