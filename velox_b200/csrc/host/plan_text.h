#pragma once
#include <string>

#include "../../abi/exec_abi.h"

namespace velox_b200 {
using namespace facebook::velox;
// Parses the plan text of the C ABI into a core::PlanNode tree. Throws VeloxRuntimeError.
core::PlanNodePtr parsePlanText(const std::string& text);
}  // namespace velox_b200
