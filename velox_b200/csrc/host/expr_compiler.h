// Compiles typed expression trees (the ExprSet of a FilterProject, a join filter ...) into
//   (a) a linear register program for the device expression VM (kernels: expr_vm.cu), with
//       common sub-expressions shared the way ExprCompiler dedups them
//       (velox/expression/ExprCompiler.cpp), and
//   (b) the canonical text used to look up an ahead-of-time fused pipeline (fused_scan.cu).
// Function names resolve through the VectorFunction registry (velox/expression/VectorFunction.h:241).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "device.h"

namespace velox_b200 {

struct CompiledProgram {
  std::vector<vb2_instr> instrs;
  std::vector<vb2_const> consts;       // str fields point into constChars after finalizeConstants()
  std::vector<std::string> constStrings;
  DeviceBufferPtr constChars;          // device copy of all string constants
  int nFilterInstrs = 0;
  int filterReg = -1;
  int nRegs = 0;
  struct Output {
    int reg = -1;            // VM register, or -1 for an identity projection
    int identityField = -1;  // input column when the projection is a plain field reference
    TypePtr type;
  };
  std::vector<Output> outputs;
  bool canRaise = false;  // contains checked integer arithmetic or casts that can fail

  vb2_program view() const {
    vb2_program p{};
    p.instrs = instrs.data();
    p.n_instrs = static_cast<int32_t>(instrs.size());
    p.n_filter_instrs = nFilterInstrs;
    p.filter_reg = filterReg;
    p.n_regs = nRegs;
    p.consts = consts.data();
    p.n_consts = static_cast<int32_t>(consts.size());
    return p;
  }
  // Per-register "may be NULL" given which input columns can hold nulls in this batch.
  std::vector<bool> nullability(const std::vector<bool>& columnMayBeNull) const;
  void uploadConstants(cudaStream_t stream);
};

// exprs[0] is the filter when hasFilter. Throws VeloxRuntimeError for unsupported shapes.
CompiledProgram compileExprs(const std::vector<core::TypedExprPtr>& exprs, bool hasFilter, const RowTypePtr& inputType);

// Registers the B200 scalar functions (plus/minus/multiply/divide/modulus/negate, lt..neq,
// between, like, not, is_null, and/or/switch/cast are special forms) in the registry.
void registerB200Functions();
int opcodeForFunction(const std::string& name);  // -1 when the name is not a registered B200 function

// ---- fused pipeline matching ---------------------------------------------------------------------
struct FusedBinding {
  std::string signature;
  std::vector<int> columns;        // input column of each renumbered expression column
  std::vector<double> pf;
  std::vector<int64_t> pl;
  std::vector<int32_t> pi;
  bool ok = false;
};
// filter may be null. `joinKeyColumn` >= 0 adds a probe on that (BIGINT) input column after the
// filter; `joinFlagExpr`, when set, is the sub-expression (over the join's build payload) that the
// projections may reference as the build-side predicate.
FusedBinding fusedSignature(const core::TypedExprPtr& filter, const std::vector<core::TypedExprPtr>& projections,
                            const RowTypePtr& inputType, int joinKeyColumn = -1,
                            const core::ITypedExpr* joinFlagExpr = nullptr);

// Replaces the references to the columns of `type` in `expr` by the given expressions (inlines a ProjectNode).
core::TypedExprPtr substituteFields(const core::TypedExprPtr& expr, const std::vector<core::TypedExprPtr>& fields, const RowTypePtr& type);

}  // namespace velox_b200
