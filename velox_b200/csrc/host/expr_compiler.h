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

// Runs the projections of a compiled program over rows sel[0..numOut) of a device batch (sel null: all
// rows): identity outputs are dictionary-wrapped, computed outputs written densely by one kernel.
B200VectorPtr evalProjections(const CompiledProgram& program, const B200VectorPtr& in, const DeviceBufferPtr& sel, int64_t numOut, cudaStream_t stream,
                              const DeviceBufferPtr& errorFlag, const RowTypePtr& outputType, memory::MemoryPool* pool);

// Base of every scalar function of the B200 engine. Inside a plan the function is an instruction of
// the fused expression program (one kernel per ExprSet, never node-at-a-time). apply() — the
// reference's node-at-a-time contract, velox/expression/VectorFunction.h:81-86 — is implemented on
// top of the same kernels: host argument vectors are uploaded, a one-call program runs on the device,
// the selected rows of the result come back (rows outside `rows` keep what `result` held).
class B200VectorFunction : public exec::VectorFunction {
 public:
  void apply(const SelectivityVector& rows, std::vector<VectorPtr>& args, const TypePtr& outputType, exec::EvalCtx& context,
             VectorPtr& result) const override;
};

// A user-defined scalar function whose device body is CUDA C++ source text:
//   exec::registerVectorFunction("hypot2", {sig}, std::make_unique<B200DeviceFunction>(
//       "hypot2", "__device__ double hypot2(double a, double b) { return sqrt(a * a + b * b); }", DOUBLE(), {DOUBLE(), DOUBLE()}));
// The expression JIT splices the source into every kernel whose ExprSet calls the function, so it
// fuses with the surrounding expression like a built-in (include/velox_b200_kernels.h
// vb2k_register_device_function). NULL arguments yield NULL (default null behaviour).
class B200DeviceFunction : public B200VectorFunction {
 public:
  B200DeviceFunction(std::string entry, std::string cudaSource, TypePtr returnType, std::vector<TypePtr> argTypes);
  int32_t id() const { return id_; }
  const TypePtr& returnType() const { return returnType_; }
  const std::vector<TypePtr>& argTypes() const { return argTypes_; }

 private:
  int32_t id_ = -1;
  TypePtr returnType_;
  std::vector<TypePtr> argTypes_;
};

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
