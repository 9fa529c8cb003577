// B200FromHost, B200ToHost, B200FilterProject and shared helpers.
#include "operators.h"

namespace velox_b200 {

void checkDeviceError(const DeviceBufferPtr& flag, cudaStream_t stream, const char* where) {
  int32_t code = 0;
  VB2_CU(cudaMemcpyAsync(&code, flag->data(), sizeof(code), cudaMemcpyDeviceToHost, stream));
  VB2_CU(cudaStreamSynchronize(stream));
  if (code == 0) return;
  VB2_CU(cudaMemsetAsync(flag->data(), 0, sizeof(int32_t), stream));
  switch (code) {
    case 1: throw VeloxUserError(std::string("integer overflow in ") + where);  // common/base/CheckedArithmetic.h:27-34
    case 2: throw VeloxUserError(std::string("division by zero in ") + where);
    case 3: throw VeloxUserError(std::string("Cannot cast value: out of range or NaN in ") + where);
    default: throw VeloxRuntimeError(std::string("device table overflow in ") + where);
  }
}

void orderAfterProducer(const B200Vector& batch, cudaStream_t stream) {
  if (batch.readyEvent()) VB2_CU(cudaStreamWaitEvent(stream, batch.readyEvent(), 0));
  cudaStream_t producer = batch.stream();
  if (!producer || producer == stream) return;
  // an event recorded now covers everything the producing stream was given so far, the batch included
  cudaEvent_t ev = nullptr;
  VB2_CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  VB2_CU(cudaEventRecord(ev, producer));
  VB2_CU(cudaStreamWaitEvent(stream, ev, 0));
  VB2_CU(cudaEventDestroy(ev));  // destruction is deferred until the event has completed
}

std::vector<vb2_column> describe(const B200Vector& v) {
  std::vector<vb2_column> cols;
  for (auto& c : v.columns()) cols.push_back(c->desc);
  return cols;
}

DeviceColumnPtr wrapColumn(const DeviceColumnPtr& col, const DeviceBufferPtr& sel, int64_t n, cudaStream_t stream) {
  auto out = std::make_shared<DeviceColumn>(*col);
  out->desc.size = n;
  if (!sel) return out;  // all rows pass
  const vb2_column& d = col->desc;
  if (d.encoding == VB2_CONSTANT) return out;
  if (d.encoding == VB2_FLAT) {
    out->desc.encoding = VB2_DICTIONARY;
    out->desc.indices = sel->as<int32_t>();
    out->desc.dict_size = d.size;
    out->desc.dict_nulls = d.nulls;
    out->desc.nulls = nullptr;
    out->owners.push_back(sel);
    return out;
  }
  // dictionary over dictionary: compose the indices, keep the base
  auto idx = allocDevice(static_cast<size_t>(n) * 4, stream);
  kernelCheck(vb2k_gather(d.indices, sel->as<int32_t>(), n, 4, idx->data(), stream));
  out->desc.indices = idx->as<int32_t>();
  out->owners.push_back(idx);
  if (d.nulls) {
    auto nb = allocDevice(bits::nbytes(n), stream);
    kernelCheck(vb2k_gather_bits(d.nulls, sel->as<int32_t>(), n, nb->as<uint64_t>(), stream));
    out->desc.nulls = nb->as<uint64_t>();
    out->owners.push_back(nb);
  }
  return out;
}

FlatColumn flattenColumn(const DeviceColumnPtr& col, const int32_t* sel, int64_t n, cudaStream_t stream) {
  // A one-instruction program (LOAD) run by the expression VM decodes any encoding.
  FlatColumn out;
  out.type = col->desc.type;
  VELOX_CHECK(out.type != VB2_VARCHAR, "flattenColumn: VARCHAR columns stay wrapped");
  vb2_instr in{VB2_OP_LOAD, out.type, 0, 0, 0, 0};
  vb2_program p{};
  p.instrs = &in;
  p.n_instrs = 1;
  p.filter_reg = -1;
  p.n_regs = 1;
  const int w = out.type == VB2_BOOLEAN ? 1 : widthOf(out.type);
  out.values = allocDevice(static_cast<size_t>(n) * w, stream);
  out.nulls = allocDevice(bits::nbytes(n), stream);
  vb2_output o{0, out.type, out.values->data(), out.nulls->as<uint64_t>()};
  auto flag = allocDeviceZeroed(8, stream);  // a LOAD cannot raise; the kernel still wants a flag
  kernelCheck(vb2k_eval_project(&p, &col->desc, 1, sel, n, &o, 1, flag->as<int32_t>(), stream));
  if (!col->mayHaveNulls()) out.nulls = nullptr;
  return out;
}

// ---- B200FromHost -----------------------------------------------------------------------------
void B200FromHost::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
}
RowVectorPtr B200FromHost::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!input_) return nullptr;
  RowVectorPtr in = std::move(input_);
  input_ = nullptr;
  if (std::dynamic_pointer_cast<B200Vector>(in)) return in;  // already resident in HBM
  auto out = toDevice(in, dev_->stream);
  // keep the source alive until its asynchronous copies have certainly completed
  inFlight_.push_back(in);
  if (inFlight_.size() > 2) {
    VB2_CU(cudaStreamSynchronize(dev_->stream));
    inFlight_.clear();
  }
  return out;
}

RowVectorPtr B200ToHost::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!input_) return nullptr;
  RowVectorPtr in = std::move(input_);
  input_ = nullptr;
  auto dev = std::dynamic_pointer_cast<B200Vector>(in);
  if (!dev) return in;
  return toHost(dev);
}

// ---- B200FilterProject ------------------------------------------------------------------------
B200FilterProject::B200FilterProject(int32_t id, exec::DriverCtx* ctx, const exec::FilterProject& cpu)
    : Operator(ctx, cpu.outputType(), id, cpu.planNodeId(), "B200FilterProject"),
      exprs_(cpu.exprsAndProjection().exprs->exprs()), hasFilter_(cpu.exprsAndProjection().hasFilter), inputType_(cpu.inputType()) {
  if (!cpu.projectNode()) {
    // filter only: every input column passes through
    for (uint32_t i = 0; i < inputType_->size(); ++i)
      exprs_.push_back(std::make_shared<core::FieldAccessTypedExpr>(inputType_->childAt(i), inputType_->nameOf(i)));
  }
  program_ = compileExprs(exprs_, hasFilter_, inputType_);
  if (hasFilter_) fastFilter_ = fusedSignature(exprs_[0], {}, inputType_);
}

void B200FilterProject::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
  errorFlag_ = allocDeviceZeroed(8, dev_->stream);
  program_.uploadConstants(dev_->stream);
  if (fastFilter_.ok && driverCtx_->queryConfig().b200FusedPipelines()) {
    fastFilterId_ = vb2k_fused_find(fastFilter_.signature.c_str());
    if (fastFilterId_ >= 0 && !vb2k_fused_has_filter(fastFilterId_)) fastFilterId_ = -1;
  }
}

B200VectorPtr B200FilterProject::apply(const B200VectorPtr& in) {
  cudaStream_t st = dev_->stream;
  const int64_t n = in->size();
  std::vector<vb2_column> cols = describe(*in);
  const vb2_program prog = program_.view();
  DeviceBufferPtr sel;  // null = every row passes
  int64_t numOut = n;
  if (hasFilter_) {
    auto bitsBuf = allocDevice(bits::nbytes(n), st);
    bool done = false;
    int64_t selCapacity = n;
    if (fastFilterId_ >= 0 && n >= (1 << 16)) {
      // flat NULL-free filter columns: the TMA-staged bitmap kernel (HBM-bound) instead of one row per thread
      vb2_fused_args fa{};
      bool flat = true;
      for (size_t i = 0; i < fastFilter_.columns.size() && flat; ++i) {
        const vb2_column& d = in->column(fastFilter_.columns[i])->desc;
        flat = d.encoding == VB2_FLAT && !d.nulls;
        fa.cols[i] = d.values;
      }
      if (flat) {
        for (size_t i = 0; i < fastFilter_.pf.size(); ++i) fa.pf[i] = fastFilter_.pf[i];
        for (size_t i = 0; i < fastFilter_.pl.size(); ++i) fa.pl[i] = fastFilter_.pl[i];
        for (size_t i = 0; i < fastFilter_.pi.size(); ++i) fa.pi[i] = fastFilter_.pi[i];
        fa.rows = n;
        auto counters = allocDeviceZeroed(16, st);
        const int rc = vb2k_fused_filter_bits(fastFilterId_, &fa, 1, bitsBuf->as<uint64_t>(), counters->as<int64_t>(), st);
        if (rc == VB2_OK) {
          done = true;
          addRuntimeStat("b200.fastFilterBatches", exec::RuntimeCounter{1});
          // the kernel counted the surviving rows: size the row-number buffer exactly (a filter that keeps
          // 1 % of 300 M rows needs 12 MB, not 1.2 GB)
          int64_t kept[2] = {0, 0};
          VB2_CU(cudaMemcpyAsync(kept, counters->data(), 16, cudaMemcpyDeviceToHost, st));
          VB2_CU(cudaStreamSynchronize(st));
          selCapacity = kept[0];
        } else if (rc != VB2_ERR_UNSUPPORTED) kernelCheck(rc);
      }
    }
    if (!done)
      kernelCheck(vb2k_eval_filter(&prog, cols.data(), static_cast<int32_t>(cols.size()), n, bitsBuf->as<uint64_t>(), errorFlag_->as<int32_t>(), st));
    if (selCapacity == 0) return nullptr;
    sel = allocDevice(static_cast<size_t>(selCapacity) * 4, st);
    auto count = allocDevice(8, st);
    const size_t wsBytes = vb2k_bits_to_indices_workspace(n);
    auto ws = allocDevice(wsBytes, st);
    kernelCheck(vb2k_bits_to_indices(bitsBuf->as<uint64_t>(), n, sel->as<int32_t>(), count->as<int64_t>(), ws->data(), wsBytes, st));
    VB2_CU(cudaMemcpyAsync(&numOut, count->data(), 8, cudaMemcpyDeviceToHost, st));
    checkDeviceError(errorFlag_, st, "filter");  // synchronises
    if (numOut == 0) return nullptr;
    if (numOut == n) sel = nullptr;
  }
  return evalProjections(program_, in, sel, numOut, st, errorFlag_, outputType_, pool());
}

B200VectorPtr evalProjections(const CompiledProgram& program, const B200VectorPtr& in, const DeviceBufferPtr& sel, int64_t numOut, cudaStream_t st,
                              const DeviceBufferPtr& errorFlag, const RowTypePtr& outputType, memory::MemoryPool* pool) {
  std::vector<vb2_column> cols = describe(*in);
  const vb2_program prog = program.view();
  // which registers can be NULL for this batch
  std::vector<bool> colNull;
  for (auto& c : in->columns()) colNull.push_back(c->mayHaveNulls());
  const std::vector<bool> regNull = program.nullability(colNull);

  std::vector<DeviceColumnPtr> outCols(program.outputs.size());
  std::vector<vb2_output> outs;
  for (size_t i = 0; i < program.outputs.size(); ++i) {
    const auto& o = program.outputs[i];
    if (o.identityField >= 0) {
      outCols[i] = wrapColumn(in->column(o.identityField), sel, numOut, st);
      continue;
    }
    auto col = std::make_shared<DeviceColumn>();
    col->type = o.type;
    col->desc.type = veloxTypeToVb2(o.type);
    col->desc.encoding = VB2_FLAT;
    col->desc.size = numOut;
    const int t = col->desc.type;
    // BOOLEAN results are written one byte per row by the VM, then packed
    const size_t bytes = static_cast<size_t>(numOut) * (t == VB2_BOOLEAN ? 1 : widthOf(t));
    auto values = allocDevice(bytes, st);
    auto nulls = allocDevice(bits::nbytes(numOut), st);
    outs.push_back(vb2_output{o.reg, t, values->data(), nulls->as<uint64_t>()});
    col->owners = {values, nulls};
    col->desc.values = values->data();
    col->desc.nulls = regNull[o.reg] ? nulls->as<uint64_t>() : nullptr;
    outCols[i] = col;
  }
  if (!outs.empty()) {
    kernelCheck(vb2k_eval_project(&prog, cols.data(), static_cast<int32_t>(cols.size()), sel ? sel->as<int32_t>() : nullptr, numOut,
                                  outs.data(), static_cast<int32_t>(outs.size()), errorFlag->as<int32_t>(), st));
    if (program.canRaise) checkDeviceError(errorFlag, st, "projection");
    // pack BOOLEAN byte results into the bit-packed layout of FlatVector<bool>
    for (size_t i = 0; i < outCols.size(); ++i) {
      auto& c = outCols[i];
      if (program.outputs[i].identityField < 0 && c->desc.type == VB2_BOOLEAN) {
        auto packed = allocDevice(bits::nbytes(numOut), st);
        kernelCheck(vb2k_pack_bools(reinterpret_cast<const uint8_t*>(c->desc.values), numOut, packed->as<uint64_t>(), st));
        c->owners.push_back(packed);
        c->desc.values = packed->data();
      }
    }
  }
  // the input batch's buffers back the wrapped columns: the owners lists keep them alive
  return std::make_shared<B200Vector>(pool, outputType, static_cast<vector_size_t>(numOut), std::move(outCols), st);
}

RowVectorPtr B200FilterProject::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!input_) return nullptr;
  auto in = std::dynamic_pointer_cast<B200Vector>(input_);
  input_ = nullptr;
  VELOX_CHECK(in != nullptr, "B200FilterProject expects device-resident input (B200FromHost missing?)");
  orderAfterProducer(*in, dev_->stream);
  return apply(in);
}

}  // namespace velox_b200
