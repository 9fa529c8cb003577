// B200HashBuild / B200HashProbe and the key-normalisation helpers shared with aggregation.
#include <map>

#include "join.h"
#include "plan_resolve.h"

namespace velox_b200 {

void columnMinMax(const DeviceColumn& col, int64_t rows, cudaStream_t stream, int64_t& lo, int64_t& hi, int64_t& nonNull) {
  auto out = allocDevice(24, stream);
  kernelCheck(vb2k_column_minmax(&col.desc, rows, out->as<int64_t>(), stream));
  int64_t h[3];
  VB2_CU(cudaMemcpyAsync(h, out->data(), 24, cudaMemcpyDeviceToHost, stream));
  VB2_CU(cudaStreamSynchronize(stream));
  lo = h[0];
  hi = h[1];
  nonNull = h[2];
}

NormalizedKeys normalizeKeys(const B200Vector& batch, const std::vector<int32_t>& keyColumns, const KeyLayout& layout,
                             const int32_t* sel, int64_t n, bool forJoin, cudaStream_t stream) {
  std::vector<vb2_column> cols;
  for (int32_t k : keyColumns) cols.push_back(batch.column(k)->desc);
  NormalizedKeys out;
  out.keys = allocDevice(static_cast<size_t>(n) * 8, stream);
  if (forJoin) out.valid = allocDevice(bits::nbytes(n), stream);
  kernelCheck(vb2k_normalize_keys(cols.data(), static_cast<int32_t>(cols.size()), layout.mins.data(), layout.mults.data(),
                                  forJoin ? layout.ranges.data() : nullptr, forJoin ? 1 : 0, sel, n, out.keys->as<uint64_t>(),
                                  forJoin ? out.valid->as<uint64_t>() : nullptr, stream));
  return out;
}

// HashBuild::addInput appends every batch to the build side (exec/HashBuild.cpp:442-598). On the device
// the batches are decoded into one flat column each at noMoreInput: fixed-width values are
// flattened and copied, validity travels as bytes (bitmaps of different batches start at arbitrary
// bit offsets) and is packed once at the end, BOOLEAN values likewise, and VARCHAR columns keep
// their dictionary form over the union of the batches' alphabets (codes remapped per batch).
B200VectorPtr concatBatches(const std::vector<B200VectorPtr>& batches, memory::MemoryPool* pool, cudaStream_t stream) {
  VELOX_CHECK(!batches.empty(), "concatBatches: no input");
  if (batches.size() == 1) return batches[0];
  int64_t total = 0;
  for (auto& b : batches) total += b->size();
  VELOX_CHECK(total < (1ll << 31), "build side above 2^31 rows");
  const size_t ncols = batches[0]->columns().size();
  std::vector<DeviceColumnPtr> cols;
  for (size_t c = 0; c < ncols; ++c) {
    const auto& first = batches[0]->column(c);
    auto col = std::make_shared<DeviceColumn>();
    col->type = first->type;
    col->desc.type = first->desc.type;
    col->desc.size = total;
    const int t = first->desc.type;
    bool anyNulls = false;
    for (auto& b : batches) anyNulls = anyNulls || b->column(c)->mayHaveNulls();
    DeviceBufferPtr validBytes = anyNulls ? allocDevice(static_cast<size_t>(total), stream) : nullptr;
    if (t == VB2_VARCHAR) {
      // union of the alphabets (first occurrence wins); codes of every batch go through its own remap
      auto merged = std::make_shared<HostAlphabet>();
      std::map<std::string, int32_t> ids;
      auto codes = allocDevice(static_cast<size_t>(total) * 4, stream);
      int64_t off = 0;
      for (auto& b : batches) {
        const DeviceColumnPtr& bc = b->column(c);
        if (bc->desc.encoding == VB2_FLAT || !bc->alphabet) VELOX_UNSUPPORTED("multi-batch build side with flat (non-dictionary) VARCHAR columns");
        std::vector<int32_t> remap(bc->alphabet->values.size());
        for (size_t i = 0; i < remap.size(); ++i) {
          if (bc->alphabet->nulls[i]) { remap[i] = 0; continue; }  // NULL entries: the validity byte decides, the code is unused
          auto it = ids.find(bc->alphabet->values[i]);
          if (it == ids.end()) {
            it = ids.emplace(bc->alphabet->values[i], static_cast<int32_t>(merged->values.size())).first;
            merged->values.push_back(bc->alphabet->values[i]);
            merged->nulls.push_back(false);
          }
          remap[i] = it->second;
        }
        const int64_t bn = b->size();
        auto raw = allocDevice(static_cast<size_t>(bn) * 4, stream);
        kernelCheck(vb2k_dictionary_codes(&bc->desc, bn, raw->as<int32_t>(), validBytes ? validBytes->as<uint8_t>() + off : nullptr, stream));
        auto lut = allocDevice(remap.size() * 4 + 4, stream);
        VB2_CU(cudaMemcpyAsync(lut->data(), remap.data(), remap.size() * 4, cudaMemcpyHostToDevice, stream));  // pageable: staged before return
        kernelCheck(vb2k_gather(lut->data(), raw->as<int32_t>(), bn, 4, codes->as<int32_t>() + off, stream));
        off += bn;
      }
      if (merged->values.empty()) { merged->values.push_back(""); merged->nulls.push_back(false); }
      DeviceBufferPtr offBuf, charBuf;
      deviceAlphabet(*merged, stream, offBuf, charBuf);
      col->desc.encoding = VB2_DICTIONARY;
      col->desc.indices = codes->as<int32_t>();
      col->desc.values = offBuf->data();
      col->desc.aux = charBuf->data();
      col->desc.dict_size = static_cast<int64_t>(merged->values.size());
      col->owners = {codes, offBuf, charBuf};
      col->alphabet = merged;
    } else {
      col->desc.encoding = VB2_FLAT;
      const int w = t == VB2_BOOLEAN ? 1 : widthOf(t);  // BOOLEAN: one byte per row until the final pack
      auto values = allocDevice(static_cast<size_t>(total) * w, stream);
      int64_t off = 0;
      for (auto& b : batches) {
        const int64_t bn = b->size();
        FlatColumn f = flattenColumn(b->column(c), nullptr, bn, stream);
        VB2_CU(cudaMemcpyAsync(values->as<uint8_t>() + off * w, f.values->data(), static_cast<size_t>(bn) * w, cudaMemcpyDeviceToDevice, stream));
        if (validBytes) {
          if (f.nulls) kernelCheck(vb2k_unpack_bits(f.nulls->as<uint64_t>(), bn, validBytes->as<uint8_t>() + off, stream));
          else VB2_CU(cudaMemsetAsync(validBytes->as<uint8_t>() + off, 1, static_cast<size_t>(bn), stream));
        }
        off += bn;
      }
      if (t == VB2_BOOLEAN) {
        auto packed = allocDevice(bits::nbytes(total), stream);
        kernelCheck(vb2k_pack_bools(values->as<uint8_t>(), total, packed->as<uint64_t>(), stream));
        values = packed;
      }
      col->desc.values = values->data();
      col->owners = {values};
    }
    if (validBytes) {
      auto bitsBuf = allocDevice(bits::nbytes(total), stream);
      kernelCheck(vb2k_pack_bools(validBytes->as<uint8_t>(), total, bitsBuf->as<uint64_t>(), stream));
      col->desc.nulls = bitsBuf->as<uint64_t>();
      col->owners.push_back(bitsBuf);
    }
    cols.push_back(col);
  }
  return std::make_shared<B200Vector>(pool, batches[0]->type(), static_cast<vector_size_t>(total), std::move(cols), stream);
}

// ---- B200HashBuild ----------------------------------------------------------------------------
B200HashBuild::B200HashBuild(int32_t id, exec::DriverCtx* ctx, const exec::HashBuild& cpu)
    : Operator(ctx, nullptr, id, cpu.planNodeId(), "B200HashBuild"), node_(cpu.node()), bridge_(cpu.joinBridge()) {}

void B200HashBuild::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
}

void B200HashBuild::addInput(RowVectorPtr input) {
  B200_NVTX_OPERATOR_RANGE("addInput");
  auto in = std::dynamic_pointer_cast<B200Vector>(input);
  VELOX_CHECK(in != nullptr, "B200HashBuild expects device-resident input");
  orderAfterProducer(*in, dev_->stream);
  batches_.push_back(std::move(in));
}

namespace {
uint64_t nextPow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}
}  // namespace

void B200HashBuild::noMoreInput() {
  B200_NVTX_OPERATOR_RANGE("noMoreInput");
  Operator::noMoreInput();
  // The last build driver to get here gathers the rows of all its peers and builds the one table the
  // probe side reads (exec/HashBuild.cpp:819-993 finishHashBuild); the others park until it has.
  exec::Task* task = driverCtx_->task;
  if (task && driverCtx_->driver && task->numDrivers(driverCtx_->pipelineId) > 1) {
    std::vector<exec::ContinuePromise> promises;
    std::vector<std::shared_ptr<exec::Driver>> peers;
    if (!task->allPeersFinished(planNodeId(), driverCtx_->driver, &peerFuture_, promises, peers)) return;
    for (auto& peer : peers) {
      auto* op = dynamic_cast<B200HashBuild*>(peer->findOperator(planNodeId()));
      VELOX_CHECK(op != nullptr, "peer driver without a B200HashBuild for this join");
      for (auto& b : op->takeBatches()) {
        orderAfterProducer(*b, dev_->stream);
        batches_.push_back(std::move(b));
      }
    }
    addRuntimeStat("b200.buildPeers", exec::RuntimeCounter{static_cast<int64_t>(peers.size()) + 1});
    try {
      buildTable();
    } catch (...) {
      for (auto& p : promises) p.setValue();
      throw;
    }
    // the peers' rows are referenced by this stream's kernels only from here on: release them
    for (auto& p : promises) p.setValue();
    return;
  }
  buildTable();
}

std::vector<vb2_column> keyedJoinColumns(const B200Vector& batch, const std::vector<int32_t>& keyColumns, JoinTableHolder& holder, bool build,
                                         std::vector<KeyedLutCache>* cache, std::vector<DeviceBufferPtr>& keep, cudaStream_t stream) {
  std::vector<vb2_column> cols;
  if (build) {
    holder.keyIsVarchar.assign(keyColumns.size(), false);
    holder.varcharIds.assign(keyColumns.size(), {});
  }
  if (cache && cache->size() < keyColumns.size()) cache->resize(keyColumns.size());
  for (size_t k = 0; k < keyColumns.size(); ++k) {
    const DeviceColumn& col = *batch.column(keyColumns[k]);
    vb2_column d = col.desc;
    if (d.type == VB2_VARCHAR) {
      if (d.encoding == VB2_FLAT || !col.alphabet)
        VELOX_UNSUPPORTED("join on flat VARCHAR keys with more than 65536 distinct values per batch (flat strings are dictionary-encoded on upload up to that size)");
      if (build) holder.keyIsVarchar[k] = true;
      VELOX_CHECK(holder.keyIsVarchar[k], "join key types differ between the build and the probe side");
      auto& ids = holder.varcharIds[k];
      DeviceBufferPtr lutBuf;
      if (cache && (*cache)[k].alphabet == col.alphabet) {
        lutBuf = (*cache)[k].lut;
      } else {
        std::vector<int32_t> lut(col.alphabet->values.size());
        for (size_t i = 0; i < lut.size(); ++i) {
          if (col.alphabet->nulls[i]) { lut[i] = 0; continue; }  // NULL entries: dict_nulls decides, the id is unused
          auto it = ids.find(col.alphabet->values[i]);
          if (it == ids.end()) {
            if (!build) { lut[i] = -1; continue; }
            it = ids.emplace(col.alphabet->values[i], static_cast<int32_t>(ids.size())).first;
          }
          lut[i] = it->second;
        }
        lutBuf = allocDevice(lut.size() * 4 + 4, stream);
        VB2_CU(cudaMemcpyAsync(lutBuf->data(), lut.data(), lut.size() * 4, cudaMemcpyHostToDevice, stream));  // pageable: staged before return
        if (cache) (*cache)[k] = KeyedLutCache{col.alphabet, lutBuf};
      }
      keep.push_back(lutBuf);
      d.type = VB2_INTEGER;
      d.values = lutBuf->data();
      d.aux = nullptr;
      if (d.encoding == VB2_CONSTANT) d.dict_size = 1;
    } else {
      VELOX_CHECK(!holder.keyIsVarchar[k], "join key types differ between the build and the probe side");
    }
    cols.push_back(d);
  }
  return cols;
}

// Keyed mode: ids of the build rows' key tuples (find-or-insert into a keyed table), then the usual chains
// over an array-mode table addressed by those ids.
void B200HashBuild::buildKeyedTable(const std::shared_ptr<JoinTableHolder>& holder, const std::vector<int32_t>& keys, int64_t n) {
  cudaStream_t st = dev_->stream;
  VELOX_CHECK(keys.size() <= VB2_KEYED_MAX_KEYS, "at most " + std::to_string(VB2_KEYED_MAX_KEYS) + " join keys in keyed mode");
  holder->keyed = true;
  std::vector<DeviceBufferPtr> keep;
  std::vector<vb2_column> cols = keyedJoinColumns(*holder->rows, keys, *holder, true, nullptr, keep, st);
  for (int32_t k : keys)
    if (holder->rows->column(k)->mayHaveNulls()) holder->hasNullKeys = true;  // conservative: a NULL-capable key column
  const int32_t nk = static_cast<int32_t>(keys.size());
  vb2_group_table& kt = holder->keyedTable;
  kt.capacity = static_cast<int64_t>(nextPow2(static_cast<uint64_t>(n) * 2 + 16));
  VELOX_CHECK(kt.capacity <= (1ll << 31), "join build side above 2^30 rows in keyed mode");
  kt.row_words = (nk + 2 + 3) / 4 * 4;
  kt.hash_mode = VB2_GROUP_KEYED;
  auto rowsBuf = allocDevice(static_cast<size_t>(kt.capacity) * kt.row_words * 8, st);
  kt.rows = rowsBuf->as<uint64_t>();
  std::vector<uint64_t> init(kt.row_words, 0);
  init[0] = VB2_EMPTY_KEY;
  kernelCheck(vb2k_group_table_init(&kt, init.data(), st));
  holder->owners.push_back(rowsBuf);
  auto ids = allocDevice(static_cast<size_t>(n) * 8, st);
  auto valid = allocDevice(bits::nbytes(n), st);
  auto flags = allocDeviceZeroed(16, st);
  kernelCheck(vb2k_keyed_key_ids(&kt, cols.data(), nk, n, 1, ids->as<uint64_t>(), valid->as<uint64_t>(), nullptr, flags->as<int32_t>(), st));
  vb2_join_table& t = holder->table;
  t.mode = 0;
  t.key_min = 0;
  t.capacity = kt.capacity;
  auto head = allocDeviceZeroed(static_cast<size_t>(t.capacity) * 4, st);
  auto next = allocDevice(static_cast<size_t>(n) * 4 + 4, st);
  t.head = head->as<int32_t>();
  t.next = next->as<int32_t>();
  t.build_rows = n;
  holder->owners.push_back(head);
  holder->owners.push_back(next);
  kernelCheck(vb2k_join_build(&t, ids->as<uint64_t>(), valid->as<uint64_t>(), n, flags->as<int32_t>() + 2, st));
  int32_t h[4];
  VB2_CU(cudaMemcpyAsync(h, flags->data(), 16, cudaMemcpyDeviceToHost, st));
  VB2_CU(cudaStreamSynchronize(st));
  VELOX_CHECK(h[0] == 0 && h[2] == 0, "join table build failed (table full)");
  holder->hasDuplicateKeys = h[3] != 0;
  addRuntimeStat("b200.joinTableMode", exec::RuntimeCounter{2});
  addRuntimeStat("b200.joinTableSlots", exec::RuntimeCounter{t.capacity});
  bridge_->setHashTable(holder, holder->hasNullKeys);
}

void B200HashBuild::buildTable() {
  cudaStream_t st = dev_->stream;
  auto holder = std::make_shared<JoinTableHolder>();
  holder->stream = st;
  const std::vector<int32_t> keys = resolveJoin(*node_).rightKeys;
  const auto& buildType = node_->sources()[1]->outputType();
  bool keyedTypes = false;  // DOUBLE / VARCHAR keys have no value-id range: keyed mode
  for (int32_t k : keys) {
    const TypeKind kind = buildType->childAt(k)->kind();
    if (kind == TypeKind::DOUBLE || kind == TypeKind::VARCHAR) keyedTypes = true;
    else if (kind != TypeKind::BIGINT && kind != TypeKind::INTEGER && kind != TypeKind::BOOLEAN)
      VELOX_UNSUPPORTED("join keys of type " + buildType->childAt(k)->toString() + " (BIGINT/INTEGER/DATE/BOOLEAN/DOUBLE/VARCHAR are supported)");
  }
  if (batches_.empty()) {
    // empty build side: a table nobody can hit
    holder->layout.mins.assign(keys.size(), 0);
    holder->layout.ranges.assign(keys.size(), 1);
    holder->layout.mults.assign(keys.size(), 1);
    holder->table.mode = 0;
    holder->table.capacity = 1;
    auto head = allocDeviceZeroed(4, st);
    auto next = allocDeviceZeroed(4, st);
    holder->table.head = head->as<int32_t>();
    holder->table.next = next->as<int32_t>();
    holder->owners = {head, next};
    // zero-row build side with the right column types (outer joins still project its columns as NULL)
    std::vector<DeviceColumnPtr> cols;
    auto dummy = allocDeviceZeroed(16, st);
    for (uint32_t c = 0; c < buildType->size(); ++c) {
      auto col = std::make_shared<DeviceColumn>();
      col->type = buildType->childAt(c);
      col->desc.type = veloxTypeToVb2(col->type);
      col->desc.encoding = VB2_FLAT;
      col->desc.size = 0;
      col->desc.values = dummy->data();
      col->desc.aux = dummy->data();
      col->owners = {dummy};
      cols.push_back(col);
    }
    holder->rows = std::make_shared<B200Vector>(pool(), buildType, 0, std::move(cols), st);
    bridge_->setHashTable(holder, false);
    return;
  }
  holder->rows = concatBatches(batches_, pool(), st);
  batches_.clear();
  const int64_t n = holder->rows->size();
  holder->numRows = n;
  // value ranges of the key columns -> layout (VectorHasher range mode, exec/VectorHasher.cpp:923)
  KeyLayout& lay = holder->layout;
  uint64_t product = 1;
  bool overflow = keyedTypes;
  for (int32_t k : keys) {
    if (keyedTypes) break;
    int64_t lo, hi, nn;
    columnMinMax(*holder->rows->column(k), n, st, lo, hi, nn);
    if (nn == 0) { lo = 0; hi = 0; }
    if (nn < n) holder->hasNullKeys = true;
    const unsigned __int128 range = static_cast<unsigned __int128>(static_cast<__int128>(hi) - lo) + 2;  // + NULL id
    lay.mins.push_back(lo);
    if (range > (static_cast<unsigned __int128>(1) << 62)) overflow = true;
    lay.ranges.push_back(static_cast<uint64_t>(range));
    if (!overflow) {
      const unsigned __int128 p = static_cast<unsigned __int128>(product) * range;
      if (p > (static_cast<unsigned __int128>(1) << 62)) overflow = true;
      else product = static_cast<uint64_t>(p);
    }
  }
  if (overflow) {
    buildKeyedTable(holder, keys, n);
    return;
  }
  lay.product = product;
  lay.mults.assign(keys.size(), 1);
  for (int i = static_cast<int>(keys.size()) - 2; i >= 0; --i) lay.mults[i] = lay.mults[i + 1] * lay.ranges[i + 1];

  // Array mode while the packed key space is at most 16x the row count and <= 2^28 slots: the
  // reference caps kArray at 2M entries for CPU caches (exec/HashTable.h:146); HBM + a 126 MB L2
  // move that limit (a 20 M-slot int32 table is 80 MB and stays L2 resident).
  vb2_join_table& t = holder->table;
  const uint64_t arrayLimit = std::min<uint64_t>(1ull << 28, std::max<uint64_t>(1ull << 16, static_cast<uint64_t>(n) * 16));
  if (product <= arrayLimit) {
    t.mode = 0;
    t.key_min = 0;
    t.capacity = static_cast<int64_t>(product);
  } else {
    t.mode = 1;
    t.capacity = static_cast<int64_t>(nextPow2(static_cast<uint64_t>(n) * 2 + 16));
    auto keysBuf = allocDevice(static_cast<size_t>(t.capacity) * 8, st);
    kernelCheck(vb2k_fill_u64(keysBuf->as<uint64_t>(), t.capacity, VB2_EMPTY_KEY, st));
    t.keys = keysBuf->as<uint64_t>();
    holder->owners.push_back(keysBuf);
  }
  auto head = allocDeviceZeroed(static_cast<size_t>(t.capacity) * 4, st);
  auto next = allocDevice(static_cast<size_t>(n) * 4 + 4, st);  // every inserted row writes its own entry; others are unreachable
  t.head = head->as<int32_t>();
  t.next = next->as<int32_t>();
  t.build_rows = n;
  holder->owners.push_back(head);
  holder->owners.push_back(next);
  auto flags = allocDeviceZeroed(8, st);
  if (t.mode == 0 && keys.size() == 1) {
    // one key column, array mode: value ids and insertion in one pass over the key column
    kernelCheck(vb2k_join_build_array_direct(t.head, t.next, t.capacity, &holder->rows->column(keys[0])->desc, lay.mins[0], n, flags->as<int32_t>(), st));
  } else {
    NormalizedKeys nk = normalizeKeys(*holder->rows, keys, lay, nullptr, n, true, st);
    kernelCheck(vb2k_join_build(&t, nk.keys->as<uint64_t>(), nk.valid->as<uint64_t>(), n, flags->as<int32_t>(), st));
  }
  int32_t h[2];
  VB2_CU(cudaMemcpyAsync(h, flags->data(), 8, cudaMemcpyDeviceToHost, st));
  VB2_CU(cudaStreamSynchronize(st));
  VELOX_CHECK(h[0] == 0, "join table build failed (table full)");
  holder->hasDuplicateKeys = h[1] != 0;
  addRuntimeStat("b200.joinTableMode", exec::RuntimeCounter{t.mode});
  addRuntimeStat("b200.joinTableSlots", exec::RuntimeCounter{t.capacity});
  bridge_->setHashTable(holder, holder->hasNullKeys);
}

// ---- B200HashProbe ----------------------------------------------------------------------------
B200HashProbe::B200HashProbe(int32_t id, exec::DriverCtx* ctx, const exec::HashProbe& cpu)
    : Operator(ctx, cpu.outputType(), id, cpu.planNodeId(), "B200HashProbe"), node_(cpu.node()), plan_(resolveJoin(*cpu.node())), bridge_(cpu.joinBridge()) {
  if (node_->filter()) {
    std::vector<std::string> names = node_->sources()[0]->outputType()->names();
    std::vector<TypePtr> types = node_->sources()[0]->outputType()->children();
    const auto& bt = node_->sources()[1]->outputType();
    for (uint32_t i = 0; i < bt->size(); ++i) { names.push_back(bt->nameOf(i)); types.push_back(bt->childAt(i)); }
    filterProgram_ = std::make_unique<CompiledProgram>(compileExprs({node_->filter()}, true, ROW(names, types)));
  }
}

void B200HashProbe::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
  errorFlag_ = allocDeviceZeroed(8, dev_->stream);
  if (filterProgram_) filterProgram_->uploadConstants(dev_->stream);
}

exec::BlockingReason B200HashProbe::isBlocked(exec::ContinueFuture* future) {
  if (table_) return exec::BlockingReason::kNotBlocked;
  auto t = bridge_->tableOrFuture(future);
  if (!t) return exec::BlockingReason::kWaitForJoinBuild;
  table_ = std::dynamic_pointer_cast<JoinTableHolder>(t->waveTable);
  VELOX_CHECK(table_ != nullptr, "the join bridge holds a table of another backend");
  return exec::BlockingReason::kNotBlocked;
}

B200VectorPtr B200HashProbe::apply(const B200VectorPtr& in) {
  cudaStream_t st = dev_->stream;
  const int64_t n = in->size();
  const JoinTableHolder& jt = *table_;
  const core::JoinType type = node_->joinType();
  // build and probe streams differ only across drivers; the bridge hand-off synchronised the build
  NormalizedKeys nk;
  if (jt.keyed) {
    // keyed mode: the probe row's key is the slot of its key tuple in the build side's keyed table (absent: no match)
    std::vector<DeviceBufferPtr> keep;
    std::vector<vb2_column> cols = keyedJoinColumns(*in, plan_.leftKeys, *table_, false, &keyedLuts_, keep, st);
    nk.keys = allocDevice(static_cast<size_t>(n) * 8, st);
    nk.valid = allocDevice(bits::nbytes(n), st);
    kernelCheck(vb2k_keyed_key_ids(&jt.keyedTable, cols.data(), static_cast<int32_t>(cols.size()), n, 0, nk.keys->as<uint64_t>(), nk.valid->as<uint64_t>(),
                                   nullptr, errorFlag_->as<int32_t>(), st));
  } else {
    nk = normalizeKeys(*in, plan_.leftKeys, jt.layout, nullptr, n, true, st);
  }
  int64_t pairs = 0;
  DeviceBufferPtr probeRows, buildRows;
  if (!jt.hasDuplicateKeys) {
    // unique build keys (every TPC-H primary-key join): ONE probe of the table per row. The matches of a
    // warp are a ballot word; the ordered expansion of that bitmap gives the matching probe rows, their
    // build rows are gathered from the per-row hits.
    auto hitBits = allocDevice(bits::nbytes(n), st);
    auto hits = allocDevice(static_cast<size_t>(n) * 4, st);
    kernelCheck(vb2k_join_probe_unique(&jt.table, nk.keys->as<uint64_t>(), nk.valid->as<uint64_t>(), n, hitBits->as<uint64_t>(), hits->as<int32_t>(), st));
    auto sel = allocDevice(static_cast<size_t>(n) * 4, st);
    auto cnt = allocDevice(8, st);
    const size_t wsb = vb2k_bits_to_indices_workspace(n);
    auto ws = allocDevice(wsb, st);
    kernelCheck(vb2k_bits_to_indices(hitBits->as<uint64_t>(), n, sel->as<int32_t>(), cnt->as<int64_t>(), ws->data(), wsb, st));
    VB2_CU(cudaMemcpyAsync(&pairs, cnt->data(), 8, cudaMemcpyDeviceToHost, st));
    VB2_CU(cudaStreamSynchronize(st));
    if (pairs > 0) {
      probeRows = sel;
      buildRows = allocDevice(static_cast<size_t>(pairs) * 4, st);
      kernelCheck(vb2k_gather(hits->data(), sel->as<int32_t>(), pairs, 4, buildRows->data(), st));
    }
    addRuntimeStat("b200.uniqueKeyProbes", exec::RuntimeCounter{1});
  } else {
    // duplicate build keys: count the chain of every probe row, scan, emit the pairs in probe order
    auto counts = allocDevice(static_cast<size_t>(n) * 4, st);
    kernelCheck(vb2k_join_probe_count(&jt.table, nk.keys->as<uint64_t>(), nk.valid->as<uint64_t>(), n, counts->as<int32_t>(), st));
    auto offsets = allocDevice(static_cast<size_t>(n) * 8, st);
    auto total = allocDevice(8, st);
    const size_t wsBytes = vb2k_scan_workspace(n);
    auto ws = allocDevice(wsBytes, st);
    kernelCheck(vb2k_exclusive_scan_i32(counts->as<int32_t>(), n, offsets->as<int64_t>(), total->as<int64_t>(), ws->data(), wsBytes, st));
    VB2_CU(cudaMemcpyAsync(&pairs, total->data(), 8, cudaMemcpyDeviceToHost, st));
    VB2_CU(cudaStreamSynchronize(st));
    VELOX_CHECK(pairs < (1ll << 31), "join output above 2^31 rows for one batch");
    if (pairs > 0) {
      probeRows = allocDevice(static_cast<size_t>(pairs) * 4, st);
      buildRows = allocDevice(static_cast<size_t>(pairs) * 4, st);
      kernelCheck(vb2k_join_probe_emit(&jt.table, nk.keys->as<uint64_t>(), nk.valid->as<uint64_t>(), n, offsets->as<int64_t>(),
                                       probeRows->as<int32_t>(), buildRows->as<int32_t>(), st));
    }
  }
  // optional join filter over (probe ++ build) columns of the candidate pairs
  if (filterProgram_ && pairs > 0) {
    std::vector<vb2_column> cols;
    std::vector<DeviceColumnPtr> keep;
    for (auto& c : in->columns()) { keep.push_back(wrapColumn(c, probeRows, pairs, st)); cols.push_back(keep.back()->desc); }
    for (auto& c : jt.rows->columns()) { keep.push_back(wrapColumn(c, buildRows, pairs, st)); cols.push_back(keep.back()->desc); }
    const vb2_program prog = filterProgram_->view();
    auto bitsBuf = allocDevice(bits::nbytes(pairs), st);
    kernelCheck(vb2k_eval_filter(&prog, cols.data(), static_cast<int32_t>(cols.size()), pairs, bitsBuf->as<uint64_t>(), errorFlag_->as<int32_t>(), st));
    auto sel = allocDevice(static_cast<size_t>(pairs) * 4, st);
    auto cnt = allocDevice(8, st);
    const size_t w2 = vb2k_bits_to_indices_workspace(pairs);
    auto ws2 = allocDevice(w2, st);
    kernelCheck(vb2k_bits_to_indices(bitsBuf->as<uint64_t>(), pairs, sel->as<int32_t>(), cnt->as<int64_t>(), ws2->data(), w2, st));
    int64_t kept = 0;
    VB2_CU(cudaMemcpyAsync(&kept, cnt->data(), 8, cudaMemcpyDeviceToHost, st));
    checkDeviceError(errorFlag_, st, "join filter");
    if (kept != pairs) {
      auto p2 = allocDevice(static_cast<size_t>(kept ? kept : 1) * 4, st);
      auto b2 = allocDevice(static_cast<size_t>(kept ? kept : 1) * 4, st);
      kernelCheck(vb2k_gather(probeRows->data(), sel->as<int32_t>(), kept, 4, p2->data(), st));
      kernelCheck(vb2k_gather(buildRows->data(), sel->as<int32_t>(), kept, 4, b2->data(), st));
      probeRows = p2;
      buildRows = b2;
      pairs = kept;
    }
  }
  int64_t numOut = pairs;
  DeviceBufferPtr outProbe = probeRows, outBuild = buildRows;
  bool buildSideNull = false;
  if (type != core::JoinType::kInner) {
    // matched[r] = probe row r has at least one surviving pair
    auto matched = allocDeviceZeroed(static_cast<size_t>(n) * 4 + 4, st);
    if (pairs > 0) {
      auto ones = allocDevice(static_cast<size_t>(pairs) * 4, st);
      kernelCheck(vb2k_fill_i32(ones->as<int32_t>(), pairs, 1, st));
      kernelCheck(vb2k_scatter(ones->data(), nullptr, probeRows->as<int32_t>(), pairs, 4, matched->data(), st));
    }
    if (type == core::JoinType::kLeft) {
      // unmatched probe rows appended after the matches (multiset semantics; build side NULL)
      auto missBits = allocDevice(bits::nbytes(n), st);
      auto m64 = allocDevice(static_cast<size_t>(n) * 8, st);
      kernelCheck(vb2k_widen_not_i32(matched->as<int32_t>(), n, m64->as<int64_t>(), st));
      kernelCheck(vb2k_positive_bits(m64->as<int64_t>(), n, missBits->as<uint64_t>(), st));
      auto missIdx = allocDevice(static_cast<size_t>(n) * 4, st);
      auto cnt = allocDevice(8, st);
      const size_t w3 = vb2k_bits_to_indices_workspace(n);
      auto ws3 = allocDevice(w3, st);
      kernelCheck(vb2k_bits_to_indices(missBits->as<uint64_t>(), n, missIdx->as<int32_t>(), cnt->as<int64_t>(), ws3->data(), w3, st));
      int64_t misses = 0;
      VB2_CU(cudaMemcpyAsync(&misses, cnt->data(), 8, cudaMemcpyDeviceToHost, st));
      VB2_CU(cudaStreamSynchronize(st));
      numOut = pairs + misses;
      if (numOut == 0) return nullptr;
      outProbe = allocDevice(static_cast<size_t>(numOut) * 4, st);
      outBuild = allocDevice(static_cast<size_t>(numOut) * 4, st);
      if (pairs) {
        VB2_CU(cudaMemcpyAsync(outProbe->data(), probeRows->data(), static_cast<size_t>(pairs) * 4, cudaMemcpyDeviceToDevice, st));
        VB2_CU(cudaMemcpyAsync(outBuild->data(), buildRows->data(), static_cast<size_t>(pairs) * 4, cudaMemcpyDeviceToDevice, st));
      }
      if (misses) {
        VB2_CU(cudaMemcpyAsync(outProbe->as<int32_t>() + pairs, missIdx->data(), static_cast<size_t>(misses) * 4, cudaMemcpyDeviceToDevice, st));
        kernelCheck(vb2k_fill_i32(outBuild->as<int32_t>() + pairs, misses, -1, st));
        buildSideNull = true;
      }
    } else {
      // semi: matched probe rows once; anti: unmatched probe rows
      auto m64 = allocDevice(static_cast<size_t>(n) * 8, st);
      if (type == core::JoinType::kAnti) kernelCheck(vb2k_widen_not_i32(matched->as<int32_t>(), n, m64->as<int64_t>(), st));
      else kernelCheck(vb2k_widen_i32(matched->as<int32_t>(), n, m64->as<int64_t>(), st));
      auto bitsBuf = allocDevice(bits::nbytes(n), st);
      kernelCheck(vb2k_positive_bits(m64->as<int64_t>(), n, bitsBuf->as<uint64_t>(), st));
      auto idx = allocDevice(static_cast<size_t>(n) * 4, st);
      auto cnt = allocDevice(8, st);
      const size_t w3 = vb2k_bits_to_indices_workspace(n);
      auto ws3 = allocDevice(w3, st);
      kernelCheck(vb2k_bits_to_indices(bitsBuf->as<uint64_t>(), n, idx->as<int32_t>(), cnt->as<int64_t>(), ws3->data(), w3, st));
      VB2_CU(cudaMemcpyAsync(&numOut, cnt->data(), 8, cudaMemcpyDeviceToHost, st));
      VB2_CU(cudaStreamSynchronize(st));
      outProbe = idx;
      outBuild = nullptr;
    }
  }
  if (numOut == 0) return nullptr;
  DeviceBufferPtr buildValid;
  if (buildSideNull) {
    // rows with build index -1 are NULL on the build side: validity = index >= 0, indices clamped to 0
    buildValid = allocDevice(bits::nbytes(numOut), st);
    auto clamped = allocDevice(static_cast<size_t>(numOut) * 4, st);
    kernelCheck(vb2k_index_validity(outBuild->as<int32_t>(), numOut, buildValid->as<uint64_t>(), clamped->as<int32_t>(), st));
    outBuild = clamped;
  }
  std::vector<DeviceColumnPtr> outCols;
  for (auto& o : plan_.outputs) {
    if (o.fromProbe) {
      outCols.push_back(wrapColumn(in->column(o.column), outProbe, numOut, st));
    } else {
      VELOX_CHECK(outBuild != nullptr, "semi/anti joins project probe columns only");
      auto col = wrapColumn(jt.rows->column(o.column), outBuild, numOut, st);
      if (buildSideNull) {
        if (col->desc.encoding == VB2_CONSTANT || col->desc.nulls) VELOX_UNSUPPORTED("left join over a constant / null-wrapped build column");
        col->desc.nulls = buildValid->as<uint64_t>();
        col->owners.push_back(buildValid);
      }
      outCols.push_back(col);
    }
  }
  return std::make_shared<B200Vector>(pool(), outputType_, static_cast<vector_size_t>(numOut), std::move(outCols), st);
}

RowVectorPtr B200HashProbe::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!input_) return nullptr;
  auto in = std::dynamic_pointer_cast<B200Vector>(input_);
  input_ = nullptr;
  VELOX_CHECK(in != nullptr, "B200HashProbe expects device-resident input");
  orderAfterProducer(*in, dev_->stream);
  return apply(in);
}

}  // namespace velox_b200
