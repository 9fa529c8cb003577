// B200OrderBy / B200TopN: ORDER BY [LIMIT n] on the device.
// Replace exec::OrderBy (velox/exec/OrderBy.cpp:60-110 + SortBuffer.cpp) and exec::TopN
// (velox/exec/TopN.cpp:60-150). The sort yields a row order (vb2k_sort_order, csrc/sort.cu); the
// output columns are dictionary wraps of the input columns over that order — the same zero-copy
// wrapChild shape FilterProject emits — so no payload column is moved by the sort.
#include <algorithm>
#include <numeric>

#include "join.h"
#include "operators.h"
#include "plan_resolve.h"

namespace velox_b200 {

namespace {

// Sort keys of `batch` for vb2k_sort_order. Non-VARCHAR keys are flattened (values + validity);
// dictionary VARCHAR keys become INTEGER rank codes: the alphabet is sorted bytewise on the host
// (StringView::compare, velox/type/StringView.h) and every entry replaced by its rank.
std::vector<vb2_sort_key> sortKeysOf(const B200Vector& batch, const std::vector<int32_t>& channels, const std::vector<core::SortOrder>& orders,
                                     std::vector<DeviceBufferPtr>& keep, cudaStream_t st) {
  std::vector<vb2_sort_key> keys;
  const int64_t n = batch.size();
  for (size_t k = 0; k < channels.size(); ++k) {
    const DeviceColumnPtr& col = batch.column(channels[k]);
    vb2_sort_key key{};
    key.ascending = orders[k].isAscending() ? 1 : 0;
    key.nulls_first = orders[k].isNullsFirst() ? 1 : 0;
    if (col->desc.type == VB2_VARCHAR) {
      if (col->desc.encoding == VB2_FLAT || !col->alphabet) VELOX_UNSUPPORTED("ORDER BY on flat (non-dictionary) VARCHAR keys");
      const auto& alpha = *col->alphabet;
      std::vector<int32_t> byRank(alpha.values.size());
      std::iota(byRank.begin(), byRank.end(), 0);
      std::sort(byRank.begin(), byRank.end(), [&](int32_t a, int32_t b) { return alpha.values[a] < alpha.values[b]; });
      std::vector<int32_t> rank(alpha.values.size() + 1, 0);
      int32_t r = -1;
      for (size_t i = 0; i < byRank.size(); ++i) {
        if (i == 0 || alpha.values[byRank[i]] != alpha.values[byRank[i - 1]]) ++r;  // equal strings share a rank (ties keep input order)
        rank[byRank[i]] = r;
      }
      auto codes = allocDevice(static_cast<size_t>(n) * 4, st);
      auto valid = allocDevice(static_cast<size_t>(n), st);
      kernelCheck(vb2k_dictionary_codes(&col->desc, n, codes->as<int32_t>(), valid->as<uint8_t>(), st));
      auto lut = allocDevice(rank.size() * 4, st);
      auto staged = acquirePinned(rank.size() * 4);
      std::memcpy(staged.get(), rank.data(), rank.size() * 4);
      VB2_CU(cudaMemcpyAsync(lut->data(), staged.get(), rank.size() * 4, cudaMemcpyHostToDevice, st));
      auto ranks = allocDevice(static_cast<size_t>(n) * 4, st);
      kernelCheck(vb2k_gather(lut->data(), codes->as<int32_t>(), n, 4, ranks->data(), st));
      key.type = VB2_INTEGER;
      key.values = ranks->data();
      int bits = 1;
      while ((1ll << bits) < static_cast<int64_t>(rank.size())) ++bits;
      key.significant_bits = bits;
      keep.push_back(ranks);
      keep.push_back(lut);
      keep.push_back(codes);
      // the pinned staging block must outlive the copy: park it with the buffers (released after the sort's sync)
      if (col->mayHaveNulls()) {
        auto bitsBuf = allocDevice(bits::nbytes(n), st);
        kernelCheck(vb2k_pack_bools(valid->as<uint8_t>(), n, bitsBuf->as<uint64_t>(), st));
        key.nulls = bitsBuf->as<uint64_t>();
        keep.push_back(bitsBuf);
      }
      keep.push_back(valid);
      VB2_CU(cudaStreamSynchronize(st));  // the staging block is reused by the pool once `staged` goes out of scope
    } else if (col->desc.encoding == VB2_FLAT) {
      key.type = col->desc.type;
      key.values = col->desc.values;
      key.nulls = col->desc.nulls;
    } else {
      FlatColumn f = flattenColumn(col, nullptr, n, st);
      key.type = f.type;
      if (f.type == VB2_BOOLEAN) {
        // flattenColumn writes BOOLEAN one byte per row; the sort reads bit-packed values
        auto packed = allocDevice(bits::nbytes(n), st);
        kernelCheck(vb2k_pack_bools(f.values->as<uint8_t>(), n, packed->as<uint64_t>(), st));
        keep.push_back(packed);
        key.values = packed->data();
      } else {
        key.values = f.values->data();
      }
      key.nulls = f.nulls ? f.nulls->as<uint64_t>() : nullptr;
      keep.push_back(f.values);
      if (f.nulls) keep.push_back(f.nulls);
    }
    keys.push_back(key);
  }
  return keys;
}

// Row order of `batch` under the sort keys (device int32[n]).
DeviceBufferPtr sortOrder(const B200Vector& batch, const std::vector<int32_t>& channels, const std::vector<core::SortOrder>& orders, cudaStream_t st) {
  const int64_t n = batch.size();
  std::vector<DeviceBufferPtr> keep;
  std::vector<vb2_sort_key> keys = sortKeysOf(batch, channels, orders, keep, st);
  auto order = allocDevice(static_cast<size_t>(std::max<int64_t>(n, 1)) * 4, st);
  const size_t wsBytes = vb2k_sort_order_workspace(n, static_cast<int32_t>(keys.size()));
  auto ws = allocDevice(wsBytes, st);
  kernelCheck(vb2k_sort_order(keys.data(), static_cast<int32_t>(keys.size()), n, order->as<int32_t>(), ws->data(), wsBytes, st));
  return order;  // `keep` and `ws` are freed stream-ordered after the kernels above
}

// Rows order[0..m) of `batch`: every column wrapped over the order (zero copy).
B200VectorPtr takeRows(const B200VectorPtr& batch, const DeviceBufferPtr& order, int64_t m, memory::MemoryPool* pool, cudaStream_t st) {
  std::vector<DeviceColumnPtr> cols;
  for (auto& c : batch->columns()) cols.push_back(wrapColumn(c, order, m, st));
  return std::make_shared<B200Vector>(pool, batch->type(), static_cast<vector_size_t>(m), std::move(cols), st);
}

// Dense copy of a (wrapped) batch: the kept top-N rows must not pin the batches they came from.
B200VectorPtr compact(const B200VectorPtr& batch, memory::MemoryPool* pool, cudaStream_t st) {
  const int64_t n = batch->size();
  std::vector<DeviceColumnPtr> cols;
  for (auto& c : batch->columns()) {
    if (c->desc.type == VB2_VARCHAR || c->desc.encoding == VB2_CONSTANT) {
      // dictionary codes over a small alphabet: the composed index array is already n-sized
      cols.push_back(c);
      continue;
    }
    FlatColumn f = flattenColumn(c, nullptr, n, st);
    auto col = std::make_shared<DeviceColumn>();
    col->type = c->type;
    col->desc.type = c->desc.type;
    col->desc.encoding = VB2_FLAT;
    col->desc.size = n;
    if (f.type == VB2_BOOLEAN) {
      auto packed = allocDevice(bits::nbytes(n), st);
      kernelCheck(vb2k_pack_bools(f.values->as<uint8_t>(), n, packed->as<uint64_t>(), st));
      f.values = packed;
    }
    col->desc.values = f.values->data();
    col->owners = {f.values};
    if (f.nulls) {
      col->desc.nulls = f.nulls->as<uint64_t>();
      col->owners.push_back(f.nulls);
    }
    cols.push_back(col);
  }
  return std::make_shared<B200Vector>(pool, batch->type(), static_cast<vector_size_t>(n), std::move(cols), st);
}

std::vector<int32_t> channelsOf(const RowTypePtr& type, const std::vector<core::FieldAccessTypedExprPtr>& keys) {
  std::vector<int32_t> out;
  for (auto& k : keys) out.push_back(channelOf(type, *k));
  VELOX_CHECK(out.size() <= VB2_SORT_MAX_KEYS, "ORDER BY with more than 8 keys");
  return out;
}

}  // namespace

// ---- B200OrderBy -------------------------------------------------------------------------------
B200OrderBy::B200OrderBy(int32_t id, exec::DriverCtx* ctx, std::shared_ptr<const core::OrderByNode> node)
    : Operator(ctx, node->outputType(), id, node->id(), "B200OrderBy"), node_(std::move(node)) {
  channels_ = channelsOf(node_->sources()[0]->outputType(), node_->sortingKeys());
}
void B200OrderBy::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
}
void B200OrderBy::addInput(RowVectorPtr input) {
  B200_NVTX_OPERATOR_RANGE("addInput");
  auto in = std::dynamic_pointer_cast<B200Vector>(input);
  VELOX_CHECK(in != nullptr, "B200OrderBy expects device-resident input");
  orderAfterProducer(*in, dev_->stream);
  batches_.push_back(std::move(in));
}
RowVectorPtr B200OrderBy::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!noMoreInput_ || finished_) return nullptr;
  finished_ = true;
  if (batches_.empty()) return nullptr;
  B200VectorPtr all = concatBatches(batches_, pool(), dev_->stream);
  batches_.clear();
  auto order = sortOrder(*all, channels_, node_->sortingOrders(), dev_->stream);
  addRuntimeStat("b200.sortedRows", exec::RuntimeCounter{static_cast<int64_t>(all->size())});
  return takeRows(all, order, all->size(), pool(), dev_->stream);
}

// ---- B200TopN ----------------------------------------------------------------------------------
B200TopN::B200TopN(int32_t id, exec::DriverCtx* ctx, std::shared_ptr<const core::TopNNode> node)
    : Operator(ctx, node->outputType(), id, node->id(), "B200TopN"), node_(std::move(node)) {
  channels_ = channelsOf(node_->sources()[0]->outputType(), node_->sortingKeys());
}
void B200TopN::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
}
void B200TopN::addInput(RowVectorPtr input) {
  B200_NVTX_OPERATOR_RANGE("addInput");
  auto in = std::dynamic_pointer_cast<B200Vector>(input);
  VELOX_CHECK(in != nullptr, "B200TopN expects device-resident input");
  orderAfterProducer(*in, dev_->stream);
  pendingRows_ += in->size();
  pending_.push_back(std::move(in));
  // fold the pending batches into the kept rows once they outweigh them: memory stays O(count + batch)
  if (pendingRows_ >= std::max<int64_t>(4 * static_cast<int64_t>(node_->count()), 1 << 22)) fold();
}
void B200TopN::fold() {
  if (pending_.empty()) return;
  std::vector<B200VectorPtr> parts;
  if (top_) parts.push_back(top_);  // the kept rows come first: on ties the earlier input wins, as in a single sort
  for (auto& b : pending_) parts.push_back(b);
  pending_.clear();
  pendingRows_ = 0;
  B200VectorPtr all = concatBatches(parts, pool(), dev_->stream);
  auto order = sortOrder(*all, channels_, node_->sortingOrders(), dev_->stream);
  const int64_t m = std::min<int64_t>(node_->count(), all->size());
  top_ = compact(takeRows(all, order, m, pool(), dev_->stream), pool(), dev_->stream);
}
RowVectorPtr B200TopN::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!noMoreInput_ || finished_) return nullptr;
  finished_ = true;
  fold();
  return top_;
}

}  // namespace velox_b200
