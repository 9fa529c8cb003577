// B200PartitionedOutput / B200Exchange: the shuffle between the fragments of a distributed plan as
// device operators.
//   reference ....... velox/exec/PartitionedOutput.cpp (partition each batch with the node's partition
//                     function, append to per-destination buffers, flush to the OutputBufferManager),
//                     velox/exec/Exchange.cpp + ExchangeClient (pull pages from every producer)
//   GPU precedent ... velox/experimental/ucx-exchange/UcxPartitionedOutput.h:29-110 (whole device
//                     tables per destination instead of serialized rows)
// One process per GPU runs the same plan; the producing and the consuming fragment meet inside the
// Task through an ExchangeQueue, and the rows cross NVLink in one grouped ncclSend/ncclRecv all-to-all
// per exchange (all columns in one launch), preceded by ONE metadata all-gather (per-destination
// row counts + VARCHAR alphabets) that is the exchange's only host synchronisation.
#include <cstring>
#include <map>
#include <mutex>

#include "exchange_impl.h"
#include "join.h"
#include "operators.h"

namespace velox_b200 {

namespace {

constexpr size_t kAlphabetBlockBytes = 12 * 1024;  // per VARCHAR column and rank inside the metadata block
constexpr size_t kEagerBytes = 16 * 1024;          // payload small enough to ride inside the metadata block

std::shared_ptr<void> makeEvent() {
  cudaEvent_t e = nullptr;
  VB2_CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return std::shared_ptr<void>(e, [](void* p) { cudaEventDestroy(static_cast<cudaEvent_t>(p)); });
}

// One column in exchange form: flat fixed-width values (VARCHAR: int32 dictionary codes, BOOLEAN:
// one byte per row) and, when the column may hold NULLs, one validity byte per row.
struct WireColumn {
  TypePtr type;
  int32_t vb2Type = 0;
  int32_t width = 8;
  const void* values = nullptr;
  const uint8_t* validBytes = nullptr;
  std::shared_ptr<const HostAlphabet> alphabet;  // VARCHAR
  std::vector<DeviceBufferPtr> keep;
};

NcclTransport* transportOf(exec::DriverCtx* ctx) {
  auto* t = ctx->task ? dynamic_cast<NcclTransport*>(ctx->task->exchangeTransport().get()) : nullptr;
  return t;
}

}  // namespace

// ---- B200PartitionedOutput -------------------------------------------------------------------------
B200PartitionedOutput::B200PartitionedOutput(int32_t id, exec::DriverCtx* ctx, const exec::PartitionedOutput& cpu)
    : Operator(ctx, cpu.node()->outputType(), id, cpu.planNodeId(), "B200PartitionedOutput"), node_(cpu.node()), queue_(cpu.queue()) {
  // The partition function comes from the node's spec (core/PlanNode.h:2728); the hash spec is the one
  // with a device implementation. Without a spec (gather, broadcast) every row has one destination rule.
  if (auto spec = node_->partitionFunctionSpecPtr()) {
    auto hash = dynamic_cast<const exec::HashPartitionFunctionSpec*>(spec.get());
    if (!hash) VELOX_UNSUPPORTED("partition function " + spec->toString() + " (HashPartitionFunctionSpec is supported)");
    if (!hash->constValues().empty()) VELOX_UNSUPPORTED("constant partition keys");
    for (auto c : hash->keyChannels()) keyChannels_.push_back(static_cast<int32_t>(c));
  } else {
    for (auto& k : node_->keys()) {
      auto f = dynamic_cast<const core::FieldAccessTypedExpr*>(k.get());
      VELOX_CHECK(f != nullptr, "partition keys must be input columns");
      keyChannels_.push_back(channelOf(node_->sources()[0]->outputType(), *f));
    }
  }
}

void B200PartitionedOutput::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
}

void B200PartitionedOutput::addInput(RowVectorPtr input) {
  B200_NVTX_OPERATOR_RANGE("addInput");
  auto in = std::dynamic_pointer_cast<B200Vector>(input);
  VELOX_CHECK(in != nullptr, "B200PartitionedOutput expects device-resident input");
  orderAfterProducer(*in, dev_->stream);
  batches_.push_back(std::move(in));
}

void B200PartitionedOutput::noMoreInput() {
  B200_NVTX_OPERATOR_RANGE("noMoreInput");
  Operator::noMoreInput();
  cudaStream_t st = dev_->stream;
  NcclTransport* tr = transportOf(driverCtx_);
  const int world = tr ? tr->world() : 1;
  const int rank = tr ? tr->rank() : 0;
  const auto& type = node_->inputType();
  const size_t ncols = type->size();
  int parts = node_->numPartitions() > 0 ? node_->numPartitions() : world;
  VELOX_CHECK(parts <= world, "more partitions than ranks in the exchange transport");
  const bool broadcast = node_->isBroadcast();

  // ---- rows of this rank in wire form -----------------------------------------------------------
  int64_t n = 0;
  for (auto& b : batches_) n += b->size();
  VELOX_CHECK(n < (1ll << 31), "exchange input above 2^31 rows per rank");
  std::vector<WireColumn> wire(ncols);
  for (size_t c = 0; c < ncols; ++c) {
    WireColumn& w = wire[c];
    w.type = type->childAt(static_cast<uint32_t>(c));
    w.vb2Type = veloxTypeToVb2(w.type);
    w.width = w.vb2Type == VB2_VARCHAR ? 4 : (w.vb2Type == VB2_BOOLEAN ? 1 : widthOf(w.vb2Type));
    bool anyNulls = false;
    for (auto& b : batches_) anyNulls = anyNulls || b->column(c)->mayHaveNulls();
    if (w.vb2Type == VB2_VARCHAR) {
      // dictionary codes travel (NULL rows and NULL dictionary entries become validity bytes); every
      // batch must bring the same host alphabet contents
      for (auto& b : batches_) {
        const DeviceColumn& col = *b->column(c);
        if (col.desc.encoding == VB2_FLAT || !col.alphabet) VELOX_UNSUPPORTED("exchange of flat (non-dictionary) VARCHAR columns");
        if (!w.alphabet) w.alphabet = col.alphabet;
        else if (w.alphabet != col.alphabet && w.alphabet->values != col.alphabet->values) VELOX_UNSUPPORTED("exchange input batches with different VARCHAR dictionaries");
      }
      if (!w.alphabet) { auto a = std::make_shared<HostAlphabet>(); w.alphabet = a; }
    }
    if (batches_.size() == 1 && !anyNulls) {
      const DeviceColumn& col = *batches_[0]->column(c);
      if (w.vb2Type == VB2_VARCHAR && col.desc.encoding == VB2_DICTIONARY) { w.values = col.desc.indices; continue; }
      if (w.vb2Type != VB2_VARCHAR && w.vb2Type != VB2_BOOLEAN && col.desc.encoding == VB2_FLAT) { w.values = col.desc.values; continue; }
    }
    // general case: decode every batch into one flat buffer (+ validity bytes)
    auto values = allocDevice(static_cast<size_t>(n ? n : 1) * w.width, st);
    DeviceBufferPtr valid = anyNulls ? allocDevice(static_cast<size_t>(n ? n : 1), st) : nullptr;
    int64_t off = 0;
    for (auto& b : batches_) {
      const DeviceColumnPtr& col = b->column(c);
      const int64_t bn = b->size();
      if (w.vb2Type == VB2_VARCHAR) {
        kernelCheck(vb2k_dictionary_codes(&col->desc, bn, values->as<int32_t>() + off, valid ? valid->as<uint8_t>() + off : nullptr, st));
      } else {
        FlatColumn f = flattenColumn(col, nullptr, bn, st);
        VB2_CU(cudaMemcpyAsync(values->as<uint8_t>() + off * w.width, f.values->data(), static_cast<size_t>(bn) * w.width, cudaMemcpyDeviceToDevice, st));
        if (valid) {
          if (f.nulls) kernelCheck(vb2k_unpack_bits(f.nulls->as<uint64_t>(), bn, valid->as<uint8_t>() + off, st));
          else VB2_CU(cudaMemsetAsync(valid->as<uint8_t>() + off, 1, static_cast<size_t>(bn), st));
        }
      }
      off += bn;
    }
    w.values = values->data();
    w.keep.push_back(values);
    if (valid) { w.validBytes = valid->as<uint8_t>(); w.keep.push_back(valid); }
  }

  // ---- partition (HashPartitionFunction: hash(keys) % partitions) ---------------------------------
  DeviceBufferPtr countsDev;            // per-destination row counts, on the device when a partition pass produced them
  std::vector<int64_t> hostCounts;      // ... or on the host when they follow from the row count alone
  DeviceBufferPtr order;
  const DeviceColumn* singleKey = nullptr;
  if (!broadcast && parts > 1 && n > 0 && batches_.size() == 1 && keyChannels_.size() == 1) {
    const DeviceColumn& kc = *batches_[0]->column(keyChannels_[0]);
    if (kc.desc.encoding == VB2_FLAT && !kc.desc.nulls && (kc.desc.type == VB2_BIGINT || kc.desc.type == VB2_INTEGER)) singleKey = &kc;
  }
  if (singleKey) {
    // one flat NULL-free integer key: hash, partition id, histogram and order in three launches, no hash / id arrays
    order = allocDevice(static_cast<size_t>(n) * 4, st);
    countsDev = allocDeviceZeroed(static_cast<size_t>(world) * 8, st);
    kernelCheck(vb2k_partition_order_key(singleKey->desc.values, singleKey->desc.type == VB2_BIGINT, n, parts, countsDev->as<int64_t>(), order->as<int32_t>(), st));
  } else if (!broadcast && parts > 1 && n > 0) {
    auto hashes = allocDevice(static_cast<size_t>(n) * 8, st);
    int64_t off = 0;
    for (auto& b : batches_) {
      std::vector<vb2_column> keyCols;
      for (int32_t c : keyChannels_) keyCols.push_back(b->column(c)->desc);
      VELOX_CHECK(!keyCols.empty(), "partitioned exchange without keys");
      kernelCheck(vb2k_hash_columns(keyCols.data(), static_cast<int32_t>(keyCols.size()), b->size(), hashes->as<uint64_t>() + off, st));
      off += b->size();
    }
    auto ids = allocDevice(static_cast<size_t>(n) * 4, st);
    kernelCheck(vb2k_partition_ids(hashes->as<uint64_t>(), n, parts, ids->as<uint32_t>(), st));
    order = allocDevice(static_cast<size_t>(n) * 4, st);
    countsDev = allocDeviceZeroed(static_cast<size_t>(world) * 8, st);
    kernelCheck(vb2k_partition_scatter_order(ids->as<uint32_t>(), n, parts, countsDev->as<int64_t>(), order->as<int32_t>(), st));
  } else if (n > 0) {
    // one partition (gather) or broadcast: every row goes to rank 0 / to every rank; no reordering.
    // The counts are known on the host: they go straight into the metadata block.
    hostCounts.assign(world, broadcast ? n : 0);
    if (!broadcast) hostCounts[0] = n;
  } else {
    hostCounts.assign(world, 0);
  }
  // wire columns in transfer order: values, then validity bytes where present
  std::vector<const void*> sendPtr;
  std::vector<int32_t> elemBytes;
  auto listColumns = [&]() {
    sendPtr.clear();
    elemBytes.clear();
    for (auto& w : wire) {
      sendPtr.push_back(w.values);
      elemBytes.push_back(w.width);
      if (w.validBytes) { sendPtr.push_back(w.validBytes); elemBytes.push_back(1); }
    }
  };
  listColumns();

  // ---- metadata round: [counts[world] | has-valid flags | alphabets] from every rank ---------------
  size_t nVarchar = 0;
  for (auto& w : wire) nVarchar += w.vb2Type == VB2_VARCHAR;
  // Gathers and broadcasts of a handful of rows (partial aggregates in front of a final aggregation)
  // ride inside the metadata block itself ("eager" payload): one round instead of two.
  const bool eagerKind = (broadcast || parts == 1) && world > 1;
  const size_t headBytes = (static_cast<size_t>(world) * 8 + ncols * 8 + 16 + nVarchar * kAlphabetBlockBytes + 63) / 64 * 64;
  const size_t blockBytes = (headBytes + (eagerKind ? kEagerBytes : 0) + 255) / 256 * 256;
  std::vector<uint8_t> myBlock(blockBytes, 0);
  // eager payload layout: per wire column values (n * width, padded to 8) then validity bytes if this rank sends them
  auto pad8 = [](size_t b) { return (b + 7) / 8 * 8; };
  size_t eagerNeed = 0;
  for (auto& w : wire) eagerNeed += pad8(static_cast<size_t>(n) * w.width) + (w.validBytes ? pad8(static_cast<size_t>(n)) : 0);
  bool eagerMine = eagerKind && eagerNeed <= kEagerBytes;
  const std::shared_ptr<const HostMirror> inMirror = batches_.size() == 1 ? batches_[0]->mirror() : nullptr;
  std::vector<std::pair<size_t, std::pair<const void*, size_t>>> eagerDeviceCopies;  // (block offset, (device source, bytes))
  if (eagerMine) {
    size_t at = headBytes;
    auto place = [&](const void* dev, size_t bytes) {
      const uint8_t* host = inMirror ? inMirror->hostOf(dev) : nullptr;
      if (host) std::memcpy(myBlock.data() + at, host, bytes);       // the producer's result is already on the host
      else if (bytes) eagerDeviceCopies.push_back({at, {dev, bytes}});  // patched into the device block below
      at += pad8(bytes);
    };
    for (auto& w : wire) {
      place(w.values, static_cast<size_t>(n) * w.width);
      if (w.validBytes) place(w.validBytes, static_cast<size_t>(n));
    }
  }
  if (!hostCounts.empty()) std::memcpy(myBlock.data(), hostCounts.data(), static_cast<size_t>(world) * 8);
  {
    int64_t* eagerFlag = reinterpret_cast<int64_t*>(myBlock.data() + static_cast<size_t>(world) * 8 + ncols * 8);
    eagerFlag[0] = eagerMine ? 1 : 0;
    eagerFlag[1] = n;
    int64_t* flags = reinterpret_cast<int64_t*>(myBlock.data() + static_cast<size_t>(world) * 8);
    for (size_t c = 0; c < ncols; ++c) flags[c] = wire[c].validBytes ? 1 : 0;
    uint8_t* ap = myBlock.data() + static_cast<size_t>(world) * 8 + ncols * 8 + 16;
    for (auto& w : wire) {
      if (w.vb2Type != VB2_VARCHAR) continue;
      // [int32 entries | int32 lengths[entries] | chars]
      size_t need = 4 + w.alphabet->values.size() * 4;
      for (auto& v : w.alphabet->values) need += v.size();
      if (need > kAlphabetBlockBytes) VELOX_UNSUPPORTED("VARCHAR dictionary too large for the exchange metadata block");
      int32_t* hp = reinterpret_cast<int32_t*>(ap);
      hp[0] = static_cast<int32_t>(w.alphabet->values.size());
      char* cp = reinterpret_cast<char*>(hp + 1 + w.alphabet->values.size());
      for (size_t i = 0; i < w.alphabet->values.size(); ++i) {
        hp[1 + i] = static_cast<int32_t>(w.alphabet->values[i].size());
        std::memcpy(cp, w.alphabet->values[i].data(), w.alphabet->values[i].size());
        cp += w.alphabet->values[i].size();
      }
      ap += kAlphabetBlockBytes;
    }
  }
  std::shared_ptr<void> allHost;
  if (world > 1) {
    VELOX_CHECK(tr != nullptr, "the plan contains an exchange but the task has no exchange transport (vb2_task_set_comm)");
    VELOX_CHECK(blockBytes <= exchangeMaxMetadataBytes(tr->comm()), "exchange metadata block too large (too many VARCHAR columns)");
    std::vector<ExchangePatch> patches;
    for (auto& e : eagerDeviceCopies) patches.push_back(ExchangePatch{e.first, e.second.first, e.second.second});
    allHost = exchangeMetadata(tr->comm(), myBlock.data(), blockBytes, countsDev ? countsDev->as<int64_t>() : nullptr, patches, st);
  } else {
    allHost = acquirePinned(blockBytes);
    std::memcpy(allHost.get(), myBlock.data(), blockBytes);
    if (countsDev) {
      VB2_CU(cudaMemcpyAsync(allHost.get(), countsDev->data(), 8, cudaMemcpyDeviceToHost, st));
      VB2_CU(cudaStreamSynchronize(st));
    }
  }
  const uint8_t* all = static_cast<const uint8_t*>(allHost.get());
  auto blockOf = [&](int r) { return all + static_cast<size_t>(r) * blockBytes; };
  std::vector<int64_t> matrix(static_cast<size_t>(world) * world), recvCounts(world);
  for (int r = 0; r < world; ++r)
    for (int p = 0; p < world; ++p) matrix[static_cast<size_t>(r) * world + p] = reinterpret_cast<const int64_t*>(blockOf(r))[p];
  int64_t total = 0, sent = 0;
  for (int p = 0; p < world; ++p) {
    recvCounts[p] = matrix[static_cast<size_t>(p) * world + rank];
    total += recvCounts[p];
    sent += matrix[static_cast<size_t>(rank) * world + p];
  }
  VELOX_CHECK(total < (1ll << 31), "exchange output above 2^31 rows per rank");
  // a column carries validity bytes if any rank sends them: ranks without NULLs send all-ones
  std::vector<bool> anyValid(ncols, false);
  for (int r = 0; r < world; ++r) {
    const int64_t* flags = reinterpret_cast<const int64_t*>(blockOf(r) + static_cast<size_t>(world) * 8);
    for (size_t c = 0; c < ncols; ++c) anyValid[c] = anyValid[c] || flags[c] != 0;
  }
  for (size_t c = 0; c < ncols; ++c) {
    WireColumn& w = wire[c];
    if (anyValid[c] && !w.validBytes) {
      auto ones = allocDevice(static_cast<size_t>(n ? n : 1), st);
      VB2_CU(cudaMemsetAsync(ones->data(), 1, static_cast<size_t>(n ? n : 1), st));
      w.validBytes = ones->as<uint8_t>();
      w.keep.push_back(ones);
    }
  }
  listColumns();
  // merged alphabets (rank order, first occurrence wins) and per-source code remaps
  struct Merged {
    std::shared_ptr<HostAlphabet> alphabet = std::make_shared<HostAlphabet>();
    std::vector<std::vector<int32_t>> remap;  // [source rank][code] -> merged code
    bool identical = true;
  };
  std::vector<Merged> merged(ncols);
  {
    size_t v = 0;
    for (size_t c = 0; c < ncols; ++c) {
      if (wire[c].vb2Type != VB2_VARCHAR) continue;
      Merged& m = merged[c];
      std::map<std::string, int32_t> ids;
      m.remap.resize(world);
      for (int r = 0; r < world; ++r) {
        const uint8_t* ap = blockOf(r) + static_cast<size_t>(world) * 8 + ncols * 8 + 16 + v * kAlphabetBlockBytes;
        const int32_t* hp = reinterpret_cast<const int32_t*>(ap);
        const int32_t entries = hp[0];
        const char* cp = reinterpret_cast<const char*>(hp + 1 + entries);
        for (int32_t i = 0; i < entries; ++i) {
          std::string sv(cp, hp[1 + i]);
          cp += hp[1 + i];
          auto it = ids.find(sv);
          if (it == ids.end()) {
            it = ids.emplace(sv, static_cast<int32_t>(m.alphabet->values.size())).first;
            m.alphabet->values.push_back(sv);
            m.alphabet->nulls.push_back(false);
          }
          if (it->second != i) m.identical = false;
          m.remap[r].push_back(it->second);
        }
      }
      ++v;
    }
  }

  // ---- payload ---------------------------------------------------------------------------------------
  std::vector<DeviceBufferPtr> recvBuf;
  std::vector<void*> recvPtr;
  bool allEager = eagerKind;
  for (int r = 0; r < world && allEager; ++r) allEager = reinterpret_cast<const int64_t*>(blockOf(r) + static_cast<size_t>(world) * 8 + ncols * 8)[0] != 0;
  std::shared_ptr<HostMirror> pageMirror;
  std::shared_ptr<void> ready;
  if (allEager) {
    // every source's rows arrived with the metadata: assemble the page in pinned host memory and
    // bring it to the device in ONE copy (it doubles as the page's host mirror); no second round.
    size_t arenaBytes = 0;
    for (size_t i = 0; i < sendPtr.size(); ++i) arenaBytes += (static_cast<size_t>(total ? total : 1) * elemBytes[i] + 255) / 256 * 256;
    auto arenaHost = acquirePinned(arenaBytes);
    auto arenaDev = allocDevice(arenaBytes, st);
    uint8_t* hb = static_cast<uint8_t*>(arenaHost.get());
    size_t colOff = 0;
    std::vector<size_t> srcAt(world, headBytes);
    size_t i = 0;
    for (size_t c = 0; c < ncols; ++c) {
      const int width = wire[c].width;
      // values of column c from every source that sends to this rank, then the validity bytes
      for (int pass = 0; pass < (anyValid[c] ? 2 : 1); ++pass) {
        const int eb = pass == 0 ? width : 1;
        size_t rowAt = 0;
        for (int r = 0; r < world; ++r) {
          const int64_t* hdr = reinterpret_cast<const int64_t*>(blockOf(r) + static_cast<size_t>(world) * 8);
          const int64_t rn = hdr[ncols + 1];           // rows the source holds (its whole payload)
          const bool srcHasValid = hdr[c] != 0;
          const int64_t mine = recvCounts[r];           // rows of it addressed to this rank (all or none)
          if (pass == 1 && !srcHasValid) {
            if (mine) std::memset(hb + colOff + rowAt, 1, static_cast<size_t>(mine));
          } else {
            if (mine) std::memcpy(hb + colOff + rowAt * eb, blockOf(r) + srcAt[r], static_cast<size_t>(mine) * eb);
            srcAt[r] += pad8(static_cast<size_t>(rn) * eb);
          }
          rowAt += static_cast<size_t>(mine);
        }
        recvPtr.push_back(arenaDev->as<uint8_t>() + colOff);
        recvBuf.push_back(arenaDev);
        colOff += (static_cast<size_t>(total ? total : 1) * eb + 255) / 256 * 256;
        ++i;
      }
    }
    VB2_CU(cudaMemcpyAsync(arenaDev->data(), arenaHost.get(), arenaBytes, cudaMemcpyHostToDevice, st));
    pageMirror = std::make_shared<HostMirror>();
    pageMirror->host = arenaHost;
    pageMirror->devBase = arenaDev->as<uint8_t>();
    pageMirror->bytes = arenaBytes;
    addRuntimeStat("b200.exchangeEager", exec::RuntimeCounter{1});
  } else {
    for (size_t i = 0; i < sendPtr.size(); ++i) {
      recvBuf.push_back(allocDevice(static_cast<size_t>(total ? total : 1) * elemBytes[i], st));
      recvPtr.push_back(recvBuf.back()->data());
    }
  }
  if (allEager) {
    // nothing more to move
  } else if (world > 1) {
    bool peerMemory = false;
    if (!countsDev) {  // a gather / broadcast too large for the eager round: the transfer kernel reads the counts on the device
      countsDev = allocDevice(static_cast<size_t>(world) * 8, st);
      VB2_CU(cudaMemcpyAsync(countsDev->data(), hostCounts.data(), static_cast<size_t>(world) * 8, cudaMemcpyHostToDevice, st));
    }
    ready = exchangePayload(tr->comm(), order ? order->as<int32_t>() : nullptr, countsDev->as<int64_t>(), matrix.data(), n, sendPtr, elemBytes, recvPtr,
                            broadcast, st, &peerMemory);
    addRuntimeStat("b200.exchangePeerMemory", exec::RuntimeCounter{peerMemory ? 1 : 0});
    // everything derived from the received buffers below runs on this operator's stream
    VB2_CU(cudaStreamWaitEvent(st, static_cast<cudaEvent_t>(ready.get()), 0));
  } else {
    for (size_t i = 0; i < sendPtr.size(); ++i) {
      if (!total) continue;
      if (order) kernelCheck(vb2k_gather(sendPtr[i], order->as<int32_t>(), total, elemBytes[i], recvPtr[i], st));
      else VB2_CU(cudaMemcpyAsync(recvPtr[i], sendPtr[i], static_cast<size_t>(total) * elemBytes[i], cudaMemcpyDeviceToDevice, st));
    }
  }
  addRuntimeStat("b200.exchangeRowsSent", exec::RuntimeCounter{sent});
  addRuntimeStat("b200.exchangeRowsReceived", exec::RuntimeCounter{total});

  // ---- received page ----------------------------------------------------------------------------------
  if (total > 0) {
    std::vector<DeviceColumnPtr> cols;
    size_t i = 0;
    for (size_t c = 0; c < ncols; ++c) {
      const WireColumn& w = wire[c];
      auto col = std::make_shared<DeviceColumn>();
      col->type = w.type;
      col->desc.type = w.vb2Type;
      col->desc.size = total;
      // recvPtr[i] lies inside recvBuf[i] (its own buffer, or the shared arena of an eager exchange)
      const void* valuesPtr = recvPtr[i];
      col->owners.push_back(recvBuf[i]);
      ++i;
      if (anyValid[c]) {
        auto bitsBuf = allocDevice(bits::nbytes(total), st);
        kernelCheck(vb2k_pack_bools(static_cast<const uint8_t*>(recvPtr[i]), total, bitsBuf->as<uint64_t>(), st));
        col->desc.nulls = bitsBuf->as<uint64_t>();
        col->owners.push_back(recvBuf[i]);
        col->owners.push_back(bitsBuf);
        ++i;
      }
      if (w.vb2Type == VB2_VARCHAR) {
        Merged& m = merged[c];
        if (!m.identical) {
          // codes of source r are rewritten through remap[r] (segments are contiguous per source)
          int64_t off = 0;
          auto fixed = allocDevice(static_cast<size_t>(total) * 4, st);
          for (int r = 0; r < world; ++r) {
            if (recvCounts[r] == 0) continue;
            auto lut = allocDevice(m.remap[r].size() * 4 + 4, st);
            VB2_CU(cudaMemcpyAsync(lut->data(), m.remap[r].data(), m.remap[r].size() * 4, cudaMemcpyHostToDevice, st));
            kernelCheck(vb2k_gather(lut->data(), static_cast<const int32_t*>(valuesPtr) + off, recvCounts[r], 4, fixed->as<int32_t>() + off, st));
            col->owners.push_back(lut);
            off += recvCounts[r];
          }
          valuesPtr = fixed->data();
          col->owners.push_back(fixed);
        }
        DeviceBufferPtr offBuf, charBuf;
        deviceAlphabet(*m.alphabet, st, offBuf, charBuf);
        col->desc.encoding = VB2_DICTIONARY;
        col->desc.indices = static_cast<const int32_t*>(valuesPtr);
        col->desc.values = offBuf->data();
        col->desc.aux = charBuf->data();
        col->desc.dict_size = static_cast<int64_t>(m.alphabet->values.size());
        col->owners.push_back(offBuf);
        col->owners.push_back(charBuf);
        col->alphabet = m.alphabet;
      } else if (w.vb2Type == VB2_BOOLEAN) {
        auto packed = allocDevice(bits::nbytes(total), st);
        kernelCheck(vb2k_pack_bools(static_cast<const uint8_t*>(valuesPtr), total, packed->as<uint64_t>(), st));
        col->desc.encoding = VB2_FLAT;
        col->desc.values = packed->data();
        col->owners.push_back(packed);
      } else {
        col->desc.encoding = VB2_FLAT;
        col->desc.values = valuesPtr;
      }
      cols.push_back(std::move(col));
    }
    auto page = std::make_shared<B200Vector>(pool(), type, static_cast<vector_size_t>(total), std::move(cols), st);
    if (pageMirror) page->setMirror(pageMirror);
    auto ev = makeEvent();
    VB2_CU(cudaEventRecord(static_cast<cudaEvent_t>(ev.get()), st));
    page->setReadyEvent(ev);
    queue_->enqueue(page);
  }
  batches_.clear();
  queue_->noMoreData();
}

// ---- B200Exchange --------------------------------------------------------------------------------------
B200Exchange::B200Exchange(int32_t id, exec::DriverCtx* ctx, const exec::Exchange& cpu)
    : SourceOperator(ctx, cpu.outputType(), id, cpu.planNodeId(), "B200Exchange"), queue_(cpu.queue()) {}

void B200Exchange::initialize() {
  Operator::initialize();
  dev_ = driverDeviceContext(driverCtx_);
}

exec::BlockingReason B200Exchange::isBlocked(exec::ContinueFuture* future) {
  if (next_ || atEnd_) return exec::BlockingReason::kNotBlocked;
  next_ = queue_->dequeue(&atEnd_, future);
  if (!next_ && !atEnd_) return exec::BlockingReason::kWaitForProducer;
  return exec::BlockingReason::kNotBlocked;
}

RowVectorPtr B200Exchange::getOutput() {
  B200_NVTX_OPERATOR_RANGE("getOutput");
  if (!next_ && !atEnd_) {
    exec::ContinueFuture f;
    next_ = queue_->dequeue(&atEnd_, &f);
  }
  if (!next_) return nullptr;
  RowVectorPtr out = std::move(next_);
  next_ = nullptr;
  if (auto page = std::dynamic_pointer_cast<B200Vector>(out)) {
    // the page was produced on the sending pipeline's stream
    if (page->readyEvent()) VB2_CU(cudaStreamWaitEvent(dev_->stream, page->readyEvent(), 0));
  }
  return out;
}

}  // namespace velox_b200
