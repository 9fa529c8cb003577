#include "device.h"

#include <cstring>
#include <map>
#include <mutex>

#include <nvtx3/nvToolsExt.h>
#include <string_view>
#include <unordered_map>

namespace velox_b200 {

void cudaCheck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw VeloxRuntimeError(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}
void kernelCheck(int rc) {
  if (rc == VB2_OK) return;
  std::string msg = vb2_last_error();
  if (rc == VB2_ERR_USER) throw VeloxUserError(msg);
  throw VeloxRuntimeError(msg);
}

namespace {
std::mutex& streamMutex() {
  static std::mutex m;
  return m;
}
std::map<cudaStream_t, std::weak_ptr<DeviceContext>>& streamOwners() {
  static std::map<cudaStream_t, std::weak_ptr<DeviceContext>> m;
  return m;
}
std::shared_ptr<void> ownerOf(cudaStream_t stream) {
  std::lock_guard<std::mutex> l(streamMutex());
  auto it = streamOwners().find(stream);
  if (it == streamOwners().end()) return nullptr;
  return it->second.lock();
}
}  // namespace

DeviceBuffer::DeviceBuffer(size_t bytes, cudaStream_t stream) : bytes_(bytes), stream_(stream), streamOwner_(ownerOf(stream)) {
  VB2_CU(cudaMallocAsync(&ptr_, bytes ? bytes : 8, stream));
}
DeviceBuffer::~DeviceBuffer() {
  if (owned_ && ptr_) cudaFreeAsync(ptr_, stream_);
}
DeviceBufferPtr allocDevice(size_t bytes, cudaStream_t stream) { return std::make_shared<DeviceBuffer>(bytes, stream); }
DeviceBufferPtr allocDeviceZeroed(size_t bytes, cudaStream_t stream) {
  auto b = allocDevice(bytes, stream);
  VB2_CU(cudaMemsetAsync(b->data(), 0, bytes ? bytes : 8, stream));
  return b;
}

namespace {
struct PinnedPool {
  std::mutex mu;
  std::map<size_t, std::vector<void*>> free;  // size class -> blocks
  ~PinnedPool() {
    for (auto& kv : free)
      for (void* p : kv.second) cudaFreeHost(p);
  }
};
PinnedPool& pinnedPool() {
  static PinnedPool* p = new PinnedPool();  // leaked on purpose: blocks may outlive static destruction order
  return *p;
}
}  // namespace

std::shared_ptr<void> acquirePinned(size_t bytes) {
  size_t cls = 4096;
  while (cls < bytes) cls <<= 1;
  void* p = nullptr;
  {
    std::lock_guard<std::mutex> l(pinnedPool().mu);
    auto& v = pinnedPool().free[cls];
    if (!v.empty()) { p = v.back(); v.pop_back(); }
  }
  if (!p) VB2_CU(cudaHostAlloc(&p, cls, cudaHostAllocDefault));
  return std::shared_ptr<void>(p, [cls](void* q) {
    std::lock_guard<std::mutex> l(pinnedPool().mu);
    pinnedPool().free[cls].push_back(q);
  });
}

namespace {
// Streams are pooled too: a Task creates one per driver, and cudaStreamCreate / Destroy cost tens
// of microseconds each.
struct StreamPool {
  std::mutex mu;
  std::map<int, std::vector<cudaStream_t>> free;  // device -> idle streams
};
StreamPool& streamPool() {
  static StreamPool* p = new StreamPool();
  return *p;
}
}  // namespace

DeviceContext::DeviceContext() {
  VB2_CU(cudaGetDevice(&device));
  {
    std::lock_guard<std::mutex> l(streamPool().mu);
    auto& v = streamPool().free[device];
    if (!v.empty()) { stream = v.back(); v.pop_back(); }
  }
  if (stream) return;
  VB2_CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  // keep freed blocks in the pool: operators allocate and free large scratch buffers per batch
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
}
DeviceContext::~DeviceContext() {
  if (stream) {
    {
      std::lock_guard<std::mutex> l(streamMutex());
      streamOwners().erase(stream);
    }
    // Buffers that were freed on this stream (cudaFreeAsync) are ordered behind its work; the next
    // user of the pooled stream is ordered behind them as well, so no synchronisation is needed.
    std::lock_guard<std::mutex> l(streamPool().mu);
    streamPool().free[device].push_back(stream);
  }
}

std::shared_ptr<DeviceContext> driverDeviceContext(exec::DriverCtx* ctx) {
  static std::mutex mu;
  static std::map<exec::DriverCtx*, std::weak_ptr<DeviceContext>> contexts;
  std::lock_guard<std::mutex> l(mu);
  auto& w = contexts[ctx];
  auto sp = w.lock();
  if (!sp) {
    const int dev = ctx && ctx->config ? ctx->config->b200DeviceId() : -1;
    if (dev >= 0) VB2_CU(cudaSetDevice(dev));
    sp = std::make_shared<DeviceContext>();
    w = sp;
    std::lock_guard<std::mutex> l2(streamMutex());
    streamOwners()[sp->stream] = sp;
  }
  return sp;
}

int32_t veloxTypeToVb2(const TypePtr& t) {
  switch (t->kind()) {
    case TypeKind::BOOLEAN: return VB2_BOOLEAN;
    case TypeKind::INTEGER: return VB2_INTEGER;
    case TypeKind::BIGINT: return VB2_BIGINT;
    case TypeKind::DOUBLE: return VB2_DOUBLE;
    case TypeKind::VARCHAR: return VB2_VARCHAR;
    default: VELOX_UNSUPPORTED("type " + t->toString() + " on the B200 path");
  }
}
int32_t widthOf(int32_t t) {
  switch (t) {
    case VB2_INTEGER: return 4;
    case VB2_BIGINT: case VB2_DOUBLE: return 8;
    default: return 0;
  }
}

namespace {
thread_local UploadCache* tlsUploadCache = nullptr;
thread_local int64_t tlsUploadedBytes = 0;
constexpr size_t kCacheMinBytes = 1 << 16;  // small buffers are often staging copies with short-lived addresses
}  // namespace
void setThreadUploadCache(UploadCache* cache) { tlsUploadCache = cache; }
int64_t threadUploadedBytes() { return tlsUploadedBytes; }

namespace {


DeviceBufferPtr upload(const void* src, size_t bytes, cudaStream_t stream) {
  UploadCache* cache = bytes >= kCacheMinBytes ? tlsUploadCache : nullptr;
  if (cache) {
    std::lock_guard<std::mutex> lock(cache->mu);
    auto it = cache->entries.find(reinterpret_cast<uint64_t>(src));
    if (it != cache->entries.end() && it->second.bytes == bytes) {
      cache->hitBytes += static_cast<int64_t>(bytes);
      // the copy may still be in flight on another task's stream
      if (it->second.stream != stream && it->second.copied)
        VB2_CU(cudaStreamWaitEvent(stream, static_cast<cudaEvent_t>(it->second.copied.get()), 0));
      return it->second.buffer;
    }
  }
  auto b = allocDevice(bytes, stream);
  if (bytes) VB2_CU(cudaMemcpyAsync(b->data(), src, bytes, cudaMemcpyHostToDevice, stream));
  tlsUploadedBytes += static_cast<int64_t>(bytes);
  if (cache) {
    std::lock_guard<std::mutex> lock(cache->mu);
    cudaEvent_t ev = nullptr;
    VB2_CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    VB2_CU(cudaEventRecord(ev, stream));
    cache->entries[reinterpret_cast<uint64_t>(src)] =
        UploadCache::Entry{bytes, b, stream, std::shared_ptr<void>(ev, [](void* e) { cudaEventDestroy(static_cast<cudaEvent_t>(e)); })};
    cache->missBytes += static_cast<int64_t>(bytes);
  }
  return b;
}

// Rewrites StringViews as int32 offsets + chars. keepAlive holds the staging until the copies ran.
struct StringStaging {
  std::vector<int32_t> offsets;
  std::string chars;
};

void stageStrings(const StringView* views, const BaseVector& v, vector_size_t n, StringStaging& out) {
  out.offsets.resize(static_cast<size_t>(n) + 1);
  out.offsets[0] = 0;
  size_t total = 0;
  for (vector_size_t i = 0; i < n; ++i)
    if (!v.isNullAt(i)) total += views[i].size();
  out.chars.reserve(total + 1);
  for (vector_size_t i = 0; i < n; ++i) {
    if (!v.isNullAt(i)) out.chars.append(views[i].data(), views[i].size());
    VELOX_CHECK(out.chars.size() < (1ull << 31), "VARCHAR column above 2 GiB of characters");
    out.offsets[i + 1] = static_cast<int32_t>(out.chars.size());
  }
}

// Uploads the value buffer(s) of a flat vector; fills values/aux of desc.
void uploadFlatValues(const BaseVector& v, vb2_column& d, std::vector<DeviceBufferPtr>& owners, cudaStream_t stream,
                      std::vector<std::shared_ptr<StringStaging>>& staging, HostAlphabet* alphabet) {
  const vector_size_t n = v.size();
  switch (v.typeKind()) {
    case TypeKind::BOOLEAN: {
      auto* f = v.as<FlatVector<bool>>();
      owners.push_back(upload(f->values()->as<uint8_t>(), bits::nbytes(n), stream));
      break;
    }
    case TypeKind::INTEGER: owners.push_back(upload(v.as<FlatVector<int32_t>>()->rawValues(), size_t(n) * 4, stream)); break;
    case TypeKind::BIGINT: owners.push_back(upload(v.as<FlatVector<int64_t>>()->rawValues(), size_t(n) * 8, stream)); break;
    case TypeKind::DOUBLE: owners.push_back(upload(v.as<FlatVector<double>>()->rawValues(), size_t(n) * 8, stream)); break;
    case TypeKind::VARCHAR: {
      auto st = std::make_shared<StringStaging>();
      stageStrings(v.as<FlatVector<StringView>>()->rawValues(), v, n, *st);
      staging.push_back(st);
      owners.push_back(upload(st->offsets.data(), st->offsets.size() * 4, stream));
      d.values = owners.back()->data();
      owners.push_back(upload(st->chars.data(), st->chars.size(), stream));
      d.aux = owners.back()->data();
      if (alphabet) {
        for (vector_size_t i = 0; i < n; ++i) {
          alphabet->nulls.push_back(v.isNullAt(i));
          alphabet->values.push_back(v.isNullAt(i) ? std::string() : st->chars.substr(st->offsets[i], st->offsets[i + 1] - st->offsets[i]));
        }
      }
      return;
    }
    default: VELOX_UNSUPPORTED("type " + v.type()->toString() + " on the B200 path");
  }
  d.values = owners.back()->data();
}

const uint64_t* uploadNulls(const BaseVector& v, std::vector<DeviceBufferPtr>& owners, cudaStream_t stream) {
  if (!v.rawNulls()) return nullptr;
  owners.push_back(upload(v.rawNulls(), bits::nbytes(v.size()), stream));
  return owners.back()->as<uint64_t>();
}

constexpr vector_size_t kMaxAlphabet = 1 << 16;

// A flat VARCHAR column with few distinct values goes up dictionary-encoded: every operator of the path
// groups, joins, sorts and exchanges dictionary strings through their int32 codes (flat strings would
// need variable-width keys on the device). The reference reaches the same end on the CPU by mapping short
// strings to numbers inside VectorHasher (exec/VectorHasher.h:377-387); TPC-H's l_returnflag /
// l_linestatus / p_type (tpch/gen/TpchGen.cpp:278-317) are such columns. One hash lookup per row on the
// host, next to a PCIe copy of the same rows. Columns with more than 65536 distinct values stay flat.
bool encodeFlatStrings(const BaseVector& v, DeviceColumn& col, std::vector<std::shared_ptr<StringStaging>>& staging, cudaStream_t stream) {
  const auto* flat = v.as<FlatVector<StringView>>();
  if (!flat) return false;
  const vector_size_t n = v.size();
  const StringView* views = flat->rawValues();
  auto codes = std::make_shared<StringStaging>();
  codes->offsets.resize(static_cast<size_t>(n));
  std::unordered_map<std::string_view, int32_t> ids;
  auto alpha = std::make_shared<HostAlphabet>();
  for (vector_size_t i = 0; i < n; ++i) {
    if (v.isNullAt(i)) { codes->offsets[i] = 0; continue; }
    const std::string_view sv(views[i].data(), views[i].size());
    auto it = ids.find(sv);
    if (it == ids.end()) {
      if (static_cast<vector_size_t>(alpha->values.size()) >= kMaxAlphabet) return false;
      it = ids.emplace(sv, static_cast<int32_t>(alpha->values.size())).first;
      alpha->values.emplace_back(sv);
      alpha->nulls.push_back(false);
    }
    codes->offsets[i] = it->second;
  }
  if (alpha->values.empty()) { alpha->values.emplace_back(); alpha->nulls.push_back(false); }  // all NULL: one unused entry
  vb2_column& d = col.desc;
  d.encoding = VB2_DICTIONARY;
  staging.push_back(codes);
  col.owners.push_back(upload(codes->offsets.data(), codes->offsets.size() * 4, stream));
  d.indices = col.owners.back()->as<int32_t>();
  DeviceBufferPtr offBuf, charBuf;
  deviceAlphabet(*alpha, stream, offBuf, charBuf);
  d.values = offBuf->data();
  d.aux = charBuf->data();
  d.dict_size = static_cast<int64_t>(alpha->values.size());
  col.owners.push_back(offBuf);
  col.owners.push_back(charBuf);
  d.nulls = uploadNulls(v, col.owners, stream);
  col.alphabet = alpha;
  return true;
}

}  // namespace

B200VectorPtr toDevice(const RowVectorPtr& host, cudaStream_t stream) {
  if (auto already = std::dynamic_pointer_cast<B200Vector>(host)) return already;
  std::vector<DeviceColumnPtr> cols;
  std::vector<std::shared_ptr<StringStaging>> staging;
  for (size_t c = 0; c < host->childrenSize(); ++c) {
    const VectorPtr& child = host->childAt(static_cast<uint32_t>(c));
    auto col = std::make_shared<DeviceColumn>();
    col->type = child->type();
    vb2_column& d = col->desc;
    d.type = veloxTypeToVb2(child->type());
    d.size = child->size();
    switch (child->encoding()) {
      case VectorEncoding::Simple::FLAT:
        if (child->typeKind() == TypeKind::VARCHAR && encodeFlatStrings(*child, *col, staging, stream)) break;
        d.encoding = VB2_FLAT;
        uploadFlatValues(*child, d, col->owners, stream, staging, nullptr);
        d.nulls = uploadNulls(*child, col->owners, stream);
        break;
      case VectorEncoding::Simple::DICTIONARY: {
        d.encoding = VB2_DICTIONARY;
        // compose nested dictionaries down to a flat base
        std::vector<vector_size_t> idx;
        std::vector<bool> wrapNull;
        const BaseVector* cur = child.get();
        VectorPtr base;
        bool first = true, anyWrapNull = false;
        while (cur->encoding() == VectorEncoding::Simple::DICTIONARY) {
          const vector_size_t* raw;
          VectorPtr next;
          switch (cur->typeKind()) {
            case TypeKind::BOOLEAN: raw = cur->as<DictionaryVector<bool>>()->rawIndices(); next = cur->as<DictionaryVector<bool>>()->valueVector(); break;
            case TypeKind::INTEGER: raw = cur->as<DictionaryVector<int32_t>>()->rawIndices(); next = cur->as<DictionaryVector<int32_t>>()->valueVector(); break;
            case TypeKind::BIGINT: raw = cur->as<DictionaryVector<int64_t>>()->rawIndices(); next = cur->as<DictionaryVector<int64_t>>()->valueVector(); break;
            case TypeKind::DOUBLE: raw = cur->as<DictionaryVector<double>>()->rawIndices(); next = cur->as<DictionaryVector<double>>()->valueVector(); break;
            case TypeKind::VARCHAR: raw = cur->as<DictionaryVector<StringView>>()->rawIndices(); next = cur->as<DictionaryVector<StringView>>()->valueVector(); break;
            default: VELOX_UNSUPPORTED("dictionary over " + cur->type()->toString());
          }
          if (first) {
            if (next->encoding() == VectorEncoding::Simple::FLAT && !cur->rawNulls()) {
              // common case: single wrapper, indices copied straight from the vector's buffer
              col->owners.push_back(upload(raw, size_t(child->size()) * 4, stream));
              d.indices = col->owners.back()->as<int32_t>();
              base = next;
              break;
            }
            idx.assign(raw, raw + child->size());
            wrapNull.assign(child->size(), false);
            for (vector_size_t i = 0; i < child->size(); ++i)
              if (cur->rawNulls() && bits::isBitNull(cur->rawNulls(), i)) { wrapNull[i] = true; anyWrapNull = true; }
            first = false;
          } else {
            for (vector_size_t i = 0; i < child->size(); ++i) {
              if (wrapNull[i]) continue;
              if (cur->rawNulls() && bits::isBitNull(cur->rawNulls(), idx[i])) { wrapNull[i] = true; anyWrapNull = true; idx[i] = 0; continue; }
              idx[i] = raw[idx[i]];
            }
          }
          base = next;
          cur = next.get();
        }
        VELOX_CHECK(base->encoding() == VectorEncoding::Simple::FLAT, "dictionary base must be flat");
        if (!d.indices) {
          auto st = std::make_shared<StringStaging>();
          st->offsets.assign(idx.begin(), idx.end());
          staging.push_back(st);
          col->owners.push_back(upload(st->offsets.data(), st->offsets.size() * 4, stream));
          d.indices = col->owners.back()->as<int32_t>();
          if (anyWrapNull) {
            auto nb = std::make_shared<StringStaging>();
            nb->chars.assign(bits::nbytes(child->size()), '\xff');
            auto* words = reinterpret_cast<uint64_t*>(nb->chars.data());
            for (vector_size_t i = 0; i < child->size(); ++i)
              if (wrapNull[i]) bits::clearBit(words, i);
            staging.push_back(nb);
            col->owners.push_back(upload(nb->chars.data(), nb->chars.size(), stream));
            d.nulls = col->owners.back()->as<uint64_t>();
          }
        }
        d.dict_size = base->size();
        std::shared_ptr<HostAlphabet> alpha;
        if (base->typeKind() == TypeKind::VARCHAR && base->size() <= kMaxAlphabet) alpha = std::make_shared<HostAlphabet>();
        uploadFlatValues(*base, d, col->owners, stream, staging, alpha.get());
        d.dict_nulls = uploadNulls(*base, col->owners, stream);
        col->alphabet = alpha;
        break;
      }
      case VectorEncoding::Simple::CONSTANT: {
        d.encoding = VB2_CONSTANT;
        const bool isNull = child->isNullAt(0);
        uint64_t word = 0;
        switch (child->typeKind()) {
          case TypeKind::BOOLEAN: word = child->as<ConstantVector<bool>>()->value() ? 1 : 0; break;
          case TypeKind::INTEGER: { int32_t x = child->as<ConstantVector<int32_t>>()->value(); std::memcpy(&word, &x, 4); break; }
          case TypeKind::BIGINT: { int64_t x = child->as<ConstantVector<int64_t>>()->value(); std::memcpy(&word, &x, 8); break; }
          case TypeKind::DOUBLE: { double x = child->as<ConstantVector<double>>()->value(); std::memcpy(&word, &x, 8); break; }
          case TypeKind::VARCHAR: {
            auto st = std::make_shared<StringStaging>();
            const StringView sv = child->as<ConstantVector<StringView>>()->value();
            if (!isNull) st->chars.assign(sv.data(), sv.size());
            st->offsets = {0, static_cast<int32_t>(st->chars.size())};
            staging.push_back(st);
            col->owners.push_back(upload(st->offsets.data(), 8, stream));
            d.values = col->owners.back()->data();
            col->owners.push_back(upload(st->chars.data(), st->chars.size(), stream));
            d.aux = col->owners.back()->data();
            auto alpha = std::make_shared<HostAlphabet>();
            alpha->values.push_back(st->chars);
            alpha->nulls.push_back(isNull);
            col->alphabet = alpha;
            break;
          }
          default: VELOX_UNSUPPORTED("constant of " + child->type()->toString());
        }
        if (child->typeKind() != TypeKind::VARCHAR) {
          auto st = std::make_shared<StringStaging>();
          st->chars.assign(reinterpret_cast<const char*>(&word), 8);
          staging.push_back(st);
          col->owners.push_back(upload(st->chars.data(), 8, stream));
          d.values = col->owners.back()->data();
        }
        if (isNull) {
          col->owners.push_back(allocDeviceZeroed(8, stream));
          d.nulls = col->owners.back()->as<uint64_t>();
        }
        break;
      }
      default: VELOX_UNSUPPORTED("vector encoding on the B200 path");
    }
    cols.push_back(col);
  }
  if (!staging.empty()) VB2_CU(cudaStreamSynchronize(stream));  // staging buffers die here
  return std::make_shared<B200Vector>(host->pool(), host->type(), host->size(), std::move(cols), stream);
}

namespace {

template <class T>
std::shared_ptr<FlatVector<T>> downloadFlat(memory::MemoryPool* pool, const TypePtr& type, const void* devValues, const uint64_t* devNulls,
                                            vector_size_t n, cudaStream_t stream) {
  BufferPtr values = AlignedBuffer::allocate<T>(n ? n : 1, pool);
  if (n) VB2_CU(cudaMemcpyAsync(values->asMutable<T>(), devValues, size_t(n) * sizeof(T), cudaMemcpyDeviceToHost, stream));
  BufferPtr nulls;
  if (devNulls) {
    nulls = allocateNulls(n, pool);
    VB2_CU(cudaMemcpyAsync(nulls->asMutable<uint8_t>(), devNulls, bits::nbytes(n), cudaMemcpyDeviceToHost, stream));
  }
  return std::make_shared<FlatVector<T>>(pool, type, nulls, n, values);
}

VectorPtr downloadValues(memory::MemoryPool* pool, const TypePtr& type, int32_t vb2type, const void* values, const void* aux,
                         const uint64_t* nulls, vector_size_t n, cudaStream_t stream) {
  switch (vb2type) {
    case VB2_INTEGER: return downloadFlat<int32_t>(pool, type, values, nulls, n, stream);
    case VB2_BIGINT: return downloadFlat<int64_t>(pool, type, values, nulls, n, stream);
    case VB2_DOUBLE: return downloadFlat<double>(pool, type, values, nulls, n, stream);
    case VB2_BOOLEAN: {
      BufferPtr v = std::make_shared<Buffer>(bits::nbytes(n ? n : 1), pool);
      if (n) VB2_CU(cudaMemcpyAsync(v->asMutable<uint8_t>(), values, bits::nbytes(n), cudaMemcpyDeviceToHost, stream));
      BufferPtr nb;
      if (nulls) {
        nb = allocateNulls(n, pool);
        VB2_CU(cudaMemcpyAsync(nb->asMutable<uint8_t>(), nulls, bits::nbytes(n), cudaMemcpyDeviceToHost, stream));
      }
      return std::make_shared<FlatVector<bool>>(pool, type, nb, n, v);
    }
    default: {  // VARCHAR: offsets + chars -> StringViews over one string buffer
      std::vector<int32_t> off(size_t(n) + 1, 0);
      if (n) VB2_CU(cudaMemcpyAsync(off.data(), values, off.size() * 4, cudaMemcpyDeviceToHost, stream));
      BufferPtr nb;
      if (nulls) {
        nb = allocateNulls(n, pool);
        VB2_CU(cudaMemcpyAsync(nb->asMutable<uint8_t>(), nulls, bits::nbytes(n), cudaMemcpyDeviceToHost, stream));
      }
      VB2_CU(cudaStreamSynchronize(stream));
      const size_t nchars = n ? off[n] : 0;
      BufferPtr chars = std::make_shared<Buffer>(nchars ? nchars : 1, pool);
      if (nchars) VB2_CU(cudaMemcpyAsync(chars->asMutable<char>(), aux, nchars, cudaMemcpyDeviceToHost, stream));
      VB2_CU(cudaStreamSynchronize(stream));
      BufferPtr views = AlignedBuffer::allocate<StringView>(n ? n : 1, pool);
      auto* sv = views->asMutable<StringView>();
      for (vector_size_t i = 0; i < n; ++i) sv[i] = StringView(chars->as<char>() + off[i], off[i + 1] - off[i]);
      return std::make_shared<FlatVector<StringView>>(pool, type, nb, n, views, std::vector<BufferPtr>{chars});
    }
  }
}

}  // namespace

namespace {
// View over a slice of a host mirror; the buffer keeps the pinned block alive.
BufferPtr mirrorView(const std::shared_ptr<const HostMirror>& m, const uint8_t* p, size_t bytes) {
  return BufferPtr(new Buffer(p, bytes), [m](Buffer* b) { delete b; });
}

// Host vector of a column whose buffers all lie inside the batch's host mirror (no device access).
VectorPtr mirroredColumn(memory::MemoryPool* pool, const DeviceColumn& col, const std::shared_ptr<const HostMirror>& m) {
  const vb2_column& d = col.desc;
  const vector_size_t n = static_cast<vector_size_t>(d.size);
  BufferPtr nulls;
  if (d.nulls) {
    const uint8_t* hn = m->hostOf(d.nulls);
    if (!hn) return nullptr;
    nulls = mirrorView(m, hn, bits::nbytes(n));
  }
  if (d.encoding == VB2_FLAT && d.type != VB2_VARCHAR) {
    const uint8_t* hv = m->hostOf(d.values);
    if (!hv) return nullptr;
    switch (d.type) {
      case VB2_INTEGER: return std::make_shared<FlatVector<int32_t>>(pool, col.type, nulls, n, mirrorView(m, hv, size_t(n) * 4));
      case VB2_BIGINT: return std::make_shared<FlatVector<int64_t>>(pool, col.type, nulls, n, mirrorView(m, hv, size_t(n) * 8));
      case VB2_DOUBLE: return std::make_shared<FlatVector<double>>(pool, col.type, nulls, n, mirrorView(m, hv, size_t(n) * 8));
      case VB2_BOOLEAN: return std::make_shared<FlatVector<bool>>(pool, col.type, nulls, n, mirrorView(m, hv, bits::nbytes(n)));
      default: return nullptr;
    }
  }
  if (d.encoding == VB2_DICTIONARY && d.type == VB2_VARCHAR && col.alphabet && !d.dict_nulls &&
      static_cast<int64_t>(col.alphabet->values.size()) == d.dict_size) {
    const uint8_t* hi = m->hostOf(d.indices);
    if (!hi) return nullptr;
    // the dictionary's strings are known on the host (the operator that made the column keeps its alphabet)
    size_t total = 0;
    for (auto& v : col.alphabet->values) total += v.size();
    BufferPtr chars = std::make_shared<Buffer>(total ? total : 1, pool);
    BufferPtr views = AlignedBuffer::allocate<StringView>(col.alphabet->values.empty() ? 1 : col.alphabet->values.size(), pool);
    char* cp = chars->asMutable<char>();
    auto* sv = views->asMutable<StringView>();
    size_t at = 0;
    for (size_t i = 0; i < col.alphabet->values.size(); ++i) {
      const std::string& v = col.alphabet->values[i];
      std::memcpy(cp + at, v.data(), v.size());
      sv[i] = StringView(cp + at, v.size());
      at += v.size();
    }
    auto base = std::make_shared<FlatVector<StringView>>(pool, col.type, nullptr, static_cast<vector_size_t>(col.alphabet->values.size()), views,
                                                         std::vector<BufferPtr>{chars});
    return BaseVector::wrapInDictionary(nulls, mirrorView(m, hi, size_t(n) * 4), n, base);
  }
  return nullptr;
}
}  // namespace

RowVectorPtr toHost(const B200VectorPtr& dev) {
  cudaStream_t stream = dev->stream();
  auto* pool = dev->pool();
  std::vector<VectorPtr> children;
  bool touchedDevice = false;
  for (auto& col : dev->columns()) {
    const vb2_column& d = col->desc;
    const vector_size_t n = static_cast<vector_size_t>(d.size);
    if (dev->mirror()) {
      if (VectorPtr v = mirroredColumn(pool, *col, dev->mirror())) {
        children.push_back(std::move(v));
        continue;
      }
    }
    touchedDevice = true;
    if (d.encoding == VB2_FLAT) {
      children.push_back(downloadValues(pool, col->type, d.type, d.values, d.aux, d.nulls, n, stream));
    } else if (d.encoding == VB2_DICTIONARY) {
      VectorPtr base = downloadValues(pool, col->type, d.type, d.values, d.aux, d.dict_nulls, static_cast<vector_size_t>(d.dict_size), stream);
      BufferPtr idx = allocateIndices(n ? n : 1, pool);
      if (n) VB2_CU(cudaMemcpyAsync(idx->asMutable<int32_t>(), d.indices, size_t(n) * 4, cudaMemcpyDeviceToHost, stream));
      BufferPtr nb;
      if (d.nulls) {
        nb = allocateNulls(n, pool);
        VB2_CU(cudaMemcpyAsync(nb->asMutable<uint8_t>(), d.nulls, bits::nbytes(n), cudaMemcpyDeviceToHost, stream));
      }
      children.push_back(BaseVector::wrapInDictionary(nb, idx, n, base));
    } else {
      // constant: materialise as a one-entry dictionary (every row index 0)
      VectorPtr base = downloadValues(pool, col->type, d.type, d.values, d.aux, d.nulls, 1, stream);
      BufferPtr idx = AlignedBuffer::allocate<vector_size_t>(n ? n : 1, pool, 0);
      children.push_back(BaseVector::wrapInDictionary(nullptr, idx, n, base));
    }
  }
  if (touchedDevice) VB2_CU(cudaStreamSynchronize(stream));
  return std::make_shared<RowVector>(pool, dev->type(), nullptr, dev->size(), std::move(children));
}

// Device copy (int32 offsets + chars) of a small alphabet, shared by every page that carries the same
// strings: exchanged dictionaries repeat from query to query, the upload happens once per content.
void deviceAlphabet(const HostAlphabet& a, cudaStream_t st, DeviceBufferPtr& offBuf, DeviceBufferPtr& charBuf) {
  static std::mutex mu;
  static std::map<std::string, std::pair<DeviceBufferPtr, DeviceBufferPtr>> cache;
  std::string key;
  std::vector<int32_t> off(a.values.size() + 1, 0);
  std::string chars;
  for (size_t k = 0; k < a.values.size(); ++k) {
    chars += a.values[k];
    off[k + 1] = static_cast<int32_t>(chars.size());
    key += std::to_string(a.values[k].size()) + ":" + a.values[k];
  }
  int dev = 0;
  cudaGetDevice(&dev);
  key = std::to_string(dev) + "|" + key;
  std::lock_guard<std::mutex> l(mu);
  auto it = cache.find(key);
  if (it == cache.end()) {
    if (cache.size() > 256) cache.clear();
    // plain cudaMalloc'd, never freed while cached: usable from any stream once the copies below are done
    auto ob = allocDevice(off.size() * 4, st);
    auto cb = allocDevice(chars.size() + 1, st);
    VB2_CU(cudaMemcpyAsync(ob->data(), off.data(), off.size() * 4, cudaMemcpyHostToDevice, st));
    if (!chars.empty()) VB2_CU(cudaMemcpyAsync(cb->data(), chars.data(), chars.size(), cudaMemcpyHostToDevice, st));
    VB2_CU(cudaStreamSynchronize(st));
    it = cache.emplace(key, std::make_pair(ob, cb)).first;
  }
  offBuf = it->second.first;
  charBuf = it->second.second;
}


DeviceColumnPtr borrowFlatColumn(TypePtr type, const void* values, int64_t size) {
  auto col = std::make_shared<DeviceColumn>();
  col->type = std::move(type);
  col->desc.type = veloxTypeToVb2(col->type);
  col->desc.encoding = VB2_FLAT;
  col->desc.size = size;
  col->desc.values = values;
  return col;
}

B200VectorPtr sliceVector(const B200VectorPtr& v, int64_t offset, int64_t length) {
  VELOX_CHECK(offset % 64 == 0 && offset >= 0 && offset + length <= v->size(), "sliceVector: bad range");
  std::vector<DeviceColumnPtr> cols;
  for (const auto& c : v->columns()) {
    auto s = std::make_shared<DeviceColumn>(*c);  // shares the owners
    vb2_column& d = s->desc;
    d.size = length;
    if (d.encoding == VB2_DICTIONARY) {
      if (d.nulls) d.nulls += offset / 64;
      d.indices += offset;
    } else if (d.encoding == VB2_FLAT) {
      if (d.nulls) d.nulls += offset / 64;
      const char* base = static_cast<const char*>(d.values);
      if (d.type == VB2_BOOLEAN) base += offset / 8;
      else if (d.type == VB2_VARCHAR) base += offset * 4;  // int32 offsets; chars (aux) stay absolute
      else base += offset * widthOf(d.type);
      d.values = base;
    }
    cols.push_back(std::move(s));
  }
  return std::make_shared<B200Vector>(v->pool(), v->type(), static_cast<vector_size_t>(length), std::move(cols), v->stream());
}

NvtxRange::NvtxRange(const char* method, const std::string& operatorType, const std::string& planNodeId) {
  nvtxEventAttributes_t a{};
  a.version = NVTX_VERSION;
  a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
  a.colorType = NVTX_COLOR_ARGB;
  a.color = 0xff76b900u ^ static_cast<uint32_t>(std::hash<std::string>{}(operatorType) & 0x00ffffffu);  // one colour per operator type
  a.messageType = NVTX_MESSAGE_TYPE_ASCII;
  const std::string msg = operatorType + "::" + method + " [" + planNodeId + "]";
  a.message.ascii = msg.c_str();  // NVTX copies the string during the call
  nvtxRangePushEx(&a);
}
NvtxRange::~NvtxRange() { nvtxRangePop(); }

}  // namespace velox_b200
