// Arrow C data interface -> RowVector, in C++ (the scan-side step in front of the
// path). Restates velox/vector/arrow/Bridge.cpp importFromArrowAsViewer / AsOwner for the types of the
// hot path; the device import pattern it feeds is velox/experimental/cudf/exec/VeloxCudfInterop.cpp:115-240
// (Arrow host buffers -> device table). Buffers are viewed, not copied, wherever Arrow's layout equals
// Velox's: fixed-width values, LSB-first validity (1 = valid) and boolean bitmaps, int32 dictionary
// indices. Copies: bitmaps of arrays with a non-zero offset (re-based to bit 0), and the 16-byte
// StringViews of utf8 columns (their characters stay in the Arrow buffer).
#include <cstring>

#include "../../abi/arrow_abi.h"
#include "device.h"

namespace facebook::velox {

namespace {

BufferPtr viewOf(const void* p, size_t bytes) { return p ? std::make_shared<Buffer>(p, bytes) : nullptr; }

// bits [offset, offset + n) of `bits`, re-based to bit 0 (viewed when offset == 0)
BufferPtr bitsAt(const void* bits, int64_t offset, int64_t n, memory::MemoryPool* pool) {
  if (!bits) return nullptr;
  if (offset == 0) return viewOf(bits, velox::bits::nbytes(n));
  auto out = std::make_shared<Buffer>(velox::bits::nbytes(n), pool);
  std::memset(out->asMutable<uint8_t>(), 0, out->capacity());
  const uint64_t* src = nullptr;
  const uint8_t* bytes = static_cast<const uint8_t*>(bits);
  (void)src;
  uint64_t* dst = out->asMutable<uint64_t>();
  for (int64_t i = 0; i < n; ++i) {
    const int64_t s = offset + i;
    if ((bytes[s >> 3] >> (s & 7)) & 1) dst[i >> 6] |= 1ull << (i & 63);
  }
  return out;
}

TypePtr typeOfFormat(const std::string& f) {
  if (f == "l") return BIGINT();
  if (f == "i") return INTEGER();
  if (f == "tdD") return DATE();
  if (f == "g") return DOUBLE();
  if (f == "b") return BOOLEAN();
  if (f == "u") return VARCHAR();
  VELOX_UNSUPPORTED("Arrow format '" + f + "' (BIGINT l, INTEGER i, DATE tdD, DOUBLE g, BOOLEAN b, VARCHAR u are supported)");
}

VectorPtr importFlat(const std::string& format, const ArrowArray& a, memory::MemoryPool* pool) {
  const int64_t n = a.length, off = a.offset;
  const vector_size_t size = static_cast<vector_size_t>(n);
  const TypePtr type = typeOfFormat(format);
  VELOX_CHECK(a.n_buffers >= 2, "Arrow array of format '" + format + "' needs validity + data buffers");
  BufferPtr nulls = a.null_count == 0 ? nullptr : bitsAt(a.buffers[0], off, n, pool);
  switch (type->kind()) {
    case TypeKind::BIGINT: return std::make_shared<FlatVector<int64_t>>(pool, type, nulls, size, viewOf(static_cast<const int64_t*>(a.buffers[1]) + off, n * 8));
    case TypeKind::INTEGER: return std::make_shared<FlatVector<int32_t>>(pool, type, nulls, size, viewOf(static_cast<const int32_t*>(a.buffers[1]) + off, n * 4));
    case TypeKind::DOUBLE: return std::make_shared<FlatVector<double>>(pool, type, nulls, size, viewOf(static_cast<const double*>(a.buffers[1]) + off, n * 8));
    case TypeKind::BOOLEAN: return std::make_shared<FlatVector<bool>>(pool, type, nulls, size, bitsAt(a.buffers[1], off, n, pool));
    default: {  // VARCHAR: int32 offsets + characters
      VELOX_CHECK(a.n_buffers >= 3, "Arrow utf8 array needs validity, offsets and data buffers");
      const int32_t* o = static_cast<const int32_t*>(a.buffers[1]) + off;
      const char* chars = static_cast<const char*>(a.buffers[2]);
      BufferPtr views = AlignedBuffer::allocate<StringView>(n ? n : 1, pool);
      auto* sv = views->asMutable<StringView>();
      for (int64_t i = 0; i < n; ++i) sv[i] = StringView(chars + o[i], o[i + 1] - o[i]);
      return std::make_shared<FlatVector<StringView>>(pool, type, nulls, size, views);
    }
  }
}

VectorPtr importColumn(const ArrowSchema& s, const ArrowArray& a, memory::MemoryPool* pool) {
  const std::string format = s.format ? s.format : "";
  if (s.dictionary) {
    // dictionary-encoded: `format` is the index type, the value type lives in s.dictionary
    VELOX_CHECK(format == "i", "Arrow dictionary indices must be int32 (format 'i'), not '" + format + "'");
    VELOX_CHECK(a.dictionary != nullptr, "Arrow dictionary array without its dictionary");
    VectorPtr base = importFlat(s.dictionary->format ? s.dictionary->format : "", *a.dictionary, pool);
    const int64_t n = a.length, off = a.offset;
    BufferPtr nulls = a.null_count == 0 ? nullptr : bitsAt(a.buffers[0], off, n, pool);
    BufferPtr indices = viewOf(static_cast<const int32_t*>(a.buffers[1]) + off, n * 4);
    if (nulls) {
      // Arrow leaves the index of a NULL slot undefined; DictionaryVector wants it in range
      auto fixed = AlignedBuffer::allocate<vector_size_t>(n ? n : 1, pool);
      const int32_t* src = static_cast<const int32_t*>(a.buffers[1]) + off;
      auto* dst = fixed->asMutable<vector_size_t>();
      for (int64_t i = 0; i < n; ++i) dst[i] = velox::bits::isBitSet(nulls->as<uint64_t>(), i) ? src[i] : 0;
      indices = fixed;
    }
    return BaseVector::wrapInDictionary(nulls, indices, static_cast<vector_size_t>(n), base);
  }
  return importFlat(format, a, pool);
}

// keeps the moved-from Arrow structs and releases them with the vector
struct ArrowOwner {
  ArrowSchema schema;
  ArrowArray array;
  ~ArrowOwner() {
    if (array.release) array.release(&array);
    if (schema.release) schema.release(&schema);
  }
};
class OwningRowVector : public RowVector {
 public:
  OwningRowVector(const RowVector& v, std::shared_ptr<ArrowOwner> owner)
      : RowVector(v.pool(), v.type(), v.nulls(), v.size(), v.children()), owner_(std::move(owner)) {}

 private:
  std::shared_ptr<ArrowOwner> owner_;
};

}  // namespace

VectorPtr importFromArrowAsViewer(const ArrowSchema& arrowSchema, const ArrowArray& arrowArray, memory::MemoryPool* pool) {
  VELOX_CHECK(arrowSchema.release != nullptr && arrowArray.release != nullptr, "Arrow schema / array was released");
  const std::string format = arrowSchema.format ? arrowSchema.format : "";
  if (format != "+s") return importColumn(arrowSchema, arrowArray, pool);
  VELOX_CHECK(arrowSchema.n_children == arrowArray.n_children, "Arrow struct: schema and array disagree on the number of children");
  VELOX_CHECK(arrowArray.offset == 0 && arrowArray.null_count <= 0, "Arrow struct arrays (record batches) with an offset or NULL rows");
  std::vector<std::string> names;
  std::vector<TypePtr> types;
  std::vector<VectorPtr> children;
  for (int64_t c = 0; c < arrowSchema.n_children; ++c) {
    const ArrowSchema& cs = *arrowSchema.children[c];
    const ArrowArray& ca = *arrowArray.children[c];
    VELOX_CHECK(ca.length == arrowArray.length, "Arrow struct: child length differs from the batch's");
    children.push_back(importColumn(cs, ca, pool));
    names.push_back(cs.name ? cs.name : "c" + std::to_string(c));
    types.push_back(children.back()->type());
  }
  return std::make_shared<RowVector>(pool, ROW(names, types), nullptr, static_cast<vector_size_t>(arrowArray.length), std::move(children));
}

VectorPtr importFromArrowAsOwner(ArrowSchema& arrowSchema, ArrowArray& arrowArray, memory::MemoryPool* pool) {
  VectorPtr viewed = importFromArrowAsViewer(arrowSchema, arrowArray, pool);
  auto owner = std::make_shared<ArrowOwner>();
  owner->schema = arrowSchema;
  owner->array = arrowArray;
  arrowSchema.release = nullptr;  // marked released: the vector owns them now
  arrowArray.release = nullptr;
  auto* row = viewed->as<RowVector>();
  VELOX_CHECK(row != nullptr, "importFromArrowAsOwner: a struct array (record batch) is expected");
  return std::make_shared<OwningRowVector>(*row, std::move(owner));
}

}  // namespace facebook::velox
