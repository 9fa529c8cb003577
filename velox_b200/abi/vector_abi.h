// Contract shim — the vector data contract of the reference, restated without its dependencies
// (folly, fmt, xsimd ...) so the B200 operators can be built and exercised where Velox itself
// cannot be compiled. Only the members the operator hot path touches are mirrored; layouts and
// meanings follow the reference:
//   Type / TypeKind ........ velox/type/Type.h (kind numbering, DATE = INTEGER days :1305)
//   StringView ............. velox/type/StringView.h:76-77 (16 B, 12 B inline)
//   Buffer / AlignedBuffer . velox/buffer/Buffer.h:352-358 (64-byte aligned)
//   nulls .................. velox/common/base/Nulls.h:26-27 (bit 1 = not null, LSB first)
//   BaseVector ............. velox/vector/BaseVector.h
//   FlatVector ............. velox/vector/FlatVector.h:604-607 (values_ / rawValues_)
//   DictionaryVector ....... velox/vector/DictionaryVector.h:275-278 (int32 indices_ + dictionaryValues_)
//   ConstantVector ......... velox/vector/ConstantVector.h
//   RowVector .............. velox/vector/ComplexVector.h
//   SelectivityVector ...... velox/vector/SelectivityVector.h:39
// When real Velox headers are available, define VELOX_B200_WITH_REAL_VELOX and include them
// instead; the operator sources only use the names declared here.
#pragma once
#ifndef VELOX_B200_WITH_REAL_VELOX

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace facebook::velox {

using vector_size_t = int32_t;

// ---- errors (velox/common/base/VeloxException.h) ------------------------------------------------
class VeloxException : public std::runtime_error {
 public:
  using std::runtime_error::runtime_error;
};
class VeloxRuntimeError : public VeloxException {
 public:
  using VeloxException::VeloxException;
};
class VeloxUserError : public VeloxException {
 public:
  using VeloxException::VeloxException;
};
#define VELOX_CHECK(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) throw ::facebook::velox::VeloxRuntimeError(std::string("VELOX_CHECK failed: ") + (msg)); \
  } while (0)
#define VELOX_USER_CHECK(cond, msg)                                                \
  do {                                                                             \
    if (!(cond)) throw ::facebook::velox::VeloxUserError(std::string(msg));        \
  } while (0)
#define VELOX_FAIL(msg) throw ::facebook::velox::VeloxRuntimeError(std::string(msg))
#define VELOX_NYI(msg) throw ::facebook::velox::VeloxRuntimeError(std::string("Not yet implemented: ") + (msg))
#define VELOX_UNSUPPORTED(msg) throw ::facebook::velox::VeloxRuntimeError(std::string("Unsupported: ") + (msg))

// ---- types ------------------------------------------------------------------------------------
enum class TypeKind : int8_t {
  BOOLEAN = 0, TINYINT = 1, SMALLINT = 2, INTEGER = 3, BIGINT = 4, REAL = 5, DOUBLE = 6, VARCHAR = 7,
  VARBINARY = 8, TIMESTAMP = 9, HUGEINT = 10, ARRAY = 30, MAP = 31, ROW = 32, UNKNOWN = 33, INVALID = 36
};

class Type;
using TypePtr = std::shared_ptr<const Type>;
class Type {
 public:
  Type(TypeKind kind, bool isDate = false) : kind_(kind), isDate_(isDate) {}
  Type(std::vector<std::string> names, std::vector<TypePtr> children)
      : kind_(TypeKind::ROW), names_(std::move(names)), children_(std::move(children)) {}
  TypeKind kind() const { return kind_; }
  bool isDate() const { return isDate_; }
  bool isRow() const { return kind_ == TypeKind::ROW; }
  uint32_t size() const { return static_cast<uint32_t>(children_.size()); }
  const TypePtr& childAt(uint32_t i) const { return children_.at(i); }
  const std::string& nameOf(uint32_t i) const { return names_.at(i); }
  const std::vector<std::string>& names() const { return names_; }
  const std::vector<TypePtr>& children() const { return children_; }
  std::optional<uint32_t> getChildIdxIfExists(const std::string& name) const {
    for (uint32_t i = 0; i < names_.size(); ++i)
      if (names_[i] == name) return i;
    return std::nullopt;
  }
  std::string toString() const {
    switch (kind_) {
      case TypeKind::BOOLEAN: return "BOOLEAN";
      case TypeKind::INTEGER: return isDate_ ? "DATE" : "INTEGER";
      case TypeKind::BIGINT: return "BIGINT";
      case TypeKind::DOUBLE: return "DOUBLE";
      case TypeKind::VARCHAR: return "VARCHAR";
      case TypeKind::ROW: {
        std::string s = "ROW<";
        for (size_t i = 0; i < children_.size(); ++i) s += (i ? "," : "") + names_[i] + ":" + children_[i]->toString();
        return s + ">";
      }
      default: return "UNKNOWN";
    }
  }
  bool equivalent(const Type& o) const {
    if (kind_ != o.kind_ || children_.size() != o.children_.size()) return false;
    for (size_t i = 0; i < children_.size(); ++i)
      if (!children_[i]->equivalent(*o.children_[i])) return false;
    return true;
  }

 private:
  TypeKind kind_;
  bool isDate_ = false;
  std::vector<std::string> names_;
  std::vector<TypePtr> children_;
};
using RowType = Type;
using RowTypePtr = std::shared_ptr<const RowType>;

inline TypePtr BOOLEAN() { static TypePtr t = std::make_shared<Type>(TypeKind::BOOLEAN); return t; }
inline TypePtr INTEGER() { static TypePtr t = std::make_shared<Type>(TypeKind::INTEGER); return t; }
inline TypePtr DATE() { static TypePtr t = std::make_shared<Type>(TypeKind::INTEGER, true); return t; }
inline TypePtr BIGINT() { static TypePtr t = std::make_shared<Type>(TypeKind::BIGINT); return t; }
inline TypePtr DOUBLE() { static TypePtr t = std::make_shared<Type>(TypeKind::DOUBLE); return t; }
inline TypePtr VARCHAR() { static TypePtr t = std::make_shared<Type>(TypeKind::VARCHAR); return t; }
inline RowTypePtr ROW(std::vector<std::string> names, std::vector<TypePtr> types) {
  return std::make_shared<const Type>(std::move(names), std::move(types));
}

// 16-byte string reference: 4-byte size, then either 12 inline bytes or 4-byte prefix + pointer.
struct StringView {
  static constexpr uint32_t kInlineSize = 12;
  uint32_t size_ = 0;
  char prefix_[4] = {0, 0, 0, 0};
  union {
    char inlined[8];
    const char* data;
  } value_ = {{0}};
  StringView() = default;
  StringView(const char* data, size_t len) : size_(static_cast<uint32_t>(len)) {
    if (isInline()) {
      std::memset(prefix_, 0, 4);
      std::memset(value_.inlined, 0, 8);
      if (len) std::memcpy(prefix_, data, len < 4 ? len : 4);
      if (len > 4) std::memcpy(value_.inlined, data + 4, len - 4);
    } else {
      std::memcpy(prefix_, data, 4);
      value_.data = data;
    }
  }
  explicit StringView(std::string_view s) : StringView(s.data(), s.size()) {}
  bool isInline() const { return size_ <= kInlineSize; }
  uint32_t size() const { return size_; }
  const char* data() const { return isInline() ? prefix_ : value_.data; }
  std::string str() const { return std::string(data(), size_); }
  operator std::string_view() const { return std::string_view(data(), size_); }
};
static_assert(sizeof(StringView) == 16, "StringView is 16 bytes (velox/type/StringView.h)");

// ---- buffers ----------------------------------------------------------------------------------
namespace memory {
class MemoryPool {  // velox/common/memory/MemoryPool.h — accounting only in the shim
 public:
  explicit MemoryPool(std::string name = "b200") : name_(std::move(name)) {}
  const std::string& name() const { return name_; }
  int64_t usedBytes() const { return used_; }
  void reserve(int64_t b) { used_ += b; }
  void release(int64_t b) { used_ -= b; }

 private:
  std::string name_;
  int64_t used_ = 0;
};
}  // namespace memory

class Buffer {
 public:
  Buffer(size_t bytes, memory::MemoryPool* pool) : size_(bytes), pool_(pool) {
    capacity_ = (bytes + 63) / 64 * 64 + 64;
    data_ = static_cast<uint8_t*>(std::aligned_alloc(64, capacity_));
    if (!data_) throw VeloxRuntimeError("allocation failed");
    if (pool_) pool_->reserve(static_cast<int64_t>(capacity_));
  }
  // Non-owning view over caller memory (BufferView in velox/buffer/Buffer.h): zero-copy import.
  Buffer(const void* borrowed, size_t bytes) : data_(const_cast<uint8_t*>(static_cast<const uint8_t*>(borrowed))), size_(bytes), capacity_(bytes), pool_(nullptr), owned_(false) {}
  ~Buffer() {
    if (!owned_) return;
    std::free(data_);
    if (pool_) pool_->release(static_cast<int64_t>(capacity_));
  }
  Buffer(const Buffer&) = delete;
  template <class T> const T* as() const { return reinterpret_cast<const T*>(data_); }
  template <class T> T* asMutable() { return reinterpret_cast<T*>(data_); }
  size_t size() const { return size_; }
  size_t capacity() const { return capacity_; }

 private:
  uint8_t* data_;
  size_t size_, capacity_;
  memory::MemoryPool* pool_;
  bool owned_ = true;
};
using BufferPtr = std::shared_ptr<Buffer>;

struct AlignedBuffer {
  template <class T>
  static BufferPtr allocate(size_t n, memory::MemoryPool* pool, std::optional<T> init = std::nullopt) {
    auto b = std::make_shared<Buffer>(n * sizeof(T), pool);
    if (init) {
      T* p = b->asMutable<T>();
      for (size_t i = 0; i < n; ++i) p[i] = *init;
    }
    return b;
  }
};

namespace bits {
inline uint64_t nwords(uint64_t bits) { return (bits + 63) / 64; }
inline uint64_t nbytes(uint64_t bits) { return nwords(bits) * 8; }
inline bool isBitSet(const uint64_t* b, uint64_t i) { return (b[i >> 6] >> (i & 63)) & 1; }
inline void setBit(uint64_t* b, uint64_t i, bool v = true) {
  if (v) b[i >> 6] |= 1ull << (i & 63);
  else b[i >> 6] &= ~(1ull << (i & 63));
}
inline void clearBit(uint64_t* b, uint64_t i) { setBit(b, i, false); }
inline constexpr bool kNull = false;     // velox/common/base/Nulls.h
inline constexpr bool kNotNull = true;
inline bool isBitNull(const uint64_t* b, uint64_t i) { return !isBitSet(b, i); }
}  // namespace bits

inline BufferPtr allocateNulls(vector_size_t size, memory::MemoryPool* pool, bool initValue = bits::kNotNull) {
  auto b = std::make_shared<Buffer>(bits::nbytes(size), pool);
  std::memset(b->asMutable<uint8_t>(), initValue ? 0xff : 0, b->capacity());
  return b;
}
inline BufferPtr allocateIndices(vector_size_t size, memory::MemoryPool* pool) {
  return AlignedBuffer::allocate<vector_size_t>(size, pool);
}

// ---- vectors ----------------------------------------------------------------------------------
namespace VectorEncoding {
enum class Simple { BIASED, CONSTANT, DICTIONARY, FLAT, SEQUENCE, ROW, MAP, ARRAY, LAZY, FUNCTION };
}

class BaseVector;
using VectorPtr = std::shared_ptr<BaseVector>;

class BaseVector {
 public:
  BaseVector(memory::MemoryPool* pool, TypePtr type, VectorEncoding::Simple encoding, BufferPtr nulls, vector_size_t length)
      : pool_(pool), type_(std::move(type)), encoding_(encoding), nulls_(std::move(nulls)), length_(length) {}
  virtual ~BaseVector() = default;
  const TypePtr& type() const { return type_; }
  TypeKind typeKind() const { return type_->kind(); }
  VectorEncoding::Simple encoding() const { return encoding_; }
  vector_size_t size() const { return length_; }
  memory::MemoryPool* pool() const { return pool_; }
  const BufferPtr& nulls() const { return nulls_; }
  const uint64_t* rawNulls() const { return nulls_ ? nulls_->as<uint64_t>() : nullptr; }
  virtual bool mayHaveNulls() const { return nulls_ != nullptr; }
  virtual bool isNullAt(vector_size_t i) const { return nulls_ && bits::isBitNull(rawNulls(), i); }
  void setNull(vector_size_t i, bool isNull) {
    if (!nulls_) {
      if (!isNull) return;
      nulls_ = allocateNulls(length_, pool_);
    }
    bits::setBit(nulls_->asMutable<uint64_t>(), i, !isNull);
  }
  bool isFlatEncoding() const { return encoding_ == VectorEncoding::Simple::FLAT; }
  bool isConstantEncoding() const { return encoding_ == VectorEncoding::Simple::CONSTANT; }
  template <class T> T* as() { return dynamic_cast<T*>(this); }
  template <class T> const T* as() const { return dynamic_cast<const T*>(this); }
  // velox/vector/BaseVector.h wrapInDictionary: zero-copy dictionary over `vector`.
  static VectorPtr wrapInDictionary(BufferPtr nulls, BufferPtr indices, vector_size_t size, VectorPtr vector);

 protected:
  memory::MemoryPool* pool_;
  TypePtr type_;
  VectorEncoding::Simple encoding_;
  BufferPtr nulls_;
  vector_size_t length_;
};

template <class T>
class FlatVector : public BaseVector {
 public:
  FlatVector(memory::MemoryPool* pool, TypePtr type, BufferPtr nulls, vector_size_t length, BufferPtr values,
             std::vector<BufferPtr> stringBuffers = {})
      : BaseVector(pool, std::move(type), VectorEncoding::Simple::FLAT, std::move(nulls), length),
        values_(std::move(values)), stringBuffers_(std::move(stringBuffers)) {}
  const BufferPtr& values() const { return values_; }
  const T* rawValues() const { return values_ ? values_->template as<T>() : nullptr; }
  T* mutableRawValues() { return values_ ? values_->template asMutable<T>() : nullptr; }
  T valueAt(vector_size_t i) const {
    if constexpr (std::is_same_v<T, bool>) return bits::isBitSet(values_->template as<uint64_t>(), i);
    else return rawValues()[i];
  }
  const std::vector<BufferPtr>& stringBuffers() const { return stringBuffers_; }

 private:
  BufferPtr values_;  // T[length]; bool is bit-packed
  std::vector<BufferPtr> stringBuffers_;
};

template <class T>
class DictionaryVector : public BaseVector {
 public:
  DictionaryVector(memory::MemoryPool* pool, BufferPtr nulls, vector_size_t length, VectorPtr dictionaryValues, BufferPtr indices)
      : BaseVector(pool, dictionaryValues->type(), VectorEncoding::Simple::DICTIONARY, std::move(nulls), length),
        indices_(std::move(indices)), dictionaryValues_(std::move(dictionaryValues)) {}
  const BufferPtr& indices() const { return indices_; }
  const vector_size_t* rawIndices() const { return indices_->as<vector_size_t>(); }
  const VectorPtr& valueVector() const { return dictionaryValues_; }
  bool mayHaveNulls() const override { return nulls_ != nullptr || dictionaryValues_->mayHaveNulls(); }
  bool isNullAt(vector_size_t i) const override {
    if (BaseVector::isNullAt(i)) return true;
    return dictionaryValues_->isNullAt(rawIndices()[i]);
  }

 private:
  BufferPtr indices_;
  VectorPtr dictionaryValues_;
};

template <class T>
class ConstantVector : public BaseVector {
 public:
  ConstantVector(memory::MemoryPool* pool, vector_size_t length, bool isNull, TypePtr type, T value)
      : BaseVector(pool, std::move(type), VectorEncoding::Simple::CONSTANT, nullptr, length), value_(std::move(value)), isNull_(isNull) {}
  const T& value() const { return value_; }
  bool mayHaveNulls() const override { return isNull_; }
  bool isNullAt(vector_size_t) const override { return isNull_; }

 private:
  T value_;
  bool isNull_;
  std::string stringStorage_;

 public:
  // owns the characters of a non-inline VARCHAR constant
  void setStringStorage(std::string s) {
    stringStorage_ = std::move(s);
    if constexpr (std::is_same_v<T, StringView>) value_ = StringView(stringStorage_.data(), stringStorage_.size());
  }
};

class RowVector : public BaseVector {
 public:
  RowVector(memory::MemoryPool* pool, TypePtr type, BufferPtr nulls, vector_size_t length, std::vector<VectorPtr> children)
      : BaseVector(pool, std::move(type), VectorEncoding::Simple::ROW, std::move(nulls), length), children_(std::move(children)) {}
  const VectorPtr& childAt(uint32_t i) const { return children_.at(i); }
  VectorPtr& childAt(uint32_t i) { return children_.at(i); }
  const std::vector<VectorPtr>& children() const { return children_; }
  size_t childrenSize() const { return children_.size(); }

 private:
  std::vector<VectorPtr> children_;
};
using RowVectorPtr = std::shared_ptr<RowVector>;

inline VectorPtr BaseVector::wrapInDictionary(BufferPtr nulls, BufferPtr indices, vector_size_t size, VectorPtr vector) {
  auto* pool = vector->pool();
  switch (vector->typeKind()) {
    case TypeKind::BOOLEAN: return std::make_shared<DictionaryVector<bool>>(pool, nulls, size, vector, indices);
    case TypeKind::INTEGER: return std::make_shared<DictionaryVector<int32_t>>(pool, nulls, size, vector, indices);
    case TypeKind::BIGINT: return std::make_shared<DictionaryVector<int64_t>>(pool, nulls, size, vector, indices);
    case TypeKind::DOUBLE: return std::make_shared<DictionaryVector<double>>(pool, nulls, size, vector, indices);
    case TypeKind::VARCHAR: return std::make_shared<DictionaryVector<StringView>>(pool, nulls, size, vector, indices);
    default: VELOX_UNSUPPORTED("wrapInDictionary: type " + vector->type()->toString());
  }
}

// Selected rows of a batch: bitmap + [begin, end) bounds (velox/vector/SelectivityVector.h:39).
class SelectivityVector {
 public:
  SelectivityVector() = default;
  explicit SelectivityVector(vector_size_t length, bool allSelected = true) { resize(length, allSelected); }
  void resize(vector_size_t length, bool value = true) {
    size_ = length;
    bits_.assign(bits::nwords(length), value ? ~0ull : 0ull);
    updateBounds();
  }
  vector_size_t size() const { return size_; }
  vector_size_t begin() const { return begin_; }
  vector_size_t end() const { return end_; }
  bool isValid(vector_size_t i) const { return bits::isBitSet(bits_.data(), i); }
  void setValid(vector_size_t i, bool v) { bits::setBit(bits_.data(), i, v); }
  void setAll() { resize(size_, true); }
  void clearAll() { resize(size_, false); }
  bool isAllSelected() const { return countSelected() == size_; }
  bool hasSelections() const { return begin_ < end_; }
  void setFromBits(const uint64_t* bits, vector_size_t size) {
    size_ = size;
    bits_.assign(bits, bits + bits::nwords(size));
    updateBounds();
  }
  void updateBounds() {
    begin_ = size_;
    end_ = 0;
    for (vector_size_t i = 0; i < size_; ++i)
      if (isValid(i)) {
        if (begin_ == size_) begin_ = i;
        end_ = i + 1;
      }
    if (begin_ == size_) begin_ = end_ = 0;
  }
  vector_size_t countSelected() const {
    vector_size_t c = 0;
    for (vector_size_t i = 0; i < size_; ++i) c += isValid(i);
    return c;
  }
  template <class F>
  void applyToSelected(F&& f) const {
    for (vector_size_t i = begin_; i < end_; ++i)
      if (isValid(i)) f(i);
  }
  const uint64_t* asRange() const { return bits_.data(); }

 private:
  std::vector<uint64_t> bits_;
  vector_size_t size_ = 0, begin_ = 0, end_ = 0;
};

}  // namespace facebook::velox

#endif  // VELOX_B200_WITH_REAL_VELOX
