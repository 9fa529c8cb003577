// TEST INFRASTRUCTURE — CPU oracle. Not part of the shipped product path.
//
// Restatement of the hashing used by Velox's key-hashing hot path.
//   * folly::hasher<int64_t>  == folly::hash::twang_mix64      (folly v2026.01.05.00, folly/hash/Hash.h;
//     the reference restates it in velox/experimental/gpu/tests/HashTableTest.cu:51-60)
//   * folly::hasher<int32_t/int16_t/int8_t> == jenkins_rev_mix32 of the sign-extended
//     value (restated in velox/experimental/gpu/tests/HashTableTest.cu:38-49)
//   * folly::hasher<double>: +0/-0 hash to 0, otherwise twang_mix64 of the bit pattern;
//     NaN canonicalised first (velox/type/FloatingPointUtil.h:100-109)
//   * bits::hashMix     velox/common/base/BitUtil.h:775-784
//   * bits::hashBytes   velox/common/base/BitUtil.cpp:177-230 (CRC32-C via SSE4.2)
//   * kNullHash = 1     velox/common/base/BitUtil.h:52
// Call sites: velox/exec/VectorHasher.cpp:62-126 (hashOne / hashValues).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <limits>
#include <nmmintrin.h>

namespace orc {

constexpr uint64_t kNullHash = 1;

inline uint64_t twang_mix64(uint64_t key) {
  key = (~key) + (key << 21);
  key = key ^ (key >> 24);
  key = key + (key << 3) + (key << 8);
  key = key ^ (key >> 14);
  key = key + (key << 2) + (key << 4);
  key = key ^ (key >> 28);
  key = key + (key << 31);
  return key;
}

inline uint32_t jenkins_rev_mix32(uint32_t key) {
  key += (key << 12);
  key ^= (key >> 22);
  key += (key << 4);
  key ^= (key >> 9);
  key += (key << 10);
  key ^= (key >> 2);
  key += (key << 7);
  key += (key << 12);
  return key;
}

inline uint64_t hash_mix(uint64_t upper, uint64_t lower) {
  const uint64_t kMul = 0x9ddfea08eb382d69ULL;
  uint64_t a = (lower ^ upper) * kMul;
  a ^= (a >> 47);
  uint64_t b = (upper ^ a) * kMul;
  b ^= (b >> 47);
  b *= kMul;
  return b;
}

inline uint64_t hash_i64(int64_t v) { return twang_mix64(static_cast<uint64_t>(v)); }
// folly integral_hasher for sizeof <= 4: sign-extends to int32 then jenkins_rev_mix32.
inline uint64_t hash_i32(int32_t v) { return jenkins_rev_mix32(static_cast<uint32_t>(v)); }
inline uint64_t hash_bool(bool v) { return v ? 1 : 0; }

inline uint64_t hash_f64_raw(double v) {
  if (v == 0.0) return 0;  // folly float_hasher: +0 and -0 hash alike
  uint64_t u;
  std::memcpy(&u, &v, 8);
  return twang_mix64(u);
}
inline uint64_t hash_f64(double v) {
  if (std::isnan(v)) return hash_f64_raw(std::numeric_limits<double>::quiet_NaN());
  return hash_f64_raw(v);
}

inline uint64_t load_partial_word(const uint8_t* data, int32_t size) {
  // velox/common/base/BitUtil.h loadPartialWord: little-endian assemble of <8 bytes.
  uint64_t r = 0;
  for (int32_t i = 0; i < size; ++i) r |= static_cast<uint64_t>(data[i]) << (i * 8);
  return r;
}

inline uint64_t crc32_u64(uint64_t seed, uint64_t v) { return _mm_crc32_u64(seed, v); }

inline uint64_t hash_bytes(uint64_t seed, const char* data, size_t size) {
  auto begin = reinterpret_cast<const uint8_t*>(data);
  const uint64_t kMul = 0x9ddfea08eb382d69ULL;
  if (size < 8) {
    uint64_t word = load_partial_word(begin, static_cast<int32_t>(size));
    uint64_t crc = crc32_u64(seed, word);
    uint64_t crc2 = crc32_u64(seed, word >> 32);
    return crc | (crc2 << 32);
  }
  uint64_t a0 = seed, a1 = seed << 32, a2 = seed >> 16;
  int32_t toGo = static_cast<int32_t>(size);
  const uint8_t* p = begin;
  auto word_at = [&](int i) { uint64_t w; std::memcpy(&w, p + 8 * i, 8); return w; };
  while (toGo >= 24) {
    a0 = crc32_u64(a0, word_at(0));
    a1 = crc32_u64(a1, word_at(1));
    a2 = crc32_u64(a2, word_at(2));
    p += 24;
    toGo -= 24;
  }
  if (toGo > 16) {
    a0 = crc32_u64(a0, word_at(0));
    a1 = crc32_u64(a1, word_at(1));
    a2 = crc32_u64(a2, load_partial_word(p + 16, toGo - 16));
  } else if (toGo > 8) {
    a0 = crc32_u64(a0, word_at(0));
    a1 = crc32_u64(a1, toGo == 16 ? word_at(1) : load_partial_word(p + 8, toGo - 8));
  } else if (toGo > 0) {
    a0 = crc32_u64(a0, toGo == 8 ? word_at(0) : load_partial_word(p, toGo));
  }
  return a0 ^ (a1 * kMul) ^ (a2 * kMul);
}

// StringView hash: velox/type/StringView.h -> bits::hashBytes(1, data, size) via folly hasher
// specialisation (velox/exec/VectorHasher.cpp:80 hashes StringView through folly::hasher<StringView>).
inline uint64_t hash_string(const char* data, size_t size) { return hash_bytes(1, data, size); }

}  // namespace orc
