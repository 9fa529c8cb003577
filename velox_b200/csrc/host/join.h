// Device join table shared between B200HashBuild and B200HashProbe through the HashJoinBridge.
#pragma once
#include <unordered_map>

#include "operators.h"

namespace velox_b200 {

// Key normalisation shared by build and probe: every key column is mapped to a value id
// (v - min + 1; 0 = NULL) and the ids are packed into one 64-bit word (exec/VectorHasher.h:523-585).
struct KeyLayout {
  std::vector<int64_t> mins;
  std::vector<uint64_t> ranges;  // ids per column, including the NULL id
  std::vector<uint64_t> mults;
  uint64_t product = 1;          // size of the packed key space (0 when it overflows 63 bits)
};

struct JoinTableHolder : wave::HashTableHolder {
  vb2_join_table table{};
  std::vector<DeviceBufferPtr> owners;
  KeyLayout layout;
  B200VectorPtr rows;           // concatenated build side (payload gathered from here)
  bool hasDuplicateKeys = false;
  bool hasNullKeys = false;     // a build row had a NULL key (null-aware anti joins ask; exec/HashJoinBridge.h:86)
  int64_t numRows = 0;
  cudaStream_t stream = nullptr;
  // Keyed mode (the reference's kHash for joins, exec/HashTable.cpp:1751-1838): DOUBLE / VARCHAR keys and key
  // tuples that do not pack into one 64-bit word. The build side's distinct key tuples live in a keyed group
  // table (rows = [state | key words | NULL mask]); a row's join key is the slot of its tuple there, and
  // `table` is an array-mode table over those slots.
  bool keyed = false;
  vb2_group_table keyedTable{};
  std::vector<bool> keyIsVarchar;
  std::vector<std::unordered_map<std::string, int32_t>> varcharIds;  // per VARCHAR key: build-side string -> id
};

// Key columns as the keyed-id kernel reads them: VARCHAR dictionary keys become INTEGER dictionary columns over
// a LUT of the build side's string ids (build: ids are assigned; probe: strings the build side never saw map to
// -1, an id no build row has).
std::vector<vb2_column> keyedJoinColumns(const B200Vector& batch, const std::vector<int32_t>& keyColumns, JoinTableHolder& holder, bool build,
                                         std::vector<KeyedLutCache>* cache, std::vector<DeviceBufferPtr>& keep, cudaStream_t stream);

// Packs key columns of `batch` into normalized keys. valid bit = no key column is NULL.
// `probe`: ids outside the layout's ranges get an invalid key (cannot match).
struct NormalizedKeys {
  DeviceBufferPtr keys, valid;
};
NormalizedKeys normalizeKeys(const B200Vector& batch, const std::vector<int32_t>& keyColumns, const KeyLayout& layout,
                             const int32_t* sel, int64_t n, bool nullsInvalid, cudaStream_t stream);
// min/max/non-null count of an integer-like key column (synchronises).
void columnMinMax(const DeviceColumn& col, int64_t rows, cudaStream_t stream, int64_t& lo, int64_t& hi, int64_t& nonNull);
// Concatenates device batches into flat columns (VARCHAR stays dictionary-coded when all batches
// share one dictionary).
B200VectorPtr concatBatches(const std::vector<B200VectorPtr>& batches, memory::MemoryPool* pool, cudaStream_t stream);

}  // namespace velox_b200
