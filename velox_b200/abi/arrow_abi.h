// Arrow C data interface (https://arrow.apache.org/docs/format/CDataInterface.html): the two structs
// are a stable C ABI, restated here exactly as the reference vendors them in
// velox/vector/arrow/Abi.h. Bridge functions: velox/vector/arrow/Bridge.h:153-173.
#pragma once
#include <cstdint>

#include "vector_abi.h"

extern "C" {
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif
}

namespace facebook::velox {

// velox/vector/arrow/Bridge.h:153 — a vector VIEWING the Arrow buffers (no copy where the layouts
// agree: fixed-width values, validity and boolean bitmaps at offset 0, dictionary indices). The
// caller keeps arrowArray alive while the vector is used. Supported formats (the path's types):
// "+s" struct -> RowVector; "l" BIGINT, "i" INTEGER, "tdD" DATE, "g" DOUBLE, "b" BOOLEAN,
// "u" VARCHAR (copied into StringViews over the Arrow characters); dictionary-encoded children
// (int32 indices) -> DictionaryVector.
VectorPtr importFromArrowAsViewer(const ArrowSchema& arrowSchema, const ArrowArray& arrowArray, memory::MemoryPool* pool);
// velox/vector/arrow/Bridge.h:170 — as above, taking ownership: the inputs are marked released and
// the returned vector calls their release callbacks when destroyed.
VectorPtr importFromArrowAsOwner(ArrowSchema& arrowSchema, ArrowArray& arrowArray, memory::MemoryPool* pool);

}  // namespace facebook::velox
