// TEST INFRASTRUCTURE — CPU oracle. Not part of the shipped product path.
//
// C ABI of the CPU restatement of Velox's vectorized operator hot path
// (FilterProject / ExprSet, HashAggregation, HashBuild/HashProbe, VectorHasher,
// HashPartitionFunction). Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// Parity status: the scalar kernels are pinned against the reference's own known-answer
// tests (tests/golden/*.json, transcribed from velox/functions/prestosql/tests/
// ArithmeticTest.cpp:192-249 and ComparisonsTest.cpp:147-260,650-720), the hash mixers
// against the reference's in-tree restatement (velox/experimental/gpu/tests/
// HashTableTest.cu:38-60) and the symbolic relations of velox/exec/tests/
// VectorHasherTest.cpp:166-262. Operator-level results in the reference are pinned only
// through DuckDB at test time (absent here): those are cross-checked against
// numpy/pyarrow re-computations instead — see DESIGN.md "Oracle".
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// TypeKind numbering follows velox/type/Type.h (BOOLEAN 0, INTEGER 3, BIGINT 4, DOUBLE 6,
// VARCHAR 7). DATE is INTEGER days since epoch (velox/type/Type.h:1305).
enum { ORC_BOOLEAN = 0, ORC_INTEGER = 3, ORC_BIGINT = 4, ORC_DOUBLE = 6, ORC_VARCHAR = 7 };
enum { ORC_FLAT = 0, ORC_DICTIONARY = 1, ORC_CONSTANT = 2 };

// One column of a batch. Layout mirrors the vector data contract of SURVEY.md §8 a20:
//   FLAT        values = T[size] (BOOLEAN: bit-packed u64 words, LSB first;
//               VARCHAR: int32 offsets[size+1] into aux), nulls = validity bitmap (1 = not null)
//   DICTIONARY  indices = int32[size], nulls = wrapper validity; values/aux/dict_nulls describe
//               the dict_size base values
//   CONSTANT    values holds one element; nulls (if set) bit 0 tells validity
typedef struct orc_column {
  int32_t type;
  int32_t encoding;
  int64_t size;
  const void* values;
  const uint64_t* nulls;
  const int32_t* indices;
  int64_t dict_size;
  const uint64_t* dict_nulls;
  const void* aux;
} orc_column;

typedef struct orc_table {
  int32_t ncols;
  int32_t reserved;
  int64_t rows;
  const orc_column* cols;
} orc_table;

// Runs a plan (S-expression text, grammar in DESIGN.md) over host tables. Returns an opaque
// result or NULL (message in err). threads = driver count, batch_rows = rows per batch.
void* orc_run_plan(const char* plan, int32_t n_sources, const orc_table* sources, int32_t threads,
                   int32_t batch_rows, char* err, int32_t errlen);
int64_t orc_result_rows(void* r);
int32_t orc_result_cols(void* r);
int32_t orc_result_type(void* r, int32_t col);
// Fixed width: copies values (BOOLEAN as one byte per row) and a byte-per-row null flag (1 = null).
void orc_result_copy(void* r, int32_t col, void* values, uint8_t* nulls);
int64_t orc_result_str_bytes(void* r, int32_t col);
void orc_result_copy_str(void* r, int32_t col, int32_t* offsets, char* chars, uint8_t* nulls);
void orc_result_free(void* r);

// VectorHasher::hash over key columns (mix = hashMix across columns, null -> kNullHash).
int32_t orc_hash_columns(const orc_column* cols, int32_t ncols, int64_t rows, uint64_t* out);
// HashPartitionFunction::partition: hash % num_partitions.
int32_t orc_partition(const orc_column* cols, int32_t ncols, int64_t rows, int32_t num_partitions,
                      uint32_t* out);

// Scalars, for golden-vector tests.
uint64_t orc_twang_mix64(uint64_t v);
uint32_t orc_jenkins_rev_mix32(uint32_t v);
uint64_t orc_hash_mix(uint64_t upper, uint64_t lower);
uint64_t orc_hash_bytes(uint64_t seed, const char* data, int64_t size);
uint64_t orc_hash_f64(double v);
// op: 0 lt, 1 lte, 2 gt, 3 gte, 4 eq, 5 neq
int32_t orc_compare_f64(int32_t op, double a, double b);
// op: 0 plus, 1 minus, 2 multiply, 3 divide, 4 modulus; returns 0 ok, 1 arithmetic error
int32_t orc_checked_i64(int32_t op, int64_t a, int64_t b, int64_t* out);

#ifdef __cplusplus
}
#endif
