/*
 * velox_b200 — kernel-level C ABI (device pointers in, device pointers out).
 *
 * These are the entry points the C++ operator layer (velox_b200/csrc/host) calls; Velox itself
 * never sees them (SURVEY.md §8b "What a C-ABI replacement must export"). Every function
 * takes raw device pointers, row counts and a cudaStream_t (as void*), returns 0 or a VB2_ERR_*
 * code, and never synchronises the stream unless stated. vb2_last_error() gives the message.
 *
 * Each entry cites the reference code whose work it takes over (paths into
 * /root/reference/velox).
 */
#ifndef VELOX_B200_KERNELS_H_
#define VELOX_B200_KERNELS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  VB2_OK = 0,
  VB2_ERR_CUDA = 1,        /* CUDA runtime / NCCL failure -> VeloxRuntimeError */
  VB2_ERR_INVALID = 2,     /* bad argument -> VeloxRuntimeError */
  VB2_ERR_UNSUPPORTED = 3, /* shape not covered by this build -> VeloxRuntimeError (never a CPU fallback) */
  VB2_ERR_USER = 4         /* data error: integer overflow, division by zero, bad cast -> VeloxUserError */
};

/* TypeKind numbering of velox/type/Type.h. DATE is INTEGER days (velox/type/Type.h:1305). */
enum { VB2_BOOLEAN = 0, VB2_INTEGER = 3, VB2_BIGINT = 4, VB2_DOUBLE = 6, VB2_VARCHAR = 7 };
/* VectorEncoding::Simple subset of velox/vector/VectorEncoding.h. */
enum { VB2_FLAT = 0, VB2_DICTIONARY = 1, VB2_CONSTANT = 2 };

/*
 * One column of a batch (mirror of FlatVector / DictionaryVector / ConstantVector buffers,
 * velox/vector/FlatVector.h:604-607, DictionaryVector.h:275-278):
 *   FLAT        values = T[size]; BOOLEAN is bit-packed (u64 words, LSB first); VARCHAR is
 *               int32 offsets[size+1] into aux (chars). nulls = validity bitmap, 1 = not null
 *               (velox/common/base/Nulls.h:26-27), or NULL when the column has no nulls.
 *   DICTIONARY  indices = int32[size]; nulls = validity of the wrapper; values/aux/dict_nulls
 *               describe the dict_size base values.
 *   CONSTANT    values holds one element; nulls (if set) bit 0 gives validity.
 * At this level every pointer is a device pointer.
 */
typedef struct vb2_column {
  int32_t type;
  int32_t encoding;
  int64_t size;
  const void* values;
  const uint64_t* nulls;
  const int32_t* indices;
  int64_t dict_size;
  const uint64_t* dict_nulls;
  const void* aux;
} vb2_column;

const char* vb2_last_error(void);
int vb2k_device_sm_count(void);
/* Kernels this library has launched in this process so far (every launch site counts itself). */
int64_t vb2k_kernel_launches(void);

/* ------------------------------------------------------------------------------------------
 * Key hashing and partitioning.
 * Replaces VectorHasher::hash (velox/exec/VectorHasher.cpp:567-594, hashValues :87-126) and
 * HashPartitionFunction::partition (velox/exec/HashPartitionFunction.cpp:75-118): bit-exact
 * folly::hasher<T> per key column, bits::hashMix across columns, kNullHash for nulls, then
 * hash % num_partitions.
 * ------------------------------------------------------------------------------------------ */
int vb2k_hash_columns(const vb2_column* cols, int32_t ncols, int64_t rows, uint64_t* hashes, void* stream);
int vb2k_partition_ids(const uint64_t* hashes, int64_t rows, int32_t num_partitions, uint32_t* ids, void* stream);
/* counts[p] = rows with id p; offsets computed on device; row_order = rows grouped by partition
 * (stable within a partition). All outputs device memory: counts int64[P], row_order int32[rows]. */
int vb2k_partition_scatter_order(const uint32_t* ids, int64_t rows, int32_t num_partitions, int64_t* counts,
                                 int32_t* row_order, void* stream);
/* Same for ONE flat NULL-free BIGINT / INTEGER key column, with hash and partition id computed on the
 * fly (identical to vb2k_hash_columns + vb2k_partition_ids on that column): no hash / id arrays. */
int vb2k_partition_order_key(const void* key_values, int32_t key_is64, int64_t rows, int32_t num_partitions, int64_t* counts, int32_t* row_order,
                             void* stream);
/* Sync-free variant for shuffles whose sizes are planned ahead (from statistics of an earlier run
 * of the same plan): rows of a non-null BIGINT key column and up to 4 fixed-width payload columns
 * go straight into fixed-capacity per-destination segments — row r to partition
 * p = twang_mix64(key) % P (= VectorHasher hash % P, as above) at seg_keys[p * segcap + rank],
 * rank = its stable position among the rows of p. Only rows r < min(rows, *rows_dev) are read when
 * rows_dev (device) is given, so a producer's device-side row count never visits the host.
 * Segment tails hold VB2_SENTINEL_KEY, a key value the shuffled tables must not contain: join
 * probes / builds downstream treat it as out of range (a miss). counts[p] (device) = rows of p;
 * *overflow (device) becomes 1 if a partition exceeds segcap (its excess rows are dropped and the
 * caller must re-plan). */
#define VB2_SENTINEL_KEY ((int64_t)0x8080808080808080ull)
int vb2k_partition_segments(const int64_t* keys, const void* const* cols, const int32_t* col_elem_bytes, int32_t ncols, int64_t rows,
                            const int64_t* rows_dev, int32_t num_partitions, int64_t segcap, int64_t* seg_keys, void* const* seg_cols,
                            int64_t* counts, int32_t* overflow, void* stream);
/* *flag (device) <- 1 if any key other than VB2_SENTINEL_KEY lies outside [lo, hi]. */
int vb2k_key_range_check(const int64_t* keys, int64_t n, int64_t lo, int64_t hi, int32_t* flag, void* stream);
/* out[i] = in[order[i]] for fixed-width columns (elem_bytes 1, 4 or 8). */
int vb2k_gather(const void* in, const int32_t* order, int64_t n, int32_t elem_bytes, void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Expression VM: evaluates a compiled ExprSet (filter + projections) in one kernel.
 * Replaces ExprSet::eval / Expr::eval / evalFlatNoNulls / evalAll / applyFunction
 * (velox/expression/Expr.cpp:2339,848,801,1513,1787), ConjunctExpr / SwitchExpr special forms
 * (expression/ConjunctExpr.cpp:93, SwitchExpr.cpp:71), the scalar kernels of
 * functions/prestosql/{Arithmetic,Comparisons}.h, CastExpr, LIKE (functions/lib/Re2Functions.cpp:710)
 * and processFilterResults (velox/exec/OperatorUtils.cpp:231-321).
 * ------------------------------------------------------------------------------------------ */
enum vb2_opcode {
  VB2_OP_LOAD = 1,    /* dst <- column a (decodes flat/dictionary/constant, nulls) */
  VB2_OP_CONST = 2,   /* dst <- constant a */
  VB2_OP_ADD = 3, VB2_OP_SUB = 4, VB2_OP_MUL = 5, VB2_OP_DIV = 6, VB2_OP_MOD = 7, VB2_OP_NEG = 8,
  VB2_OP_LT = 9, VB2_OP_LTE = 10, VB2_OP_GT = 11, VB2_OP_GTE = 12, VB2_OP_EQ = 13, VB2_OP_NEQ = 14,
  VB2_OP_BETWEEN = 15, /* a between b and c */
  VB2_OP_AND = 16, VB2_OP_OR = 17, VB2_OP_NOT = 18, VB2_OP_IS_NULL = 19,
  VB2_OP_SELECT = 20,  /* dst <- (a is true and not null) ? b : c  — CASE/IF */
  VB2_OP_CAST = 21,    /* dst <- cast(a); type = target, b = source type */
  VB2_OP_LIKE = 22,    /* dst <- column a LIKE constant b (VARCHAR column, pattern constant) */
  VB2_OP_STRCMP = 23,  /* dst <- column a <cmp c> constant b, c = vb2 compare code (0 lt .. 5 neq) */
  VB2_OP_NULL = 24,    /* dst <- NULL of `type` */
  VB2_OP_CALL = 25     /* dst <- registered device function (vb2k_register_device_function) over registers a, b, c
                          (unused ones -1); type = result type | function id << 8. NULL in -> NULL out. */
};
#define VB2_CALL_TYPE(t) ((t) & 0xff)
#define VB2_CALL_FN(t) ((t) >> 8)

typedef struct vb2_instr {
  int32_t op;
  int32_t type; /* operand type for arithmetic/compare, result type otherwise */
  int32_t dst;
  int32_t a, b, c;
} vb2_instr;

typedef struct vb2_const {
  int32_t type;
  int32_t is_null;
  int64_t i; /* BOOLEAN / INTEGER / BIGINT payload */
  double d;  /* DOUBLE payload */
  const char* str; /* VARCHAR payload (device pointer) */
  int32_t len;
  int32_t pad;
} vb2_const;

typedef struct vb2_program {
  const vb2_instr* instrs; /* host pointer; copied to the kernel as an argument block */
  int32_t n_instrs;        /* <= 256 */
  int32_t n_filter_instrs; /* instrs [0, n_filter_instrs) compute the filter register */
  int32_t filter_reg;      /* -1 when there is no filter */
  int32_t n_regs;          /* <= 64 */
  const vb2_const* consts; /* host pointer */
  int32_t n_consts;        /* <= 32 */
  int32_t pad;
} vb2_program;

typedef struct vb2_output {
  int32_t reg;
  int32_t type;
  void* values;    /* T[n_out]; BOOLEAN written one byte per row */
  uint64_t* nulls; /* validity bitmap of n_out bits, always written */
} vb2_output;

/* Expression JIT (expr_jit.cu): programs are compiled to straight-line sm_100a code with NVRTC and
 * cached per (program, column layout); vb2k_eval_filter / vb2k_eval_project use it when available
 * and fall back to the interpreter kernels otherwise. On by default (VB2_EXPR_JIT=0 disables). */
void vb2k_set_expression_jit(int32_t enabled);
/* Generates and compiles (does not launch; no GPU needed) the kernel of a program for the given
 * column layout: 1 = a JIT kernel is available, 0 = the interpreter would run. source_out receives
 * the generated source (or the compiler log on failure). */
int32_t vb2k_expression_jit_compiles(const vb2_program* prog, const vb2_column* cols, int32_t ncols, int32_t filter, const vb2_output* outs,
                                     int32_t nouts, char* source_out, int32_t source_len);

/* ------------------------------------------------------------------------------------------
 * High-cardinality GROUP BY in shared memory (slice_agg.cu). The buffered input (chunks of normalized
 * or raw BIGINT keys + up to 3 eight-byte payload columns) is radix-partitioned twice by the top bits
 * of a Fibonacci hash of the key into up to 65536 slices; every slice is aggregated in one CTA's shared-memory
 * hash table and its groups are appended to rows_out as group rows in the vb2_group_table layout
 * (word 0 = normalized key, accumulator words at ops[].word, the rest from row_init).
 * Replaces HashTable::groupProbe + the accumulator scatter (velox/exec/HashTable.cpp:470-519,
 * functions/lib/aggregates/SimpleNumericAggregate.h:94-150) when the group table would not fit L2.
 *   vb2k_slice_agg_partition  level-1 histogram (+ HyperLogLog sketch, copied to hll_host[4096]; the
 *                             call synchronises when hll_host is given) and scatter of every chunk
 *   vb2k_slice_agg_finish     level 2 (sized from distinct_estimate), aggregation. num_groups / reserved_rows /
 *                             error_flag / overflow_slices are zeroed device words: groups written, rows handed out,
 *                             1 = SUM(BIGINT) overflow or 100 = rows_out full, slices whose keys did
 *                             not fit their table (then the result is incomplete: use the table path).
 * VB2_ERR_UNSUPPORTED: more distinct keys than 65536 slices can hold.
 * ------------------------------------------------------------------------------------------ */
#define VB2_SLICE_MAX_COLS 3
#define VB2_SLICE_MAX_OPS 8
typedef struct vb2_slice_chunk {
  const uint64_t* norm_keys; /* normalized keys, or NULL with raw_keys */
  const int64_t* raw_keys;   /* flat NULL-free BIGINT key column: normalized key = raw - key_min + 1 */
  int64_t key_min;
  const void* cols[VB2_SLICE_MAX_COLS]; /* 8-byte payload columns (BIGINT / DOUBLE), flat, NULL-free */
  int64_t rows;
} vb2_slice_chunk;
typedef struct vb2_slice_op {
  int32_t kind; /* vb2_agg_kind */
  int32_t col;  /* payload column, -1 for COUNT */
  int32_t word; /* word of the group row that accumulates */
} vb2_slice_op;
int32_t vb2k_slice_agg_hll_registers(void);
size_t vb2k_slice_agg_workspace(int64_t total_rows, int32_t ncols);
int vb2k_slice_agg_partition(const vb2_slice_chunk* chunks, int32_t nchunks, int32_t ncols, int64_t total_rows, void* workspace, size_t workspace_bytes,
                             int32_t* hll_host, void* stream);
/* rows_out must hold vb2k_slice_agg_output_rows(min(distinct_estimate, total_rows)) rows, pre-filled with row_init and
 * VB2_EMPTY_KEY in word 0 (vb2k_group_table_init): the blocks reserve output rows in chunks
 * (reserved_rows, a zeroed device word) and leave the unused tail of a chunk EMPTY — the result reads
 * like a hash-mode group table with num_groups occupied rows. */
int64_t vb2k_slice_agg_output_rows(int64_t distinct_estimate);
int vb2k_slice_agg_finish(int64_t total_rows, int32_t ncols, int64_t distinct_estimate, const vb2_slice_op* ops, int32_t nops, int32_t row_words,
                          const uint64_t* row_init, uint64_t* rows_out, int64_t rows_capacity, int64_t* num_groups, int64_t* reserved_rows,
                          int32_t* error_flag, int32_t* overflow_slices, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * ORDER BY: stable multi-key sort producing the row order. Replaces the sort of exec::OrderBy /
 * SortBuffer (velox/exec/OrderBy.cpp, SortBuffer.cpp) and the ordering of exec::TopN
 * (velox/exec/TopN.cpp); semantics of core::SortOrder (velox/core/PlanNode.h:64-95): NULLs first or
 * last independent of direction, NaN largest, -0 == +0. Keys are flat columns (VARCHAR keys are
 * passed as INTEGER rank codes of their sorted dictionary). order[i] = input row at output position i.
 * ------------------------------------------------------------------------------------------ */
#define VB2_SORT_MAX_KEYS 8
typedef struct vb2_sort_key {
  const void* values;     /* INTEGER int32 / BIGINT int64 / DOUBLE double / BOOLEAN bit-packed */
  const uint64_t* nulls;  /* validity bitmap or NULL */
  int32_t type;
  int32_t ascending;
  int32_t nulls_first;
  int32_t significant_bits; /* > 0: INTEGER / BIGINT values are known to lie in [0, 2^bits): fewer radix passes */
} vb2_sort_key;
size_t vb2k_sort_order_workspace(int64_t n, int32_t nkeys);
int vb2k_sort_order(const vb2_sort_key* keys, int32_t nkeys, int64_t n, int32_t* order, void* workspace, size_t workspace_bytes, void* stream);

/* Registry of user-supplied scalar device functions — the device half of exec::registerVectorFunction
 * (velox/expression/VectorFunction.h:241): `cuda_source` is CUDA C++ text defining
 *   __device__ RET entry(ARG0 [, ARG1 [, ARG2]])      with BIGINT = long long, INTEGER = int, DOUBLE = double, BOOLEAN = bool
 * which the expression JIT splices into every kernel whose program calls it (VB2_OP_CALL), so a
 * registered function fuses with the rest of the expression tree like a built-in. Default NULL
 * behaviour: the function is not called on rows with a NULL argument. Returns the function id (>= 0)
 * or a negative VB2_ERR_*. Programs that call registered functions need the JIT (NVRTC): without it
 * vb2k_eval_* returns VB2_ERR_UNSUPPORTED — there is no CPU fallback. */
int32_t vb2k_register_device_function(const char* entry, const char* cuda_source, int32_t ret_type, const int32_t* arg_types, int32_t nargs);
int32_t vb2k_device_function_count(void);

/* Pass 1: evaluates the filter over `rows` input rows. Writes the selection bitmap (1 = row kept:
 * predicate true and not null) and per-block popcounts for the compaction that follows. */
int vb2k_eval_filter(const vb2_program* prog, const vb2_column* cols, int32_t ncols, int64_t rows,
                     uint64_t* sel_bits, int32_t* error_flag, void* stream);
/* Expands a selection bitmap into ascending row numbers (`selectedIndices` of
 * exec/OperatorUtils.cpp:291). count_out is a device int64. */
int vb2k_bits_to_indices(const uint64_t* sel_bits, int64_t rows, int32_t* indices, int64_t* count_out,
                         void* workspace, size_t workspace_bytes, void* stream);
size_t vb2k_bits_to_indices_workspace(int64_t rows);
/* Pass 2: evaluates the projections for rows sel[0..n) (sel == NULL: rows 0..n) and writes them
 * densely. error_flag receives the first VB2 user-error code hit by a live row. */
int vb2k_eval_project(const vb2_program* prog, const vb2_column* cols, int32_t ncols, const int32_t* sel,
                      int64_t n, const vb2_output* outs, int32_t nouts, int32_t* error_flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused scan -> filter -> project -> aggregate pipelines (ahead-of-time specialised).
 * One pass over null-free flat input columns: the expression DAG is a C++ expression template,
 * the registry key is the canonical text the same template prints, so code and key cannot drift.
 * Replaces, for one batch, FilterProject::getOutput (velox/exec/FilterProject.cpp:200-259) +
 * GroupingSet::addInputForActiveRows in array mode (velox/exec/GroupingSet.cpp:288,
 * HashTable::arrayGroupProbe exec/HashTable.cpp:560) + SUM/AVG/COUNT addRawInput
 * (functions/lib/aggregates/SimpleNumericAggregate.h:94-150).
 * ------------------------------------------------------------------------------------------ */
#define VB2_FUSED_MAX_COLS 8
#define VB2_FUSED_MAX_PARAMS 12
#define VB2_FUSED_MAX_KEYS 2

typedef struct vb2_fused_args {
  const void* cols[VB2_FUSED_MAX_COLS]; /* expression inputs, renumbered in first-use order */
  double pf[VB2_FUSED_MAX_PARAMS];      /* DOUBLE constants in first-use order */
  int64_t pl[VB2_FUSED_MAX_PARAMS];     /* BIGINT constants */
  int32_t pi[VB2_FUSED_MAX_PARAMS];     /* INTEGER/DATE constants */
  int64_t rows;
  int32_t nkeys;                        /* 0 = global aggregation */
  int32_t ngroups;                      /* size of the group-id space (product of key ranges) */
  const void* key[VB2_FUSED_MAX_KEYS];  /* int32 (dictionary indices / INTEGER) or int64 key values */
  int32_t key_is64[VB2_FUSED_MAX_KEYS];
  int32_t key_mult[VB2_FUSED_MAX_KEYS]; /* gid = sum_k id_k * key_mult[k] */
  int64_t key_min[VB2_FUSED_MAX_KEYS];  /* id_k = lut ? lut[v - min] : v - min */
  const int32_t* key_lut[VB2_FUSED_MAX_KEYS];
  /* Optional inner-join probe fused between filter and projection (Q14 shape): the probe key
   * column named by the pipeline signature (";J:l<col>") is looked up in a dense array table of
   * one byte per key slot (slot = key - join_min): 0 = no build row (row dropped), 1 = match,
   * 2 = match and the build-side predicate (e.g. p_type LIKE 'PROMO%') holds. Build keys must be
   * unique; vb2k_join_slot_flags prepares the table from an array-mode vb2_join_table. */
  const uint8_t* join_slot_flags;
  int64_t join_min, join_range;
} vb2_fused_args;

/* head: int32[range] build row + 1 (0 = empty); codes: int32[build rows] dictionary code of the
 * payload column or NULL (code = row); flag: uint8[codes] predicate per code or NULL (all false). */
int vb2k_join_slot_flags(const int32_t* head, const int32_t* codes, const uint8_t* flag, int64_t range, uint8_t* out,
                         void* stream);
int vb2k_fused_find(const char* signature);  /* kernel id >= 0, or -1 when no specialisation matches */
int vb2k_fused_count(void);
const char* vb2k_fused_signature(int32_t id);
int32_t vb2k_fused_nproj(int32_t id);
size_t vb2k_fused_workspace_bytes(int32_t id, int32_t ngroups);
/* Adds this batch into the persistent accumulators: sums is double[ngroups][nproj] (one running
 * sum per projection and group, accumulated in a fixed, run-to-run deterministic order),
 * counts is int64[ngroups] (rows that passed the filter per group). */
int vb2k_fused_scan_agg(int32_t id, const vb2_fused_args* args, double* sums, int64_t* counts, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Pipelines without an ahead-of-time specialisation are instantiated from the same expression
 * templates with NVRTC when vb2k_fused_find misses (fused_jit.cu; VB2_PIPELINE_JIT=0 or
 * vb2k_set_pipeline_jit(0) disables it — the generic kernels run then). */
void vb2k_set_pipeline_jit(int32_t enabled);
/* Compiles (no GPU needed, nothing is loaded or launched) one kernel of the pipeline `signature`
 * describes: kind 0 TMA scan-aggregate, 1 direct scan-aggregate, 2 filter bitmap, 3 gather-aggregate;
 * max_groups 1 / 4 / 8 register accumulators or 0 shared-memory accumulators. 1 = compiled. */
int32_t vb2k_pipeline_jit_compiles(const char* signature, int32_t kind, int32_t max_groups, int32_t key64, char* log_out, int32_t log_len);
/* Late materialisation for selective filters (FilterProject evaluates the filter first and the
 * projections only on surviving rows, velox/exec/FilterProject.cpp:200-259):
 *   vb2k_fused_has_filter    1 when pipeline `id` has a filter that can run on its own;
 *   vb2k_fused_filter_bits   streams only the filter's columns and writes the selection bitmap of
 *                            args->rows rows (tile_stride 1), or visits every tile_stride-th 1024-row
 *                            tile and only counts (sel_bits NULL): counters = device int64[2]
 *                            {rows kept, rows evaluated}, accumulated;
 *   vb2k_fused_gather_agg    join probe + projections + aggregation over the selected row numbers
 *                            sel[0 .. *nsel_dev) (ascending), same accumulators as vb2k_fused_scan_agg;
 *                            at most 4 groups. nsel_hint sizes the grid.
 * A signature "F:<filter>;P:" (no projections) names the filter alone. */
int32_t vb2k_fused_has_filter(int32_t id);
int vb2k_fused_filter_bits(int32_t id, const vb2_fused_args* args, int32_t tile_stride, uint64_t* sel_bits, int64_t* counters, void* stream);
int vb2k_fused_gather_agg(int32_t id, const vb2_fused_args* args, const int32_t* sel, const int64_t* nsel_dev, int64_t nsel_hint, double* sums,
                          int64_t* counts, void* workspace, size_t workspace_bytes, void* stream);
/* Scan -> filter -> project -> compact pipelines (signatures "F:...;C:..."): writes the
 * projections of the surviving rows densely into outs[p] (element width
 * vb2k_fused_output_width). count: device int64 running total (rows written so far);
 * *error_flag = 100 when capacity is exceeded. Feeds the hash-partitioned exchange. */
int vb2k_fused_scan_compact(int32_t id, const vb2_fused_args* args, void* const* outs, int32_t nouts, int64_t capacity,
                            int64_t* count, int32_t* error_flag, void* stream);
int32_t vb2k_fused_output_width(int32_t id, int32_t out);

/* ------------------------------------------------------------------------------------------
 * Generic aggregation over materialised columns.
 * Array mode (small group-id space): replaces HashTable::arrayGroupProbe + Aggregate::addRawInput /
 * addIntermediateResults scatter loops. Hash mode: open-addressing table over 64-bit normalized
 * keys (exec/HashTable.cpp:470-523 groupProbe / groupNormalizedKeyProbe, insertEntry :337).
 * ------------------------------------------------------------------------------------------ */
enum vb2_agg_kind { VB2_AGG_SUM_F64 = 1, VB2_AGG_SUM_I64 = 2, VB2_AGG_COUNT = 3, VB2_AGG_MIN_F64 = 4,
                    VB2_AGG_MAX_F64 = 5, VB2_AGG_MIN_I64 = 6, VB2_AGG_MAX_I64 = 7, VB2_AGG_COUNT_MERGE = 8 };

/* Group table: row-wise group storage, the role of exec::RowContainer under exec::HashTable for
 * GROUP BY (velox/exec/RowContainer.h, velox/exec/HashTable.cpp:706-772 groupProbe). `capacity` rows
 * of `row_words` 8-byte words; word 0 is the occupancy word — the 64-bit normalized key in hash
 * mode (VB2_EMPTY_KEY = free; open addressing from the home slot twang_mix64(key) >> (64 - log2(capacity)),
 * linear probing, capacity a power of two), "rows seen" in array mode (slot = normalized key, 0 = free; capacity = key space) — the
 * remaining words are accumulators / non-null counters. Rows of more than two words should be
 * padded to a multiple of four words (whole 32-byte sectors). */
#define VB2_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define VB2_MAX_ROW_WORDS 64
typedef struct vb2_group_table {
  uint64_t* rows;
  int64_t capacity;
  int32_t row_words;
  int32_t hash_mode;   /* 0 array, 1 normalized-key hash, 2 keyed hash (VB2_GROUP_KEYED, below) */
} vb2_group_table;
/* Keyed hash mode — the reference's kHash mode (velox/exec/HashTable.cpp:1751-1838: keys that do not
 * fit one normalized 64-bit word, DOUBLE keys): the group row stores its key columns verbatim.
 *   word 0            state: VB2_EMPTY_KEY = free, else (hash63 << 1) | ready; a row is claimed by CAS
 *                     with ready = 0, its key words are written, then ready is published
 *   words 1 .. K      one 64-bit word per key column (integers sign-extended, DOUBLE as canonical
 *                     bits: one NaN, +0 for -0 — the equality of velox/type/FloatingPointUtil.h)
 *   word K + 1        NULL mask (bit k = key k is NULL; NULL keys form a group, GroupingSet.cpp:448-455)
 *   words K + 2 ..    accumulators / non-null counters
 * Probing compares the 63-bit hash first and the key words on a hash match (ProbeState::fullProbe,
 * exec/HashTable.cpp:138, compares a 7-bit tag, then the row's keys). */
#define VB2_GROUP_KEYED 2
#define VB2_KEYED_MAX_KEYS 4
typedef struct vb2_group_row_init { uint64_t words[VB2_MAX_ROW_WORDS]; } vb2_group_row_init;

typedef struct vb2_agg_update {
  int32_t kind;
  int32_t input_type;       /* VB2_DOUBLE / VB2_BIGINT / VB2_INTEGER (converted as the reference does) */
  const void* input;        /* input values (NULL for COUNT(*)): row i reads input[indices ? indices[i] : i] */
  const uint64_t* nulls;    /* validity bitmap over rows or NULL */
  const uint64_t* mask;     /* optional aggregate mask bitmap (exec/AggregationMasks.cpp), 1 = use row */
  const int32_t* indices;   /* optional: the input is a dictionary wrap (FilterProject output) over `input` */
  const uint64_t* base_nulls; /* validity of the wrapped values (indexed like `input`) or NULL */
  int32_t acc_word;         /* word of the accumulator (double or int64) inside the group row, >= 1 */
  int32_t nonnull_word;     /* word counting non-null inputs (drives NULL results and AVG counts), -1 = not tracked */
} vb2_agg_update;

/* Fills every row with row_init[0 .. row_words) (host array: VB2_EMPTY_KEY or 0, then accumulator identities). */
int vb2k_group_table_init(const vb2_group_table* t, const uint64_t* row_init, void* stream);
/* One batch of GroupingSet::addInput (velox/exec/GroupingSet.cpp:236-358): finds or inserts each
 * row's group (row_keys from vb2k_normalize_keys; NULL = global aggregation, every row in row 0;
 * rows whose row_valid bit is clear are skipped) and applies every aggregate update to that row in
 * the same kernel. Array-mode tables of <= 8 rows take register-accumulator kernels (one launch
 * per aggregate) because same-address atomics serialise. num_groups (device, optional) is
 * incremented per inserted key; SUM(BIGINT) overflow sets *error_flag = 1, a full table 100. */
int vb2k_group_update(const vb2_group_table* t, const uint64_t* row_keys, const uint64_t* row_valid, int64_t n,
                      const vb2_agg_update* aggs, int32_t naggs, int64_t* num_groups, int32_t* error_flag, void* stream);
/* vb2k_group_update for a keyed table: the key columns of the batch (BOOLEAN / INTEGER / BIGINT /
 * DOUBLE, any encoding, NULLs allowed) are read in place; nkeys must equal the table's key count
 * (rows hold nkeys + 2 leading words). A full table sets *error_flag = 100. */
int vb2k_group_update_keyed(const vb2_group_table* t, const vb2_column* keys, int32_t nkeys, int64_t n, const vb2_agg_update* aggs, int32_t naggs,
                            int64_t* num_groups, int32_t* error_flag, void* stream);
/* Moves the listed groups of a keyed table into another (bigger) keyed table. */
int vb2k_group_move_keyed(const vb2_group_table* from, const int32_t* slots, int64_t n, int32_t nkeys, const vb2_group_table* to, int64_t* num_groups,
                          int32_t* error_flag, void* stream);
/* Moves the listed groups of an array / normalized-key table into a keyed table: key k of a group is
 * decoded from the normalized key (value = id - 1 + mins[k], id = (key / mults[k]) % ranges[k], id 0 =
 * NULL when null_reserved[k]); accumulator word w of the source lands at w + word_shift. */
int vb2k_group_move_to_keyed(const vb2_group_table* from, const int32_t* slots, int64_t n, int32_t nkeys, const int64_t* mins, const uint64_t* mults,
                             const uint64_t* ranges, const int32_t* null_reserved, int32_t word_shift, const vb2_group_table* to,
                             int64_t* num_groups, int32_t* error_flag, void* stream);
/* Value ids of key tuples through a keyed table — the join side of the reference's kHash mode
 * (velox/exec/HashTable.cpp:1751-1838, HashTable::insertForJoin :1518 / joinProbe :610 compare the stored
 * keys of a row): ids[r] = slot of row r's key tuple in the keyed table, valid bit r = the row has an id.
 * insert != 0 (build side): absent tuples are inserted; insert == 0 (probe side): absent tuples clear
 * the valid bit. Rows with a NULL key column never get an id (NULL join keys never match,
 * exec/HashBuild.cpp:475-479). The ids are dense enough (slots of a table of <= 2 x tuples + 16 rows,
 * rounded to a power of two) to address an array-mode vb2_join_table of the same capacity directly.
 * valid: u64 words covering n bits. A full table sets *error_flag = 100. */
int vb2k_keyed_key_ids(const vb2_group_table* t, const vb2_column* keys, int32_t nkeys, int64_t n, int32_t insert, uint64_t* ids, uint64_t* valid,
                       int64_t* num_groups, int32_t* error_flag, void* stream);
/* Radix partitioning in front of a high-cardinality aggregation (radix_partition.cu): a hash-mode
 * group table places key k at slot twang_mix64(k) >> (64 - log2(capacity)), so rows ordered by the
 * top 8 bits of that hash walk the table slice by slice (1/256 of it at a time, L2 resident).
 * Keys: norm_keys (from vb2k_normalize_keys), or NULL and ONE flat NULL-free integer column
 * (key_values, key_is64, key_min: normalized key = v - key_min + 1).
 *   vb2k_radix_histogram  per-chunk partition histograms into `workspace` + a HyperLogLog sketch of the
 *                         keys whose hash ends in 000 (hll_out: device int32[vb2k_radix_hll_registers()];
 *                         distinct keys of the batch ~ 8 x the sketch's estimate) — sizes the table
 *                         before any row is inserted (HashTable::checkSize, exec/HashTable.cpp:772, grows by rehashing instead)
 *   vb2k_radix_scatter    keys_out / cols_out = keys and up to 4 payload columns (4 or 8 bytes wide) in
 *                         partition order; part_start_out: device int64[257] */
size_t vb2k_radix_workspace_bytes(int64_t rows);
int32_t vb2k_radix_hll_registers(void);
int vb2k_radix_histogram(const uint64_t* norm_keys, const void* key_values, int32_t key_is64, int64_t key_min, int64_t rows, void* workspace,
                         size_t workspace_bytes, int32_t* hll_out, void* stream);
int vb2k_radix_scatter(const uint64_t* norm_keys, const void* key_values, int32_t key_is64, int64_t key_min, int64_t rows, void* workspace,
                       size_t workspace_bytes, const void* const* cols, void* const* cols_out, const int32_t* col_bytes, int32_t ncols,
                       uint64_t* keys_out, int64_t* part_start_out, void* stream);
/* vb2k_group_update over rows in radix-partition order (row_keys from vb2k_radix_scatter, part_start =
 * its int64[nparts + 1] output, inputs of the updates reordered the same way): the grid folds partition
 * p into table slice p while slice p + 1 is prefetched into L2, in lock step (cooperative launch;
 * barrier_word: device uint32 scratch). Hash-mode tables of >= 65536 rows. */
int vb2k_group_update_partitioned(const vb2_group_table* t, const uint64_t* row_keys, const int64_t* part_start, int32_t nparts, int64_t n,
                                  const vb2_agg_update* aggs, int32_t naggs, int64_t* num_groups, int32_t* error_flag, uint32_t* barrier_word,
                                  void* stream);
/* Compacts occupied rows: slot_list int32[<=capacity] ascending, count device int64. */
int vb2k_group_occupied(const vb2_group_table* t, int32_t* slot_list, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);
size_t vb2k_group_occupied_workspace(int64_t capacity);
/* word <- value in every occupied row. */
int vb2k_group_set_word(const vb2_group_table* t, int32_t word, uint64_t value, void* stream);
/* Re-encodes the keys of the listed slots for a new layout after value ranges grew, and moves the
 * groups into a new table (the rehash of HashTable::checkSize / decideHashMode,
 * velox/exec/HashTable.cpp:772,1751). */
int vb2k_group_rekey(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t ncols, const int64_t* old_mins,
                     const uint64_t* old_mults, const uint64_t* old_ranges, const int32_t* old_null_reserved, const int64_t* new_mins,
                     const uint64_t* new_mults, uint64_t* keys_out, void* stream);
int vb2k_group_move(const vb2_group_table* from, const int32_t* slots, const uint64_t* new_keys, int64_t n, const vb2_group_table* to,
                    int64_t* num_groups, int32_t* error_flag, void* stream);
/* Inverse of vb2k_normalize_keys for one key column over the listed slots: value = id - 1 + min
 * with id = (key / mult) % range; id 0 is NULL when null_reserved (a layout for a column that
 * never held NULLs uses min + 1 and keeps id 0 for the smallest value). values: T[n] (BOOLEAN one
 * byte per row). */
int vb2k_group_keys(const vb2_group_table* t, const int32_t* slots, int64_t n, int64_t min, uint64_t mult, uint64_t range,
                    int32_t null_reserved, int32_t type, void* values, uint64_t* valid, void* stream);
/* Aggregate result extraction (Aggregate::extractValues, velox/exec/Aggregate.h:281-302) over the
 * listed slots: one word as a dense 8-byte column; validity bit = count word > 0; avg = sum / count. */
int vb2k_group_gather(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t word, void* out, void* stream);
int vb2k_group_valid(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t count_word, uint64_t* valid, void* stream);
int vb2k_group_avg(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t sum_word, int32_t count_word, double* out, void* stream);
/* Every output column of an aggregation in ONE launch (Aggregate::extractValues / extractAccumulators,
 * velox/exec/Aggregate.h:281-302, and the key columns RowContainer::extractColumn returns,
 * velox/exec/GroupingSet.cpp:843 extractGroups). Row i of the output is table row slots[i].
 *   slots != NULL  the n listed slots (any table size);
 *   slots == NULL  small tables (capacity <= VB2_EXTRACT_SMALL_CAPACITY): the kernel itself lists the
 *                  occupied rows in ascending slot order into scratch_slots (int32[capacity]) and
 *                  writes their count to header[0] — no host round trip before the extraction.
 * header (device int64[2], optional): [0] = rows written, [1] = *error_flag as seen by the kernel.
 * Column kinds: KEY decodes one key column from the normalized key (value = id - 1 + min,
 * id = (key / mult) % range, id 0 = NULL when null_reserved; BOOLEAN values bit-packed), WORD copies an
 * 8-byte accumulator word, WORD_I32 narrows it to int32, AVG writes sum / count
 * (functions/lib/aggregates/AverageAggregateBase.h:86-107). count_word >= 0: validity bit = count
 * word > 0. valid bitmaps (optional) are written as whole 64-bit words. */
#define VB2_EXTRACT_SMALL_CAPACITY 16384
#define VB2_EXTRACT_MAX_COLS 40
/* KEYWORD (keyed tables): key column `word` - 1 of the row, typed by `type` (BOOLEAN / INTEGER / BIGINT /
 * DOUBLE), NULL when bit `null_reserved` (reused as the key's bit index) of the row's NULL-mask word
 * `count_word` is set. */
enum vb2_extract_kind { VB2_EXTRACT_KEY = 1, VB2_EXTRACT_WORD = 2, VB2_EXTRACT_WORD_I32 = 3, VB2_EXTRACT_AVG = 4, VB2_EXTRACT_KEYWORD = 5 };
typedef struct vb2_extract_col {
  int32_t kind;
  int32_t type;        /* KEY: VB2_INTEGER / VB2_BIGINT / VB2_BOOLEAN of the written values */
  int32_t word;        /* WORD / WORD_I32 / AVG: accumulator word */
  int32_t count_word;  /* validity source (and AVG's divisor), -1 = always valid */
  int64_t min;         /* KEY decode */
  uint64_t mult, range;
  int32_t null_reserved;
  int32_t pad;
  void* values;
  uint64_t* valid;
} vb2_extract_col;
int vb2k_group_extract(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t* scratch_slots, const vb2_extract_col* cols,
                       int32_t ncols, int64_t* header, const int32_t* error_flag, void* stream);
/* Adds the per-group partials of a fused scan (sums[g * nproj + p] doubles, counts[g]) into rows
 * 0 .. ngroups of an array-mode table: target i adds sums[.., target_projs[i]] to word
 * target_words[i], or counts[g] when target_projs[i] < 0. Groups with counts[g] == 0 are untouched. */
int vb2k_group_merge_partials(const vb2_group_table* t, const double* sums, const int64_t* counts, int32_t ngroups, int32_t nproj,
                              const int32_t* target_words, const int32_t* target_projs, int32_t ntargets, void* stream);
/* Packs up to 4 key columns into one 64-bit normalized key per row:
 * key = sum_k id_k * mult_k with id_k = v_k - min_k + 1 and id 0 reserved for NULL (VectorHasher
 * value ids, velox/exec/VectorHasher.h:523-585). Columns may be flat/dictionary/constant of
 * BOOLEAN/INTEGER/BIGINT type. valid_out (optional bitmap): cleared for rows with a NULL key when
 * nulls_invalid (joins never match NULL keys, exec/HashBuild.cpp:475-479) and for ids outside
 * [1, ranges[k]) when ranges != NULL (probe keys the build side cannot contain). */
int vb2k_normalize_keys(const vb2_column* cols, int32_t ncols, const int64_t* mins, const uint64_t* mults, const uint64_t* ranges,
                        int32_t nulls_invalid, const int32_t* sel, int64_t n, uint64_t* keys_out, uint64_t* valid_out, void* stream);
/* min/max of an integer column over non-null rows: out = {min, max, nonnull_count} (device int64[3]). */
int vb2k_column_minmax(const vb2_column* col, int64_t rows, int64_t* out3, void* stream);
/* out = a & b over n bits (aggregate masks combined with validity) */
int vb2k_and_bits(const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out, void* stream);
/* ------------------------------------------------------------------------------------------
 * Hash join. Build: replaces HashBuild::addInput row store + HashTable::prepareJoinTable /
 * insertForJoin (velox/exec/HashBuild.cpp:442-598, exec/HashTable.cpp:1989,1518): key -> first
 * build row, duplicates chained through next[]. Probe: replaces HashTable::joinProbe +
 * listJoinResults (exec/HashTable.cpp:610-725,2133-2350): emits (probe row, build row) pairs in
 * probe-row order. Null keys never match (exec/HashBuild.cpp:475-479).
 * ------------------------------------------------------------------------------------------ */
typedef struct vb2_join_table {
  int32_t mode;        /* 0 = array (dense key range), 1 = hash */
  int32_t pad;
  int64_t key_min;     /* array mode: slot = key - key_min */
  int64_t capacity;    /* slots (array: range; hash: power of two) */
  uint64_t* keys;      /* hash mode: uint64[capacity], VB2_EMPTY_KEY = free */
  int32_t* head;       /* int32[capacity]: first build row + 1, 0 = empty */
  int32_t* next;       /* int32[build_rows]: next build row + 1 with the same key, 0 = end */
  int64_t build_rows;
} vb2_join_table;

int vb2k_join_build(const vb2_join_table* t, const uint64_t* build_keys, const uint64_t* valid, int64_t n,
                    int32_t* error_flag, void* stream);
/* Probe of a table WITHOUT duplicate build keys in one pass (HashTable::joinProbe, exec/HashTable.cpp:610-725,
 * for the unique-key case): hit_bits = bitmap of the probe rows that matched (warp ballots, LSB first;
 * expand with vb2k_bits_to_indices), hits[r] = matched build row of probe row r or -1. */
int vb2k_join_probe_unique(const vb2_join_table* t, const uint64_t* probe_keys, const uint64_t* valid, int64_t n, uint64_t* hit_bits,
                           int32_t* hits, void* stream);
/* codes[i] = dictionary index of row i of a DICTIONARY column (0 for CONSTANT); valid[i] (optional,
 * bytes) = 0 for NULL wrapper rows and NULL dictionary entries, whose code is written as 0. */
int vb2k_dictionary_codes(const vb2_column* col, int64_t n, int32_t* codes, uint8_t* valid, void* stream);
/* Array-mode build straight from ONE integer-typed key column (flat / dictionary / constant, NULLs
 * skipped): slot = v - lo + 1, i.e. vb2k_normalize_keys with min = lo followed by vb2k_join_build,
 * in one pass over the keys. flags: device int32[2] = {error (101: key outside the table), duplicates seen}. */
int vb2k_join_build_array_direct(int32_t* head, int32_t* next, int64_t capacity, const vb2_column* key, int64_t lo, int64_t n,
                                 int32_t* flags, void* stream);
/* Counts matches per probe row (hit_counts int32[n]), then after an exclusive scan the caller
 * asks for the pairs. total_out: device int64. */
int vb2k_join_probe_count(const vb2_join_table* t, const uint64_t* probe_keys, const uint64_t* valid, int64_t n,
                          int32_t* hit_counts, void* stream);
int vb2k_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, int64_t* total_out, void* workspace,
                            size_t workspace_bytes, void* stream);
size_t vb2k_scan_workspace(int64_t n);
int vb2k_join_probe_emit(const vb2_join_table* t, const uint64_t* probe_keys, const uint64_t* valid, int64_t n,
                         const int64_t* offsets, int32_t* probe_rows, int32_t* build_rows, void* stream);

/* ------------------------------------------------------------------------------------------
 * Peer-memory exchange (exchange_p2p.cu): the transfer of PartitionedOutput -> Exchange between
 * GPUs of one NVLink domain as plain stores into the destination's exchange heap (mapped with CUDA
 * IPC by vb2_comm), fused with the partition gather. Replaces the serialise / OutputBuffer / pull
 * path of velox/exec/PartitionedOutput.cpp + velox/exec/ExchangeClient.cpp for co-located ranks.
 * A segment holding `rows` rows stores column c at vb2k_p2p_segment_bytes(widths, c, rows).
 * ------------------------------------------------------------------------------------------ */
int64_t vb2k_p2p_segment_bytes(const int32_t* widths, int32_t ncols, int64_t rows);
/* `bytes` (multiple of 16) of src into peer_dst[p] for every p < world (host array of device pointers). */
int vb2k_p2p_put_block(void* const* peer_dst, int32_t world, const void* src, int64_t bytes, void* stream);
/* Release-stores `epoch` into *peer_flags[p] for every p (system scope), ordered after earlier puts of the stream. */
int vb2k_p2p_signal(void* const* peer_flags, int32_t world, uint64_t epoch, void* stream);
/* Spins (acquire, system scope) until flags[r * stride_words] >= epoch for every r < world; after
 * timeout_ns sets *error_flag = 200 + r and returns. */
int vb2k_p2p_wait(const uint64_t* flags, int32_t stride_words, int32_t world, uint64_t epoch, int32_t* error_flag, uint64_t timeout_ns, void* stream);
/* Rows grouped by destination (order[j] = source row, NULL = identity; counts_dev[p] rows for
 * destination p, device) -> column-major segments peer_segments[p]. broadcast: every destination gets all n rows. */
int vb2k_p2p_put_rows(const int32_t* order, const int64_t* counts_dev, int32_t world, int64_t n, const void* const* cols, const int32_t* widths,
                      int32_t ncols, void* const* peer_segments, int32_t broadcast, void* stream);
/* Segments received from every source (local_segments[s], counts[s] rows; host arrays) -> contiguous columns outs[c]. */
int vb2k_p2p_collect(const void* const* local_segments, const int64_t* counts, int32_t world, const int32_t* widths, int32_t ncols, void* const* outs,
                     void* stream);

/* Misc building blocks of the operator layer */
int vb2k_fill_u64(uint64_t* p, int64_t n, uint64_t v, void* stream);
int vb2k_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream);
int vb2k_iota_i32(int32_t* p, int64_t n, void* stream);
/* out[i] = (int32) in[i] */
int vb2k_narrow_i64(const int64_t* in, int64_t n, int32_t* out, void* stream);
/* out[i] = in[i] (widen) / out[i] = in[i] == 0 (widen_not): matched flags of outer/semi/anti joins */
int vb2k_widen_i32(const int32_t* in, int64_t n, int64_t* out, void* stream);
int vb2k_widen_not_i32(const int32_t* in, int64_t n, int64_t* out, void* stream);
/* valid bit k = idx[k] >= 0, clamped[k] = max(idx[k], 0): build side of unmatched left-join rows */
int vb2k_index_validity(const int32_t* idx, int64_t n, uint64_t* valid, int32_t* clamped, void* stream);
/* out bit k = in bit sel[k] (validity bitmaps under a selection) */
int vb2k_gather_bits(const uint64_t* in, const int32_t* sel, int64_t n, uint64_t* out, void* stream);
/* bit-packed values / validity -> one byte per row (0/1): the form bitmaps take inside an exchange */
int vb2k_unpack_bits(const uint64_t* in, int64_t n, uint8_t* out, void* stream);
/* one byte per row (0/1) -> bit-packed BOOLEAN values (FlatVector<bool> layout) */
int vb2k_pack_bools(const uint8_t* in, int64_t n, uint64_t* out, void* stream);
/* out[dst[i]] = in[src ? src[i] : i] for 4- or 8-byte elements (accumulator moves on rehash) */
int vb2k_scatter(const void* in, const int32_t* src, const int32_t* dst, int64_t n, int32_t elem_bytes, void* out, void* stream);
/* bit k = counts[k] > 0 */
int vb2k_positive_bits(const int64_t* counts, int64_t n, uint64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VELOX_B200_KERNELS_H_ */
