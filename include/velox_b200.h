/*
 * velox_b200 — operator-level C ABI: the drop-in boundary a host application binds.
 *
 * A task is a Velox plan fragment (values -> filter -> project -> aggregation / hash join ...)
 * executed by an unmodified Driver loop (velox/exec/Driver.cpp:538-850) whose FilterProject,
 * HashAggregation, HashBuild and HashProbe operators have been replaced through
 * DriverFactory::registerAdapter (velox/exec/Driver.h:789-847) by the B200 operators of this
 * library. Inputs are column batches in the layout of velox_b200_kernels.h (`vb2_column`), in
 * host memory (copied to the device by the inserted B200FromHost operator, as
 * velox/experimental/cudf/exec/CudfConversion.h:32 does) or already resident in HBM.
 *
 * Error behaviour mirrors the reference (SURVEY.md §8b "Errors"): data errors (integer overflow,
 * division by zero, failed cast — VeloxUserError) return VB2_ERR_USER, everything else
 * (VeloxRuntimeError: bad plan, CUDA failure, unsupported shape) a VB2_ERR_* code; the message
 * is written to `err`. There is no CPU fallback: an unsupported shape is an error.
 */
#ifndef VELOX_B200_H_
#define VELOX_B200_H_

#include "velox_b200_kernels.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vb2_task vb2_task;

enum { VB2_HOST = 0, VB2_DEVICE = 1 };

/* plan_text: grammar in DESIGN.md / velox_b200/plan.py. config: "key=value;key=value" query
 * config (core::QueryConfig, velox/core/QueryConfig.h): b200.enabled, b200.fused_pipelines,
 * b200.device_id. Returns NULL on error. */
vb2_task* vb2_task_create(const char* plan_text, const char* config, char* err, int32_t errlen);
/* Queues one batch for ValuesNode `source_id` (exec::Values, velox/exec/Values.h:21). The column
 * memory is borrowed: it must stay valid until vb2_task_run returns. */
int32_t vb2_task_add_input(vb2_task* task, int32_t source_id, const vb2_column* cols, int32_t ncols, int64_t rows,
                           int32_t location, char* err, int32_t errlen);
/* Runs the task to completion (Task::start + drivers; serial execution mode). */
/* Adds one input batch given through the Arrow C data interface (struct ArrowArray / ArrowSchema of a
 * record batch: format "+s", children l / i / tdD / g / b / u, optionally dictionary-encoded with int32
 * indices). The import happens in C++ (csrc/host/arrow_bridge.cpp: importFromArrowAsOwner,
 * velox/vector/arrow/Bridge.h:170): buffers are viewed, not copied; ownership of both structs moves
 * to the task, which calls their release callbacks when it is freed. */
struct ArrowArray;
struct ArrowSchema;
int32_t vb2_task_add_arrow(vb2_task* task, int32_t source_id, struct ArrowArray* array, struct ArrowSchema* schema, char* err, int32_t errlen);
int32_t vb2_task_run(vb2_task* task, char* err, int32_t errlen);

/* Runs several prepared tasks concurrently (one host thread each) and returns when all have finished:
 * the first non-zero status, with that task's message in err. Tasks whose plans exchange rows must
 * use different communicators. */
int32_t vb2_tasks_run(vb2_task* const* tasks, int32_t ntasks, char* err, int32_t errlen);

/* Scan-side device residency for host tables that several tasks read (the
 * Values / scan-side step in front of the path): while a cache is attached to a task, host buffers
 * it uploads are remembered by (address, bytes); a later task attached to the same cache reuses the
 * resident device copy instead of crossing PCIe again. The caller guarantees the host buffers are
 * neither modified nor freed while the cache lives, and runs the sharing tasks one after another.
 * The stats text of a task reports `task.h2dBytes`, the bytes it actually copied. */
/* Function registries (velox/expression/VectorFunction.h:241 registerVectorFunction,
 * velox/exec/Aggregate.h:525-575 registerAggregateFunction) for hosts that bind the C ABI instead of
 * the C++ classes (B200DeviceFunction, B200Aggregate in csrc/host/expr_compiler.h / operators.h).
 * vb2_register_scalar_function: `cuda_source` defines `__device__ RET entry(ARGS...)` over BIGINT =
 * long long, INTEGER = int, DOUBLE = double, BOOLEAN = bool; plans may then call (name e ...). The
 * body is spliced into the JIT-compiled kernel of every ExprSet that calls it (NULL in -> NULL out).
 * vb2_register_aggregate_function: `name(x)` aggregates input_function(x) with the device accumulator
 * family ("sum" "avg" "count" "min" "max"), and applies final_function to the final value;
 * input_function / final_function are names of registered scalar functions or NULL / "". */
int32_t vb2_register_scalar_function(const char* name, const char* entry, const char* cuda_source, int32_t ret_type, const int32_t* arg_types,
                                     int32_t nargs, char* err, int32_t errlen);
int32_t vb2_register_aggregate_function(const char* name, const char* family, const char* input_function, const char* final_function, char* err,
                                        int32_t errlen);
/* Node-at-a-time call of a registered scalar function — VectorFunction::apply
 * (velox/expression/VectorFunction.h:81-86) — over HOST argument columns: rows whose bit is set in
 * `selected` (LSB-first; NULL = all rows) are computed on the device and written to out_values
 * (BOOLEAN one byte per row) / out_nulls (one byte per row, 1 = NULL); the other rows of the caller's
 * buffers are left untouched (the result-reuse rule). */
int32_t vb2_scalar_function_apply(const char* name, const vb2_column* args, int32_t nargs, int64_t rows, const uint64_t* selected, int32_t ret_type,
                                  void* out_values, uint8_t* out_nulls, char* err, int32_t errlen);

/* Diagnostic, no GPU needed: the expression programs of every Filter / Project node of a plan are
 * compiled (expression compiler) and handed to the expression JIT for a flat NULL-free input;
 * reports the number of programs, of kernels (filter pass + projection pass) and of kernels that
 * generate and NVRTC-compile for sm_100a (the rest would run on the interpreter). */
int32_t vb2_plan_jit_report(const char* plan_text, int32_t* programs, int32_t* jit_kernels, int32_t* total_kernels, char* err, int32_t errlen);
/* Diagnostic (no GPU needed): the register program the expression compiler (ExprCompiler's role,
 * velox/expression/ExprCompiler.cpp) produces for the ordinal-th Filter / Project node of a plan, in caller buffers:
 * header = {n_instrs, n_consts, n_filter_instrs, filter_reg, n_regs, n_outputs, node is a filter}; per output its
 * register (-1: identity projection of input column out_identity[i]) and type; VARCHAR constants point into `strings`. */
int32_t vb2_plan_expression_program(const char* plan_text, int32_t ordinal, vb2_instr* instrs, int32_t instrs_cap, vb2_const* consts, int32_t consts_cap,
                                    char* strings, int32_t strings_cap, int32_t* header, int32_t* out_regs, int32_t* out_types, int32_t* out_identity,
                                    int32_t outs_cap, char* err, int32_t errlen);

typedef struct vb2_upload_cache vb2_upload_cache;
vb2_upload_cache* vb2_upload_cache_create(void);
void vb2_upload_cache_free(vb2_upload_cache* cache);
int32_t vb2_task_set_upload_cache(vb2_task* task, vb2_upload_cache* cache);

/* Result batches concatenated; copy-out into caller buffers. */
int64_t vb2_result_rows(vb2_task* task);
int32_t vb2_result_cols(vb2_task* task);
int32_t vb2_result_type(vb2_task* task, int32_t col);
/* Fixed width: values (BOOLEAN one byte per row) and one null flag byte per row (1 = NULL). */
void vb2_result_copy(vb2_task* task, int32_t col, void* values, uint8_t* nulls);
int64_t vb2_result_str_bytes(vb2_task* task, int32_t col);
void vb2_result_copy_str(vb2_task* task, int32_t col, int32_t* offsets, char* chars, uint8_t* nulls);
/* Config "b200.result_on_device=true": no B200ToHost is planted in front of the sink; the result
 * batches stay in HBM (vb2_result_rows still counts them, vb2_result_copy* return nothing) and their
 * buffers are lent here: cols[c] receives the device-side vb2_column of batch `batch`; returns its
 * row count. Valid until the task is freed. */
int32_t vb2_result_device_batches(vb2_task* task);
int64_t vb2_result_device_columns(vb2_task* task, int32_t batch, vb2_column* cols, int32_t ncols);
/* The whole result in two calls (bindings whose per-call cost matters): vb2_result_layout fills
 * layout[c * 4 ..] = {type, value bytes, offset bytes, char bytes} for every column and returns the
 * blob size; vb2_result_copy_all writes, per column, values | int32 offsets (VARCHAR) | chars
 * (VARCHAR) | one null flag byte per row, each region padded to 8 bytes. */
int64_t vb2_result_layout(vb2_task* task, int64_t* layout);
void vb2_result_copy_all(vb2_task* task, void* blob);
/* Host mirrors of device-resident VARCHAR dictionaries are remembered by buffer identity (device
 * dictionaries passed to vb2_task_add_input must not change while tasks use them); this forgets them. */
void vb2_dictionary_cache_clear(void);
/* Operator runtime stats as "pipeline.operator.type.name=value\n" lines (OperatorStats,
 * velox/exec/OperatorStats.h:93). Valid until the task is freed. */
const char* vb2_task_stats(vb2_task* task);
void vb2_task_free(vb2_task* task);

/* Hash-partitioned exchange across the GPUs of one node: partition ids follow
 * HashPartitionFunction (hash % world), payload moves with one grouped ncclSend/ncclRecv
 * all-to-all over NVLink. The communicator is created from an ncclUniqueId the caller broadcasts
 * (e.g. with torch.distributed). */
typedef struct vb2_comm vb2_comm;
int32_t vb2_comm_unique_id(uint8_t out[128]);
vb2_comm* vb2_comm_create(const uint8_t unique_id[128], int32_t world, int32_t rank, char* err, int32_t errlen);
void vb2_comm_free(vb2_comm* comm);
/* send: world contiguous segments of `elem_bytes`-wide elements described by send_counts (host
 * int64[world]); recv_counts (host int64[world]) is filled first through a count exchange when
 * recv == NULL, otherwise the payload is exchanged. Device pointers; stream-ordered. */
int32_t vb2_comm_exchange_counts(vb2_comm* comm, const int64_t* send_counts, int64_t* recv_counts, void* stream);
/* Same with the send counts already on the device (int64[world], e.g. from
 * vb2k_partition_scatter_order): exchanges them and returns both count vectors with one sync. */
int32_t vb2_comm_exchange_counts_dev(vb2_comm* comm, const int64_t* dev_send_counts, int64_t* send_counts_host, int64_t* recv_counts_host,
                                     void* stream);
int32_t vb2_comm_all_to_all(vb2_comm* comm, const void* send, const int64_t* send_counts, void* recv, const int64_t* recv_counts,
                            int32_t elem_bytes, void* stream);
/* All columns of a partitioned row set in one NCCL group (one launch): send[c] / recv[c] are
 * device arrays of elem_bytes[c]-wide elements segmented by send_counts / recv_counts (rows). */
int32_t vb2_comm_all_to_all_columns(vb2_comm* comm, int32_t ncols, const void* const* send, void* const* recv, const int32_t* elem_bytes,
                                    const int64_t* send_counts, const int64_t* recv_counts, void* stream);
/* Every rank's `bytes`-byte block to every rank: recv holds world blocks in rank order (the
 * metadata round of PartitionedOutput: per-destination row counts and dictionaries). */
int32_t vb2_comm_all_gather(vb2_comm* comm, const void* send, void* recv, int64_t bytes, void* stream);
/* Broadcast flavour of the exchange (PartitionedOutputNode::Kind::kBroadcast): every rank sends all
 * `send_rows` rows of every column to every rank; recv[c] is segmented by recv_counts (rows per source). */
int32_t vb2_comm_all_gather_columns(vb2_comm* comm, int32_t ncols, const void* const* send, void* const* recv, const int32_t* elem_bytes,
                                    int64_t send_rows, const int64_t* recv_counts, void* stream);
/* 1 when the ranks exchange rows through CUDA-IPC mapped peer memory over NVLink (exchange_p2p.cu; set
 * up at vb2_comm_create, VB2_EXCHANGE=nccl disables it), 0 when every exchange goes through NCCL. */
int32_t vb2_comm_peer_memory(vb2_comm* comm);
/* Exchanges moved so far through peer memory (peer_memory = 1) / through NCCL (0). */
int64_t vb2_comm_exchanges(vb2_comm* comm, int32_t peer_memory);
int32_t vb2_comm_world(vb2_comm* comm);
int32_t vb2_comm_rank(vb2_comm* comm);
/* Attaches the communicator to a task whose plan contains exchange nodes: B200PartitionedOutput /
 * B200Exchange move their pages through it (the reference selects an ExchangeSource by task URI,
 * velox/exec/ExchangeSource.h:139-145). The communicator must outlive the task. */
int32_t vb2_task_set_comm(vb2_task* task, vb2_comm* comm);
/* Sum-reduces n doubles / int64s in place across ranks (merge of per-GPU partial aggregates). */
int32_t vb2_comm_all_reduce_f64(vb2_comm* comm, double* data, int64_t n, void* stream);
int32_t vb2_comm_all_reduce_i64(vb2_comm* comm, int64_t* data, int64_t n, void* stream);
/* Element-wise maximum over ranks (exchange planning statistics). */
int32_t vb2_comm_all_reduce_max_i64(vb2_comm* comm, int64_t* data, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VELOX_B200_H_ */
