// Operator-level C ABI (include/velox_b200.h): plan text -> Task over the shim Driver with the
// B200 adapter installed; host / device column batches in, host result columns out.
#include <atomic>
#include <chrono>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <sstream>

#include "../../../include/velox_b200.h"
#include "../../abi/arrow_abi.h"
#include "operators.h"
#include "plan_text.h"
#include "task.h"

using namespace velox_b200;

struct vb2_upload_cache {
  velox_b200::UploadCache cache;
};

struct vb2_task {
  std::shared_ptr<exec::Task> task;
  vb2_upload_cache* uploadCache = nullptr;
  core::PlanNodePtr plan;
  memory::MemoryPool pool{"capi"};
  std::vector<RowVectorPtr> results;
  std::vector<B200VectorPtr> deviceResults;  // b200.result_on_device: result batches left in HBM
  // results concatenated per column for copy-out
  struct OutCol {
    int32_t type;
    std::vector<uint8_t> values;  // fixed width (BOOLEAN one byte per row)
    std::vector<int32_t> offsets;
    std::string chars;
    std::vector<uint8_t> nulls;
  };
  std::vector<OutCol> out;
  int64_t rows = 0;
  int64_t h2dBytes = 0;  // host -> device bytes of the last run (resident copies from an upload cache are not counted)
  int64_t parseNanos = 0, runNanos = 0, resultNanos = 0;
  std::string stats;
};

namespace {

void setErr(char* err, int32_t errlen, const std::string& msg) {
  if (!err || errlen <= 0) return;
  std::strncpy(err, msg.c_str(), errlen - 1);
  err[errlen - 1] = 0;
}

template <class F>
int32_t guarded(char* err, int32_t errlen, F&& f) {
  try {
    f();
    return VB2_OK;
  } catch (const VeloxUserError& e) {
    setErr(err, errlen, std::string("VeloxUserError: ") + e.what());
    return VB2_ERR_USER;
  } catch (const VeloxRuntimeError& e) {
    setErr(err, errlen, std::string("VeloxRuntimeError: ") + e.what());
    const std::string w = e.what();
    if (w.find("Unsupported") != std::string::npos || w.find("Not yet implemented") != std::string::npos) return VB2_ERR_UNSUPPORTED;
    if (w.find("CUDA") != std::string::npos) return VB2_ERR_CUDA;
    return VB2_ERR_INVALID;
  } catch (const std::exception& e) {
    setErr(err, errlen, std::string("VeloxRuntimeError: ") + e.what());
    return VB2_ERR_INVALID;
  }
}

TypePtr typeOf(int32_t t) {
  switch (t) {
    case VB2_BOOLEAN: return BOOLEAN();
    case VB2_INTEGER: return INTEGER();
    case VB2_BIGINT: return BIGINT();
    case VB2_DOUBLE: return DOUBLE();
    case VB2_VARCHAR: return VARCHAR();
    default: throw VeloxRuntimeError("unknown column type " + std::to_string(t));
  }
}

BufferPtr view(const void* p, size_t bytes) { return p ? std::make_shared<Buffer>(p, bytes) : nullptr; }

// Flat host values -> FlatVector over borrowed memory (VARCHAR: StringViews over the caller's chars).
VectorPtr importFlat(memory::MemoryPool* pool, int32_t type, const void* values, const void* aux, const uint64_t* nulls, int64_t n) {
  BufferPtr nb = view(nulls, bits::nbytes(n));
  const vector_size_t size = static_cast<vector_size_t>(n);
  switch (type) {
    case VB2_BOOLEAN: return std::make_shared<FlatVector<bool>>(pool, BOOLEAN(), nb, size, view(values, bits::nbytes(n)));
    case VB2_INTEGER: return std::make_shared<FlatVector<int32_t>>(pool, INTEGER(), nb, size, view(values, n * 4));
    case VB2_BIGINT: return std::make_shared<FlatVector<int64_t>>(pool, BIGINT(), nb, size, view(values, n * 8));
    case VB2_DOUBLE: return std::make_shared<FlatVector<double>>(pool, DOUBLE(), nb, size, view(values, n * 8));
    case VB2_VARCHAR: {
      const int32_t* off = static_cast<const int32_t*>(values);
      const char* chars = static_cast<const char*>(aux);
      BufferPtr views = AlignedBuffer::allocate<StringView>(n ? n : 1, pool);
      auto* sv = views->asMutable<StringView>();
      for (int64_t i = 0; i < n; ++i) sv[i] = StringView(chars + off[i], off[i + 1] - off[i]);
      return std::make_shared<FlatVector<StringView>>(pool, VARCHAR(), nb, size, views);
    }
    default: throw VeloxRuntimeError("unknown column type");
  }
}

VectorPtr importHostColumn(memory::MemoryPool* pool, const vb2_column& c, int64_t rows) {
  if (c.size != rows) throw VeloxRuntimeError("column size differs from the batch's row count");
  const vector_size_t n = static_cast<vector_size_t>(rows);
  if (c.encoding == VB2_FLAT) return importFlat(pool, c.type, c.values, c.aux, c.nulls, rows);
  if (c.encoding == VB2_DICTIONARY) {
    VectorPtr base = importFlat(pool, c.type, c.values, c.aux, c.dict_nulls, c.dict_size);
    return BaseVector::wrapInDictionary(view(c.nulls, bits::nbytes(rows)), view(c.indices, rows * 4), n, base);
  }
  if (c.encoding == VB2_CONSTANT) {
    const bool isNull = c.nulls && !bits::isBitSet(c.nulls, 0);
    switch (c.type) {
      case VB2_BOOLEAN: return std::make_shared<ConstantVector<bool>>(pool, n, isNull, BOOLEAN(), !isNull && bits::isBitSet(static_cast<const uint64_t*>(c.values), 0));
      case VB2_INTEGER: return std::make_shared<ConstantVector<int32_t>>(pool, n, isNull, INTEGER(), isNull ? 0 : *static_cast<const int32_t*>(c.values));
      case VB2_BIGINT: return std::make_shared<ConstantVector<int64_t>>(pool, n, isNull, BIGINT(), isNull ? 0 : *static_cast<const int64_t*>(c.values));
      case VB2_DOUBLE: return std::make_shared<ConstantVector<double>>(pool, n, isNull, DOUBLE(), isNull ? 0 : *static_cast<const double*>(c.values));
      case VB2_VARCHAR: {
        auto v = std::make_shared<ConstantVector<StringView>>(pool, n, isNull, VARCHAR(), StringView());
        if (!isNull) {
          const int32_t* off = static_cast<const int32_t*>(c.values);
          v->setStringStorage(std::string(static_cast<const char*>(c.aux) + off[0], off[1] - off[0]));
        }
        return v;
      }
      default: throw VeloxRuntimeError("unknown column type");
    }
  }
  throw VeloxRuntimeError("unknown column encoding");
}

// Host mirrors of device-resident dictionaries, remembered by the identity of their buffers: a table
// whose batches are fed to many tasks would otherwise pay two blocking device->host copies per
// dictionary column and task. Device dictionaries handed to vb2_task_add_input are immutable while in
// use (as Arrow buffers are); vb2_dictionary_cache_clear() forgets the mirrors.
struct DictKey {
  const void* values; const void* aux; const void* nulls; int64_t entries;
  bool operator<(const DictKey& o) const { return std::tie(values, aux, nulls, entries) < std::tie(o.values, o.aux, o.nulls, o.entries); }
};
std::mutex g_dictMu;
std::map<DictKey, std::shared_ptr<const HostAlphabet>> g_dictCache;

// Device-resident batch: borrow the pointers; small VARCHAR alphabets are mirrored on the host.
DeviceColumnPtr importDeviceColumn(const vb2_column& c, int64_t rows) {
  if (c.size != rows) throw VeloxRuntimeError("column size differs from the batch's row count");
  auto col = std::make_shared<DeviceColumn>();
  col->type = typeOf(c.type);
  col->desc = c;
  if (c.type == VB2_VARCHAR && c.encoding != VB2_FLAT) {
    const int64_t entries = c.encoding == VB2_DICTIONARY ? c.dict_size : 1;
    const uint64_t* dn0 = c.encoding == VB2_DICTIONARY ? c.dict_nulls : c.nulls;
    const DictKey key{c.values, c.aux, dn0, entries};
    {
      std::lock_guard<std::mutex> l(g_dictMu);
      auto it = g_dictCache.find(key);
      if (it != g_dictCache.end()) { col->alphabet = it->second; return col; }
    }
    if (entries <= (1 << 16)) {
      std::vector<int32_t> off(entries + 1);
      VB2_CU(cudaMemcpy(off.data(), c.values, off.size() * 4, cudaMemcpyDeviceToHost));
      std::string chars(off[entries], '\0');
      if (!chars.empty()) VB2_CU(cudaMemcpy(chars.data(), c.aux, chars.size(), cudaMemcpyDeviceToHost));
      std::vector<uint64_t> nb;
      const uint64_t* dn = c.encoding == VB2_DICTIONARY ? c.dict_nulls : c.nulls;
      if (dn) {
        nb.resize(bits::nwords(entries));
        VB2_CU(cudaMemcpy(nb.data(), dn, nb.size() * 8, cudaMemcpyDeviceToHost));
      }
      auto alpha = std::make_shared<HostAlphabet>();
      for (int64_t i = 0; i < entries; ++i) {
        alpha->values.push_back(chars.substr(off[i], off[i + 1] - off[i]));
        alpha->nulls.push_back(dn ? !bits::isBitSet(nb.data(), i) : false);
      }
      col->alphabet = alpha;
      std::lock_guard<std::mutex> l(g_dictMu);
      if (g_dictCache.size() > 4096) g_dictCache.clear();
      g_dictCache[key] = alpha;
    }
  }
  return col;
}

void appendResult(vb2_task& t, const RowVectorPtr& batch) {
  const vector_size_t n = batch->size();
  if (t.out.empty()) {
    t.out.resize(batch->childrenSize());
    for (size_t c = 0; c < t.out.size(); ++c) {
      t.out[c].type = veloxTypeToVb2(batch->childAt(static_cast<uint32_t>(c))->type());
      t.out[c].offsets.push_back(0);
    }
  }
  for (size_t c = 0; c < t.out.size(); ++c) {
    auto& o = t.out[c];
    const BaseVector* v = batch->childAt(static_cast<uint32_t>(c)).get();
    // decode dictionary wrapping
    const vector_size_t* idx = nullptr;
    const BaseVector* base = v;
    if (v->encoding() == VectorEncoding::Simple::DICTIONARY) {
      switch (v->typeKind()) {
        case TypeKind::BOOLEAN: idx = v->as<DictionaryVector<bool>>()->rawIndices(); base = v->as<DictionaryVector<bool>>()->valueVector().get(); break;
        case TypeKind::INTEGER: idx = v->as<DictionaryVector<int32_t>>()->rawIndices(); base = v->as<DictionaryVector<int32_t>>()->valueVector().get(); break;
        case TypeKind::BIGINT: idx = v->as<DictionaryVector<int64_t>>()->rawIndices(); base = v->as<DictionaryVector<int64_t>>()->valueVector().get(); break;
        case TypeKind::DOUBLE: idx = v->as<DictionaryVector<double>>()->rawIndices(); base = v->as<DictionaryVector<double>>()->valueVector().get(); break;
        default: idx = v->as<DictionaryVector<StringView>>()->rawIndices(); base = v->as<DictionaryVector<StringView>>()->valueVector().get();
      }
    }
    // fast path: flat fixed-width column (the common shape of aggregation / projection results)
    if (!idx && v->encoding() == VectorEncoding::Simple::FLAT && (o.type == VB2_INTEGER || o.type == VB2_BIGINT || o.type == VB2_DOUBLE)) {
      const size_t w = o.type == VB2_INTEGER ? 4 : 8;
      const uint8_t* src = o.type == VB2_INTEGER ? reinterpret_cast<const uint8_t*>(base->as<FlatVector<int32_t>>()->rawValues())
                         : o.type == VB2_BIGINT ? reinterpret_cast<const uint8_t*>(base->as<FlatVector<int64_t>>()->rawValues())
                                                : reinterpret_cast<const uint8_t*>(base->as<FlatVector<double>>()->rawValues());
      const size_t at = o.values.size();
      o.values.resize(at + static_cast<size_t>(n) * w);
      std::memcpy(o.values.data() + at, src, static_cast<size_t>(n) * w);
      const size_t nat = o.nulls.size();
      o.nulls.resize(nat + n, 0);
      if (const uint64_t* raw = v->rawNulls())
        for (vector_size_t i = 0; i < n; ++i) o.nulls[nat + i] = bits::isBitNull(raw, i);
      continue;
    }
    for (vector_size_t i = 0; i < n; ++i) {
      const bool isNull = v->isNullAt(i);
      const vector_size_t s = idx ? idx[i] : i;
      o.nulls.push_back(isNull);
      switch (o.type) {
        case VB2_BOOLEAN: o.values.push_back(isNull ? 0 : base->as<FlatVector<bool>>()->valueAt(s)); break;
        case VB2_INTEGER: { int32_t x = isNull ? 0 : base->as<FlatVector<int32_t>>()->rawValues()[s]; auto* p = reinterpret_cast<uint8_t*>(&x); o.values.insert(o.values.end(), p, p + 4); break; }
        case VB2_BIGINT: { int64_t x = isNull ? 0 : base->as<FlatVector<int64_t>>()->rawValues()[s]; auto* p = reinterpret_cast<uint8_t*>(&x); o.values.insert(o.values.end(), p, p + 8); break; }
        case VB2_DOUBLE: { double x = isNull ? 0 : base->as<FlatVector<double>>()->rawValues()[s]; auto* p = reinterpret_cast<uint8_t*>(&x); o.values.insert(o.values.end(), p, p + 8); break; }
        default: {
          if (!isNull) {
            const StringView sv = base->as<FlatVector<StringView>>()->rawValues()[s];
            o.chars.append(sv.data(), sv.size());
          }
          o.offsets.push_back(static_cast<int32_t>(o.chars.size()));
        }
      }
    }
  }
  t.rows += n;
}

}  // namespace

namespace {
// VB2_DEBUG_SEGV=1: print the native stack on SIGSEGV (development aid; no debugger on the GPU boxes)
void segvHandler(int sig) {
  void* frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "\n[velox_b200] fatal signal, native stack:\n";
  (void)!write(2, msg, sizeof(msg) - 1);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}
struct SegvInstaller {
  SegvInstaller() {
    const char* e = std::getenv("VB2_DEBUG_SEGV");
    if (e && e[0] == '1') signal(SIGSEGV, segvHandler);
    const char* t = std::getenv("VB2_SYNC_TIMING");  // attribute kernel time to operators in the wall-time stats
    if (t && t[0] == '1') facebook::velox::exec::g_timingSync = []() { cudaDeviceSynchronize(); };
  }
} g_segvInstaller;
}  // namespace

extern "C" {

vb2_task* vb2_task_create(const char* plan_text, const char* config, char* err, int32_t errlen) {
  vb2_task* t = nullptr;
  const int32_t rc = guarded(err, errlen, [&] {
    registerB200();
    std::unordered_map<std::string, std::string> kv;
    if (config) {
      std::stringstream ss(config);
      std::string item;
      while (std::getline(ss, item, ';')) {
        const auto eq = item.find('=');
        if (eq != std::string::npos) kv[item.substr(0, eq)] = item.substr(eq + 1);
      }
    }
    auto task = std::make_unique<vb2_task>();
    const auto t0 = std::chrono::steady_clock::now();
    task->plan = parsePlanText(plan_text ? plan_text : "");
    task->parseNanos = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    task->task = std::make_shared<exec::Task>(task->plan, core::QueryConfig(std::move(kv)));
    t = task.release();
  });
  return rc == VB2_OK ? t : nullptr;
}

int32_t vb2_task_add_input(vb2_task* task, int32_t source_id, const vb2_column* cols, int32_t ncols, int64_t rows,
                           int32_t location, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    VELOX_CHECK(task && cols, "null task or columns");
    VELOX_CHECK(rows >= 0 && rows < (1ll << 31), "a batch holds fewer than 2^31 rows (vector_size_t, velox/vector/TypeAliases.h:29)");
    std::vector<std::string> names;
    std::vector<TypePtr> types;
    for (int32_t c = 0; c < ncols; ++c) { names.push_back("c" + std::to_string(c)); types.push_back(typeOf(cols[c].type)); }
    RowTypePtr type = ROW(names, types);
    if (rows == 0) return;
    if (location == VB2_DEVICE) {
      std::vector<DeviceColumnPtr> dc;
      for (int32_t c = 0; c < ncols; ++c) dc.push_back(importDeviceColumn(cols[c], rows));
      task->task->addInput(source_id, std::make_shared<B200Vector>(&task->pool, type, static_cast<vector_size_t>(rows), std::move(dc), nullptr));
    } else {
      std::vector<VectorPtr> children;
      for (int32_t c = 0; c < ncols; ++c) children.push_back(importHostColumn(&task->pool, cols[c], rows));
      task->task->addInput(source_id, std::make_shared<RowVector>(&task->pool, type, nullptr, static_cast<vector_size_t>(rows), std::move(children)));
    }
  });
}

int32_t vb2_task_add_arrow(vb2_task* task, int32_t source_id, struct ArrowArray* array, struct ArrowSchema* schema, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    VELOX_CHECK(task && array && schema, "null task, array or schema");
    VELOX_CHECK(array->length < (1ll << 31), "a batch holds fewer than 2^31 rows (vector_size_t, velox/vector/TypeAliases.h:29)");
    auto batch = std::dynamic_pointer_cast<RowVector>(importFromArrowAsOwner(*schema, *array, &task->pool));
    VELOX_CHECK(batch != nullptr, "Arrow input must be a struct array (record batch)");
    if (batch->size() > 0) task->task->addInput(source_id, batch);
  });
}

int32_t vb2_task_run(vb2_task* task, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    VELOX_CHECK(task != nullptr, "null task");
    struct Attach {  // the serial Task runs its drivers on this thread
      explicit Attach(vb2_upload_cache* c) { velox_b200::setThreadUploadCache(c ? &c->cache : nullptr); }
      ~Attach() { velox_b200::setThreadUploadCache(nullptr); }
    } attach(task->uploadCache);
    const int64_t uploadedBefore = velox_b200::threadUploadedBytes();
    // driver threads (task.max_drivers > 1) inherit this thread's device and upload cache, and report their copies
    int device = 0;
    cudaGetDevice(&device);
    std::atomic<int64_t> threadBytes{0};
    vb2_upload_cache* uc = task->uploadCache;
    thread_local int64_t threadStart = 0;
    task->task->setDriverThreadHooks(
        [device, uc] {
          cudaSetDevice(device);
          velox_b200::setThreadUploadCache(uc ? &uc->cache : nullptr);
          threadStart = velox_b200::threadUploadedBytes();
        },
        [&threadBytes] {
          threadBytes += velox_b200::threadUploadedBytes() - threadStart;
          velox_b200::setThreadUploadCache(nullptr);
        });
    const auto tRun = std::chrono::steady_clock::now();
    task->results = task->task->run();
    const auto tRan = std::chrono::steady_clock::now();
    task->runNanos = std::chrono::duration_cast<std::chrono::nanoseconds>(tRan - tRun).count();
    task->h2dBytes = velox_b200::threadUploadedBytes() - uploadedBefore + threadBytes.load();
    task->out.clear();
    task->rows = 0;
    task->deviceResults.clear();
    bool allDevice = !task->results.empty();
    for (auto& b : task->results) allDevice = allDevice && std::dynamic_pointer_cast<B200Vector>(b) != nullptr;
    if (allDevice) {
      // b200.result_on_device: batches stay in HBM; vb2_result_device_columns lends their buffers
      for (auto& b : task->results) {
        auto dv = std::dynamic_pointer_cast<B200Vector>(b);
        VB2_CU(cudaStreamSynchronize(dv->stream()));
        task->rows += dv->size();
        task->deviceResults.push_back(dv);
      }
      const auto& type = task->plan->outputType();
      task->out.resize(type->size());
      for (uint32_t c = 0; c < type->size(); ++c) { task->out[c].type = veloxTypeToVb2(type->childAt(c)); task->out[c].offsets.push_back(0); }
      task->results.clear();
    }
    for (auto& b : task->results) appendResult(*task, b);
    if (task->out.empty()) {
      const auto& type = task->plan->outputType();
      task->out.resize(type->size());
      for (uint32_t c = 0; c < type->size(); ++c) { task->out[c].type = veloxTypeToVb2(type->childAt(c)); task->out[c].offsets.push_back(0); }
    }
    task->resultNanos = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - tRan).count();
    std::ostringstream os;
    for (auto& kv : task->task->stats()) os << kv.first << "=" << kv.second << "\n";
    os << "task.h2dBytes=" << task->h2dBytes << "\n";
    os << "task.parseWallNanos=" << task->parseNanos << "\n";
    os << "task.runWallNanos=" << task->runNanos << "\n";
    os << "task.resultWallNanos=" << task->resultNanos << "\n";
    task->stats = os.str();
    task->results.clear();
  });
}

// Runs several prepared tasks at once, one host thread each (the reference runs the drivers of
// concurrent tasks on its executor's threads, velox/exec/Task.cpp Task::start): independent queries
// overlap their host-side setup, build sides and exchanges with each other's scans. Returns the first
// non-zero status; err receives that task's message.
int32_t vb2_tasks_run(vb2_task* const* tasks, int32_t ntasks, char* err, int32_t errlen) {
  if (ntasks <= 0) return VB2_OK;
  if (ntasks == 1) return vb2_task_run(tasks[0], err, errlen);
  int device = 0;
  cudaGetDevice(&device);
  std::vector<int32_t> rc(ntasks, VB2_OK);
  std::vector<std::string> msgs(ntasks, std::string(1024, '\0'));
  std::vector<std::thread> threads;
  for (int32_t i = 1; i < ntasks; ++i)
    threads.emplace_back([&, i] {
      cudaSetDevice(device);
      rc[i] = vb2_task_run(tasks[i], msgs[i].data(), 1024);
    });
  rc[0] = vb2_task_run(tasks[0], msgs[0].data(), 1024);
  for (auto& t : threads) t.join();
  for (int32_t i = 0; i < ntasks; ++i)
    if (rc[i] != VB2_OK) {
      setErr(err, errlen, msgs[i].c_str());
      return rc[i];
    }
  return VB2_OK;
}

int32_t vb2_register_scalar_function(const char* name, const char* entry, const char* cuda_source, int32_t ret_type, const int32_t* arg_types,
                                     int32_t nargs, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    VELOX_CHECK(name && entry && cuda_source && nargs >= 1 && nargs <= 3, "vb2_register_scalar_function: name, entry, source and 1-3 argument types");
    registerB200Functions();
    std::vector<TypePtr> args;
    auto sig = std::make_shared<exec::FunctionSignature>();
    auto nameOf = [](int32_t t) -> std::string {
      std::string n = typeOf(t)->toString();
      for (auto& c : n) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
      return n;
    };
    for (int32_t i = 0; i < nargs; ++i) {
      args.push_back(typeOf(arg_types[i]));
      sig->argTypes.push_back(nameOf(arg_types[i]));
    }
    sig->returnType = nameOf(ret_type);
    exec::registerVectorFunction(name, {sig}, std::make_unique<B200DeviceFunction>(entry, cuda_source, typeOf(ret_type), args));
  });
}

int32_t vb2_register_aggregate_function(const char* name, const char* family, const char* input_function, const char* final_function, char* err,
                                        int32_t errlen) {
  return guarded(err, errlen, [&] {
    VELOX_CHECK(name && family, "vb2_register_aggregate_function: name and family");
    registerB200Aggregates();
    const std::string in = input_function ? input_function : "", fin = final_function ? final_function : "";
    if (!in.empty()) scalarFunctionReturnType(in, DOUBLE());   // throws when the function is not registered
    if (!fin.empty()) scalarFunctionReturnType(fin, DOUBLE());
    registerB200Aggregate(name, family, in, fin);
  });
}

int32_t vb2_scalar_function_apply(const char* name, const vb2_column* args, int32_t nargs, int64_t rows, const uint64_t* selected, int32_t ret_type,
                                  void* out_values, uint8_t* out_nulls, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    registerB200Functions();
    auto fn = exec::getVectorFunction(name);
    VELOX_CHECK(fn != nullptr, std::string("scalar function '") + name + "' is not registered");
    static memory::MemoryPool pool("b200.apply");
    std::vector<VectorPtr> argv;
    for (int32_t i = 0; i < nargs; ++i) argv.push_back(importHostColumn(&pool, args[i], rows));
    const vector_size_t n = static_cast<vector_size_t>(rows);
    SelectivityVector sel(n, true);
    if (selected) sel.setFromBits(selected, n);
    // a pre-allocated result holding the caller's current values: apply() may only overwrite selected rows
    const TypePtr type = typeOf(ret_type);
    VectorPtr result;
    auto prefill = [&](auto tag) {
      using T = decltype(tag);
      auto values = AlignedBuffer::allocate<T>(n ? n : 1, &pool);
      std::memcpy(values->template asMutable<T>(), out_values, static_cast<size_t>(n) * sizeof(T));
      result = std::make_shared<FlatVector<T>>(&pool, type, nullptr, n, values);
    };
    if (ret_type == VB2_BOOLEAN) {
      auto values = std::make_shared<Buffer>(bits::nbytes(n), &pool);
      std::memset(values->asMutable<uint8_t>(), 0, values->capacity());
      for (vector_size_t i = 0; i < n; ++i) bits::setBit(values->asMutable<uint64_t>(), i, static_cast<const uint8_t*>(out_values)[i] != 0);
      result = std::make_shared<FlatVector<bool>>(&pool, type, nullptr, n, values);
    } else if (ret_type == VB2_INTEGER) prefill(int32_t{});
    else if (ret_type == VB2_BIGINT) prefill(int64_t{});
    else if (ret_type == VB2_DOUBLE) prefill(double{});
    else VELOX_UNSUPPORTED("apply: result type");
    for (vector_size_t i = 0; i < n; ++i)
      if (out_nulls[i]) result->setNull(i, true);
    exec::EvalCtx ctx(&pool);
    fn->apply(sel, argv, type, ctx, result);
    VELOX_CHECK(result && result->size() >= n && result->isFlatEncoding(), "apply: function returned an unexpected vector");
    for (vector_size_t i = 0; i < n; ++i) {
      out_nulls[i] = result->isNullAt(i) ? 1 : 0;
      if (out_nulls[i]) continue;
      switch (ret_type) {
        case VB2_BOOLEAN: static_cast<uint8_t*>(out_values)[i] = result->as<FlatVector<bool>>()->valueAt(i); break;
        case VB2_INTEGER: static_cast<int32_t*>(out_values)[i] = result->as<FlatVector<int32_t>>()->valueAt(i); break;
        case VB2_BIGINT: static_cast<int64_t*>(out_values)[i] = result->as<FlatVector<int64_t>>()->valueAt(i); break;
        default: static_cast<double*>(out_values)[i] = result->as<FlatVector<double>>()->valueAt(i);
      }
    }
  });
}

// Diagnostic (no GPU needed): compiles the expression programs of every Filter / Project node of a
// plan with the expression compiler and reports how many of their kernels the expression JIT can
// generate and NVRTC-compile for a flat, NULL-free input (VARCHAR as a dictionary column).
int32_t vb2_plan_jit_report(const char* plan_text, int32_t* programs, int32_t* jit_kernels, int32_t* total_kernels, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    registerB200Functions();
    auto plan = parsePlanText(plan_text);
    int np = 0, nk = 0, nj = 0;
    std::function<void(const core::PlanNodePtr&)> walk = [&](const core::PlanNodePtr& node) {
      for (auto& s : node->sources()) walk(s);
      std::vector<core::TypedExprPtr> exprs;
      bool hasFilter = false;
      RowTypePtr inType;
      if (auto f = std::dynamic_pointer_cast<const core::FilterNode>(node)) {
        inType = f->sources()[0]->outputType();
        exprs.push_back(f->filter());
        hasFilter = true;
        for (uint32_t i = 0; i < inType->size(); ++i)
          exprs.push_back(std::make_shared<core::FieldAccessTypedExpr>(inType->childAt(i), inType->nameOf(i)));
      } else if (auto p = std::dynamic_pointer_cast<const core::ProjectNode>(node)) {
        inType = p->sources()[0]->outputType();
        exprs = p->projections();
      } else {
        return;
      }
      CompiledProgram prog = compileExprs(exprs, hasFilter, inType);
      ++np;
      std::vector<vb2_column> cols(inType->size());
      for (uint32_t i = 0; i < inType->size(); ++i) {
        vb2_column& c = cols[i];
        c = vb2_column{};
        c.type = veloxTypeToVb2(inType->childAt(i));
        c.encoding = c.type == VB2_VARCHAR ? VB2_DICTIONARY : VB2_FLAT;
        c.size = 1;
      }
      const vb2_program view = prog.view();
      std::vector<vb2_output> outs;
      for (auto& o : prog.outputs)
        if (o.reg >= 0) outs.push_back(vb2_output{o.reg, veloxTypeToVb2(o.type), nullptr, nullptr});
      if (hasFilter) {
        ++nk;
        nj += vb2k_expression_jit_compiles(&view, cols.data(), static_cast<int32_t>(cols.size()), 1, nullptr, 0, nullptr, 0);
      }
      if (!outs.empty()) {
        ++nk;
        nj += vb2k_expression_jit_compiles(&view, cols.data(), static_cast<int32_t>(cols.size()), 0, outs.data(), static_cast<int32_t>(outs.size()), nullptr, 0);
      }
    };
    walk(plan);
    if (programs) *programs = np;
    if (jit_kernels) *jit_kernels = nj;
    if (total_kernels) *total_kernels = nk;
  });
}

// Diagnostic (no GPU needed): the register program the expression compiler produces for the `ordinal`-th Filter /
// Project node of a plan (post-order over the plan, sources first — the order vb2_plan_jit_report walks), copied into
// caller buffers: instructions, constants (VARCHAR constants point into `strings`), and per output its register (-1: an
// identity projection of input column out_identity[i]) and type. header = {n_instrs, n_consts, n_filter_instrs,
// filter_reg, n_regs, n_outputs, node is a filter}. tests/test_jit_source_on_host.py runs such programs through the
// JIT's generated source on the CPU.
int32_t vb2_plan_expression_program(const char* plan_text, int32_t ordinal, vb2_instr* instrs, int32_t instrs_cap, vb2_const* consts, int32_t consts_cap,
                                    char* strings, int32_t strings_cap, int32_t* header, int32_t* out_regs, int32_t* out_types, int32_t* out_identity,
                                    int32_t outs_cap, char* err, int32_t errlen) {
  return guarded(err, errlen, [&] {
    VELOX_CHECK(plan_text && instrs && consts && strings && header && out_regs && out_types && out_identity, "null argument");
    registerB200Functions();
    auto plan = parsePlanText(plan_text);
    int seen = 0;
    bool found = false;
    std::function<void(const core::PlanNodePtr&)> walk = [&](const core::PlanNodePtr& node) {
      for (auto& s : node->sources()) walk(s);
      if (found) return;
      std::vector<core::TypedExprPtr> exprs;
      bool hasFilter = false;
      RowTypePtr inType;
      if (auto f = std::dynamic_pointer_cast<const core::FilterNode>(node)) {
        inType = f->sources()[0]->outputType();
        exprs.push_back(f->filter());
        hasFilter = true;
      } else if (auto p = std::dynamic_pointer_cast<const core::ProjectNode>(node)) {
        inType = p->sources()[0]->outputType();
        exprs = p->projections();
      } else {
        return;
      }
      if (seen++ != ordinal) return;
      found = true;
      CompiledProgram prog = compileExprs(exprs, hasFilter, inType);
      VELOX_CHECK(static_cast<int32_t>(prog.instrs.size()) <= instrs_cap && static_cast<int32_t>(prog.consts.size()) <= consts_cap &&
                      static_cast<int32_t>(prog.outputs.size()) <= outs_cap,
                  "program does not fit the caller's buffers");
      std::copy(prog.instrs.begin(), prog.instrs.end(), instrs);
      size_t at = 0;
      std::vector<size_t> offs;
      for (auto& str : prog.constStrings) {
        VELOX_CHECK(at + str.size() + 1 <= static_cast<size_t>(strings_cap), "string constants do not fit the caller's buffer");
        offs.push_back(at);
        std::memcpy(strings + at, str.data(), str.size());
        at += str.size();
        strings[at++] = 0;
      }
      for (size_t i = 0; i < prog.consts.size(); ++i) {
        consts[i] = prog.consts[i];
        if (consts[i].type == VB2_VARCHAR && !consts[i].is_null && consts[i].pad > 0) consts[i].str = strings + offs[consts[i].pad - 1];
      }
      for (size_t i = 0; i < prog.outputs.size(); ++i) {
        out_regs[i] = prog.outputs[i].reg;
        out_identity[i] = prog.outputs[i].identityField;
        out_types[i] = veloxTypeToVb2(prog.outputs[i].type);
      }
      header[0] = static_cast<int32_t>(prog.instrs.size());
      header[1] = static_cast<int32_t>(prog.consts.size());
      header[2] = prog.nFilterInstrs;
      header[3] = prog.filterReg;
      header[4] = prog.nRegs;
      header[5] = static_cast<int32_t>(prog.outputs.size());
      header[6] = hasFilter ? 1 : 0;
    };
    walk(plan);
    VELOX_CHECK(found, "the plan has no Filter / Project node with that ordinal");
  });
}

int32_t vb2_task_set_comm(vb2_task* task, vb2_comm* comm) {
  if (!task || !task->task) return VB2_ERR_INVALID;
  task->task->setExchangeTransport(comm ? std::make_shared<velox_b200::NcclTransport>(comm) : nullptr);
  return VB2_OK;
}

vb2_upload_cache* vb2_upload_cache_create(void) { return new vb2_upload_cache(); }
void vb2_upload_cache_free(vb2_upload_cache* cache) { delete cache; }
int32_t vb2_task_set_upload_cache(vb2_task* task, vb2_upload_cache* cache) {
  if (!task) return VB2_ERR_INVALID;
  task->uploadCache = cache;
  return VB2_OK;
}

int64_t vb2_result_rows(vb2_task* task) { return task->rows; }
int32_t vb2_result_cols(vb2_task* task) { return static_cast<int32_t>(task->out.size()); }
int32_t vb2_result_type(vb2_task* task, int32_t col) { return task->out[col].type; }
void vb2_result_copy(vb2_task* task, int32_t col, void* values, uint8_t* nulls) {
  auto& o = task->out[col];
  if (values && !o.values.empty()) std::memcpy(values, o.values.data(), o.values.size());
  if (nulls && !o.nulls.empty()) std::memcpy(nulls, o.nulls.data(), o.nulls.size());
}
int64_t vb2_result_str_bytes(vb2_task* task, int32_t col) { return static_cast<int64_t>(task->out[col].chars.size()); }
void vb2_result_copy_str(vb2_task* task, int32_t col, int32_t* offsets, char* chars, uint8_t* nulls) {
  auto& o = task->out[col];
  std::memcpy(offsets, o.offsets.data(), o.offsets.size() * 4);
  if (!o.chars.empty()) std::memcpy(chars, o.chars.data(), o.chars.size());
  if (nulls && !o.nulls.empty()) std::memcpy(nulls, o.nulls.data(), o.nulls.size());
}
void vb2_dictionary_cache_clear(void) {
  std::lock_guard<std::mutex> l(g_dictMu);
  g_dictCache.clear();
}

// Whole result in two calls: layout[c * 4 ..] = {type, value bytes, offset bytes, char bytes}; the blob
// holds, per column and in this order, values (fixed width; BOOLEAN one byte per row), int32 offsets
// (VARCHAR), chars (VARCHAR), null flags (one byte per row), each region padded to 8 bytes.
int64_t vb2_result_layout(vb2_task* task, int64_t* layout) {
  int64_t total = 0;
  auto pad = [](size_t b) { return static_cast<int64_t>((b + 7) / 8 * 8); };
  for (size_t c = 0; c < task->out.size(); ++c) {
    auto& o = task->out[c];
    const int64_t offBytes = o.type == VB2_VARCHAR ? static_cast<int64_t>(o.offsets.size() * 4) : 0;
    if (layout) {
      layout[c * 4 + 0] = o.type;
      layout[c * 4 + 1] = static_cast<int64_t>(o.values.size());
      layout[c * 4 + 2] = offBytes;
      layout[c * 4 + 3] = static_cast<int64_t>(o.chars.size());
    }
    total += pad(o.values.size()) + pad(offBytes) + pad(o.chars.size()) + pad(task->rows);
  }
  return total;
}
void vb2_result_copy_all(vb2_task* task, void* blob) {
  uint8_t* p = static_cast<uint8_t*>(blob);
  auto pad = [](size_t b) { return (b + 7) / 8 * 8; };
  for (auto& o : task->out) {
    if (!o.values.empty()) std::memcpy(p, o.values.data(), o.values.size());
    p += pad(o.values.size());
    if (o.type == VB2_VARCHAR) {
      std::memcpy(p, o.offsets.data(), o.offsets.size() * 4);
      p += pad(o.offsets.size() * 4);
      if (!o.chars.empty()) std::memcpy(p, o.chars.data(), o.chars.size());
      p += pad(o.chars.size());
    }
    if (!o.nulls.empty()) std::memcpy(p, o.nulls.data(), o.nulls.size());
    else std::memset(p, 0, static_cast<size_t>(task->rows));
    p += pad(static_cast<size_t>(task->rows));
  }
}

int32_t vb2_result_device_batches(vb2_task* task) { return task ? static_cast<int32_t>(task->deviceResults.size()) : 0; }
int64_t vb2_result_device_columns(vb2_task* task, int32_t batch, vb2_column* cols, int32_t ncols) {
  if (!task || batch < 0 || batch >= static_cast<int32_t>(task->deviceResults.size())) return -1;
  const auto& dv = task->deviceResults[batch];
  for (int32_t c = 0; c < ncols && c < static_cast<int32_t>(dv->columns().size()); ++c) cols[c] = dv->column(c)->desc;
  return dv->size();
}

const char* vb2_task_stats(vb2_task* task) { return task->stats.c_str(); }
void vb2_task_free(vb2_task* task) { delete task; }

}  // extern "C"
