# Builds libvelox_b200.so (CUDA kernels for sm_100a + C++ operator layer + C ABI) in-tree,
# and the CPU oracle used by the tests (oracle/liboracle.so).
NVCC ?= nvcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS = -std=c++20 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unknown-pragmas -Iinclude -Ivelox_b200 -I$(BUILD)
CXXFLAGS = -std=c++20 -O2 -fPIC -Wall -Iinclude -Ivelox_b200 -I/usr/local/cuda/include
BUILD = build
LIB = velox_b200/lib/libvelox_b200.so

CU_SRCS = $(wildcard velox_b200/csrc/*.cu)
CPP_SRCS = $(wildcard velox_b200/csrc/host/*.cpp)
OBJS = $(patsubst velox_b200/csrc/%.cu,$(BUILD)/%.o,$(CU_SRCS)) \
       $(patsubst velox_b200/csrc/host/%.cpp,$(BUILD)/host_%.o,$(CPP_SRCS))

all: $(LIB) oracle

# vm_ops.inc as a string: the expression JIT hands the same text to NVRTC
$(BUILD)/vm_ops_str.h: velox_b200/csrc/vm_ops.inc
	@mkdir -p $(BUILD)
	( echo 'static const char kVmOpsSource[] = R"VMOPS('; cat $<; echo ')VMOPS";' ) > $@

# the headers the pipeline JIT hands to NVRTC, as strings (fused_jit.cu)
$(BUILD)/common_str.h: velox_b200/csrc/common.cuh
	@mkdir -p $(BUILD)
	( echo 'static const char kCommonCuhSource[] = R"VB2SRC('; cat $<; echo ')VB2SRC";' ) > $@
$(BUILD)/fused_scan_str.h: velox_b200/csrc/fused_scan.cuh
	@mkdir -p $(BUILD)
	( echo 'static const char kFusedScanCuhSource[] = R"VB2SRC('; cat $<; echo ')VB2SRC";' ) > $@

$(BUILD)/%.o: velox_b200/csrc/%.cu $(wildcard velox_b200/csrc/*.cuh) $(wildcard velox_b200/csrc/*.h) $(wildcard velox_b200/csrc/*.inc) $(wildcard include/*.h) $(BUILD)/vm_ops_str.h $(BUILD)/common_str.h $(BUILD)/fused_scan_str.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(BUILD)/host_%.o: velox_b200/csrc/host/%.cpp $(wildcard velox_b200/csrc/host/*.h) $(wildcard velox_b200/abi/*.h) $(wildcard include/*.h)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p velox_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -l:libnccl.so.2 -ldl

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -rf $(BUILD) $(LIB)
	$(MAKE) -s -C oracle clean

.PHONY: all oracle clean
