# Top-level build of the B200-native QuEST backend.
#
#   make kernels   -> quest_b200/lib/libquest_b200.so   hand-written sm_100a CUDA + the C ABI (include/quest_b200.h)
#   make quest     -> quest_b200/lib/libQuEST.so        the drop-in: QuEST v4.1.0's unmodified host layers
#                                                        (api/ core/ cpu/, compiled from $(REF) where they lie)
#                                                        + our shim defining gpu_* / comm_* (quest_b200/shim)
#   make oracle    -> oracle/_ref/libQuEST.so           the unmodified reference CPU/OpenMP build (parity oracle)
#   make           -> all of the above
#
# $(REF) only exists in the build container. On the GPU box the prebuilt .so files (git-ignored, shipped
# by gpurun) are used as they are; `make quest` / `make oracle` are no-ops there.

REF      ?= /root/reference
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
LIBDIR   := quest_b200/lib
BUILD    := build

NVFLAGS  := $(ARCH) -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-fvisibility=default -Iinclude
NCCL_INC ?= /usr/include
NCCL_LIB ?= /usr/lib/x86_64-linux-gnu

CU_SRCS  := $(wildcard quest_b200/csrc/*.cu)
CU_OBJS  := $(patsubst quest_b200/csrc/%.cu,$(BUILD)/csrc/%.o,$(CU_SRCS))
CU_HDRS  := $(wildcard quest_b200/csrc/*.cuh) include/quest_b200.h

SHIM_SRCS := $(wildcard quest_b200/shim/*.cpp)
SHIM_OBJS := $(patsubst quest_b200/shim/%.cpp,$(BUILD)/shim/%.o,$(SHIM_SRCS))
SHIM_DEFS := -DFLOAT_PRECISION=2 -DCOMPILE_OPENMP=1 -DCOMPILE_MPI=1 -DCOMPILE_CUDA=1 -DCOMPILE_CUQUANTUM=0
SHIM_FLAGS := -std=c++17 -O2 -fPIC -fopenmp -Wno-unknown-pragmas -I$(REF) -Iinclude $(SHIM_DEFS)

HOSTOBJ_DIR := oracle/_ref/hostobj
HAVE_REF := $(wildcard $(REF)/quest/src/api/qureg.cpp)
# comm_* and localiser_* come from quest_b200/shim; the reference objects for those files are NOT linked
EXTRA_REF_OBJS ?=
LOCALISER_FILTER ?= ! -name localiser.o

.PHONY: all kernels quest oracle clean
all: kernels oracle quest

kernels: $(LIBDIR)/libquest_b200.so

$(BUILD)/csrc/%.o: quest_b200/csrc/%.cu $(CU_HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -I$(NCCL_INC) -c $< -o $@

$(LIBDIR)/libquest_b200.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -L$(NCCL_LIB) -lnccl

oracle:
	$(MAKE) -C oracle

ifeq ($(HAVE_REF),)
quest:
	@echo "quest: $(REF) not present; using prebuilt $(LIBDIR)/libQuEST.so (if any)"
else
quest: $(LIBDIR)/libQuEST.so

# host-layer objects of the reference are shared with the oracle build (oracle/Makefile explains why that is sound)
HOST_OBJS = $(filter-out %/core/localiser.o,$(shell find $(HOSTOBJ_DIR) -name '*.o' 2>/dev/null))

$(BUILD)/shim/%.o: quest_b200/shim/%.cpp include/quest_b200.h $(wildcard quest_b200/shim/*.hpp)
	@mkdir -p $(dir $@)
	$(CXX) $(SHIM_FLAGS) -c $< -o $@

$(LIBDIR)/libQuEST.so: $(SHIM_OBJS) $(LIBDIR)/libquest_b200.so | oracle
	$(MAKE) -C oracle hostobjs
	$(CXX) -shared -fopenmp -o $@ $(SHIM_OBJS) $$(find $(HOSTOBJ_DIR) -name '*.o' $(LOCALISER_FILTER)) $(EXTRA_REF_OBJS) \
	    -L$(LIBDIR) -lquest_b200 -Wl,-rpath,'$$ORIGIN'
endif

clean:
	rm -rf $(BUILD) $(LIBDIR)/*.so
	$(MAKE) -C oracle clean
