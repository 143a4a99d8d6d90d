# Top-level build of the B200-native QuEST backend.
#
#   make kernels   -> quest_b200/lib/libquest_b200.so   hand-written sm_100a CUDA + the C ABI (include/quest_b200.h)
#   make quest     -> quest_b200/lib/libQuEST.so        the drop-in: QuEST v4.1.0's unmodified host layers
#                                                        (api/ core/ cpu/, compiled from $(REF) into build/hostobj)
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

# The drop-in links the reference's UNMODIFIED host layers (api/ core/ cpu/ -- the caller of the replaced backend),
# compiled here from $(REF) where they lie into $(BUILD)/hostobj with the product's own flags; nothing under oracle/
# (the parity checker) is part of the product build.  core/localiser.cpp, gpu/* and comm/* are NOT compiled: the shim
# replaces them.  (accelerator.cpp dispatches to cpu_* for CPU-deployed Quregs, so the reference's CPU backend is
# necessarily linked; the B200 backend itself has no CPU path -- see DESIGN.md.)
HOSTOBJ_DIR := $(BUILD)/hostobj
HAVE_REF := $(wildcard $(REF)/quest/src/api/qureg.cpp)
HOST_SRCS := $(wildcard $(REF)/quest/src/api/*.cpp) $(filter-out %/localiser.cpp,$(wildcard $(REF)/quest/src/core/*.cpp)) $(wildcard $(REF)/quest/src/cpu/*.cpp)
HOST_OBJS := $(patsubst $(REF)/quest/src/%.cpp,$(HOSTOBJ_DIR)/%.o,$(HOST_SRCS))
HOST_FLAGS := -std=c++17 -O3 -fPIC -fopenmp -Wno-unknown-pragmas -I$(REF) $(SHIM_DEFS)

.PHONY: all kernels selftest quest oracle clean kernels32 quest32 oracle32 timing
all: kernels oracle quest kernels32 oracle32 quest32

kernels: $(LIBDIR)/libquest_b200.so selftest

# test-only twin of the kernel library: the two translation units that hold host-side self-tests are recompiled with
# -DQB_SELFTEST (include/quest_b200_selftest.h); the product library above is built WITHOUT them
selftest: $(LIBDIR)/libquest_b200_selftest.so
SELFTEST_UNITS := qb_tile qb_runtime qb_pauli_group
SELFTEST_OBJS  := $(patsubst %,$(BUILD)/selftest/%.o,$(SELFTEST_UNITS))
$(BUILD)/selftest/%.o: quest_b200/csrc/%.cu $(CU_HDRS) include/quest_b200_selftest.h
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -DQB_SELFTEST -I$(NCCL_INC) -c $< -o $@
$(LIBDIR)/libquest_b200_selftest.so: $(SELFTEST_OBJS) $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(SELFTEST_OBJS) $(filter-out $(patsubst %,$(BUILD)/csrc/%.o,$(SELFTEST_UNITS)),$(CU_OBJS)) -L$(NCCL_LIB) -lnccl

$(BUILD)/csrc/%.o: quest_b200/csrc/%.cu $(CU_HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -I$(NCCL_INC) -c $< -o $@

$(LIBDIR)/libquest_b200.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -L$(NCCL_LIB) -lnccl

oracle:
	$(MAKE) -C oracle

# probe-only twin with cycle accounting compiled into the tile kernel (-DQB_TILE_TIMING): tools/tile_timing_probe.py,
# tools/tile_round_probe.py.  Not part of `all`; never shipped as the product.
timing: $(BUILD)/timing/libquest_b200_timing.so
$(BUILD)/timing/qb_tile.o: quest_b200/csrc/qb_tile.cu $(CU_HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -DQB_TILE_TIMING -I$(NCCL_INC) -c $< -o $@
$(BUILD)/timing/libquest_b200_timing.so: $(BUILD)/timing/qb_tile.o $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(BUILD)/timing/qb_tile.o $(filter-out $(BUILD)/csrc/qb_tile.o,$(CU_OBJS)) -L$(NCCL_LIB) -lnccl

# ---- single precision (QuEST's FLOAT_PRECISION=1, quest/include/precision.h:80-96): the same sources, compiled with
# -DQB_PRECISION=1 (kernels) / -DFLOAT_PRECISION=1 (shim + the reference's host layers), into *_f32 twins of the three
# libraries -- one library per precision, as in the reference
kernels32: $(LIBDIR)/libquest_b200_f32.so
CU_OBJS32 := $(patsubst quest_b200/csrc/%.cu,$(BUILD)/f32/%.o,$(CU_SRCS))
$(BUILD)/f32/%.o: quest_b200/csrc/%.cu $(CU_HDRS)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -DQB_PRECISION=1 -I$(NCCL_INC) -c $< -o $@
$(LIBDIR)/libquest_b200_f32.so: $(CU_OBJS32)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -L$(NCCL_LIB) -lnccl

oracle32:
	$(MAKE) -C oracle PREC=1 OUT=_ref_f32

ifeq ($(HAVE_REF),)
quest:
	@echo "quest: $(REF) not present; using prebuilt $(LIBDIR)/libQuEST.so (if any)"
else
quest: $(LIBDIR)/libQuEST.so

$(HOSTOBJ_DIR)/%.o: $(REF)/quest/src/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(HOST_FLAGS) -c $< -o $@

$(BUILD)/shim/%.o: quest_b200/shim/%.cpp include/quest_b200.h $(wildcard quest_b200/shim/*.hpp)
	@mkdir -p $(dir $@)
	$(CXX) $(SHIM_FLAGS) -c $< -o $@

$(LIBDIR)/libQuEST.so: $(SHIM_OBJS) $(HOST_OBJS) $(LIBDIR)/libquest_b200.so
	$(CXX) -shared -fopenmp -o $@ $(SHIM_OBJS) $(HOST_OBJS) -L$(LIBDIR) -lquest_b200 -Wl,-rpath,'$$ORIGIN'

quest32: $(LIBDIR)/libQuEST_f32.so
SHIM_OBJS32 := $(patsubst quest_b200/shim/%.cpp,$(BUILD)/shim_f32/%.o,$(SHIM_SRCS))
HOST_OBJS32 := $(patsubst $(REF)/quest/src/%.cpp,$(BUILD)/hostobj_f32/%.o,$(HOST_SRCS))
$(BUILD)/hostobj_f32/%.o: $(REF)/quest/src/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(subst FLOAT_PRECISION=2,FLOAT_PRECISION=1,$(HOST_FLAGS)) -c $< -o $@
$(BUILD)/shim_f32/%.o: quest_b200/shim/%.cpp include/quest_b200.h $(wildcard quest_b200/shim/*.hpp)
	@mkdir -p $(dir $@)
	$(CXX) $(subst FLOAT_PRECISION=2,FLOAT_PRECISION=1,$(SHIM_FLAGS)) -c $< -o $@
$(LIBDIR)/libQuEST_f32.so: $(SHIM_OBJS32) $(HOST_OBJS32) $(LIBDIR)/libquest_b200_f32.so
	$(CXX) -shared -fopenmp -o $@ $(SHIM_OBJS32) $(HOST_OBJS32) -L$(LIBDIR) -lquest_b200_f32 -Wl,-rpath,'$$ORIGIN'
endif
ifeq ($(HAVE_REF),)
quest32:
	@echo "quest32: $(REF) not present; using prebuilt $(LIBDIR)/libQuEST_f32.so (if any)"
endif

clean:
	rm -rf $(BUILD) $(LIBDIR)/*.so
	$(MAKE) -C oracle clean
