# Build of the B200-native `-pt` renderer (libfermat_b200.so + headless CLI) and of the CPU oracle.
# Everything is built in-tree; nothing here needs a GPU (nvcc cross-compiles sm_100a).
CUDA_HOME ?= /usr/local/cuda
NVCC      := $(CUDA_HOME)/bin/nvcc
CXX       := /usr/bin/g++

HOST_DIR  := fermat_b200/csrc/host
KERN_DIR  := fermat_b200/csrc/kernels
BUILD     := build

CXXFLAGS  := -O2 -g -std=c++17 -fPIC -fopenmp -Wall -Wno-unused-function -Wno-sign-compare -I$(CUDA_HOME)/include -Iinclude
# -fmad=false: the kernels state their fused operations explicitly (fmaf) so that the arithmetic the
# parity tests pin is the arithmetic that runs (DESIGN.md "Numerics").
NVFLAGS   := -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
             -Xcompiler -fPIC -Iinclude -I$(HOST_DIR) --expt-relaxed-constexpr -Xptxas -v

HOST_SRC  := $(wildcard $(HOST_DIR)/*.cpp)
HOST_OBJ  := $(filter-out $(BUILD)/main.o,$(patsubst $(HOST_DIR)/%.cpp,$(BUILD)/%.o,$(HOST_SRC)))
KERN_SRC  := $(wildcard $(KERN_DIR)/*.cu)
KERN_OBJ  := $(patsubst $(KERN_DIR)/%.cu,$(BUILD)/%.cu.o,$(KERN_SRC))

LIB       := fermat_b200/libfermat_b200.so
CLI       := fermat_b200/fermat_pt

all: $(LIB) $(CLI) oracle

$(BUILD):
	mkdir -p $(BUILD)

$(BUILD)/%.o: $(HOST_DIR)/%.cpp $(wildcard $(HOST_DIR)/*.h) $(wildcard $(KERN_DIR)/*.h) include/fermat_b200.h | $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(BUILD)/%.cu.o: $(KERN_DIR)/%.cu $(wildcard $(KERN_DIR)/*.cuh) $(wildcard $(KERN_DIR)/*.h) $(wildcard $(HOST_DIR)/*.h) include/fermat_b200.h | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(LIB): $(HOST_OBJ) $(KERN_OBJ)
	$(NVCC) -shared -o $@ $^ -cudart static -Xlinker --no-undefined -ldl -lpthread -lgomp

$(CLI): $(BUILD)/main.o $(LIB)
	$(CXX) -o $@ $(BUILD)/main.o -Lfermat_b200 -lfermat_b200 -Wl,-rpath,'$$ORIGIN' -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) $(LIB) $(CLI)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
