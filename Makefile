# stencil_b200 build: hand-written sm_100a CUDA, no cmake needed.
#   make            -> stencil_b200/libstencil_b200.so (C ABI + kernels), lib/libstencil.a (C++ API)
#   make oracle     -> oracle/_build/liboracle.so (test infrastructure)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
INC       := -Iinclude -Istencil_b200/csrc

CSRC      := stencil_b200/csrc/box_copy.cu stencil_b200/csrc/jacobi.cu stencil_b200/csrc/capi.cu
COBJ      := $(patsubst stencil_b200/csrc/%.cu,build/csrc/%.o,$(CSRC))
SO        := stencil_b200/libstencil_b200.so

all: $(SO)

build/csrc/%.o: stencil_b200/csrc/%.cu $(wildcard stencil_b200/csrc/*.cuh) include/stencil_b200.h $(wildcard include/stencil/*.hpp)
	@mkdir -p $(dir $@)
	$(NVCC) $(NVCCFLAGS) $(INC) -c $< -o $@

build/numeric.o: src/numeric.cpp include/stencil/numeric.hpp
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) $(INC) -c $< -o $@

$(SO): $(COBJ) build/numeric.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart shared

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(SO) lib

.PHONY: all oracle clean
