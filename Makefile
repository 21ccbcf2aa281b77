# Builds libdxmc_b200.so (CUDA sm_100a product library) and the CPU oracle (test infrastructure).
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := opendxmc_b200/csrc
LIBDIR := opendxmc_b200/lib
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -fmad=false -ftz=true -prec-div=false -prec-sqrt=false -Xcompiler -fPIC,-Wall,-Wno-unused-function
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall

LIB := $(LIBDIR)/libdxmc_b200.so
OBJS := build/atom.o build/material.o build/tube.o build/beams.o build/capi.o build/transport.o build/transport_mux.o build/transport_pool.o build/context.o build/exchange.o build/icrp.o build/import_kernels.o build/h5mini.o build/h5_capi.o build/scene_io.o

all: $(LIB) oracle ref

$(LIB): $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lz

CSRC_HDRS := $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh)

build/%.o: $(CSRC)/%.cpp $(CSRC_HDRS) include/dxb.h Makefile
	@mkdir -p build
	$(CXX) $(CXXFLAGS) -c $< -o $@

# identity of the shipped transport kernel: its sources + the compiler flags (bench.py only trusts an ncu traffic capture
# whose id equals the one compiled into the loaded library, dxb_kernel_build_id)
KERNEL_ID := $(shell cat $(CSRC)/transport_pool.cu $(CSRC)/transport_common.cuh $(CSRC)/device_types.cuh | sha256sum | cut -c1-16)-$(shell echo '$(NVFLAGS)' | sha256sum | cut -c1-8)

build/%.o: $(CSRC)/%.cu $(CSRC_HDRS) include/dxb.h Makefile
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -DDXB_KERNEL_BUILD_ID='"$(KERNEL_ID)"' -c $< -o $@

oracle: oracle/liboracle.so
oracle/liboracle.so: oracle/oracle.cpp oracle/oracle.h include/dxb.h Makefile
	$(CXX) -O2 -std=c++17 -fPIC -Wall -shared -pthread -o $@ oracle/oracle.cpp

# OpenDXMC's own translation units for this boundary, compiled where they lie (only where the reference tree is mounted)
ref: $(LIB)
	@if [ -f /root/reference/src/libopendxmc/dxmc_specialization.cpp ]; then $(MAKE) --no-print-directory -f oracle/Makefile.ref; fi

clean:
	rm -rf build $(LIB) oracle/liboracle.so oracle/_ref

.PHONY: all oracle ref clean
