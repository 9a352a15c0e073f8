# b200mpi native build (sm_100a only). `make` builds everything in-tree; the
# built artefacts are git-ignored but travel to the GPU box with gpurun.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Icsrc/include
CXXFLAGS  := -std=c++17 -O2 -fPIC -Wall -Icsrc/include
LIBDIR    := mpi_operator_b200/lib
BINDIR    := mpi_operator_b200/bin

RUNTIME_SRCS := csrc/kernels/collectives.cu csrc/runtime/comm.cc csrc/runtime/rendezvous.cc
RUNTIME_HDRS := csrc/include/b200mpi.h csrc/kernels/device.cuh csrc/kernels/kernels.h csrc/runtime/rendezvous.h

all: $(LIBDIR)/libb200mpi.so native

$(LIBDIR)/libb200mpi.so: $(RUNTIME_SRCS) $(RUNTIME_HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -shared -x cu $(RUNTIME_SRCS) -o $@ -lrt -lpthread

native:
	@true

sass: $(LIBDIR)/libb200mpi.so
	@mkdir -p profiles
	/usr/local/cuda/bin/cuobjdump -sass $(LIBDIR)/libb200mpi.so > profiles/libb200mpi.sass

clean:
	rm -rf $(LIBDIR)/*.so $(BINDIR)/*

.PHONY: all native sass clean
