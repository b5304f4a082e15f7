# viso-b200: native build without Python.
#   make            libviso_b200/libviso_b200.so   (nvcc, sm_100a)
#   make oracle     oracle/libviso_oracle.so       (CPU checker, test infrastructure only)
#   make host-test  build/test_host                (C++ drop-in test of libviso_b200/host; needs a GPU to RUN)
NVCC ?= nvcc
CXX ?= g++
CSRC := libviso_b200/csrc
SRCS := $(CSRC)/detect.cu $(CSRC)/match.cu $(CSRC)/sort_circle.cu $(CSRC)/estimation.cu $(CSRC)/geometry.cu $(CSRC)/capi.cu $(CSRC)/capi_seq.cu
HDRS := $(CSRC)/viso_dev.h $(CSRC)/common.cuh $(CSRC)/introsort.h $(CSRC)/capi_internal.h $(CSRC)/glibc_sincos.h $(CSRC)/glibc_sincostab.inc include/viso_b200.h
# -fmad=false: the FP64 estimation kernels evaluate the reference's expressions with separate multiplies and adds
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -O3 -std=c++17 -Xcompiler -fPIC,-fno-builtin-sin,-fno-builtin-cos -shared

all: libviso_b200/libviso_b200.so

libviso_b200/libviso_b200.so: $(SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) $(SRCS) -o $@

oracle:
	$(MAKE) -C oracle libviso_oracle.so

host-test: libviso_b200/libviso_b200.so oracle
	mkdir -p build
	$(CXX) -std=c++17 -O1 -Wall tests/host/test_host.cpp libviso_b200/host/viso.cpp \
	    -Llibviso_b200 -lviso_b200 -Loracle -lviso_oracle \
	    -Wl,-rpath,$(CURDIR)/libviso_b200 -Wl,-rpath,$(CURDIR)/oracle -o build/test_host

clean:
	rm -f libviso_b200/libviso_b200.so build/test_host
	$(MAKE) -C oracle clean

.PHONY: all oracle host-test clean
