# Builds libfestim_b200.so (sm_100a) in-tree and the oracle helpers.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v
SRC := festim_b200/csrc/api.cu festim_b200/csrc/assemble.cu festim_b200/csrc/linalg.cu festim_b200/csrc/newton.cu
HDR := $(wildcard festim_b200/csrc/*.cuh) include/festim_b200.h
LIB := festim_b200/libfestim_b200.so

all: $(LIB)

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) -ldl 2> build_ptxas.log || (cat build_ptxas.log; false)

clean:
	rm -f $(LIB) build_ptxas.log
