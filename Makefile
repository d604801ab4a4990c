# Builds the product library (sm_100a only) in-tree, plus the test-infrastructure checkers.
NVCC   ?= /usr/local/cuda/bin/nvcc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall -cudart static
CSRC   := gpusimilarity_b200/csrc
LIB    := gpusimilarity_b200/libgpusim_b200.so

all: $(LIB) oracle

$(LIB): $(CSRC)/gsb_api.cu $(CSRC)/gsb_kernels.cuh include/gpusim_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/gsb_api.cu

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /tmp/gsb_api.o $(CSRC)/gsb_api.cu

oracle:
	$(MAKE) -s -C oracle all

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all oracle clean ptxas-info
