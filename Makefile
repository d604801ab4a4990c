# Builds the product library (sm_100a only) in-tree, plus the test-infrastructure checkers.
NVCC   ?= /usr/local/cuda/bin/nvcc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-ffp-contract=off -cudart static
CSRC   := gpusimilarity_b200/csrc
LIB    := gpusimilarity_b200/libgpusim_b200.so

all: $(LIB) adapter oracle

# Two CUDA translation units (they compile in parallel: make -j), host-only sources alongside.
OBJ    := build/obj
KHDRS  := $(CSRC)/gsb_kernels.cuh $(CSRC)/gsb_batch.cuh $(CSRC)/gsb_sliced.cuh $(CSRC)/gsb_sliced_math.h \
          $(CSRC)/gsb_tensor_params.h $(CSRC)/gsb_tensor_math.h $(CSRC)/gsb_internal.h include/gpusim_b200.h
$(OBJ)/gsb_api.o: $(CSRC)/gsb_api.cu $(KHDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $(CSRC)/gsb_api.cu
$(OBJ)/gsb_tensor.o: $(CSRC)/gsb_tensor.cu $(CSRC)/gsb_tensor.cuh $(KHDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $(CSRC)/gsb_tensor.cu
$(OBJ)/%.o: $(CSRC)/%.cpp $(CSRC)/gsb_internal.h include/gpusim_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(LIB): $(OBJ)/gsb_api.o $(OBJ)/gsb_tensor.o $(OBJ)/fsim_reader.o $(OBJ)/gpusim_server.o
	$(NVCC) $(NVFLAGS) -shared -o $@ $^ -lz

# gpusimserver without Qt: same command line as the reference's main.cpp (--cpu_only, --gpu_bitcount)
SERVER := gpusimilarity_b200/gpusimserver_b200
$(SERVER): $(CSRC)/gpusimserver_main.cpp $(LIB)
	g++ -std=c++14 -O2 -Wall -Iinclude -o $@ $(CSRC)/gpusimserver_main.cpp -Lgpusimilarity_b200 -lgpusim_b200 \
	    -Wl,-rpath,'$$ORIGIN'

# gpusim::FingerprintDB adapter (the reference's C++ surface over the C ABI).  Qt5 is not installed
# in this image, so this target compile-checks and tests it against the header-only Qt stand-ins
# in oracle/qt_shims; with real Qt, build it with -I<Qt include dirs> instead.
ADAPTER := gpusimilarity_b200/libgpusim_adapter.so
$(ADAPTER): $(CSRC)/fingerprintdb_adapter.cpp include/gpusim/fingerprintdb_cuda.h include/gpusim/calculation_functors.h \
            oracle/qt_shims/gsb_qt_shim_core.h $(LIB)
	g++ -std=c++14 -O2 -fPIC -Wall -Werror -shared -Iinclude -Ioracle/qt_shims -o $@ $(CSRC)/fingerprintdb_adapter.cpp \
	    -Lgpusimilarity_b200 -lgpusim_b200 -Wl,-rpath,'$$ORIGIN'

tests/cpp/test_adapter: tests/cpp/test_adapter.cpp $(ADAPTER) oracle/qt_shims/gsb_qt_shim_core.h
	g++ -std=c++14 -O2 -Wall -Iinclude -Ioracle/qt_shims -o $@ tests/cpp/test_adapter.cpp \
	    -Lgpusimilarity_b200 -lgpusim_adapter -lgpusim_b200 -Wl,-rpath,'$$ORIGIN/../../gpusimilarity_b200'

# CPU emulation of one warp of the bit-sliced multi-query kernel (layout, transpose, counting)
tests/cpp/test_sliced_math: tests/cpp/test_sliced_math.cpp $(CSRC)/gsb_sliced_math.h
	g++ -std=c++14 -O2 -Wall -Werror -o $@ tests/cpp/test_sliced_math.cpp

# CPU emulation of the operands of the tensor-core multi-query kernel (layouts, dot products, filter)
tests/cpp/test_tensor_math: tests/cpp/test_tensor_math.cpp $(CSRC)/gsb_tensor_math.h $(CSRC)/gsb_sliced_math.h
	g++ -std=c++14 -O2 -Wall -Werror -o $@ tests/cpp/test_tensor_math.cpp

adapter: $(ADAPTER) tests/cpp/test_adapter tests/cpp/test_sliced_math tests/cpp/test_tensor_math $(SERVER)

ptxas-info:
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /tmp/gsb_api.o $(CSRC)/gsb_api.cu
	$(NVCC) $(NVFLAGS) -Xptxas -v -c -o /tmp/gsb_tensor.o $(CSRC)/gsb_tensor.cu

oracle:
	$(MAKE) -s -C oracle all

clean:
	rm -rf $(LIB) $(OBJ)
	$(MAKE) -C oracle clean
.PHONY: all oracle clean ptxas-info adapter
