# Builds libkgpu.so (the product: sm_100a kernels + C ABI) and the CPU oracle
# (test infrastructure).  `python -c "import __graft_entry__ as g; g.build()"`
# runs exactly this.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
EXTRA ?=
NVCCFLAGS ?= $(EXTRA) -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
CSRC := kubegpu_b200/csrc
LIB := kubegpu_b200/lib/libkgpu.so

all: $(LIB) oracle

$(LIB): $(CSRC)/kgpu.cu $(CSRC)/score_pairs.cuh $(CSRC)/subset_dp_gen.cuh $(CSRC)/multi_device.cc $(CSRC)/multi_device.h include/kgpu.h
	@mkdir -p kubegpu_b200/lib
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/kgpu.cu $(CSRC)/multi_device.cc -ldl

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -rf kubegpu_b200/lib oracle/_build

.PHONY: all oracle clean
