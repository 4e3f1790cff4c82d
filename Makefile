# Builds librufus_gpu.so (sm_100a only) in-tree, the CLI shims and the oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  := $(firstword $(wildcard /usr/bin/g++) g++)
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Irufus_b200/csrc -ccbin $(CXX) $(EXTRA_NVFLAGS)
CSRC := $(wildcard rufus_b200/csrc/*.cu)
OBJS := $(CSRC:.cu=.o)
LIB  := rufus_b200/librufus_gpu.so

.PHONY: all lib tools oracle clean
all: lib tools oracle

lib: $(LIB)
rufus_b200/csrc/%.o: rufus_b200/csrc/%.cu $(wildcard rufus_b200/csrc/*.cuh) include/rufus_gpu.h
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -ccbin $(CXX)

tools: lib
	@if [ -f tools/Makefile ]; then $(MAKE) -C tools; fi

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(OBJS) rufus_b200/csrc/*.ptxas.log $(LIB)
	$(MAKE) -C oracle clean
