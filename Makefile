# Builds the C-ABI shared library of the B200 hot path (in-tree, so that it travels with gpurun snapshots)
# and the checker artefacts under oracle/. sm_100a only.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
PKG       := gnn-recsys_b200
SRCS      := $(wildcard $(PKG)/csrc/*.cu)
OBJS      := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRCS))
LIB       := $(PKG)/libgnn_recsys_b200.so

all: $(LIB)

HDRS      := $(wildcard $(PKG)/csrc/*.cuh) include/gnn_recsys_b200.h

build/%.o: $(PKG)/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -rf build $(LIB)

.PHONY: all clean
