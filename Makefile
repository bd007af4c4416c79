# Builds dformer_b200/libdformer_b200.so (sm_100a only) and the C oracle helpers.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr -Iinclude
SRC := $(wildcard dformer_b200/csrc/*.cu)
OBJ := $(patsubst dformer_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := dformer_b200/libdformer_b200.so

all: $(LIB)

build/%.o: dformer_b200/csrc/%.cu dformer_b200/csrc/common.cuh dformer_b200/csrc/tile_common.cuh dformer_b200/csrc/dfb200_internal.h include/dfb200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart_static -lpthread -ldl -lrt

clean:
	rm -rf build $(LIB)
.PHONY: all clean
