# Builds the C-ABI shared library of hand-written sm_100a kernels (in-tree, so it travels to the GPU box).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC       := $(wildcard lavie_b200/csrc/*.cu)
OBJ       := $(patsubst lavie_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := lavie_b200/liblavie_b200.so

all: $(LIB)

build/%.o: lavie_b200/csrc/%.cu lavie_b200/csrc/common.cuh include/lavie_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)
