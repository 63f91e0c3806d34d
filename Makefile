# Top-level build: the product library (sm_100a only), the CPU oracle (test infrastructure)
# and, when /root/reference is present, the reference's own sources compiled against
# stand-in headers into oracle/_ref/ (test infrastructure).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off \
             -Xptxas -v -Iinclude -Igrid_vision_b200/csrc
NCCL      ?= 1
ifeq ($(NCCL),1)
NVCCFLAGS += -DGV_WITH_NCCL
LIBS      += -lnccl
endif

LIB := grid_vision_b200/lib/libgridvision_b200.so
SRC := grid_vision_b200/csrc/gv_api.cu grid_vision_b200/csrc/gv_microbench.cu
HDR := $(wildcard grid_vision_b200/csrc/*.cuh) include/gridvision_b200.h

all: lib oracle

lib: $(LIB)

$(LIB): $(SRC) $(HDR)
	@mkdir -p grid_vision_b200/lib
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRC) $(LIBS) 2> grid_vision_b200/lib/ptxas.log || (cat grid_vision_b200/lib/ptxas.log; exit 1)
	@grep -E "error|warning" grid_vision_b200/lib/ptxas.log | grep -v "Function properties" || true

oracle:
	$(MAKE) -C oracle

ref:
	$(MAKE) -C oracle/ref_build

clean:
	rm -rf grid_vision_b200/lib oracle/_build oracle/_ref

.PHONY: all lib oracle ref clean
