# Build libkoreb200.so (sm_100a only) and nothing else.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread
SRC := kore_b200/csrc/libkoreb200.cu
DEPS := $(wildcard kore_b200/csrc/*.cu kore_b200/csrc/*.cuh kore_b200/csrc/*.hpp include/*.h)
LIB := kore_b200/libkoreb200.so

all: $(LIB)

$(LIB): $(DEPS)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) $(EXTRA)

ptxas: $(DEPS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o /tmp/kb_ptxas.so $(SRC)

# the same library with the in-kernel cycle counters compiled in (tools/dev_factor_timing.py,
# tools/dev_fold_timing.py: run them with KB_LIB_PATH=kore_b200/libkoreb200_timing.so)
timing: $(DEPS)
	$(NVCC) $(NVFLAGS) -DKB_FACTOR_TIMING -DKB_SWEEP_TICKS -shared -o kore_b200/libkoreb200_timing.so $(SRC)

clean:
	rm -f $(LIB)
