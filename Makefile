# fnft_b200 -- builds the drop-in shared library fnft_b200/lib/libfnft_b200.so
#   host side : C (gcc), fnft_b200/csrc/host/*.c
#   device    : CUDA for sm_100a (nvcc), fnft_b200/csrc/cuda/*.cu
# `make` builds the product; `make oracle` builds the test-only checkers
# (oracle/_ref needs /root/reference and is skipped when that tree is absent).
NVCC      ?= nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC
CFLAGS    := -std=gnu11 -O2 -fPIC -Wall -Wextra -Iinclude
BUILD     := build
LIBDIR    := fnft_b200/lib
LIB       := $(LIBDIR)/libfnft_b200.so

HOST_SRC  := $(wildcard fnft_b200/csrc/host/*.c)
HOST_OBJ  := $(patsubst fnft_b200/csrc/host/%.c,$(BUILD)/host_%.o,$(HOST_SRC))
CUDA_SRC  := $(wildcard fnft_b200/csrc/cuda/*.cu)
CUDA_OBJ  := $(patsubst fnft_b200/csrc/cuda/%.cu,$(BUILD)/cuda_%.o,$(CUDA_SRC))
CUDA_HDR  := $(wildcard fnft_b200/csrc/cuda/*.cuh) $(wildcard fnft_b200/csrc/cuda/*.h)
HOST_HDR  := $(wildcard fnft_b200/csrc/host/*.h) include/fnft_b200.h fnft_b200/csrc/cuda/fnftb_device.h

.PHONY: all clean oracle emul
all: $(LIB)

$(BUILD)/host_%.o: fnft_b200/csrc/host/%.c $(HOST_HDR)
	@mkdir -p $(BUILD)
	$(CC) $(CFLAGS) -c $< -o $@

# dependencies of every CUDA object come from nvcc -MMD (build/cuda_*.d), so touching one header only rebuilds
# the translation units that include it
$(BUILD)/cuda_%.o: fnft_b200/csrc/cuda/%.cu
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -MMD -MF $(BUILD)/cuda_$*.d -c $< -o $@
-include $(wildcard $(BUILD)/cuda_*.d)

$(LIB): $(HOST_OBJ) $(CUDA_OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -Xlinker -soname=libfnft.so.0 -o $@ $^ -lm
	@ln -sf libfnft_b200.so $(LIBDIR)/libfnft.so.0 && ln -sf libfnft_b200.so $(LIBDIR)/libfnft.so

emul:
	g++ -O2 -std=c++17 -DFNFTB_EMUL -shared -fPIC -o tests/emul/libfnftb_emul.so tests/emul/emul_lib.cpp

oracle:
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref; else echo "no /root/reference: keeping prebuilt oracle/_ref"; fi

clean:
	rm -rf $(BUILD) $(LIB)
