# Top-level build: the CUDA library (sm_100a only), the C host, the oracles.
#
#   make            -> graphical-edmd_b200/libedmd_cuda.so + edmd_host + oracle libs
#   make cuda       -> only the CUDA library
#   make host       -> only the C host program (links the CUDA library)
#   make oracle     -> oracle/liboracle.so (+ oracle/_ref when /root/reference exists)
NVCC      ?= /usr/local/cuda/bin/nvcc
PKG        = graphical-edmd_b200
CSRC       = $(PKG)/csrc
ARCH       = -gencode arch=compute_100a,code=sm_100a
# -fmad=false: no implicit multiply-add contraction anywhere in the library; the
# parity-critical arithmetic additionally uses explicit __d*_rn intrinsics.
TILE_ROWS ?= 8
NVCCFLAGS  = -DEDMD_TILE_ROWS=$(TILE_ROWS) $(ARCH) -O3 -lineinfo -fmad=false -std=c++17 -Iinclude -I$(CSRC) \
             -Xcompiler -fPIC,-Wall,-Wno-unused-function
CU_SRCS    = $(CSRC)/edmd_cuda.cu $(CSRC)/cell_index.cu $(CSRC)/predict.cu $(CSRC)/analysis.cu $(CSRC)/halo.cu \
             $(CSRC)/lean_index.cu $(CSRC)/predict_lean.cu $(CSRC)/cell_sweep.cu $(CSRC)/calendar.cu $(CSRC)/analysis_weighted.cu $(CSRC)/analysis_pcf_sorted.cu $(CSRC)/analysis_voronoi.cu $(CSRC)/thermostat.cu $(CSRC)/multi_gpu.cu
CU_OBJS    = $(CU_SRCS:.cu=.o)
LIB        = $(PKG)/libedmd_cuda.so

all: cuda host oracle

cuda: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/edmd_internal.cuh $(CSRC)/rowstage.cuh $(CSRC)/pairmath.cuh $(CSRC)/lean.cuh include/edmd_cuda.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS)

host: $(LIB)
	@if [ -f $(PKG)/host/Makefile ]; then $(MAKE) --no-print-directory -C $(PKG)/host; fi

oracle:
	$(MAKE) --no-print-directory -C oracle all

clean:
	rm -f $(CU_OBJS) $(LIB)
	$(MAKE) --no-print-directory -C oracle clean

.PHONY: all cuda host oracle clean
