# Top-level build: libsq.so (CUDA, sm_100a), the drop-in host binary ./tauhost.o, the oracle.
# `make` here is what __graft_entry__.build() runs.  CC is pinned: the environment exports a
# CC without libgomp.
CC    := gcc
NVCC  ?= nvcc
ARCH  := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
CSRC  := stochquant_b200/csrc
OBJS  := $(CSRC)/sq_api.o $(CSRC)/sq_compat1d.o $(CSRC)/sq_lattice.o $(CSRC)/sq_rowres.o $(CSRC)/sq_march.o $(CSRC)/sq_tile.o $(CSRC)/sq_slab.o $(CSRC)/sq_session.o
HDRS  := include/sq.h $(CSRC)/sq_kernels.h $(CSRC)/sq_lcg.cuh $(CSRC)/sq_noise.cuh $(CSRC)/sq_site.cuh $(CSRC)/sq_ctx.h $(CSRC)/sq_session.h $(CSRC)/sq_lattice_common.cuh $(CSRC)/sq_pair.cuh $(CSRC)/sq_strip_slow.cuh

all: stochquant_b200/libsq.so tauhost.o oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

stochquant_b200/libsq.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static -lrt

# the drop-in executable keeps the reference's name (README.md:8 of the reference)
tauhost.o: host/tauhost.c host/tauhost_io.c host/tauhost_io.h include/sq.h stochquant_b200/libsq.so
	$(CC) -O2 -Wall -Iinclude -o $@ host/tauhost.c host/tauhost_io.c -Lstochquant_b200 -lsq -lm \
	    -Wl,-rpath,'$$ORIGIN/stochquant_b200' -Wl,-rpath,'$$ORIGIN'

host/libtauhost_io.so: host/tauhost_io.c host/tauhost_io.h
	$(CC) -O2 -Wall -fPIC -shared -o $@ host/tauhost_io.c -lm

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log stochquant_b200/libsq.so tauhost.o host/*.so
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
