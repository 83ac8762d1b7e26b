# Builds libsla_b200.so (the product: CUDA for sm_100a + a C host layer) and, for tests only, the
# host-simulator build of the same kernels.  `python -c "import __graft_entry__ as g; g.build()"`
# runs `make all`.
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      ?= g++
CC       ?= gcc
CSRC     := sla_b200/csrc
LIBDIR   := sla_b200/lib
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC -Iinclude -I$(CSRC) $(NVEXTRA)
CUFILES  := $(CSRC)/slab_ctx.cu $(CSRC)/slab_decode.cu $(CSRC)/slab_decode_fused.cu $(CSRC)/slab_encode.cu $(CSRC)/slab_pcm.cu
HDRS     := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: product oracle hostsim cli

product: $(LIBDIR)/libsla_b200.so

$(LIBDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIBDIR)/slab_host.o: $(CSRC)/slab_host.c $(HDRS)
	@mkdir -p $(LIBDIR)
	$(CC) -std=c99 -O2 -fPIC -Wall -Wextra -Iinclude -I$(CSRC) -c -o $@ $<

$(LIBDIR)/libsla_b200.so: $(LIBDIR)/slab_ctx.o $(LIBDIR)/slab_decode.o $(LIBDIR)/slab_decode_fused.o $(LIBDIR)/slab_encode.o $(LIBDIR)/slab_pcm.o $(LIBDIR)/slab_host.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -Xlinker -Bsymbolic -lpthread

HS := tests/hostsim
oracle:
	$(MAKE) -C oracle all

# ---- command-line tool (reference UX, src/main.c) over the PCM / batch entry points ----
CLISRC := sla_b200/cli/sla_b200_cli.c
cli: $(LIBDIR)/sla_b200_cli $(HS)/sla_hostsim_cli

$(LIBDIR)/sla_b200_cli: $(CLISRC) $(LIBDIR)/libsla_b200.so include/sla_b200.h
	$(CC) -std=c99 -O2 -Wall -Wextra -Iinclude -o $@ $(CLISRC) -L$(LIBDIR) -lsla_b200 -Wl,-rpath,'$$ORIGIN'

# tests only: the same tool on the host-simulator build of the kernels
$(HS)/sla_hostsim_cli: $(CLISRC) $(HS)/libsla_hostsim.so include/sla_b200.h
	$(CC) -std=c99 -O2 -Wall -Wextra -Iinclude -o $@ $(CLISRC) -L$(HS) -lsla_hostsim -Wl,-rpath,'$$ORIGIN'

# ---- tests only: the same kernels compiled for the fibre-based host simulator ----
HS := tests/hostsim
hostsim: $(HS)/libsla_hostsim.so

$(HS)/libsla_hostsim.so: $(CUFILES) $(CSRC)/slab_host.c $(HS)/cuda_emul.cpp $(HS)/cuda_emul.h $(HDRS)
	$(CC) -std=c99 -O2 -fPIC -Iinclude -I$(CSRC) -c -o $(HS)/slab_host.o $(CSRC)/slab_host.c
	$(CXX) -std=c++17 -O2 -g -fPIC -ffp-contract=off -DSLAB_EMUL -I$(HS) -Iinclude -I$(CSRC) -Wno-unused-function \
	  -shared -o $@ $(foreach f,$(CUFILES),-x c++ $(f)) -x c++ $(HS)/cuda_emul.cpp -x none $(HS)/slab_host.o -Wl,-Bsymbolic -lpthread

# ---- memory check of the kernel code without a GPU: the simulator build under AddressSanitizer ----
hostsim-asan: $(HS)/libsla_hostsim_asan.so
ASAN_CC  ?= /usr/bin/gcc          # a toolchain that ships libasan
ASAN_CXX ?= /usr/bin/g++
$(HS)/libsla_hostsim_asan.so: $(CUFILES) $(CSRC)/slab_host.c $(HS)/cuda_emul.cpp $(HS)/cuda_emul.h $(HDRS)
	$(ASAN_CC) -std=c99 -O1 -g -fPIC -fsanitize=address -Iinclude -I$(CSRC) -c -o $(HS)/slab_host_asan.o $(CSRC)/slab_host.c
	$(ASAN_CXX) -std=c++17 -O1 -g -fPIC -fsanitize=address -ffp-contract=off -DSLAB_EMUL -I$(HS) -Iinclude -I$(CSRC) -Wno-unused-function \
	  -shared -o $@ $(foreach f,$(CUFILES),-x c++ $(f)) -x c++ $(HS)/cuda_emul.cpp -x none $(HS)/slab_host_asan.o -Wl,-Bsymbolic -lpthread

clean:
	rm -rf $(LIBDIR) $(HS)/*.so $(HS)/*.o $(HS)/sla_hostsim_cli
	$(MAKE) -C oracle clean

.PHONY: all product oracle hostsim hostsim-asan cli clean
