# Top-level build of libbooster_b200.so (the product) — sm_100a only, in-tree so that it travels to the GPU box.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC      := booster_b200/csrc
OBJDIR    := build
LIB       := booster_b200/libbooster_b200.so

OBJS := $(OBJDIR)/engine.o $(OBJDIR)/gguf.o $(OBJDIR)/bridge.o $(OBJDIR)/tokenizer.o $(OBJDIR)/janus.o

all: $(LIB)

$(OBJDIR)/engine.o: $(CSRC)/engine.cu $(CSRC)/kernels.cuh $(CSRC)/token_kernel.cuh $(CSRC)/prefill.cuh $(CSRC)/prefill_mma.cuh $(CSRC)/prefill_umma.cuh $(CSRC)/gguf.hpp $(CSRC)/tokenizer.hpp include/booster_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/engine.ptxas.log || (cat $(OBJDIR)/engine.ptxas.log; false)

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(CSRC)/gguf.hpp $(CSRC)/tokenizer.hpp $(CSRC)/janus.hpp $(CSRC)/unicode_tables.hpp include/bridge.h include/booster_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(ARCH) -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-O3 -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -ldl

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(OBJDIR) $(LIB)

.PHONY: all oracle clean
