# Build of the B200-native vGPU worker + provider (sm_100a only, no fallbacks).
# Usage: make            -> all libraries + binaries (cross-compiles without a GPU)
#        make oracle     -> test oracle (C restatements + reference provider build)
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-Wall,-ffp-contract=off --fmad=false -Iinclude -Itensor-fusion_b200/csrc
CXXFLAGS  := -O2 -std=c++17 -fPIC -fvisibility=hidden -Wall -Wextra -ffp-contract=off -Iinclude -Itensor-fusion_b200/csrc
SRC       := tensor-fusion_b200/csrc
OUT       := tensor-fusion_b200/lib
OBJ       := build/obj

WORKER_CU  := $(SRC)/kernels.cu $(SRC)/worker.cu $(SRC)/gate.cu $(SRC)/native_replay.cu $(SRC)/vram.cu
WORKER_CC  := $(SRC)/shm_quota.cc $(SRC)/quota_bridge.cc $(SRC)/tracegen.cc
WORKER_OBJ := $(patsubst $(SRC)/%.cu,$(OBJ)/%.cu.o,$(WORKER_CU)) $(patsubst $(SRC)/%.cc,$(OBJ)/%.cc.o,$(WORKER_CC))

all: $(OUT)/libtfw_b200.so $(OUT)/libaccelerator_b200.so $(OUT)/tensor-fusion-worker $(OUT)/hypervisor_harness $(OUT)/libtfc_client.so \
     $(OUT)/libcuda_limiter.so $(OUT)/libcuda_remote.so build/mock/libcuda.so.1 build/mock/hook_probe build/mock/null_worker build/mock/libnvidia-ml.so.1 build/mock/ring_lock_probe build/mock/transport_lab build/mock/bridge_pacing_sim \
     build/stub/libcuda.so.1 build/mock/cuda_remote_probe build/mock/cuda_api_probe build/mock/cuda_user_probe build/mock/cuda_user_probe_native \
     build/mock/user_kernels.cubin build/mock/user_kernels.ptx build/mock/user_kernels.fatbin

$(OBJ)/%.cu.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/%.cc.o: $(SRC)/%.cc $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(OUT)/libtfw_b200.so: $(WORKER_OBJ)
	@mkdir -p $(OUT)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $^ -lpthread -ldl -lrt

# The provider: plain C++ (NVML is dlopen()ed at run time, no CUDA context, no CUDA link).
# Installed on the node as libaccelerator_nvidia.so (pkg/constants/vendors.go:77-91).
PROVIDER_CC := $(SRC)/provider.cc $(SRC)/limiter_api.cc $(SRC)/shm_quota.cc $(SRC)/erl.cc
$(OUT)/libaccelerator_b200.so: $(PROVIDER_CC) $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OUT)
	$(CXX) $(CXXFLAGS) -I/usr/local/cuda/include -shared -Wl,--exclude-libs,ALL -o $@ $(PROVIDER_CC) -lpthread -ldl
	ln -sf libaccelerator_b200.so $(OUT)/libaccelerator_nvidia.so

# The worker executable the operator starts (`./tensor-fusion-worker -p 8000`).
$(OUT)/tensor-fusion-worker: $(SRC)/worker_main.cc $(SRC)/hv_handshake.h $(OUT)/libtfw_b200.so include/tfw_worker.h
	$(CXX) -O2 -std=c++17 -Wall -Iinclude -I$(SRC) -o $@ $(SRC)/worker_main.cc -L$(OUT) -ltfw_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

# Client side of the TFCS transport (host only, no CUDA): include/tfc_client.h
$(OUT)/libtfc_client.so: $(SRC)/client.cc include/tfc_client.h include/tfw_wire.h include/tfw_shm_ring.h
	@mkdir -p $(OUT)
	$(CXX) $(CXXFLAGS) -shared -Wl,--exclude-libs,ALL -o $@ $(SRC)/client.cc -lpthread

# LD_PRELOAD limiter of local soft mode (/home/app/libcuda_limiter.so, pkg/constants/env.go:123-131): host only,
# no link against libcuda (the real driver is dlopen()ed), libstdc++ linked statically and hidden so that it can
# be preloaded into any process.
LIMITER_CC := $(SRC)/cuda_hook.cc $(SRC)/limiter_api.cc $(SRC)/shm_quota.cc $(SRC)/erl.cc
$(OUT)/libcuda_limiter.so: $(LIMITER_CC) $(SRC)/cuda_hook.map $(wildcard $(SRC)/*.h) $(wildcard include/*.h)
	@mkdir -p $(OUT)
	$(CXX) $(CXXFLAGS) -shared -static-libstdc++ -static-libgcc -Wl,--exclude-libs,ALL \
	    -Wl,--version-script=$(SRC)/cuda_hook.map -o $@ $(LIMITER_CC) -lpthread -ldl

# Client stub of remote mode: CUDA driver-API facade over the TFCS client (install as libcuda.so.1 in /tensor-fusion).
$(OUT)/libcuda_remote.so: $(SRC)/cuda_remote.cc $(SRC)/client.cc $(SRC)/cuda_remote.map $(SRC)/hv_handshake.h $(wildcard include/*.h)
	@mkdir -p $(OUT)
	$(CXX) $(CXXFLAGS) -shared -static-libstdc++ -static-libgcc -Wl,--exclude-libs,ALL \
	    -Wl,--version-script=$(SRC)/cuda_remote.map -o $@ $(SRC)/cuda_remote.cc $(SRC)/client.cc -lpthread -ldl

# CPU test doubles for the limiter: a counting libcuda.so.1 and an "application" that uses it like libcudart does.
build/mock/libcuda.so.1: tools/mock_cuda.c
	@mkdir -p build/mock
	gcc -O2 -fPIC -fvisibility=hidden -shared -Wall -Wextra -o $@ $<
# A driver-API application and the stub directory it finds "libcuda.so.1" in (remote mode: /tensor-fusion).
build/stub/libcuda.so.1: $(OUT)/libcuda_remote.so
	@mkdir -p build/stub
	ln -sf ../../$(OUT)/libcuda_remote.so $@
build/mock/cuda_remote_probe: tools/cuda_remote_probe.c build/stub/libcuda.so.1
	@mkdir -p build/mock
	gcc -O2 -Wall -o $@ $< -Lbuild/stub -l:libcuda.so.1
build/mock/cuda_api_probe: tools/cuda_api_probe.c build/stub/libcuda.so.1
	@mkdir -p build/mock
	gcc -O2 -Wall -o $@ $< -Lbuild/stub -l:libcuda.so.1
# An application that ships its own kernels (user modules through the stub); the same source linked against the
# real driver's stub library is the native comparator (runs only where libcuda.so.1 exists: the GPU box).
build/mock/cuda_user_probe: tools/cuda_user_probe.c build/stub/libcuda.so.1
	@mkdir -p build/mock
	gcc -O2 -Wall -o $@ $< -Lbuild/stub -l:libcuda.so.1
build/mock/cuda_user_probe_native: tools/cuda_user_probe.c
	@mkdir -p build/mock
	gcc -O2 -Wall -o $@ $< -L/usr/local/cuda/lib64/stubs -lcuda
build/mock/user_kernels.cubin: tools/user_kernels.cu
	@mkdir -p build/mock
	$(NVCC) $(ARCH) -lineinfo -O3 -cubin -o $@ $<
build/mock/user_kernels.ptx: tools/user_kernels.cu
	@mkdir -p build/mock
	$(NVCC) -arch=compute_100a -O3 -ptx -o $@ $<
build/mock/user_kernels.fatbin: tools/user_kernels.cu
	@mkdir -p build/mock
	$(NVCC) $(ARCH) -O3 -fatbin -o $@ $<
# A stand-in NVML (prototypes from the real nvml.h) so that the provider's device paths run in CPU tests.
build/mock/libnvidia-ml.so.1: tools/mock_nvml.c
	@mkdir -p build/mock
	gcc -O2 -fPIC -fvisibility=hidden -shared -Wall -Wextra -I/usr/local/cuda/include -o $@ $<
build/mock/ring_lock_probe: tools/ring_lock_probe.c include/tfw_shm_ring.h
	@mkdir -p build/mock
	gcc -O2 -Wall -Iinclude -o $@ $<
build/mock/transport_lab: tools/transport_lab.c $(OUT)/libtfc_client.so include/tfc_client.h
	@mkdir -p build/mock
	gcc -O2 -Wall -Iinclude -o $@ $< -L$(OUT) -ltfc_client -Wl,-rpath,'$$ORIGIN/../../$(OUT)'
build/mock/null_worker: tools/null_worker.c include/tfw_shm_ring.h include/tfw_wire.h
	@mkdir -p build/mock
	gcc -O2 -Wall -Iinclude -o $@ $<
build/mock/bridge_pacing_sim: tools/bridge_pacing_sim.cc $(SRC)/bridge_pacing.h
	@mkdir -p build/mock
	$(CXX) -O2 -std=c++17 -Wall -I$(SRC) -o $@ $<
build/mock/hook_probe: tools/hook_probe.c
	@mkdir -p build/mock
	gcc -O2 -Wall -D_GNU_SOURCE -o $@ $< -ldl

# Compiled stand-in for the Go hypervisor's purego call sequence (tools/hypervisor_harness.c).
$(OUT)/hypervisor_harness: tools/hypervisor_harness.c include/tf_provider_abi.h
	@mkdir -p $(OUT)
	gcc -O2 -std=gnu11 -Wall -o $@ tools/hypervisor_harness.c -ldl

clean:
	rm -rf build $(OUT)

.PHONY: all clean
