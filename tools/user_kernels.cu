// user_kernels.cu -- the "application code" of the remote-mode tests: integer kernels a client ships to the
// worker as a code image (cubin / PTX / fatbin built by the Makefile), launched through libcuda_remote.so
// with TFCS_OP_MODULE_LOAD / LAUNCH_USER.  Integer arithmetic only, so results are bit-exact by construction.
#include <stdint.h>

// y[i] = a * x[i] + y[i]  (mod 2^32)
extern "C" __global__ void saxpy_u32(const uint32_t* x, uint32_t* y, uint32_t a, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i] + y[i];
}

// Pointers travel inside a by-value struct (what PyTorch's TensorIterator kernels do): the worker must find and
// translate them there too.
struct VecArgs {
  uint32_t n;
  uint32_t bias;
  const uint32_t* a;
  const uint32_t* b;
  uint32_t* out;
};
extern "C" __global__ void vec_add_struct(VecArgs v) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) v.out[i] = v.a[i] + v.b[i] + v.bias;
}

// which SMs run this grid: smid of every block (hard compute isolation test)
extern "C" __global__ void where_am_i(uint32_t* smids, uint64_t spin_ns) {
  uint32_t id;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < spin_ns);
  if (threadIdx.x == 0) smids[blockIdx.x] = id;
}
