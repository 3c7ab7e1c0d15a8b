// gate.h -- internal include: the gate C-ABI plus the bucket layout shared by
// gate.cu and the quota-file bridge.
#pragma once
#include "tfw_gate.h"

namespace tfw {

// Lives in HBM.  `tokens` / `capacity` are float64 bit patterns, exactly like
// erl_current_tokens / erl_token_capacity of the quota file
// (pkg/hypervisor/worker/state/soft_limiter_shm.go:150-161).
struct DevBucket {
  unsigned long long tokens;
  unsigned long long capacity;
  unsigned long long admitted;
  unsigned long long denied;
  unsigned long long blocked;
  unsigned long long wait_ns;
  unsigned long long timeouts;   // take kernels that gave up their bounded spin (spin fallback only)
  unsigned long long max_wait_ns;
  unsigned long long forced;     // gates released by the host watchdog (fail-open); NOT part of the completed-gate count
};

}  // namespace tfw

#include <cuda_runtime.h>
namespace tfw { cudaError_t preload_gate_kernels(); }
