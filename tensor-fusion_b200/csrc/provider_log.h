// provider_log.h -- log sink shared by the provider translation units.
#pragma once
#include <string>
namespace tfprov {
// level: "DEBUG" | "INFO" | "WARN" | "ERROR" (never "FATAL": klog.Fatal would kill
// the hypervisor, pkg/hypervisor/device/accelerator_unix.go:147-148)
void log(const char* level, const char* msg);
// base directory of the quota files as given to LimiterInit ("" before that call)
std::string limiter_base();
// limits of this process's own quota file (TF_SHM_PATH) for one device; false when the limiter is
// not configured or the device is not part of the pod
bool self_limits(const char* uuid, uint64_t* mem_limit, uint64_t* mem_used, uint32_t* up_limit);
// fast path of the LD_PRELOAD limiter: resolve once, then charge with one lock-free FetchSub per launch.
// self_bucket() returns an opaque handle (nullptr = not limited); self_charge() == CheckAndRecordComputeOps' arithmetic.
void* self_bucket(const char* uuid, int* device_index);
double self_charge(void* bucket, int device_index, double cost);
double clamp_cost(void* bucket, int device_index, double cost);  // min(cost, the bucket's capacity)
}
