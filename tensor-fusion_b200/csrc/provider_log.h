// provider_log.h -- log sink shared by the provider translation units.
#pragma once
#include <string>
namespace tfprov {
// level: "DEBUG" | "INFO" | "WARN" | "ERROR" (never "FATAL": klog.Fatal would kill
// the hypervisor, pkg/hypervisor/device/accelerator_unix.go:147-148)
void log(const char* level, const char* msg);
// base directory of the quota files as given to LimiterInit ("" before that call)
std::string limiter_base();
}
