// shm_quota.h -- the soft-limiter quota file, product side (C++).
//
// Restates the parts of pkg/hypervisor/worker/state/soft_limiter_shm.go that
// the limiter side needs (open, atomics, PID registry) and the parts the
// hypervisor-facing Limiter* ABI (provider/limiter.h:89-106) needs (create,
// heartbeat, pod memory).  Independent of oracle/shm_oracle.c on purpose: the
// tests diff the two implementations byte for byte.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "tf_shm_layout.h"

namespace tfq {

enum Status { kOk = 0, kInvalid = 1, kNotFound = 2, kIo = 5, kLegacyLayout = 100, kBadSize = 101, kBadDiscriminant = 102 };

struct DeviceConfig {
  uint32_t device_idx = 0;
  std::string uuid;
  uint32_t up_limit = 0;
  uint64_t mem_limit = 0;
  uint32_t total_cuda_cores = 0;
};

// Go math.Max / math.Min (NaN-propagating, signed-zero aware).
double go_max(double a, double b);
double go_min(double a, double b);

// single path component: non-empty, no '/', '\\' or ".." (soft_limiter_shm.go:900-909)
bool valid_component(const std::string& s);
// {base}/{namespace}/{name}/shm  ->  namespace, name   (soft_limiter_shm.go:70-108)
Status pod_from_shm_path(const std::string& path, std::string* ns, std::string* name);

class QuotaFile {
 public:
  // CreateSharedMemoryHandle (soft_limiter_shm.go:891-964): mkdir -p, O_TRUNC, init.
  static Status create(const std::string& base, const std::string& ns, const std::string& pod,
                       const std::vector<DeviceConfig>& cfgs, QuotaFile** out, std::string* err);
  // OpenSharedMemoryHandle (soft_limiter_shm.go:968-1034): size + discriminant checks.
  static Status open(const std::string& base, const std::string& ns, const std::string& pod, QuotaFile** out,
                     std::string* err);
  // Open by full file path (what TF_SHM_PATH names inside the pod, pkg/constants/env.go:133-138).
  static Status open_file(const std::string& file, QuotaFile** out, std::string* err);
  ~QuotaFile();

  tf_shm_file* raw() { return f_; }
  const std::string& path() const { return path_; }

  bool has_device(uint32_t idx) const;   // :501-503
  uint32_t device_count() const;
  // float64-in-u64 accessors (:654-697)
  double rate(uint32_t idx) const;
  double capacity(uint32_t idx) const;
  double tokens(uint32_t idx) const;
  double last_update(uint32_t idx) const;
  void set_rate(uint32_t idx, double v);
  void set_capacity(uint32_t idx, double v);
  void set_tokens(uint32_t idx, double v);
  void set_last_update(uint32_t idx, double v);
  double fetch_sub(uint32_t idx, double cost);    // :715-731, returns the value found
  double fetch_add(uint32_t idx, double amount);  // :734-748
  // Limiter-side helper: take min(current, want) tokens; returns the amount taken.
  double take_up_to(uint32_t idx, double want);
  // Limiter-side helper: return unspent prepaid tokens WITHOUT the capacity cap (the
  // hypervisor's rebalance drains a balance above capacity smoothly, quota_controller.go:363-366).
  void give_back(uint32_t idx, double amount);

  void update_heartbeat(uint64_t unix_secs);      // :511-513
  uint64_t last_heartbeat() const;
  bool is_healthy(uint64_t timeout_secs, uint64_t now_secs) const;  // :521-534
  bool set_pod_memory_used(uint32_t idx, uint64_t bytes);          // :639-652
  uint64_t pod_memory_used(uint32_t idx) const;

  bool add_pid(uint64_t pid);      // InsertIfAbsent under the PID-owned spin lock (:537-541, :791-812)
  bool remove_pid(uint64_t pid);   // :815-829
  std::vector<uint64_t> pids();    // :832-840
  void cleanup_orphaned_lock();    // :873-878

  // Cleanup (:1058-1068): unmap, remove the file, prune empty parents up to stop_at.
  Status cleanup(const std::string& stop_at);

 private:
  QuotaFile() = default;
  void lock();
  void unlock();
  tf_shm_file* f_ = nullptr;
  int fd_ = -1;
  std::string path_;
  uint64_t self_pid_ = 0;
};

}  // namespace tfq
