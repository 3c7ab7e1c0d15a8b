// worker_ctl.h -- the control channel from the node-level libraries (provider: AccelSnapshot / AccelResume;
// limiter: FreezeWorker / ResumeWorker / AutoFreeze / AutoResume, provider/limiter.h:77-81) to the vGPU workers:
// the control words of the record every worker publishes next to its quota file (include/tfw_stats_file.h).
// The reference leaves this transport open ("send snapshot command to worker via shared memory",
// pkg/hypervisor/server/handlers/worker.go:94-129).  Header-only: shared by libaccelerator_b200.so and
// libcuda_limiter.so.
#pragma once
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tfw_stats_file.h"

namespace tfctl {

// Calls fn(pod_dir, stats_file, record) for every live worker record under <base>/<namespace>/<pod>/.
template <typename Fn>
void for_each_worker_record(const std::string& base, Fn fn) {
  DIR* d1 = opendir(base.c_str());
  if (!d1) return;
  const uint64_t now = (uint64_t)time(nullptr);
  while (dirent* ns = readdir(d1)) {
    if (ns->d_name[0] == '.') continue;
    const std::string nsdir = base + "/" + ns->d_name;
    DIR* d2 = opendir(nsdir.c_str());
    if (!d2) continue;
    while (dirent* pod = readdir(d2)) {
      if (pod->d_name[0] == '.') continue;
      const std::string poddir = nsdir + "/" + pod->d_name;
      const std::string f = poddir + "/" + TFW_STATS_FILE_NAME;
      int fd = ::open(f.c_str(), O_RDONLY);
      if (fd < 0) continue;
      tfw_stats_record r{};
      bool ok = false;
      for (int attempt = 0; attempt < 4 && !ok; ++attempt) {  // seqlock read
        if (pread(fd, &r, sizeof r, 0) != (ssize_t)sizeof r) break;
        uint64_t seq2 = 0;
        ok = !(r.seq & 1) && pread(fd, &seq2, sizeof seq2, offsetof(tfw_stats_record, seq)) == (ssize_t)sizeof seq2 && seq2 == r.seq;
      }
      ::close(fd);
      if (!ok || r.magic != TFW_STATS_MAGIC || r.version != TFW_STATS_VERSION) continue;
      if (now > r.updated_unix_secs + TFW_STATS_STALE_SECS) continue;
      r.device_uuid[sizeof(r.device_uuid) - 1] = 0;
      fn(poddir, f, r);
    }
    closedir(d2);
  }
  closedir(d1);
}


// Does the record under <base>/<ns>/<pod>/ belong to `worker_id`?  The hypervisor names workers by pod UID
// (WorkerUID) -- which the worker learns as $POD_UID / $TF_WORKER_ID -- or "<namespace>/<pod>", or the pod name.
inline bool record_is_worker(const std::string& poddir, const tfw_stats_record& r, const char* worker_id) {
  if (!worker_id || !*worker_id) return false;
  char id[sizeof r.worker_id + 1];
  std::memcpy(id, r.worker_id, sizeof r.worker_id);
  id[sizeof r.worker_id] = 0;
  if (*id && std::strcmp(id, worker_id) == 0) return true;
  const size_t p1 = poddir.find_last_of('/');
  if (p1 == std::string::npos) return false;
  const std::string pod = poddir.substr(p1 + 1);
  if (pod == worker_id) return true;
  const size_t p0 = poddir.find_last_of('/', p1 - 1);
  return p0 != std::string::npos && poddir.substr(p0 + 1) == worker_id;  // "<namespace>/<pod>"
}

// Write `cmd` (TFW_CTL_FREEZE / TFW_CTL_RESUME) into the control word of each record file and wait for the workers'
// acknowledgements ($TF_SNAPSHOT_TIMEOUT_MS, default 30 s).  Returns 0, or a tfw_status-like code: 4 = a worker
// answered "resource exhausted" (resume without HBM), 5 = no acknowledgement / failure.
inline int send_control(const std::vector<std::string>& files, uint32_t cmd, uint64_t arg = 0) {
  struct Pending { tfw_stats_record* rec; uint64_t req; };
  std::vector<Pending> pend;
  for (const std::string& f : files) {
    int fd = ::open(f.c_str(), O_RDWR);
    if (fd < 0) continue;
    void* m = mmap(nullptr, sizeof(tfw_stats_record), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) continue;
    tfw_stats_record* r = static_cast<tfw_stats_record*>(m);
    const uint64_t req = (((__atomic_load_n(&r->ctl_request, __ATOMIC_ACQUIRE) >> 8) + 1) << 8) | cmd;
    r->ctl_arg = arg;
    __atomic_store_n(&r->ctl_request, req, __ATOMIC_RELEASE);
    pend.push_back({r, req});
  }
  if (pend.empty()) return 5;
  long timeout_ms = 30000;
  if (const char* t = getenv("TF_SNAPSHOT_TIMEOUT_MS")) timeout_ms = atol(t) > 0 ? atol(t) : timeout_ms;
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  int out = 0;
  for (Pending& p : pend) {
    bool acked = false;
    for (;;) {
      if (__atomic_load_n(&p.rec->ctl_ack, __ATOMIC_ACQUIRE) == p.req) { acked = true; break; }
      timespec now;
      clock_gettime(CLOCK_MONOTONIC, &now);
      if ((now.tv_sec - t0.tv_sec) * 1000 + (now.tv_nsec - t0.tv_nsec) / 1000000 > timeout_ms) break;
      usleep(500);
    }
    if (!acked) out = 5;
    else if (p.rec->ctl_status != 0 && out == 0) out = p.rec->ctl_status == 4 ? 4 : p.rec->ctl_status == 3 ? 3 : 5;
    munmap(p.rec, sizeof(tfw_stats_record));
  }
  return out;
}

// Every live record that belongs to `worker_id`: (file, record) pairs.
inline std::vector<std::pair<std::string, tfw_stats_record>> find_worker(const std::string& base, const char* worker_id) {
  std::vector<std::pair<std::string, tfw_stats_record>> out;
  for_each_worker_record(base, [&](const std::string& poddir, const std::string& file, const tfw_stats_record& r) {
    if (record_is_worker(poddir, r, worker_id)) out.emplace_back(file, r);
  });
  return out;
}

}  // namespace tfctl
