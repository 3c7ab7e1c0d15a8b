// shm_quota.cc -- see shm_quota.h.  Plain C++17 + POSIX, no CUDA.
#include "shm_quota.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <cstring>
#include <limits>

namespace tfq {

namespace {
inline uint64_t bits_of(double v) { uint64_t b; std::memcpy(&b, &v, 8); return b; }
inline double f64_of(uint64_t b) { double v; std::memcpy(&v, &b, 8); return v; }
inline uint64_t aload(const uint64_t* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
inline void astore(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
inline bool acas(uint64_t* p, uint64_t expect, uint64_t want) {
  return __atomic_compare_exchange_n(p, &expect, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
}
bool mkdir_p(const std::string& dir) {
  std::string cur;
  size_t i = 0;
  while (i <= dir.size()) {
    size_t j = dir.find('/', i);
    if (j == std::string::npos) j = dir.size();
    cur = dir.substr(0, j);
    if (!cur.empty() && mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) return false;
    i = j + 1;
  }
  return true;
}
bool alive(uint64_t pid) { return pid != 0 && kill((pid_t)pid, 0) == 0; }
std::string join(const std::string& a, const std::string& b) {
  if (a.empty()) return b;
  return a.back() == '/' ? a + b : a + "/" + b;
}
uint64_t bit_mask(uint32_t i) { return (uint64_t(1) << 63) >> (i & 63u); }
}  // namespace

double go_max(double a, double b) {
  if (a != a || b != b) return f64_of(0x7FF8000000000001ull);
  if (a == 0.0 && b == 0.0) return (bits_of(a) >> 63) ? b : a;
  return a > b ? a : b;
}
double go_min(double a, double b) {
  if (a != a || b != b) return f64_of(0x7FF8000000000001ull);
  if (a == 0.0 && b == 0.0) return (bits_of(a) >> 63) ? a : b;
  return a < b ? a : b;
}

bool valid_component(const std::string& s) {
  return !s.empty() && s.find('/') == std::string::npos && s.find('\\') == std::string::npos &&
         s.find("..") == std::string::npos;
}

Status pod_from_shm_path(const std::string& path, std::string* ns, std::string* name) {
  // filepath.Clean + split, keeping non-empty components; "." and ".." are resolved like Clean does
  std::vector<std::string> comp;
  size_t i = 0;
  while (i <= path.size()) {
    size_t j = path.find('/', i);
    if (j == std::string::npos) j = path.size();
    std::string c = path.substr(i, j - i);
    if (c == "..") { if (!comp.empty() && comp.back() != "..") comp.pop_back(); else if (path.empty() || path[0] != '/') comp.push_back(c); }
    else if (!c.empty() && c != ".") comp.push_back(c);
    i = j + 1;
  }
  if (comp.size() < 3) return kInvalid;
  if (comp.back() != TF_SHM_FILE_NAME) return kInvalid;
  if (ns) *ns = comp[comp.size() - 3];
  if (name) *name = comp[comp.size() - 2];
  return kOk;
}

static Status map_file(int fd, tf_shm_file** out) {
  void* p = mmap(nullptr, TF_SHM_FILE_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) return kIo;
  *out = static_cast<tf_shm_file*>(p);
  return kOk;
}

Status QuotaFile::create(const std::string& base, const std::string& ns, const std::string& pod,
                         const std::vector<DeviceConfig>& cfgs, QuotaFile** out, std::string* err) {
  auto e = [&](Status s, const char* m) { if (err) *err = m; return s; };
  if (!out) return kInvalid;
  *out = nullptr;
  if (ns.empty() || pod.empty()) return e(kInvalid, "pod identifier must include namespace and name");
  if (!valid_component(ns)) return e(kInvalid, "invalid namespace path component");
  if (!valid_component(pod)) return e(kInvalid, "invalid pod name path component");
  for (const auto& c : cfgs)
    if (c.device_idx >= TF_SHM_MAX_DEVICES) return e(kInvalid, "device index exceeds maximum devices");
  const std::string dir = join(join(base, ns), pod);
  if (!mkdir_p(dir)) return e(kIo, "failed to create directory");
  const std::string file = join(dir, TF_SHM_FILE_NAME);
  int fd = ::open(file.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
  if (fd < 0) return e(kIo, "failed to create file");
  if (ftruncate(fd, TF_SHM_FILE_BYTES) != 0) { ::close(fd); return e(kIo, "failed to truncate file"); }
  tf_shm_file* f = nullptr;
  if (map_file(fd, &f) != kOk) { ::close(fd); return e(kIo, "failed to mmap"); }
  const uint64_t now = (uint64_t)time(nullptr);
  std::memset(f, 0, TF_SHM_FILE_BYTES);
  f->device_count = (uint32_t)cfgs.size();
  f->last_heartbeat = now;
  f->pids.creator_pid = (uint64_t)getpid();
  for (const auto& c : cfgs) {
    tf_shm_device_entry& d = f->devices[c.device_idx];
    std::memset(d.uuid, 0, sizeof(d.uuid));
    std::memcpy(d.uuid, c.uuid.data(), c.uuid.size() < TF_SHM_UUID_LEN - 1 ? c.uuid.size() : TF_SHM_UUID_LEN - 1);
    d.total_cuda_cores = c.total_cuda_cores;
    d.up_limit = c.up_limit;
    d.mem_limit = c.mem_limit;
    d.erl_token_capacity = bits_of(100.0);
    d.erl_token_refill_rate = bits_of(10.0);
    d.erl_current_tokens = bits_of(100.0);
    d.erl_last_token_update = bits_of((double)now);
    __atomic_store_n(&d.is_active, 1u, __ATOMIC_SEQ_CST);
  }
  __atomic_store_n(&f->discriminant, TF_SHM_DISCRIMINANT_V2, __ATOMIC_SEQ_CST);
  QuotaFile* q = new QuotaFile();
  q->f_ = f; q->fd_ = fd; q->path_ = file; q->self_pid_ = (uint64_t)getpid();
  *out = q;
  return kOk;
}

Status QuotaFile::open_file(const std::string& file, QuotaFile** out, std::string* err) {
  auto e = [&](Status s, const char* m) { if (err) *err = m; return s; };
  if (!out) return kInvalid;
  *out = nullptr;
  int fd = ::open(file.c_str(), O_RDWR, 0666);
  if (fd < 0) return e(errno == ENOENT ? kNotFound : kIo, "failed to open file");
  struct stat st{};
  if (fstat(fd, &st) != 0) { ::close(fd); return e(kIo, "failed to stat file"); }
  if ((uint64_t)st.st_size != TF_SHM_FILE_BYTES) {
    ::close(fd);
    if ((uint64_t)st.st_size == TF_SHM_LEGACY_BYTES) return e(kLegacyLayout, "legacy shared memory layout detected");
    return e(kBadSize, "unexpected shared memory size");
  }
  tf_shm_file* f = nullptr;
  if (map_file(fd, &f) != kOk) { ::close(fd); return e(kIo, "failed to mmap"); }
  if (f->discriminant != TF_SHM_DISCRIMINANT_V2) {
    munmap(f, TF_SHM_FILE_BYTES);
    ::close(fd);
    return e(kBadDiscriminant, "unsupported shared memory discriminant");
  }
  QuotaFile* q = new QuotaFile();
  q->f_ = f; q->fd_ = fd; q->path_ = file; q->self_pid_ = (uint64_t)getpid();
  *out = q;
  return kOk;
}

Status QuotaFile::open(const std::string& base, const std::string& ns, const std::string& pod, QuotaFile** out,
                       std::string* err) {
  auto e = [&](Status s, const char* m) { if (err) *err = m; return s; };
  if (ns.empty() || pod.empty()) return e(kInvalid, "pod identifier must include namespace and name");
  if (!valid_component(ns)) return e(kInvalid, "invalid namespace path component");
  if (!valid_component(pod)) return e(kInvalid, "invalid pod name path component");
  return open_file(join(join(join(base, ns), pod), TF_SHM_FILE_NAME), out, err);
}

QuotaFile::~QuotaFile() {
  if (f_) munmap(f_, TF_SHM_FILE_BYTES);
  if (fd_ >= 0) ::close(fd_);
}

bool QuotaFile::has_device(uint32_t idx) const {
  return idx < TF_SHM_MAX_DEVICES && __atomic_load_n(&f_->devices[idx].is_active, __ATOMIC_SEQ_CST) != 0;
}
uint32_t QuotaFile::device_count() const { return __atomic_load_n(&f_->device_count, __ATOMIC_SEQ_CST); }

double QuotaFile::rate(uint32_t i) const { return f64_of(aload(&f_->devices[i].erl_token_refill_rate)); }
double QuotaFile::capacity(uint32_t i) const { return f64_of(aload(&f_->devices[i].erl_token_capacity)); }
double QuotaFile::tokens(uint32_t i) const { return f64_of(aload(&f_->devices[i].erl_current_tokens)); }
double QuotaFile::last_update(uint32_t i) const { return f64_of(aload(&f_->devices[i].erl_last_token_update)); }
void QuotaFile::set_rate(uint32_t i, double v) { astore(&f_->devices[i].erl_token_refill_rate, bits_of(v)); }
void QuotaFile::set_capacity(uint32_t i, double v) { astore(&f_->devices[i].erl_token_capacity, bits_of(v)); }
void QuotaFile::set_tokens(uint32_t i, double v) { astore(&f_->devices[i].erl_current_tokens, bits_of(v)); }
void QuotaFile::set_last_update(uint32_t i, double v) { astore(&f_->devices[i].erl_last_token_update, bits_of(v)); }

double QuotaFile::fetch_sub(uint32_t i, double cost) {
  uint64_t* w = &f_->devices[i].erl_current_tokens;
  for (;;) {
    const uint64_t cb = aload(w);
    const double cur = f64_of(cb);
    if (cur < cost) return cur;
    if (acas(w, cb, bits_of(go_max(0.0, cur - cost)))) return cur;
  }
}

double QuotaFile::fetch_add(uint32_t i, double amount) {
  const double cap = capacity(i);
  uint64_t* w = &f_->devices[i].erl_current_tokens;
  for (;;) {
    const uint64_t cb = aload(w);
    const double cur = f64_of(cb);
    if (acas(w, cb, bits_of(go_max(0.0, go_min(cap, cur + amount))))) return cur;
  }
}

double QuotaFile::take_up_to(uint32_t i, double want) {
  if (!(want > 0.0)) return 0.0;
  uint64_t* w = &f_->devices[i].erl_current_tokens;
  for (;;) {
    const uint64_t cb = aload(w);
    const double cur = f64_of(cb);
    if (!(cur > 0.0)) return 0.0;
    const double take = cur < want ? cur : want;
    if (acas(w, cb, bits_of(go_max(0.0, cur - take)))) return take;
  }
}

void QuotaFile::give_back(uint32_t i, double amount) {
  if (!(amount > 0.0)) return;
  uint64_t* w = &f_->devices[i].erl_current_tokens;
  for (;;) {
    const uint64_t cb = aload(w);
    if (acas(w, cb, bits_of(go_max(0.0, f64_of(cb) + amount)))) return;
  }
}

void QuotaFile::update_heartbeat(uint64_t s) { astore(&f_->last_heartbeat, s); }
uint64_t QuotaFile::last_heartbeat() const { return aload(&f_->last_heartbeat); }
bool QuotaFile::is_healthy(uint64_t timeout, uint64_t now) const {
  const uint64_t hb = last_heartbeat();
  if (hb == 0 || hb > now) return false;
  return now - hb <= timeout;
}
bool QuotaFile::set_pod_memory_used(uint32_t i, uint64_t bytes) {
  if (!has_device(i)) return false;
  astore(&f_->devices[i].pod_memory_used, bytes);
  return true;
}
uint64_t QuotaFile::pod_memory_used(uint32_t i) const { return has_device(i) ? aload(&f_->devices[i].pod_memory_used) : 0; }

void QuotaFile::lock() {
  uint64_t* l = &f_->pids.lock;
  for (;;) {
    if (acas(l, 0, self_pid_)) return;
    const uint64_t holder = aload(l);
    if (holder != 0 && holder != self_pid_ && !alive(holder)) { acas(l, holder, 0); continue; }
    sched_yield();
  }
}
void QuotaFile::unlock() { acas(&f_->pids.lock, self_pid_, 0); }
void QuotaFile::cleanup_orphaned_lock() {
  const uint64_t holder = aload(&f_->pids.lock);
  if (holder != 0 && !alive(holder)) acas(&f_->pids.lock, holder, 0);
}

bool QuotaFile::add_pid(uint64_t pid) {
  lock();
  tf_shm_pid_registry& r = f_->pids;
  bool ok = false;
  bool present = false;
  for (uint32_t i = 0; i < TF_SHM_MAX_PROCESSES; ++i)
    if ((r.bitmap[i / 64] & bit_mask(i)) && r.values[i] == pid) { present = true; break; }
  if (!present && r.len < TF_SHM_MAX_PROCESSES) {
    for (uint32_t i = 0; i < TF_SHM_MAX_PROCESSES; ++i) {
      if (r.bitmap[i / 64] & bit_mask(i)) continue;
      r.values[i] = pid;
      r.bitmap[i / 64] |= bit_mask(i);
      r.len++;
      ok = true;
      break;
    }
  }
  unlock();
  return ok;
}

bool QuotaFile::remove_pid(uint64_t pid) {
  lock();
  tf_shm_pid_registry& r = f_->pids;
  bool ok = false;
  for (uint32_t i = 0; i < TF_SHM_MAX_PROCESSES; ++i) {
    if ((r.bitmap[i / 64] & bit_mask(i)) && r.values[i] == pid) {
      r.bitmap[i / 64] &= ~bit_mask(i);
      r.values[i] = 0;
      if (r.len > 0) r.len--;
      ok = true;
      break;
    }
  }
  unlock();
  return ok;
}

std::vector<uint64_t> QuotaFile::pids() {
  lock();
  std::vector<uint64_t> v;
  const tf_shm_pid_registry& r = f_->pids;
  for (uint32_t i = 0; i < TF_SHM_MAX_PROCESSES; ++i)
    if (r.bitmap[i / 64] & bit_mask(i)) v.push_back(r.values[i]);
  unlock();
  return v;
}

Status QuotaFile::cleanup(const std::string& stop_at) {
  if (f_) { munmap(f_, TF_SHM_FILE_BYTES); f_ = nullptr; }
  if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
  if (unlink(path_.c_str()) != 0 && errno != ENOENT) return kIo;
  std::string p = path_;
  for (;;) {  // CleanupEmptyParentDirectories (:111-137)
    const size_t k = p.find_last_of('/');
    if (k == std::string::npos || k == 0) break;
    p = p.substr(0, k);
    if (!stop_at.empty() && p == stop_at) break;
    if (rmdir(p.c_str()) != 0) break;  // not empty (or gone): stop
  }
  return kOk;
}

}  // namespace tfq
