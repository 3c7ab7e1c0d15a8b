// client.cc -- libtfc_client.so: TFCS client over TCP or the shared-memory rings of
// include/tfw_shm_ring.h (see include/tfc_client.h).  Host only.
#include <arpa/inet.h>
#include <emmintrin.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "tfc_client.h"
#include "tfw_shm_ring.h"
#include "tfw_wire.h"

struct tfc_conn {
  int fd = -1;
  // shared-memory transport (fd < 0)
  tfsr_header* shm = nullptr;
  uint8_t *c2w = nullptr, *w2c = nullptr;
  uint64_t shm_bytes = 0;
  int shm_fd = -1;  // kept open: carries the liveness lock (tfsr_client_lock)
  uint32_t session = 0;  // the worker counts its clients; "closed" flags carry the session they refer to
  uint64_t tx_tail_seen = 0;  // last c2w_tail read: re-read only when the ring looks full
  uint32_t call_id = 0, next_handle = 1;
  int first_err = 0, last_err = 0;
  uint32_t last_err_call = 0;
  std::vector<uint8_t> out;  // coalesces small frames; flushed before a blocking call or when large
  // handle ids are a small space (TFCS_MAX_HANDLES): ids of freed buffers are reused, oldest first
  std::vector<uint32_t> free_handles;
  std::vector<bool> live;                       // indexed by handle
  std::map<uint32_t, uint32_t> pending_malloc;  // call_id -> handle, until a later response proves the MALLOC was accepted
  // page-locked memory shared with the worker: arena k = file <ring file>.a<k>, carved first-fit
  struct ArenaMap {
    uint8_t* base = nullptr;
    uint64_t size = 0;
    std::map<uint64_t, uint64_t> free_;  // offset -> bytes
    std::map<uint64_t, uint64_t> used;   // offset -> bytes
  };
  ArenaMap arenas[TFCS_MAX_ARENAS + 1];
  std::string ring_path;
  uint32_t next_module = 1, next_function = 1;
  int lifeline_fd = -1;        // the TCP connection a session on upgraded rings keeps open
  std::string upgraded_ring;   // ring file this client created for the upgrade (unlinked on close)
};

namespace {

bool send_all(int fd, const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  while (n) {
    ssize_t k = send(fd, b, n, MSG_NOSIGNAL);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    b += k; n -= (size_t)k;
  }
  return true;
}
bool recv_all(int fd, void* p, size_t n) {
  uint8_t* b = static_cast<uint8_t*>(p);
  while (n) {
    ssize_t k = recv(fd, b, n, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) return false;
    b += k; n -= (size_t)k;
  }
  return true;
}
// ---- shared-memory rings ----------------------------------------------------------------
// Waiting: spin briefly (the peer is usually a few microseconds away), then sleep 20 us at a time.
struct Waiter {
  int spins = 0;
  void pause() {
    if (++spins < 2000) { __builtin_ia32_pause(); return; }
    timespec ts{0, 20000};
    nanosleep(&ts, nullptr);
  }
};

// Copy into the client -> worker ring with non-temporal stores: nothing on this side reads those bytes again and
// the consumer of a payload is the GPU's copy engine, so pulling the ring's lines into this core's cache (a
// read-for-ownership per line) only costs bandwidth.  dst is 16-byte aligned by construction (frames are multiples
// of 16 bytes); the caller issues an sfence before publishing the cursor.
inline void copy_nt(uint8_t* dst, const uint8_t* src, size_t n) {
  static const bool off = [] { const char* e = getenv("TFC_NT_STORES"); return e && *e == '0'; }();
  if (off || n < 2048 || (reinterpret_cast<uintptr_t>(dst) & 15u)) { std::memcpy(dst, src, n); return; }
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  if (i < n) std::memcpy(dst + i, src + i, n - i);
  _mm_sfence();
}

// Large copies into / out of the rings are split over a few threads: one core moves 5-10 GB/s, the copy
// engine behind the ring 55 GB/s.  TFC_COPY_THREADS (default min(16, cores/4), 1 = off); pieces below 1 MiB stay on the caller.
class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool* p = new CopyPool();  // leaked on purpose: threads may outlive static destructors
    return *p;
  }
  // nt: destination is the client -> worker ring (see copy_nt)
  void copy(uint8_t* dst, const uint8_t* src, size_t n, bool nt = false) {
    const size_t parts = n >= (1u << 20) ? std::min<size_t>(threads_, n >> 18) : 1;
    if (parts <= 1) { one(dst, src, n, nt); return; }
    const size_t slice = (((n + parts - 1) / parts) + 63) & ~(size_t)63;
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (size_t i = 1; i < parts; ++i) {
        const size_t off = i * slice;
        if (off >= n) break;
        jobs_.push_back({dst + off, src + off, std::min(slice, n - off), nt});
        ++pending_;
      }
    }
    cv_.notify_all();
    one(dst, src, std::min(slice, n), nt);
    // the helpers finish within microseconds of this thread: look a few times before paying for a futex sleep + wake
    for (int spin = 0; spin < 4000; ++spin) {
      { std::lock_guard<std::mutex> lk(mu_); if (pending_ == 0) return; }
      for (int k = 0; k < 16; ++k) __builtin_ia32_pause();
    }
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  struct Job { uint8_t* d; const uint8_t* s; size_t n; bool nt; };
  static void one(uint8_t* d, const uint8_t* s, size_t n, bool nt) {
    if (nt) copy_nt(d, s, n);
    else std::memcpy(d, s, n);
  }
  CopyPool() {
    const char* e = getenv("TFC_COPY_THREADS");
    const long hw = (long)std::thread::hardware_concurrency();
    long t = e && *e ? atol(e) : std::min<long>(16, std::max<long>(2, hw / 4));  // 16 on a GPU server, 2 on a small box
    if (hw > 0 && t > hw) t = hw;
    threads_ = (size_t)std::max<long>(1, t);
    for (size_t i = 1; i < threads_; ++i) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu_);
        // a bulk transfer hands out pieces every few tens of microseconds: look again for a short while before
        // going to sleep, a futex wake costs as much as copying half a megabyte
        for (int spin = 0; jobs_.empty() && spin < 200; ++spin) {
          lk.unlock();
          for (int k = 0; k < 64; ++k) __builtin_ia32_pause();
          lk.lock();
        }
        cv_.wait(lk, [&] { return !jobs_.empty(); });
        j = jobs_.back();
        jobs_.pop_back();
      }
      one(j.d, j.s, j.n, j.nt);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::vector<Job> jobs_;
  size_t pending_ = 0, threads_ = 1;
};

bool shm_write(tfc_conn* c, const void* p, size_t n) {
  tfsr_header* h = c->shm;
  const uint8_t* b = static_cast<const uint8_t*>(p);
  const uint64_t size = h->c2w_size;
  uint64_t head = h->c2w_head;  // only this side writes it
  Waiter w;
  while (n) {
    const uint64_t tail = __atomic_load_n(&h->c2w_tail, __ATOMIC_ACQUIRE);
    const uint64_t free_b = size - (head - tail);
    if (!free_b) {
      if (__atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) >= c->session) return false;
      w.pause();
      continue;
    }
    const uint64_t pos = head % size;
    // publish in pieces of at most 16 MiB so the worker's DMA overlaps the rest of a large copy
    const uint64_t k = std::min<uint64_t>(std::min<uint64_t>(n, free_b), std::min<uint64_t>(size - pos, 16u << 20));
    CopyPool::get().copy(c->c2w + pos, b, k, true);
    head += k;
    // a session the worker has already closed (it failed, or was told to stop) must not publish into the ring any
    // more: the worker hands the ring to the next client with the cursors where it left them
    if (__atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) >= c->session) return false;
    __atomic_store_n(&h->c2w_head, head, __ATOMIC_RELEASE);
    b += k;
    n -= k;
    w.spins = 0;
  }
  return true;
}

bool shm_read(tfc_conn* c, void* p, size_t n) {
  tfsr_header* h = c->shm;
  uint8_t* b = static_cast<uint8_t*>(p);
  const uint64_t size = h->w2c_size;
  uint64_t tail = h->w2c_tail;
  Waiter w;
  while (n) {
    const uint64_t head = __atomic_load_n(&h->w2c_head, __ATOMIC_ACQUIRE);
    const uint64_t avail = head - tail;
    if (!avail) {
      if (__atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) >= c->session && __atomic_load_n(&h->w2c_head, __ATOMIC_ACQUIRE) == tail) return false;
      w.pause();
      continue;
    }
    const uint64_t pos = tail % size;
    const uint64_t k = std::min<uint64_t>(std::min<uint64_t>(n, avail), size - pos);
    CopyPool::get().copy(b, c->w2c + pos, k);
    tail += k;
    __atomic_store_n(&h->w2c_tail, tail, __ATOMIC_RELEASE);
    b += k;
    n -= k;
    w.spins = 0;
  }
  return true;
}

// One small frame = one reservation: a single look at the consumer's cursor (a cached one when it already shows
// enough room), header + payload + padding copied, a single publish.
bool shm_write_frame(tfc_conn* c, const tfcs_frame_hdr& hdr, const void* payload, size_t n) {
  tfsr_header* h = c->shm;
  const uint64_t size = h->c2w_size, padded = tfcs_pad16(n), total = sizeof hdr + padded;
  if (total > (256u << 10) || total > size / 2) {  // large: stream it, the worker starts on the front while we copy the rest
    static const uint8_t zeros[16] = {0};
    return shm_write(c, &hdr, sizeof hdr) && shm_write(c, payload, n) && shm_write(c, zeros, padded - n);
  }
  const uint64_t head = h->c2w_head;
  Waiter w;
  while (size - (head - c->tx_tail_seen) < total) {
    c->tx_tail_seen = __atomic_load_n(&h->c2w_tail, __ATOMIC_ACQUIRE);
    if (size - (head - c->tx_tail_seen) >= total) break;
    if (__atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) >= c->session) return false;
    w.pause();
  }
  uint64_t pos = head % size;
  auto put_bytes = [&](const void* src, uint64_t k) {
    const uint64_t first = std::min<uint64_t>(k, size - pos);
    copy_nt(c->c2w + pos, static_cast<const uint8_t*>(src), first);
    if (k > first) copy_nt(c->c2w, static_cast<const uint8_t*>(src) + first, k - first);
    pos += k;
    if (pos >= size) pos -= size;
  };
  put_bytes(&hdr, sizeof hdr);
  if (n) put_bytes(payload, n);
  if (padded > n) {
    const uint8_t zeros[16] = {0};
    put_bytes(zeros, padded - n);
  }
  if (__atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) >= c->session) return false;  // (see shm_write)
  __atomic_store_n(&h->c2w_head, head + total, __ATOMIC_RELEASE);
  return true;
}

bool tx(tfc_conn* c, const void* p, size_t n) { return c->fd >= 0 ? send_all(c->fd, p, n) : shm_write(c, p, n); }
bool rx(tfc_conn* c, void* p, size_t n) { return c->fd >= 0 ? recv_all(c->fd, p, n) : shm_read(c, p, n); }

bool flush(tfc_conn* c) {
  if (c->out.empty()) return true;
  const bool ok = tx(c, c->out.data(), c->out.size());
  c->out.clear();
  return ok;
}
tfcs_frame_hdr mk(tfc_conn* c, uint16_t op) {
  tfcs_frame_hdr h{};
  h.magic = TFCS_MAGIC; h.version = TFCS_VERSION; h.opcode = op; h.call_id = c->call_id++;
  return h;
}
bool put(tfc_conn* c, const tfcs_frame_hdr& h) {
  if (c->fd < 0) return shm_write_frame(c, h, nullptr, 0);  // no system call to amortise: write through
  const uint8_t* p = reinterpret_cast<const uint8_t*>(&h);
  c->out.insert(c->out.end(), p, p + sizeof h);
  return c->out.size() < (1u << 20) || flush(c);
}
// read responses until the one answering `want_call` (opcode want_op) arrives.  A RESP_D2H payload must be exactly
// `n` bytes; any other payload (RESP_FUNCTION) may be shorter: *got tells its length.
int wait_for(tfc_conn* c, uint32_t want_call, uint16_t want_op, void* payload, uint64_t n, uint64_t* got = nullptr, tfcs_frame_hdr* out_hdr = nullptr) {
  for (;;) {
    tfcs_frame_hdr r;
    if (!rx(c, &r, sizeof r) || r.magic != TFCS_MAGIC) return 7;
    if (r.opcode == TFCS_OP_RESP_ERROR) {
      c->last_err = (int)r.arg0; c->last_err_call = r.call_id;
      const auto pm = c->pending_malloc.find(r.call_id);
      if (pm != c->pending_malloc.end()) {  // the worker refused this MALLOC: the id never became a buffer
        if (pm->second < c->live.size() && c->live[pm->second]) {  // (not already given back by a tfc_free)
          c->live[pm->second] = false;
          c->free_handles.push_back(pm->second);
        }
        c->pending_malloc.erase(pm);
      }
      if (r.call_id == want_call) return (int)r.arg0;  // reported to the caller directly
      if (!c->first_err) c->first_err = (int)r.arg0;     // fire-and-forget call: reported by the next tfc_sync
      continue;
    }
    const uint64_t padded = tfcs_pad16(tfcs_has_payload(r.opcode) ? r.length : 0);
    if (r.call_id == want_call && r.opcode == want_op) {
      // responses arrive in call order: every MALLOC issued before this call has been answered if it failed
      c->pending_malloc.erase(c->pending_malloc.begin(), c->pending_malloc.lower_bound(want_call));
      if (out_hdr) *out_hdr = r;
      if (got) *got = r.length;
      if (padded) {
        if (want_op == TFCS_OP_RESP_D2H ? r.length != n : r.length > n) return 7;
        if (!rx(c, payload, r.length)) return 7;
        uint8_t pad[16];
        if (padded > r.length && !rx(c, pad, padded - r.length)) return 7;
      }
      return 0;
    }
    std::vector<uint8_t> skip(padded);
    if (padded && !rx(c, skip.data(), padded)) return 7;
  }
}

// header + payload (+ padding) of a frame that carries bytes the worker assembles on the host
bool put_with_payload(tfc_conn* c, const tfcs_frame_hdr& h, const void* payload, size_t n) {
  static const uint8_t zeros[16] = {0};
  if (c->fd < 0) return shm_write_frame(c, h, payload, n);
  const uint8_t* p = reinterpret_cast<const uint8_t*>(&h);
  c->out.insert(c->out.end(), p, p + sizeof h);
  if (n >= (256u << 10)) return flush(c) && tx(c, payload, n) && tx(c, zeros, tfcs_pad16(n) - n);
  const uint8_t* b = static_cast<const uint8_t*>(payload);
  c->out.insert(c->out.end(), b, b + n);
  c->out.insert(c->out.end(), zeros, zeros + (tfcs_pad16(n) - n));
  return c->out.size() < (1u << 20) || flush(c);
}

// ---- arenas --------------------------------------------------------------------------------------------------
std::string arena_path(const tfc_conn* c, uint32_t id) { return c->ring_path + ".a" + std::to_string(id); }

// [p, p+n) inside one arena?  -> id and offset
bool find_arena(const tfc_conn* c, const void* p, uint64_t n, uint32_t* id, uint64_t* off) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (uint32_t a = 1; a <= TFCS_MAX_ARENAS; ++a) {
    const tfc_conn::ArenaMap& m = c->arenas[a];
    if (m.base && b >= m.base && b < m.base + m.size && n <= (uint64_t)(m.base + m.size - b)) { *id = a; *off = (uint64_t)(b - m.base); return true; }
  }
  return false;
}

int round_trip(tfc_conn* c) {  // SYNC + wait; errors of the calls before it are left in first_err
  tfcs_frame_hdr h = mk(c, TFCS_OP_SYNC);
  if (!put(c, h) || !flush(c)) return 5;
  return wait_for(c, h.call_id, TFCS_OP_RESP_SYNC, nullptr, 0);
}

void drop_arena(tfc_conn* c, uint32_t id, bool tell_worker) {
  tfc_conn::ArenaMap& m = c->arenas[id];
  if (!m.base) return;
  if (tell_worker) {  // the worker drains the vGPU stream before it unmaps; wait for that before the pages go away
    tfcs_frame_hdr h = mk(c, TFCS_OP_HOST_UNREGISTER);
    h.h0 = id;
    if (put(c, h)) round_trip(c);
  }
  munmap(m.base, m.size);
  unlink(arena_path(c, id).c_str());
  m = tfc_conn::ArenaMap{};
}

}  // namespace

extern "C" {

// "shmem+<name>+<MiB>+<initVersion>": attach to the rings the worker created in /dev/shm (or $TFC_SHM_DIR)
static int connect_shm(const std::string& u, tfc_conn** out) {
  const size_t a = 6, b = u.find('+', a);
  const std::string name = u.substr(a, b == std::string::npos ? std::string::npos : b - a);
  if (name.empty() || name.find('/') != std::string::npos) return 1;
  // fields after the name: size in MiB (the file is authoritative) and the layout's initVersion
  // ("protocol+identifier+size+initVersion", internal/webhook/v1/pod_webhook.go:583)
  if (b != std::string::npos) {
    const size_t c3 = u.find('+', b + 1);
    if (c3 != std::string::npos && c3 + 1 < u.size() && (uint32_t)atoi(u.c_str() + c3 + 1) != TFSR_VERSION) return 3;
  }
  const char* dir = getenv("TFC_SHM_DIR");
  const std::string path = std::string(dir && *dir ? dir : "/dev/shm") + "/" + name;
  long wait_ms = 10000;
  if (const char* w = getenv("TFC_CONNECT_TIMEOUT_MS")) wait_ms = atol(w) > 0 ? atol(w) : wait_ms;
  timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  auto expired = [&] {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (t.tv_sec - t0.tv_sec) * 1000 + (t.tv_nsec - t0.tv_nsec) / 1000000 > wait_ms;
  };
  for (;;) {  // the operator `touch`es the file before the worker has sized it: wait for a ready header
    int fd = open(path.c_str(), O_RDWR);
    struct stat st{};
    if (fd >= 0 && fstat(fd, &st) == 0 && (uint64_t)st.st_size >= TFSR_MIN_BYTES) {
      // MAP_POPULATE: the worker has already page-locked every page; build this process's page tables now
      // instead of one fault per 4 KiB during the first lap of the ring
      void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
      if (m == MAP_FAILED) { close(fd); return 5; }
      tfsr_header* h = static_cast<tfsr_header*>(m);
      for (;;) {
        if (__atomic_load_n(&h->magic, __ATOMIC_ACQUIRE) == TFSR_MAGIC && h->version == TFSR_VERSION &&
            __atomic_load_n(&h->worker_ready, __ATOMIC_ACQUIRE) == 1 && h->total_bytes == (uint64_t)st.st_size) {
          uint32_t expect = 0;
          if (__atomic_compare_exchange_n(&h->client_pid, &expect, (uint32_t)getpid(), false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
            tfc_conn* c = new tfc_conn();
            c->shm = h;
            c->shm_bytes = (uint64_t)st.st_size;
            c->session = h->session;
            c->shm_fd = fd;
            c->ring_path = path;
#ifdef TFSR_HAVE_LIVENESS
            if (tfsr_client_lock(fd, 1) == 0) __atomic_store_n(&h->client_lock_session, c->session, __ATOMIC_RELEASE);
#endif
            c->tx_tail_seen = __atomic_load_n(&h->c2w_tail, __ATOMIC_ACQUIRE);  // cursors keep counting across sessions
            c->c2w = static_cast<uint8_t*>(m) + h->c2w_off;
            c->w2c = static_cast<uint8_t*>(m) + h->w2c_off;
            *out = c;
            return 0;
          }
        }
        if (expired()) { munmap(m, (size_t)st.st_size); close(fd); return 5; }  // no worker, or the one session is taken
        timespec ts{0, 1000000};
        nanosleep(&ts, nullptr);
        struct stat st2{};
        if (stat(path.c_str(), &st2) != 0 || st2.st_size != st.st_size) { munmap(m, (size_t)st.st_size); close(fd); break; }  // re-sized: map again
      }
      continue;
    }
    if (fd >= 0) close(fd);
    if (expired()) return 5;
    timespec ts{0, 1000000};
    nanosleep(&ts, nullptr);
  }
}

static int connect_shm(const std::string& u, tfc_conn** out);

// A worker on this very node (loopback address): propose to carry the session over shared-memory rings instead of the
// socket (TFCS_OP_UPGRADE_SHM).  The client creates the ring file; the worker maps and page-locks it if it shares this
// /dev/shm and acknowledges.  Returns the ring connection, or nullptr (stay on TCP; nothing has been consumed but the refusal).
static tfc_conn* try_shm_upgrade(int fd) {
  static int seq = 0;
  const char* off = getenv("TFC_NO_SHM_UPGRADE");
  if (off && *off && *off != '0') return nullptr;
  const char* dir = getenv("TFC_SHM_DIR");
  const std::string base = dir && *dir ? dir : "/dev/shm";
  long mib = 1024;
  if (const char* e = getenv("TFC_UPGRADE_MIB")) mib = atol(e) >= 2 ? atol(e) : mib;
  const std::string name = "tfw-up-" + std::to_string((long)getpid()) + "-" + std::to_string(++seq);
  const std::string path = base + "/" + name;
  const uint64_t total = (uint64_t)mib << 20;
  const int sfd = open(path.c_str(), O_RDWR | O_CREAT | O_EXCL, 0666);
  if (sfd < 0) return nullptr;
  fchmod(sfd, 0666);
  if (ftruncate(sfd, (off_t)total) != 0) { close(sfd); unlink(path.c_str()); return nullptr; }
  close(sfd);
  tfcs_frame_hdr h{};
  h.magic = TFCS_MAGIC; h.version = TFCS_VERSION; h.opcode = TFCS_OP_UPGRADE_SHM; h.off0 = total; h.length = name.size();
  uint8_t frame[TFCS_HDR_BYTES + 272] = {0};
  std::memcpy(frame, &h, sizeof h);
  std::memcpy(frame + sizeof h, name.data(), name.size());
  tfcs_frame_hdr r{};
  if (!send_all(fd, frame, sizeof h + (size_t)tfcs_pad16(name.size())) || !recv_all(fd, &r, sizeof r) || r.magic != TFCS_MAGIC || r.opcode != TFCS_OP_RESP_ACK) {
    unlink(path.c_str());
    return nullptr;
  }
  tfc_conn* c = nullptr;
  const std::string saved = dir ? dir : "";
  setenv("TFC_SHM_DIR", base.c_str(), 1);
  const int rc = connect_shm("shmem+" + name + "+" + std::to_string(mib) + "+" + std::to_string(TFSR_VERSION), &c);
  if (dir) setenv("TFC_SHM_DIR", saved.c_str(), 1); else unsetenv("TFC_SHM_DIR");
  if (rc != 0 || !c) { unlink(path.c_str()); return nullptr; }
  c->lifeline_fd = fd;
  c->upgraded_ring = path;
  return c;
}

int tfc_connect(const char* url, tfc_conn** out) {
  if (!url || !out) return 1;
  std::string u(url), ip;
  int port = 8000;
  if (u.rfind("shmem+", 0) == 0) return connect_shm(u, out);
  if (u.rfind("native+", 0) == 0) {  // native+<ip>+<port>+<name>-<rv>
    const size_t a = 7, b = u.find('+', a);
    if (b == std::string::npos) return 1;
    ip = u.substr(a, b - a);
    port = atoi(u.c_str() + b + 1);
  } else {
    const size_t k = u.rfind(':');
    if (k == std::string::npos) return 1;
    ip = u.substr(0, k);
    port = atoi(u.c_str() + k + 1);
  }
  int fd = socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return 5;
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  if (inet_pton(AF_INET, ip.c_str(), &a.sin_addr) != 1 || connect(fd, (sockaddr*)&a, sizeof a) != 0) { close(fd); return 5; }
  if ((ntohl(a.sin_addr.s_addr) >> 24) == 127) {  // the worker is on this node: bulk bytes need not cross the TCP stack
    if (tfc_conn* up = try_shm_upgrade(fd)) { *out = up; return 0; }
  }
  tfc_conn* c = new tfc_conn();
  c->fd = fd;
  *out = c;
  return 0;
}

void tfc_close(tfc_conn* c) {
  if (!c) return;
  if (c->fd < 0) {
    tfsr_header* h = c->shm;
    for (uint32_t a = 1; a <= TFCS_MAX_ARENAS; ++a) drop_arena(c, a, __atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) < c->session);
    __atomic_store_n(&h->client_closed, c->session, __ATOMIC_RELEASE);
    // let the worker drain and answer (responses nobody waits for are dropped), bounded
    Waiter w;
    uint64_t tail = h->w2c_tail;
    for (int i = 0; i < 500000 && __atomic_load_n(&h->worker_closed, __ATOMIC_ACQUIRE) < c->session; ++i) {
      const uint64_t head = __atomic_load_n(&h->w2c_head, __ATOMIC_ACQUIRE);
      if (head != tail) { tail = head; __atomic_store_n(&h->w2c_tail, tail, __ATOMIC_RELEASE); }
      w.pause();
    }
    munmap(h, c->shm_bytes);
    if (c->shm_fd >= 0) close(c->shm_fd);  // releases the liveness lock
    if (c->lifeline_fd >= 0) close(c->lifeline_fd);
    if (!c->upgraded_ring.empty()) unlink(c->upgraded_ring.c_str());
    delete c;
    return;
  }
  flush(c);
  shutdown(c->fd, SHUT_WR);
  uint8_t buf[4096];
  while (recv(c->fd, buf, sizeof buf, 0) > 0) {}
  close(c->fd);
  delete c;
}

int tfc_malloc(tfc_conn* c, uint64_t bytes, uint32_t* handle) {
  if (!c || !handle) return 1;
  uint32_t id;
  if (!c->free_handles.empty()) {
    id = c->free_handles.front();
    c->free_handles.erase(c->free_handles.begin());
  } else {
    if (c->next_handle >= TFCS_MAX_HANDLES) return 4;  // every id is a live buffer
    id = c->next_handle++;
  }
  if (id >= c->live.size()) c->live.resize(id + 1, false);
  c->live[id] = true;
  tfcs_frame_hdr h = mk(c, TFCS_OP_MALLOC);
  h.h0 = *handle = id;
  h.length = bytes;
  c->pending_malloc[h.call_id] = id;
  return put(c, h) ? 0 : 5;
}
int tfc_free(tfc_conn* c, uint32_t handle) {
  if (!c) return 1;
  if (handle < c->live.size() && c->live[handle]) {  // ours: the id may be handed out again (FREE precedes the next MALLOC on the wire)
    c->live[handle] = false;
    c->free_handles.push_back(handle);
  }
  tfcs_frame_hdr h = mk(c, TFCS_OP_FREE);
  h.h0 = handle;
  return put(c, h) ? 0 : 5;
}
int tfc_memcpy_h2d(tfc_conn* c, uint32_t dst, uint64_t off, const void* src, uint64_t n) {
  if (!c || (!src && n)) return 1;
  uint32_t arena = 0;
  uint64_t aoff = 0;
  if (n && find_arena(c, src, n, &arena, &aoff)) {  // page-locked memory the worker maps too: no payload, the copy engine reads these pages
    tfcs_frame_hdr r = mk(c, TFCS_OP_MEMCPY_H2D_REF);
    r.h0 = dst; r.off0 = off; r.h1 = arena; r.off1 = aoff; r.length = n;
    return put(c, r) ? 0 : 5;
  }
  tfcs_frame_hdr h = mk(c, TFCS_OP_MEMCPY_H2D);
  h.h0 = dst; h.off0 = off; h.length = n;
  static const uint8_t zeros[16] = {0};
  if (c->fd < 0) return shm_write_frame(c, h, src, n) ? 0 : 5;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(&h);
  c->out.insert(c->out.end(), p, p + sizeof h);
  if (n >= (256u << 10)) {  // large payload: do not copy it through the coalescing buffer
    if (!flush(c) || !tx(c, src, n) || !tx(c, zeros, tfcs_pad16(n) - n)) return 5;
    return 0;
  }
  const uint8_t* s = static_cast<const uint8_t*>(src);
  c->out.insert(c->out.end(), s, s + n);
  c->out.insert(c->out.end(), zeros, zeros + (tfcs_pad16(n) - n));
  return c->out.size() < (1u << 20) || flush(c) ? 0 : 5;
}
int tfc_memcpy_d2h(tfc_conn* c, void* dst, uint32_t src, uint64_t off, uint64_t n) {
  if (!c || (!dst && n)) return 1;
  uint32_t arena = 0;
  uint64_t aoff = 0;
  if (n && find_arena(c, dst, n, &arena, &aoff)) {  // the copy engine writes the client's own pages; only an acknowledgement comes back
    tfcs_frame_hdr r = mk(c, TFCS_OP_MEMCPY_D2H_REF);
    r.h0 = src; r.off0 = off; r.h1 = arena; r.off1 = aoff; r.length = n; r.flags = TFCS_F_ACK;
    if (!put(c, r) || !flush(c)) return 5;
    return wait_for(c, r.call_id, TFCS_OP_RESP_ACK, nullptr, 0);
  }
  tfcs_frame_hdr h = mk(c, TFCS_OP_MEMCPY_D2H);
  h.h0 = src; h.off0 = off; h.length = n;
  if (!put(c, h) || !flush(c)) return 5;
  return wait_for(c, h.call_id, TFCS_OP_RESP_D2H, dst, n);
}
int tfc_memcpy_d2h_async(tfc_conn* c, void* dst, uint32_t src, uint64_t off, uint64_t n) {
  if (!c || (!dst && n)) return 1;
  uint32_t arena = 0;
  uint64_t aoff = 0;
  if (!n) return 0;
  if (!find_arena(c, dst, n, &arena, &aoff)) return 1;
  tfcs_frame_hdr r = mk(c, TFCS_OP_MEMCPY_D2H_REF);
  r.h0 = src; r.off0 = off; r.h1 = arena; r.off1 = aoff; r.length = n;
  return put(c, r) ? 0 : 5;
}

int tfc_host_alloc(tfc_conn* c, uint64_t bytes, void** out) {
  if (!c || !out || !bytes) return 1;
  *out = nullptr;
  if (c->fd >= 0) return 3;  // TCP: client and worker share no memory
  const uint64_t need = (bytes + 4095) & ~(uint64_t)4095;
  for (int pass = 0; pass < 2; ++pass) {
    for (uint32_t a = 1; a <= TFCS_MAX_ARENAS; ++a) {  // first fit in the arenas we have
      tfc_conn::ArenaMap& m = c->arenas[a];
      if (!m.base) continue;
      for (auto it = m.free_.begin(); it != m.free_.end(); ++it) {
        if (it->second < need) continue;
        const uint64_t o = it->first, rest = it->second - need;
        m.free_.erase(it);
        if (rest) m.free_[o + need] = rest;
        m.used[o] = need;
        *out = m.base + o;
        return 0;
      }
    }
    if (pass) break;
    // a new arena: at least 64 MiB so that small allocations share one registration
    uint32_t id = 0;
    for (uint32_t a = 1; a <= TFCS_MAX_ARENAS && !id; ++a) if (!c->arenas[a].base) id = a;
    if (!id) return 4;
    const uint64_t size = std::max<uint64_t>(need, 64ull << 20);
    const std::string path = arena_path(c, id);
    unlink(path.c_str());  // a leftover of a crashed client of this ring
    const int fd = open(path.c_str(), O_RDWR | O_CREAT | O_EXCL, 0666);
    if (fd < 0) return 4;
    fchmod(fd, 0666);  // the worker container may run as another user (cf. compose.go:1316)
    if (posix_fallocate(fd, 0, (off_t)size) != 0) { close(fd); unlink(path.c_str()); return 4; }  // fail now, not with SIGBUS on first touch
    void* mm = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
    close(fd);
    if (mm == MAP_FAILED) { unlink(path.c_str()); return 4; }
    tfc_conn::ArenaMap& m = c->arenas[id];
    m.base = static_cast<uint8_t*>(mm);
    m.size = size;
    m.free_[0] = size;
    tfcs_frame_hdr h = mk(c, TFCS_OP_HOST_REGISTER);
    h.h0 = id; h.length = size;
    const uint32_t call = h.call_id;
    int rc = put(c, h) ? round_trip(c) : 5;
    if (rc == 0 && c->last_err && c->last_err_call == call) { rc = c->last_err; if (c->first_err == rc) c->first_err = 0; }
    if (rc) { drop_arena(c, id, false); return rc; }
  }
  return 4;
}
int tfc_host_free(tfc_conn* c, void* p) {
  if (!c || !p) return 1;
  uint32_t a = 0;
  uint64_t o = 0;
  if (!find_arena(c, p, 1, &a, &o)) return 2;
  tfc_conn::ArenaMap& m = c->arenas[a];
  auto it = m.used.find(o);
  if (it == m.used.end()) return 2;
  uint64_t lo = o, n = it->second;
  m.used.erase(it);
  auto nx = m.free_.lower_bound(lo);
  if (nx != m.free_.end() && nx->first == lo + n) { n += nx->second; nx = m.free_.erase(nx); }
  if (nx != m.free_.begin()) { auto pv = std::prev(nx); if (pv->first + pv->second == lo) { lo = pv->first; n += pv->second; m.free_.erase(pv); } }
  m.free_[lo] = n;
  if (m.used.empty()) drop_arena(c, a, true);  // nothing left in it: give the pages back (the worker drains first)
  else if (round_trip(c) != 0) return 5;       // memory may be handed out again: copies that still use it must have completed
  return 0;
}

int tfc_module_load(tfc_conn* c, const void* image, uint64_t bytes, uint32_t* module) {
  if (!c || !image || !bytes || !module) return 1;
  if (c->next_module > TFCS_MAX_MODULES) return 4;
  tfcs_frame_hdr h = mk(c, TFCS_OP_MODULE_LOAD);
  h.h0 = *module = c->next_module++;
  h.length = bytes;
  const uint32_t call = h.call_id;
  if (!put_with_payload(c, h, image, bytes)) return 5;
  const int rc = round_trip(c);  // cuModuleLoadData is synchronous in CUDA too: a bad image is reported here
  if (rc) return rc;
  if (c->last_err && c->last_err_call == call) { const int e = c->last_err; if (c->first_err == e) c->first_err = 0; return e; }
  return 0;
}
int tfc_module_unload(tfc_conn* c, uint32_t module) {
  if (!c) return 1;
  tfcs_frame_hdr h = mk(c, TFCS_OP_MODULE_UNLOAD);
  h.h0 = module;
  return put(c, h) ? 0 : 5;
}
int tfc_module_get_function(tfc_conn* c, uint32_t module, const char* name, uint32_t* function, uint32_t* nparams, uint32_t* offsets,
                            uint32_t* sizes, uint32_t cap, uint32_t* param_bytes) {
  if (!c || !name || !*name || !function) return 1;
  const size_t len = strlen(name);
  if (len > 1024 || c->next_function >= TFCS_MAX_FUNCTIONS) return 1;
  tfcs_frame_hdr h = mk(c, TFCS_OP_MODULE_GET_FUNCTION);
  h.h0 = module;
  h.h1 = *function = c->next_function++;
  h.length = len;
  if (!put_with_payload(c, h, name, len) || !flush(c)) return 5;
  uint32_t table[2 * 512];
  uint64_t got = 0;
  tfcs_frame_hdr r{};
  const int rc = wait_for(c, h.call_id, TFCS_OP_RESP_FUNCTION, table, sizeof table, &got, &r);
  if (rc) return rc;
  const uint32_t count = (uint32_t)(got / 8);
  if (nparams) *nparams = count;
  if (param_bytes) *param_bytes = r.arg1;
  for (uint32_t i = 0; i < count && i < cap; ++i) {
    if (offsets) offsets[i] = table[2 * i];
    if (sizes) sizes[i] = table[2 * i + 1];
  }
  return 0;
}
int tfc_launch_user(tfc_conn* c, uint32_t function, const uint32_t grid[3], const uint32_t block[3], uint32_t shared_bytes,
                    const void* params, uint32_t param_bytes, uint32_t cost_tokens) {
  if (!c || !grid || !block || (!params && param_bytes) || param_bytes > TFCS_MAX_PARAM_BYTES) return 1;
  uint8_t buf[sizeof(tfcs_launch_params) + TFCS_MAX_PARAM_BYTES];
  tfcs_launch_params lp{};
  for (int i = 0; i < 3; ++i) { lp.grid[i] = grid[i]; lp.block[i] = block[i]; }
  lp.shared_bytes = shared_bytes;
  lp.param_bytes = param_bytes;
  std::memcpy(buf, &lp, sizeof lp);
  if (param_bytes) std::memcpy(buf + sizeof lp, params, param_bytes);
  tfcs_frame_hdr h = mk(c, TFCS_OP_LAUNCH_USER);
  h.h1 = function;
  h.arg3 = cost_tokens;
  h.length = sizeof lp + param_bytes;
  return put_with_payload(c, h, buf, sizeof lp + param_bytes) ? 0 : 5;
}
int tfc_memcpy_d2d(tfc_conn* c, uint32_t dst, uint64_t doff, uint32_t src, uint64_t soff, uint64_t n) {
  if (!c) return 1;
  tfcs_frame_hdr h = mk(c, TFCS_OP_MEMCPY_D2D);
  h.h0 = dst; h.off0 = doff; h.h1 = src; h.off1 = soff; h.length = n;
  return put(c, h) ? 0 : 5;
}
int tfc_memset(tfc_conn* c, uint32_t dst, uint64_t off, int value, uint64_t n) {
  if (!c) return 1;
  tfcs_frame_hdr h = mk(c, TFCS_OP_MEMSET);
  h.h0 = dst; h.off0 = off; h.length = n; h.arg0 = (uint32_t)(value & 0xff);
  return put(c, h) ? 0 : 5;
}
int tfc_launch(tfc_conn* c, uint32_t kernel_id, uint32_t grid, uint32_t block, uint32_t handle, uint64_t off, uint64_t n,
               uint64_t scalar, uint32_t cost_tokens) {
  if (!c) return 1;
  tfcs_frame_hdr h = mk(c, TFCS_OP_LAUNCH);
  h.arg0 = kernel_id; h.arg1 = grid; h.arg2 = block; h.arg3 = cost_tokens;
  h.h0 = handle; h.off0 = off; h.length = n; h.off1 = scalar;
  return put(c, h) ? 0 : 5;
}
int tfc_sync(tfc_conn* c) {
  if (!c) return 1;
  tfcs_frame_hdr h = mk(c, TFCS_OP_SYNC);
  if (!put(c, h) || !flush(c)) return 5;
  const int rc = wait_for(c, h.call_id, TFCS_OP_RESP_SYNC, nullptr, 0);
  if (rc) return rc;
  const int e = c->first_err;
  c->first_err = 0;
  return e;
}
int tfc_last_error_code(const tfc_conn* c) { return c ? c->last_err : 1; }
uint32_t tfc_last_error_call(const tfc_conn* c) { return c ? c->last_err_call : 0; }

}  // extern "C"
