// worker_main.cc -- `tensor-fusion-worker`: the process the operator starts in the worker
// container (reference: internal/utils/compose.go:1304-1325)
//     ./tensor-fusion-worker -p 8000                        TCP transport, port "remote-vgpu"
//     ./tensor-fusion-worker -n shmem -m tf_shm -M 1024     shared-memory transport (local sidecar)
// Each accepted connection is one vGPU session: bytes read from the socket land in a pinned
// ring and go to the GPU through the C-ABI of libtfw_b200.so (include/tfw_worker.h); response
// frames go back on the same socket.  Environment contract: SURVEY.md App. D
// (TF_SHM_PATH, TF_CUDA_MEMORY_LIMIT [MiB], DISABLE_GPU_LIMITER, HYPERVISOR_IP/PORT, POD_NAME, ...).
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <signal.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "hv_handshake.h"
#include "tfw_worker.h"

namespace {

std::atomic<bool> g_stop{false};
bool g_log = false;

void logf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void logf(const char* fmt, ...) {
  if (!g_log) return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[tensor-fusion-worker] ");
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
}

bool send_all(int fd, const uint8_t* p, size_t n) {
  while (n) {
    ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    p += k;
    n -= (size_t)k;
  }
  return true;
}

// Bootstrap handshake with the hypervisor (hv_handshake.h), best effort.
uint64_t g_vram_limit_from_hypervisor = 0;

void hypervisor_handshake() {
  const tfhv::Result r = tfhv::handshake("tensorfusion-worker");  // pkg/constants/env.go:61
  if (!getenv("HYPERVISOR_IP")) return;
  if (!r.reached) { logf("hypervisor not reachable (continuing with env limits)"); return; }
  logf("hypervisor /api/v1/pod -> %.80s", r.pod_reply.c_str());
  g_vram_limit_from_hypervisor = r.vram_limit;
  logf("hypervisor /api/v1/process -> %.80s", r.process_reply.c_str());
}

void serve(int fd, int device) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  tfw_config cfg{};
  cfg.struct_size = sizeof cfg;
  cfg.device = device;
  const char* shm = getenv("TF_SHM_PATH");
  const char* nolim = getenv("DISABLE_GPU_LIMITER");
  if (shm && *shm && !(nolim && *nolim) && access(shm, R_OK | W_OK) == 0) cfg.shm_path = shm;
  if (const char* m = getenv("TF_CUDA_MEMORY_LIMIT")) cfg.vram_limit_bytes = strtoull(m, nullptr, 10) << 20;  // MiB (compose.go:1287-1295)
  else if (g_vram_limit_from_hypervisor) cfg.vram_limit_bytes = g_vram_limit_from_hypervisor;                   // RemotePodInfo.vram_limit, bytes
  tfw_worker* w = nullptr;
  tfw_status rc = tfw_worker_create(&cfg, &w);
  if (rc != TFW_OK) {
    fprintf(stderr, "[tensor-fusion-worker] tfw_worker_create failed: %d%s\n", rc, rc == TFW_ERR_NO_DEVICE ? " (no CUDA device; there is no CPU fallback)" : "");
    close(fd);
    return;
  }
  // Two pinned rings: while the GPU still reads ring A (in-place DMA), the socket fills ring B.
  const size_t ring = 64u << 20;
  uint8_t* bufs[2] = {nullptr, nullptr};
  uint64_t tickets[2] = {0, 0};
  std::vector<uint8_t> out(16u << 20);
  if (tfw_host_alloc(ring, reinterpret_cast<void**>(&bufs[0])) != TFW_OK || tfw_host_alloc(ring, reinterpret_cast<void**>(&bufs[1])) != TFW_OK) {
    if (bufs[0]) tfw_host_free(bufs[0]);
    tfw_worker_destroy(w);
    close(fd);
    return;
  }
  int cur = 0;
  size_t fill = 0;
  uint64_t total = 0;
  auto drain = [&](bool block) {
    for (;;) {
      size_t m = 0;
      if (tfw_poll_responses(w, out.data(), out.size(), &m) != TFW_OK) return false;
      if (m && !send_all(fd, out.data(), m)) return false;
      if (!m || !block) return true;
    }
  };
  for (;;) {
    uint8_t* buf = bufs[cur];
    // AccelSnapshot / AccelResume of the provider arrive through the stats record.  A frozen vGPU stops
    // reading its socket, so the client is back-pressured by TCP until the resume.
    int frozen = 0;
    tfw_worker_poll_control(w, &frozen);
    if (frozen) {
      if (!drain(false)) break;
      usleep(2000);
      continue;
    }
    // Responses (D2H payloads, SYNC acks) become ready asynchronously: while the client is quiet,
    // keep delivering them instead of blocking in recv().
    pollfd pf{fd, POLLIN, 0};
    int pr = poll(&pf, 1, 2);
    if (pr == 0) { if (!drain(false)) break; continue; }
    if (pr < 0 && errno == EINTR) continue;
    ssize_t n = recv(fd, buf + fill, ring - fill, 0);
    if (n < 0 && errno == EINTR) continue;
    if (n <= 0) break;
    total += (uint64_t)n;
    size_t have = fill + (size_t)n, used = 0;
    for (;;) {
      rc = tfw_submit(w, buf, have, &used);
      if (rc != TFW_ERR_EXHAUSTED) break;  // response arena full: ship responses, then resume where we stopped
      tfw_flush(w);
      if (!drain(true)) { rc = TFW_ERR_FAILED; break; }
      std::memmove(buf, buf + used, have - used);
      have -= used;
    }
    if (rc != TFW_OK) { logf("submit failed: %d (%s)", rc, tfw_last_error(w)); break; }
    if (tfw_fence(w, &tickets[cur]) != TFW_OK) break;
    const int nxt = cur ^ 1;
    if (tickets[nxt] && tfw_fence_wait(w, tickets[nxt]) != TFW_OK) break;  // the other ring is free again
    const size_t rest = have - used;  // < 64 bytes: a partial header, carried over to the other ring
    std::memcpy(bufs[nxt], buf + used, rest);
    fill = rest;
    cur = nxt;
    if (!drain(false)) break;
  }
  tfw_flush(w);
  drain(true);
  tfw_stats st{};
  tfw_get_stats(w, &st);
  logf("session closed: %llu bytes in, %llu frames, %llu payload bytes, %llu mover launches", (unsigned long long)total,
       (unsigned long long)st.frames, (unsigned long long)st.payload_bytes, (unsigned long long)st.mover_launches);
  tfw_host_free(bufs[0]);
  tfw_host_free(bufs[1]);
  tfw_worker_destroy(w);
  close(fd);
}

}  // namespace

int main(int argc, char** argv) {
  int port = 8000;  // pkg/constants/env.go:155-156
  std::string transport = "native", shm_name;
  long shm_mb = 0;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "-p") && i + 1 < argc) port = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-n") && i + 1 < argc) transport = argv[++i];
    else if (!strcmp(argv[i], "-m") && i + 1 < argc) shm_name = argv[++i];
    else if (!strcmp(argv[i], "-M") && i + 1 < argc) shm_mb = atol(argv[++i]);
    else if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) {
      printf("usage: tensor-fusion-worker -p <port> | -n shmem -m <name> -M <MiB>\n");
      return 0;
    }
  }
  g_log = getenv("TF_ENABLE_LOG") != nullptr;
  if (transport == "shmem") {
    // The client half of the shmem transport is closed source and its ring layout unpublished
    // (internal/webhook/v1/pod_webhook.go:580-590 only fixes the URL "shmem+tf_shm+1024+1").
    fprintf(stderr, "[tensor-fusion-worker] shmem transport (%s, %ld MiB) is not implemented in this round; use -p <port>\n", shm_name.c_str(), shm_mb);
    return 3;
  }
  signal(SIGPIPE, SIG_IGN);
  hypervisor_handshake();
  int ls = socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  const char* bind_ip = getenv("TFW_BIND");
  a.sin_addr.s_addr = bind_ip ? inet_addr(bind_ip) : htonl(INADDR_ANY);
  if (bind(ls, (sockaddr*)&a, sizeof a) != 0 || listen(ls, 16) != 0) { perror("tensor-fusion-worker: bind/listen"); return 2; }
  socklen_t al = sizeof a;
  getsockname(ls, (sockaddr*)&a, &al);
  printf("tensor-fusion-worker listening on port %d\n", ntohs(a.sin_port));
  fflush(stdout);
  const char* once = getenv("TFW_ONESHOT");  // serve N connections, then exit (tests)
  int budget = once ? atoi(once) : -1;
  std::vector<std::thread> sessions;
  while (budget != 0) {
    int fd = accept(ls, nullptr, nullptr);
    if (fd < 0) { if (errno == EINTR) continue; break; }
    sessions.emplace_back(serve, fd, 0);
    if (budget > 0) --budget;
  }
  for (auto& t : sessions) t.join();
  close(ls);
  return 0;
}
