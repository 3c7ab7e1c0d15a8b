// hv_handshake.h -- bootstrap handshake of a GPU process with the node's hypervisor, best effort
// (pkg/hypervisor/server/handlers/legacy.go:191-262, 319-384):
//   GET  /api/v1/pod?container_name=...                     -> RemotePodInfo {gpu_uuids, tflops_limit, vram_limit, ...}
//   POST /api/v1/process?container_name=...&container_pid=N -> the hypervisor maps the container PID to the host
//                                                              PID and adds it to the pod's quota file (legacy.go:576)
// Authorization: Bearer <service-account JWT> (parsed, not verified, by the hypervisor: legacy.go:394-413).
// Shared by the worker executable (worker_main.cc) and the LD_PRELOAD limiter (cuda_hook.cc).
#pragma once
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace tfhv {

struct Result {
  bool reached = false;         // the hypervisor answered /api/v1/pod
  bool registered = false;      // /api/v1/process answered 2xx
  uint64_t vram_limit = 0;      // RemotePodInfo.vram_limit, bytes (0 = not reported)
  double tflops_limit = 0;      // RemotePodInfo.tflops_limit (0 = not reported)
  bool hard_isolation = false;  // RemotePodInfo.isolation == "hard"
  long auto_freeze_ttl_ms = 0;  // RemotePodInfo.auto_freeze {enable, freeze_to_mem_ttl} -> milliseconds (0 = off)
  std::string pod_reply, process_reply;  // first line + body, for logs
};

// Go duration ("90s", "5m", "1h30m", "250ms") -> milliseconds; 0 if it does not parse
inline long go_duration_ms(const std::string& d) {
  double total = 0;
  size_t i = 0;
  while (i < d.size()) {
    char* end = nullptr;
    const double v = strtod(d.c_str() + i, &end);
    if (end == d.c_str() + i) return 0;
    i = (size_t)(end - d.c_str());
    if (d.compare(i, 2, "ms") == 0) { total += v; i += 2; }
    else if (d.compare(i, 2, "us") == 0) { total += v / 1000.0; i += 2; }
    else if (d.compare(i, 1, "s") == 0) { total += v * 1000.0; i += 1; }
    else if (d.compare(i, 1, "m") == 0) { total += v * 60000.0; i += 1; }
    else if (d.compare(i, 1, "h") == 0) { total += v * 3600000.0; i += 1; }
    else return 0;
  }
  return (long)total;
}

// Position just behind `"key" :` in a JSON text (white space tolerated; Go's encoder writes none), npos if absent.
inline size_t json_value_pos(const std::string& js, const char* key, size_t from = 0) {
  const std::string k = std::string("\"") + key + "\"";
  for (size_t at = js.find(k, from); at != std::string::npos; at = js.find(k, at + 1)) {
    size_t i = at + k.size();
    while (i < js.size() && (js[i] == ' ' || js[i] == '\t' || js[i] == '\n' || js[i] == '\r')) ++i;
    if (i < js.size() && js[i] == ':') {
      ++i;
      while (i < js.size() && (js[i] == ' ' || js[i] == '\t' || js[i] == '\n' || js[i] == '\r')) ++i;
      return i;
    }
  }
  return std::string::npos;
}
inline std::string json_string_at(const std::string& js, size_t pos) {
  if (pos == std::string::npos || pos >= js.size() || js[pos] != '"') return std::string();
  const size_t q = js.find('"', pos + 1);
  return q == std::string::npos ? std::string() : js.substr(pos + 1, q - pos - 1);
}

// "auto_freeze":{"freeze_to_mem_ttl":"5m","enable":true} of RemotePodInfo (api/http_types.go:82-100)
inline long parse_auto_freeze(const std::string& reply) {
  const size_t a = reply.find("\"auto_freeze\"");
  if (a == std::string::npos) return 0;
  const size_t close = reply.find('}', a);
  const std::string obj = reply.substr(a, close == std::string::npos ? std::string::npos : close - a);
  const size_t en = json_value_pos(obj, "enable");
  if (en == std::string::npos || obj.compare(en, 4, "true") != 0) return 0;
  return go_duration_ms(json_string_at(obj, json_value_pos(obj, "freeze_to_mem_ttl")));
}

inline std::string http_call(const char* ip, int port, const std::string& request) {
  std::string reply;
  int fd = socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return reply;
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  timeval tv{2, 0};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
  if (inet_pton(AF_INET, ip, &a.sin_addr) == 1 && connect(fd, (sockaddr*)&a, sizeof a) == 0) {
    const char* p = request.data();
    size_t n = request.size();
    bool ok = true;
    while (n) {
      ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
      if (k < 0) { if (errno == EINTR) continue; ok = false; break; }
      p += k;
      n -= (size_t)k;
    }
    if (ok) {
      char buf[4096];
      ssize_t k;
      while ((k = recv(fd, buf, sizeof buf, 0)) > 0 && reply.size() < (1u << 20)) reply.append(buf, (size_t)k);
    }
  }
  close(fd);
  return reply;
}

// `default_container`: CONTAINER_NAME fallback (pkg/constants/env.go:61 for the worker).
inline Result handshake(const char* default_container) {
  Result out;
  const char* ip = getenv("HYPERVISOR_IP");
  if (!ip || !*ip) return out;
  const char* port_s = getenv("HYPERVISOR_PORT");
  const int port = atoi(port_s && *port_s ? port_s : "8001");
  const char* cname = getenv("CONTAINER_NAME");
  const std::string container = cname && *cname ? cname : default_container;
  std::string token;
  const char* tf = getenv("TFW_SA_TOKEN_FILE");
  if (FILE* f = fopen(tf && *tf ? tf : "/var/run/secrets/kubernetes.io/serviceaccount/token", "r")) {
    char buf[8192];
    size_t n = fread(buf, 1, sizeof buf - 1, f);
    buf[n] = 0;
    token = buf;
    while (!token.empty() && (token.back() == '\n' || token.back() == '\r')) token.pop_back();
    fclose(f);
  }
  const std::string common =
      std::string(" HTTP/1.1\r\nHost: ") + ip + "\r\nAuthorization: Bearer " + token + "\r\nConnection: close\r\n";
  out.pod_reply = http_call(ip, port, "GET /api/v1/pod?container_name=" + container + common + "\r\n");
  if (out.pod_reply.empty()) return out;
  out.reached = true;
  const size_t k = json_value_pos(out.pod_reply, "vram_limit");
  if (k != std::string::npos) out.vram_limit = strtoull(out.pod_reply.c_str() + k, nullptr, 10);
  out.auto_freeze_ttl_ms = parse_auto_freeze(out.pod_reply);
  const size_t tfl = json_value_pos(out.pod_reply, "tflops_limit");
  if (tfl != std::string::npos) out.tflops_limit = strtod(out.pod_reply.c_str() + tfl, nullptr);
  out.hard_isolation = json_string_at(out.pod_reply, json_value_pos(out.pod_reply, "isolation")) == "hard";
  out.process_reply = http_call(ip, port,
                                "POST /api/v1/process?container_name=" + container +
                                    "&container_pid=" + std::to_string((long)getpid()) + common + "Content-Length: 0\r\n\r\n");
  out.registered = out.process_reply.compare(0, 10, "HTTP/1.1 2") == 0 || out.process_reply.compare(0, 10, "HTTP/1.0 2") == 0;
  return out;
}

}  // namespace tfhv
