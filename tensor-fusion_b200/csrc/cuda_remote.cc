// cuda_remote.cc -- libcuda_remote.so: the client stub of remote vGPU mode.
//
// In a remote-mode pod the operator's init container fills /tensor-fusion with stub libraries and an
// ld.so.preload (internal/utils/compose.go:330-390); the application's CUDA calls end in a stub that
// forwards them to the worker named by
//   TENSOR_FUSION_OPERATOR_CONNECTION_INFO      "shmem+tf_shm+1024+1" (local sidecar, pod_webhook.go:580-586), or
//   TENSOR_FUSION_OPERATOR_GET_CONNECTION_URL   operator endpoint answering the connection URL as text/plain
//                                               ("native+<ip>+<port>+<name>-<rv>", internal/server/router/
//                                               connection.go:46-100, bearer = service-account token).
// The reference's stub is closed.  This one is a CUDA *driver API* facade over libtfc_client (TFCS,
// include/tfw_wire.h): device memory, copies, memsets, synchronisation and the worker's built-in
// kernels.  Install it as libcuda.so.1 in the stub directory.  Device pointers are synthetic:
//   bit 62 set | handle << 40 | offset      (65 535 buffers of up to 1 TiB)
// so pointer arithmetic inside a buffer keeps working on the client.
//
// User modules are forwarded: cuModuleLoadData ships the image (cubin / PTX / fatbin; its size is read from the
// image's own header) with TFCS_OP_MODULE_LOAD, cuModuleGetFunction fetches the kernel's parameter layout from the
// worker, cuLaunchKernel packs the parameter block and the worker swaps the synthetic device pointers in it for
// real addresses.  cuMemAllocHost returns memory from an arena the worker page-locks too, so copies from / to it
// are zero-copy DMA (same-node transport).  Not forwarded (CUDA_ERROR_NOT_SUPPORTED): module globals, textures,
// graphs, IPC, peer access, cuGetExportTable -- the private tables libcudart needs, so runtime-API programs
// (PyTorch included) cannot run on this driver-level stub; they need a runtime-level interposer.
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "hv_handshake.h"
#include "tfc_client.h"
#include "tfw_wire.h"

#define CU_EXPORT extern "C" __attribute__((visibility("default")))

typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef struct CUctx_st* CUcontext;
typedef struct CUstream_st* CUstream;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUevent_st* CUevent;
struct CUuuid { unsigned char bytes[16]; };

enum : CUresult {
  OK = 0, INVALID_VALUE = 1, OUT_OF_MEMORY = 2, NOT_INITIALIZED = 3, NO_DEVICE = 100, INVALID_DEVICE = 101,
  INVALID_CONTEXT = 201, NOT_FOUND = 500, LAUNCH_FAILED = 719, NOT_SUPPORTED = 801, UNKNOWN = 999
};

namespace {

constexpr CUdeviceptr kPtrTag = 1ull << 62;
constexpr int kOffBits = 40;

std::mutex g_mu;            // one connection, calls serialised (TFCS is one ordered stream)
tfc_conn* g_conn = nullptr;
bool g_init_tried = false;
CUresult g_init_rc = NOT_INITIALIZED;
std::map<uint32_t, uint64_t> g_sizes;  // handle -> bytes
uint64_t g_used = 0, g_total = 180ull << 30;
int g_ctx_token, g_mod_token;          // addresses serve as the one context / module
thread_local CUcontext t_ctx = nullptr;
bool g_log = false;

void rlog(const char* fmt, const char* a) {
  if (g_log) { fprintf(stderr, "[libcuda_remote] "); fprintf(stderr, fmt, a); fputc('\n', stderr); }
}

// tfw_status -> CUresult
CUresult map_rc(int rc) {
  switch (rc) {
    case 0: return OK;
    case 1: return INVALID_VALUE;
    case 2: return INVALID_VALUE;   // unknown handle = invalid device pointer
    case 4: return OUT_OF_MEMORY;   // TFW_ERR_EXHAUSTED: the vGPU's VRAM quota
    case 3: return NOT_SUPPORTED;
    default: return UNKNOWN;
  }
}

bool split(CUdeviceptr p, uint32_t* handle, uint64_t* off) {
  if (!(p & kPtrTag)) return false;
  *handle = (uint32_t)((p >> kOffBits) & 0xffff);
  *off = p & ((1ull << kOffBits) - 1);
  return g_sizes.count(*handle) != 0;
}

// "http://host:port/path?query" -> body of the 200 answer, "" otherwise
std::string http_get_text(const std::string& url, const std::string& bearer) {
  if (url.compare(0, 7, "http://") != 0) return "";
  const size_t slash = url.find('/', 7);
  const std::string hostport = url.substr(7, slash == std::string::npos ? std::string::npos : slash - 7);
  const std::string path = slash == std::string::npos ? "/" : url.substr(slash);
  const size_t colon = hostport.rfind(':');
  const std::string host = colon == std::string::npos ? hostport : hostport.substr(0, colon);
  const int port = colon == std::string::npos ? 80 : atoi(hostport.c_str() + colon + 1);
  const std::string req = "GET " + path + " HTTP/1.1\r\nHost: " + hostport + "\r\nAuthorization: Bearer " + bearer +
                          "\r\nConnection: close\r\n\r\n";
  const std::string r = tfhv::http_call(host.c_str(), port, req);
  if (r.compare(0, 12, "HTTP/1.1 200") != 0 && r.compare(0, 12, "HTTP/1.0 200") != 0) return "";
  const size_t body = r.find("\r\n\r\n");
  if (body == std::string::npos) return "";
  std::string b = r.substr(body + 4);
  if (r.find("chunked") != std::string::npos && r.find("chunked") < body) {  // gin streams short strings un-chunked; be safe
    const size_t nl = b.find("\r\n");
    if (nl != std::string::npos) {
      const size_t len = strtoul(b.c_str(), nullptr, 16);
      b = b.substr(nl + 2, len);
    }
  }
  while (!b.empty() && (b.back() == '\n' || b.back() == '\r' || b.back() == ' ')) b.pop_back();
  return b;
}

std::string connection_url() {
  if (const char* e = getenv("TENSOR_FUSION_OPERATOR_CONNECTION_INFO"))  // pkg/constants/env.go:72
    if (*e) return e;
  const char* get = getenv("TENSOR_FUSION_OPERATOR_GET_CONNECTION_URL");  // env.go:71
  if (!get || !*get) return "";
  std::string token;
  const char* tf = getenv("TFW_SA_TOKEN_FILE");
  if (FILE* f = fopen(tf && *tf ? tf : "/var/run/secrets/kubernetes.io/serviceaccount/token", "r")) {
    char buf[8192];
    const size_t n = fread(buf, 1, sizeof buf - 1, f);
    buf[n] = 0;
    token = buf;
    while (!token.empty() && (token.back() == '\n' || token.back() == '\r')) token.pop_back();
    fclose(f);
  }
  return http_get_text(get, token);
}

CUresult ensure_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_init_tried) return g_init_rc;
  g_init_tried = true;
  const char* lg = getenv("TF_ENABLE_LOG");
  g_log = lg && *lg && strcmp(lg, "0") != 0;
  const std::string url = connection_url();
  if (url.empty()) { rlog("no connection URL (%s)", "TENSOR_FUSION_OPERATOR_CONNECTION_INFO / _GET_CONNECTION_URL"); return g_init_rc = NO_DEVICE; }
  if (tfc_connect(url.c_str(), &g_conn) != 0) { rlog("cannot reach the worker at %s", url.c_str()); return g_init_rc = NO_DEVICE; }
  if (const char* m = getenv("TF_CUDA_MEMORY_LIMIT")) g_total = strtoull(m, nullptr, 10) << 20;  // MiB, compose.go:1287-1295
  rlog("connected to %s", url.c_str());
  return g_init_rc = OK;
}

#define NEED_INIT()                                   \
  do {                                                \
    if (!g_conn || g_init_rc != OK) return g_init_tried ? g_init_rc : NOT_INITIALIZED; \
  } while (0)

// ---- user modules ------------------------------------------------------------------------------------------
struct RemoteModule { uint32_t id; };
struct RemoteFunction {
  uint32_t id = 0, param_bytes = 0;
  std::vector<uint32_t> off, size;
};
std::set<RemoteModule*> g_modules;
std::set<RemoteFunction*> g_functions;
std::set<void*> g_plain_host;  // cuMemAllocHost memory that is ordinary memory (TCP transport: nothing to share)

// Bytes of a code image handed to cuModuleLoadData (the API carries no length): ELF cubin, fatbin container,
// fatbin wrapper (what nvcc emits around embedded fatbins), else NUL-terminated PTX text.  0 = not recognised.
uint64_t image_bytes(const void** image) {
  const uint8_t* p = static_cast<const uint8_t*>(*image);
  uint32_t magic;
  memcpy(&magic, p, 4);
  if (magic == 0x466243B1u) {  // __fatBinC_Wrapper_t {int magic; int version; const void* data; void* filename}
    const void* data;
    memcpy(&data, p + 8, sizeof data);
    if (!data) return 0;
    *image = data;
    p = static_cast<const uint8_t*>(data);
    memcpy(&magic, p, 4);
  }
  if (magic == 0xBA55ED50u) {  // fatBinaryHeader {u32 magic; u16 version; u16 headerSize; u64 fatSize}
    uint16_t hs;
    uint64_t fs;
    memcpy(&hs, p + 6, 2);
    memcpy(&fs, p + 8, 8);
    return (uint64_t)hs + fs;
  }
  if (memcmp(p, "\177ELF", 4) == 0 && p[4] == 2) {  // ELF64: the furthest byte any header or section reaches
    uint64_t shoff, phoff;
    uint16_t shentsize, shnum, phentsize, phnum;
    memcpy(&phoff, p + 0x20, 8); memcpy(&shoff, p + 0x28, 8);
    memcpy(&phentsize, p + 0x36, 2); memcpy(&phnum, p + 0x38, 2);
    memcpy(&shentsize, p + 0x3A, 2); memcpy(&shnum, p + 0x3C, 2);
    uint64_t end = std::max<uint64_t>(64, std::max(shoff + (uint64_t)shentsize * shnum, phoff + (uint64_t)phentsize * phnum));
    for (uint16_t i = 0; i < shnum; ++i) {
      const uint8_t* sh = p + shoff + (uint64_t)i * shentsize;
      uint32_t type;
      uint64_t off, size;
      memcpy(&type, sh + 4, 4); memcpy(&off, sh + 0x18, 8); memcpy(&size, sh + 0x20, 8);
      if (type != 8 /* SHT_NOBITS */) end = std::max(end, off + size);
    }
    return end;
  }
  const size_t n = strnlen(reinterpret_cast<const char*>(p), 64u << 20);
  if (n > 16 && n < (64u << 20) && (strstr(reinterpret_cast<const char*>(p), ".version") || strstr(reinterpret_cast<const char*>(p), ".target"))) return n;
  return 0;
}

struct Builtin { const char* name; uint32_t id; };
const Builtin kBuiltins[] = {{"tfw_noop", TFCS_KERNEL_NOOP}, {"tfw_spin", TFCS_KERNEL_SPIN}, {"tfw_add_u8", TFCS_KERNEL_ADD_U8}, {"tfw_xor_idx", TFCS_KERNEL_XOR_IDX}};

// Positions in the one ordered stream, for events: everything recorded in epoch E is complete once a synchronise that
// started after it has returned (which opens epoch E + 1).
uint64_t g_record_epoch = 1;  // the epoch an event recorded now belongs to
uint64_t g_sync_epoch = 1;    // events of earlier epochs are complete
CUresult sync_locked() {      // caller holds g_mu
  const int rc = tfc_sync(g_conn);
  if (rc != 0) return map_rc(rc);
  g_sync_epoch = ++g_record_epoch;
  return OK;
}

}  // namespace

// ---------------------------------------------------------------- initialisation, device, context
CU_EXPORT CUresult cuInit(unsigned flags) { return flags ? INVALID_VALUE : ensure_init(); }
CU_EXPORT CUresult cuDriverGetVersion(int* v) { if (!v) return INVALID_VALUE; *v = 12090; return OK; }
CU_EXPORT CUresult cuDeviceGetCount(int* n) { if (!n) return INVALID_VALUE; NEED_INIT(); *n = 1; return OK; }
CU_EXPORT CUresult cuDeviceGet(CUdevice* d, int ordinal) { if (!d) return INVALID_VALUE; NEED_INIT(); if (ordinal != 0) return INVALID_DEVICE; *d = 0; return OK; }
CU_EXPORT CUresult cuDeviceGetName(char* name, int len, CUdevice d) {
  if (!name || len <= 0) return INVALID_VALUE;
  NEED_INIT();
  if (d != 0) return INVALID_DEVICE;
  snprintf(name, (size_t)len, "NVIDIA B200 (tensor-fusion remote vGPU)");
  return OK;
}
CU_EXPORT CUresult cuDeviceTotalMem_v2(size_t* bytes, CUdevice d) { if (!bytes) return INVALID_VALUE; NEED_INIT(); if (d != 0) return INVALID_DEVICE; *bytes = g_total; return OK; }
CU_EXPORT CUresult cuDeviceGetUuid_v2(CUuuid* u, CUdevice d) { if (!u) return INVALID_VALUE; NEED_INIT(); if (d != 0) return INVALID_DEVICE; memset(u, 0, sizeof *u); memcpy(u->bytes, "tf-remote-vgpu", 14); return OK; }
CU_EXPORT CUresult cuDeviceGetUuid(CUuuid* u, CUdevice d) { return cuDeviceGetUuid_v2(u, d); }
CU_EXPORT CUresult cuDeviceGetAttribute(int* v, int attrib, CUdevice d) {
  if (!v) return INVALID_VALUE;
  NEED_INIT();
  if (d != 0) return INVALID_DEVICE;
  switch (attrib) {  // CUdevice_attribute
    case 1: *v = 1024; return OK;             // MAX_THREADS_PER_BLOCK
    case 2: *v = 1024; return OK;             // MAX_BLOCK_DIM_X
    case 5: *v = 2147483647; return OK;       // MAX_GRID_DIM_X
    case 10: *v = 32; return OK;              // WARP_SIZE
    case 16: *v = 148; return OK;             // MULTIPROCESSOR_COUNT
    case 75: *v = 10; return OK;              // COMPUTE_CAPABILITY_MAJOR
    case 76: *v = 0; return OK;               // COMPUTE_CAPABILITY_MINOR
    case 41: *v = 1; return OK;               // UNIFIED_ADDRESSING
    default: *v = 0; return OK;
  }
}
CU_EXPORT CUresult cuCtxCreate_v2(CUcontext* c, unsigned, CUdevice d) { if (!c) return INVALID_VALUE; NEED_INIT(); if (d != 0) return INVALID_DEVICE; *c = t_ctx = reinterpret_cast<CUcontext>(&g_ctx_token); return OK; }
CU_EXPORT CUresult cuDevicePrimaryCtxRetain(CUcontext* c, CUdevice d) { return cuCtxCreate_v2(c, 0, d); }
CU_EXPORT CUresult cuDevicePrimaryCtxRelease_v2(CUdevice) { return OK; }
CU_EXPORT CUresult cuCtxDestroy_v2(CUcontext) { t_ctx = nullptr; return OK; }
CU_EXPORT CUresult cuCtxSetCurrent(CUcontext c) { t_ctx = c; return OK; }
CU_EXPORT CUresult cuCtxGetCurrent(CUcontext* c) { if (!c) return INVALID_VALUE; *c = t_ctx; return OK; }
CU_EXPORT CUresult cuCtxPushCurrent_v2(CUcontext c) { t_ctx = c; return OK; }
CU_EXPORT CUresult cuCtxPopCurrent_v2(CUcontext* c) { if (c) *c = t_ctx; t_ctx = nullptr; return OK; }
CU_EXPORT CUresult cuCtxGetDevice(CUdevice* d) { if (!d) return INVALID_VALUE; if (!t_ctx) return INVALID_CONTEXT; *d = 0; return OK; }
CU_EXPORT CUresult cuCtxSynchronize(void) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  return sync_locked();
}
CU_EXPORT CUresult cuCtxGetApiVersion(CUcontext, unsigned* v) { if (!v) return INVALID_VALUE; *v = 3020; return OK; }
CU_EXPORT CUresult cuCtxGetFlags(unsigned* f) { if (!f) return INVALID_VALUE; if (!t_ctx) return INVALID_CONTEXT; *f = 0; return OK; }
CU_EXPORT CUresult cuCtxSetLimit(int, size_t) { return OK; }           // stack / heap / printf limits live on the worker
CU_EXPORT CUresult cuCtxGetLimit(size_t* v, int) { if (!v) return INVALID_VALUE; *v = 0; return OK; }
CU_EXPORT CUresult cuDeviceComputeCapability(int* major, int* minor, CUdevice d) {
  if (!major || !minor) return INVALID_VALUE;
  NEED_INIT();
  if (d != 0) return INVALID_DEVICE;
  *major = 10; *minor = 0;
  return OK;
}

// ---------------------------------------------------------------- memory
CU_EXPORT CUresult cuMemAlloc_v2(CUdeviceptr* p, size_t bytes) {
  if (!p || !bytes) return INVALID_VALUE;
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  if (bytes >> kOffBits) return OUT_OF_MEMORY;
  uint32_t h = 0;
  if (tfc_malloc(g_conn, bytes, &h) != 0) return UNKNOWN;
  if (h > 0xffff) { tfc_free(g_conn, h); return OUT_OF_MEMORY; }
  // the allocation itself is fire-and-forget; a quota refusal must surface here, like cuMemAlloc's own OOM
  const CUresult sr = sync_locked();
  if (sr != OK) return sr;
  g_sizes[h] = bytes;
  g_used += bytes;
  *p = kPtrTag | ((CUdeviceptr)h << kOffBits);
  return OK;
}
CU_EXPORT CUresult cuMemFree_v2(CUdeviceptr p) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h;
  uint64_t off;
  if (!split(p, &h, &off) || off) return INVALID_VALUE;
  g_used -= g_sizes[h];
  g_sizes.erase(h);
  return tfc_free(g_conn, h) == 0 ? OK : UNKNOWN;
}
CU_EXPORT CUresult cuMemGetInfo_v2(size_t* free_b, size_t* total_b) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  if (total_b) *total_b = g_total;
  if (free_b) *free_b = g_total > g_used ? g_total - g_used : 0;
  return OK;
}
CU_EXPORT CUresult cuMemGetAddressRange_v2(CUdeviceptr* base, size_t* size, CUdeviceptr p) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h;
  uint64_t off;
  if (!split(p, &h, &off)) return INVALID_VALUE;
  if (base) *base = p - off;
  if (size) *size = g_sizes[h];
  return OK;
}
// Page-locked host memory: an arena the worker maps and page-locks too (tfc_host_alloc), so cuMemcpy*
// from / to it is DMA on the client's own pages.  Over TCP there is nothing to share: ordinary memory.
CU_EXPORT CUresult cuMemAllocHost_v2(void** pp, size_t bytes) {
  if (!pp || !bytes) return INVALID_VALUE;
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  const int rc = tfc_host_alloc(g_conn, bytes, pp);
  if (rc == 0) return OK;
  if (rc != 3) return OUT_OF_MEMORY;
  if (posix_memalign(pp, 4096, bytes) != 0) return OUT_OF_MEMORY;
  g_plain_host.insert(*pp);
  return OK;
}
CU_EXPORT CUresult cuMemHostAlloc(void** pp, size_t bytes, unsigned) { return cuMemAllocHost_v2(pp, bytes); }
CU_EXPORT CUresult cuMemFreeHost(void* p) {
  if (!p) return INVALID_VALUE;
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_plain_host.erase(p)) { free(p); return OK; }
  return tfc_host_free(g_conn, p) == 0 ? OK : INVALID_VALUE;
}
CU_EXPORT CUresult cuMemHostRegister_v2(void*, size_t, unsigned) { return NOT_SUPPORTED; }  // the worker cannot map arbitrary client pages
CU_EXPORT CUresult cuMemHostUnregister(void*) { return NOT_SUPPORTED; }

CU_EXPORT CUresult cuMemcpyHtoD_v2(CUdeviceptr dst, const void* src, size_t n) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h;
  uint64_t off;
  if (!split(dst, &h, &off) || (!src && n) || off + n > g_sizes[h]) return INVALID_VALUE;
  return n ? map_rc(tfc_memcpy_h2d(g_conn, h, off, src, n)) : OK;
}
CU_EXPORT CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr dst, const void* src, size_t n, CUstream) { return cuMemcpyHtoD_v2(dst, src, n); }
CU_EXPORT CUresult cuMemcpyDtoH_v2(void* dst, CUdeviceptr src, size_t n) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h;
  uint64_t off;
  if (!split(src, &h, &off) || (!dst && n) || off + n > g_sizes[h]) return INVALID_VALUE;
  return n ? map_rc(tfc_memcpy_d2h(g_conn, dst, h, off, n)) : OK;
}
CU_EXPORT CUresult cuMemcpyDtoHAsync_v2(void* dst, CUdeviceptr src, size_t n, CUstream) {
  NEED_INIT();
  {  // into page-locked (arena) memory the copy really is asynchronous: the next synchronize covers it
    std::lock_guard<std::mutex> lk(g_mu);
    uint32_t h;
    uint64_t off;
    if (!split(src, &h, &off) || (!dst && n) || off + n > g_sizes[h]) return INVALID_VALUE;
    if (!n) return OK;
    const int rc = tfc_memcpy_d2h_async(g_conn, dst, h, off, n);
    if (rc != 1) return map_rc(rc);
  }
  return cuMemcpyDtoH_v2(dst, src, n);
}
CU_EXPORT CUresult cuMemcpyDtoD_v2(CUdeviceptr dst, CUdeviceptr src, size_t n) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t hd, hs;
  uint64_t od, os;
  if (!split(dst, &hd, &od) || !split(src, &hs, &os) || od + n > g_sizes[hd] || os + n > g_sizes[hs]) return INVALID_VALUE;
  return n ? map_rc(tfc_memcpy_d2d(g_conn, hd, od, hs, os, n)) : OK;
}
CU_EXPORT CUresult cuMemcpyDtoDAsync_v2(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream) { return cuMemcpyDtoD_v2(dst, src, n); }
CU_EXPORT CUresult cuMemsetD8_v2(CUdeviceptr dst, unsigned char v, size_t n) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h;
  uint64_t off;
  if (!split(dst, &h, &off) || off + n > g_sizes[h]) return INVALID_VALUE;
  return n ? map_rc(tfc_memset(g_conn, h, off, v, n)) : OK;
}
CU_EXPORT CUresult cuMemsetD8Async(CUdeviceptr dst, unsigned char v, size_t n, CUstream) { return cuMemsetD8_v2(dst, v, n); }
// TFCS MEMSET is a byte fill.  A 16- or 32-bit pattern whose bytes differ is laid down with operations the wire has:
// one H2D of a seed block (the pattern repeated, at most 64 KiB) and then D2D copies that double the filled prefix --
// log2(n / 64 KiB) frames, every one a non-overlapping copy inside the buffer.
static CUresult pattern_fill(CUdeviceptr dst, const void* pat, size_t width, size_t count) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h;
  uint64_t off;
  const uint64_t total = (uint64_t)count * width;
  if (!split(dst, &h, &off) || (off % width) || off + total > g_sizes[h]) return INVALID_VALUE;
  if (!total) return OK;
  const uint64_t seed = std::min<uint64_t>(total, 64u << 10);
  std::vector<uint8_t> blk(seed);
  for (uint64_t i = 0; i < seed; i += width) memcpy(blk.data() + i, pat, width);
  int rc = tfc_memcpy_h2d(g_conn, h, off, blk.data(), seed);
  for (uint64_t done = seed; rc == 0 && done < total; ) {
    const uint64_t k = std::min(done, total - done);
    rc = tfc_memcpy_d2d(g_conn, h, off + done, h, off, k);
    done += k;
  }
  return map_rc(rc);
}
CU_EXPORT CUresult cuMemsetD32_v2(CUdeviceptr dst, unsigned v, size_t n) {
  const unsigned char b = (unsigned char)(v & 0xff);
  if (((v >> 8) & 0xff) == b && ((v >> 16) & 0xff) == b && (v >> 24) == b) return cuMemsetD8_v2(dst, b, n * 4);
  return pattern_fill(dst, &v, 4, n);
}
CU_EXPORT CUresult cuMemsetD32Async(CUdeviceptr dst, unsigned v, size_t n, CUstream) { return cuMemsetD32_v2(dst, v, n); }
CU_EXPORT CUresult cuMemsetD16_v2(CUdeviceptr dst, unsigned short v, size_t n) {
  if ((v >> 8) == (v & 0xff)) return cuMemsetD8_v2(dst, (unsigned char)(v & 0xff), n * 2);
  return pattern_fill(dst, &v, 2, n);
}
CU_EXPORT CUresult cuMemsetD16Async(CUdeviceptr dst, unsigned short v, size_t n, CUstream) { return cuMemsetD16_v2(dst, v, n); }

// Unified addressing: the direction is in the pointers themselves (device pointers carry the stub's tag).
CU_EXPORT CUresult cuMemcpy(CUdeviceptr dst, CUdeviceptr src, size_t n) {
  const bool dd = (dst & kPtrTag) != 0, sd = (src & kPtrTag) != 0;
  if (dd && sd) return cuMemcpyDtoD_v2(dst, src, n);
  if (dd) return cuMemcpyHtoD_v2(dst, reinterpret_cast<const void*>(src), n);
  if (sd) return cuMemcpyDtoH_v2(reinterpret_cast<void*>(dst), src, n);
  if (n) memmove(reinterpret_cast<void*>(dst), reinterpret_cast<const void*>(src), n);  // host to host
  return OK;
}
CU_EXPORT CUresult cuMemcpyAsync(CUdeviceptr dst, CUdeviceptr src, size_t n, CUstream st) {
  const bool dd = (dst & kPtrTag) != 0, sd = (src & kPtrTag) != 0;
  if (!dd && sd) return cuMemcpyDtoHAsync_v2(reinterpret_cast<void*>(dst), src, n, st);
  return cuMemcpy(dst, src, n);
}
// CUpointer_attribute: CONTEXT 1, MEMORY_TYPE 2 (HOST 1 / DEVICE 2), DEVICE_POINTER 3, HOST_POINTER 4, IS_MANAGED 8, DEVICE_ORDINAL 9,
// RANGE_START_ADDR 11, RANGE_SIZE 12
CU_EXPORT CUresult cuPointerGetAttribute(void* data, int attribute, CUdeviceptr p) {
  if (!data) return INVALID_VALUE;
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h = 0;
  uint64_t off = 0;
  const bool dev = split(p, &h, &off);
  if (!dev && (p & kPtrTag)) return INVALID_VALUE;  // tagged, but not a live allocation
  switch (attribute) {
    case 1: *static_cast<CUcontext*>(data) = reinterpret_cast<CUcontext>(&g_ctx_token); return OK;
    case 2: if (!dev) return INVALID_VALUE; *static_cast<unsigned*>(data) = 2; return OK;  // (plain host memory is unknown to CUDA: INVALID_VALUE, as the driver answers)
    case 3: if (!dev) return INVALID_VALUE; *static_cast<CUdeviceptr*>(data) = p; return OK;
    case 4: return INVALID_VALUE;  // device memory has no host address
    case 8: *static_cast<unsigned*>(data) = 0; return OK;
    case 9: *static_cast<int*>(data) = 0; return OK;
    case 11: if (!dev) return INVALID_VALUE; *static_cast<CUdeviceptr*>(data) = p - off; return OK;
    case 12: if (!dev) return INVALID_VALUE; *static_cast<size_t*>(data) = g_sizes[h]; return OK;
    default: return NOT_SUPPORTED;
  }
}

// ---------------------------------------------------------------- streams (one ordered stream on the worker)
CU_EXPORT CUresult cuStreamCreate(CUstream* s, unsigned) { if (!s) return INVALID_VALUE; NEED_INIT(); *s = reinterpret_cast<CUstream>(&g_ctx_token); return OK; }
CU_EXPORT CUresult cuStreamCreateWithPriority(CUstream* s, unsigned f, int) { return cuStreamCreate(s, f); }
CU_EXPORT CUresult cuStreamDestroy_v2(CUstream) { return OK; }
CU_EXPORT CUresult cuStreamSynchronize(CUstream) { return cuCtxSynchronize(); }
CU_EXPORT CUresult cuStreamQuery(CUstream) { return cuCtxSynchronize(); }
CU_EXPORT CUresult cuStreamGetFlags(CUstream, unsigned* f) { if (!f) return INVALID_VALUE; *f = 1; return OK; }  // CU_STREAM_NON_BLOCKING
CU_EXPORT CUresult cuStreamGetPriority(CUstream, int* p) { if (!p) return INVALID_VALUE; *p = 0; return OK; }
// "after everything before it in the stream": the stream is synchronised, then the function runs on the calling thread
CU_EXPORT CUresult cuLaunchHostFunc(CUstream, void (*fn)(void*), void* user) {
  if (!fn) return INVALID_VALUE;
  const CUresult r = cuCtxSynchronize();
  if (r == OK) fn(user);
  return r;
}

// ---------------------------------------------------------------- events
// Every stream of the application is the vGPU's one ordered stream, so an event is a position in that stream: anything
// recorded is complete once the stream has been synchronised after the record.  Waiting on an event from another stream is
// a no-op (already ordered).  Elapsed time needs timestamps taken on the GPU, which the wire does not carry: NOT_SUPPORTED.
namespace {
struct RemoteEvent { uint64_t recorded_at = 0; bool recorded = false; };
std::set<RemoteEvent*> g_events;
}  // namespace
CU_EXPORT CUresult cuEventCreate(CUevent* e, unsigned) {
  if (!e) return INVALID_VALUE;
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  RemoteEvent* ev = new RemoteEvent();
  g_events.insert(ev);
  *e = reinterpret_cast<CUevent>(ev);
  return OK;
}
CU_EXPORT CUresult cuEventDestroy_v2(CUevent e) {
  std::lock_guard<std::mutex> lk(g_mu);
  RemoteEvent* ev = reinterpret_cast<RemoteEvent*>(e);
  if (!g_events.erase(ev)) return 400;  // CUDA_ERROR_INVALID_HANDLE
  delete ev;
  return OK;
}
CU_EXPORT CUresult cuEventRecord(CUevent e, CUstream) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  RemoteEvent* ev = reinterpret_cast<RemoteEvent*>(e);
  if (!g_events.count(ev)) return 400;
  ev->recorded = true;
  ev->recorded_at = g_record_epoch;
  return OK;
}
CU_EXPORT CUresult cuEventRecordWithFlags(CUevent e, CUstream s, unsigned) { return cuEventRecord(e, s); }
CU_EXPORT CUresult cuEventSynchronize(CUevent e) {
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  RemoteEvent* ev = reinterpret_cast<RemoteEvent*>(e);
  if (!g_events.count(ev)) return 400;
  if (!ev->recorded || ev->recorded_at < g_sync_epoch) return OK;  // never recorded, or a synchronise has covered it
  return sync_locked();
}
CU_EXPORT CUresult cuEventQuery(CUevent e) { return cuEventSynchronize(e); }  // (a round trip, never NOT_READY: the stub has no completion feed)
CU_EXPORT CUresult cuStreamWaitEvent(CUstream, CUevent e, unsigned) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_events.count(reinterpret_cast<RemoteEvent*>(e)) ? OK : 400;
}
CU_EXPORT CUresult cuEventElapsedTime(float*, CUevent, CUevent) { return NOT_SUPPORTED; }

// ---------------------------------------------------------------- modules and launches (built-in kernels by name)
CU_EXPORT CUresult cuModuleLoadData(CUmodule* m, const void* image) {
  if (!m) return INVALID_VALUE;
  NEED_INIT();
  // the worker's own kernel table: an image with "tfw_builtin" in front (or none) selects it
  if (!image || memcmp(image, "tfw_builtin", 11) == 0) { *m = reinterpret_cast<CUmodule>(&g_mod_token); return OK; }
  const uint64_t bytes = image_bytes(&image);
  if (!bytes) return 200;  // CUDA_ERROR_INVALID_IMAGE
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t id = 0;
  const int rc = tfc_module_load(g_conn, image, bytes, &id);
  if (rc != 0) return rc == 1 ? 200 : map_rc(rc);
  RemoteModule* rm = new RemoteModule{id};
  g_modules.insert(rm);
  *m = reinterpret_cast<CUmodule>(rm);
  return OK;
}
CU_EXPORT CUresult cuModuleLoadDataEx(CUmodule* m, const void* image, unsigned, void*, void**) { return cuModuleLoadData(m, image); }
CU_EXPORT CUresult cuModuleLoadFatBinary(CUmodule* m, const void* fatbin) { return cuModuleLoadData(m, fatbin); }
CU_EXPORT CUresult cuModuleLoad(CUmodule* m, const char* fname) {
  if (!m || !fname) return INVALID_VALUE;
  FILE* f = fopen(fname, "rb");
  if (!f) return 301;  // CUDA_ERROR_FILE_NOT_FOUND
  std::vector<char> img;
  char buf[65536];
  size_t k;
  while ((k = fread(buf, 1, sizeof buf, f)) > 0) img.insert(img.end(), buf, buf + k);
  fclose(f);
  img.push_back(0);
  return cuModuleLoadData(m, img.data());
}
CU_EXPORT CUresult cuModuleUnload(CUmodule m) {
  if (m == reinterpret_cast<CUmodule>(&g_mod_token)) return OK;
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  RemoteModule* rm = reinterpret_cast<RemoteModule*>(m);
  if (!g_modules.erase(rm)) return INVALID_VALUE;
  const int rc = tfc_module_unload(g_conn, rm->id);
  delete rm;
  return rc == 0 ? OK : UNKNOWN;
}
CU_EXPORT CUresult cuModuleGetFunction(CUfunction* f, CUmodule m, const char* name) {
  if (!f || !name) return INVALID_VALUE;
  if (m == reinterpret_cast<CUmodule>(&g_mod_token)) {
    for (const Builtin& b : kBuiltins)
      if (strcmp(b.name, name) == 0) { *f = reinterpret_cast<CUfunction>(const_cast<Builtin*>(&b)); return OK; }
    return NOT_FOUND;
  }
  NEED_INIT();
  std::lock_guard<std::mutex> lk(g_mu);
  RemoteModule* rm = reinterpret_cast<RemoteModule*>(m);
  if (!g_modules.count(rm)) return INVALID_VALUE;
  RemoteFunction* rf = new RemoteFunction();
  uint32_t n = 0, off[512], size[512];
  const int rc = tfc_module_get_function(g_conn, rm->id, name, &rf->id, &n, off, size, 512, &rf->param_bytes);
  if (rc != 0 || n > 512) { delete rf; return rc == 2 ? NOT_FOUND : rc ? map_rc(rc) : NOT_SUPPORTED; }
  rf->off.assign(off, off + n);
  rf->size.assign(size, size + n);
  g_functions.insert(rf);
  *f = reinterpret_cast<CUfunction>(rf);
  return OK;
}
CU_EXPORT CUresult cuModuleGetGlobal_v2(CUdeviceptr*, size_t*, CUmodule, const char*) { return NOT_SUPPORTED; }
CU_EXPORT CUresult cuFuncSetAttribute(CUfunction, int, int) { return OK; }  // the worker raises the dynamic shared-memory limit itself
CU_EXPORT CUresult cuFuncSetCacheConfig(CUfunction, int) { return OK; }

// user kernels: the parameter block is packed from kernelParams with the layout the worker reported (or taken
// from `extra`); built-in kernels take (CUdeviceptr data, uint64_t n, uint64_t scalar)
CU_EXPORT CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned shmem,
                                  CUstream, void** params, void** extra) {
  NEED_INIT();
  const Builtin* b = reinterpret_cast<const Builtin*>(f);
  if (b < kBuiltins || b >= kBuiltins + sizeof(kBuiltins) / sizeof(kBuiltins[0])) {
    std::lock_guard<std::mutex> lk(g_mu);
    RemoteFunction* rf = reinterpret_cast<RemoteFunction*>(f);
    if (!g_functions.count(rf)) return 400;  // CUDA_ERROR_INVALID_HANDLE
    if (!gx || !gy || !gz || !bx || !by || !bz || (uint64_t)bx * by * bz > 1024) return INVALID_VALUE;
    alignas(16) uint8_t block[TFCS_MAX_PARAM_BYTES] = {0};
    uint32_t nbytes = rf->param_bytes;
    if (params) {
      for (size_t i = 0; i < rf->off.size(); ++i) {
        if (!params[i]) return INVALID_VALUE;
        memcpy(block + rf->off[i], params[i], rf->size[i]);
      }
    } else if (extra) {  // {CU_LAUNCH_PARAM_BUFFER_POINTER, buf, CU_LAUNCH_PARAM_BUFFER_SIZE, &size, CU_LAUNCH_PARAM_END}
      const void* buf = nullptr;
      size_t size = 0;
      for (int i = 0; extra[i] != nullptr && i < 8; i += 2) {
        if (extra[i] == reinterpret_cast<void*>(1)) buf = extra[i + 1];
        else if (extra[i] == reinterpret_cast<void*>(2)) size = *static_cast<size_t*>(extra[i + 1]);
      }
      if (!buf || size > TFCS_MAX_PARAM_BYTES) return INVALID_VALUE;
      memcpy(block, buf, size);
      nbytes = std::max<uint32_t>(nbytes, (uint32_t)size);
    } else if (!rf->off.empty()) return INVALID_VALUE;
    const uint32_t grid[3] = {gx, gy, gz}, blk[3] = {bx, by, bz};
    const uint64_t tokens = (uint64_t)gx * gy * gz * (((uint64_t)bx * by * bz + 31) / 32);
    return map_rc(tfc_launch_user(g_conn, rf->id, grid, blk, shmem, block, nbytes, (uint32_t)std::min<uint64_t>(tokens, 0xffffffffu)));
  }
  if (gy != 1 || gz != 1 || by != 1 || bz != 1 || !gx || !bx) return INVALID_VALUE;
  CUdeviceptr p = 0;
  uint64_t n = 0, scalar = 0;
  if (b->id != TFCS_KERNEL_NOOP) {
    if (!params || !params[0] || !params[1] || !params[2]) return INVALID_VALUE;
    p = *static_cast<CUdeviceptr*>(params[0]);
    n = *static_cast<uint64_t*>(params[1]);
    scalar = *static_cast<uint64_t*>(params[2]);
  }
  std::lock_guard<std::mutex> lk(g_mu);
  uint32_t h = 0;
  uint64_t off = 0;
  if (b->id != TFCS_KERNEL_NOOP && b->id != TFCS_KERNEL_SPIN && (!split(p, &h, &off) || off + n > g_sizes[h])) return INVALID_VALUE;
  const uint64_t tokens = (uint64_t)gx * ((bx + 31) / 32);  // blocks x warps, the unit libcuda_limiter.so charges
  return map_rc(tfc_launch(g_conn, b->id, gx, bx, h, off, n, scalar, (uint32_t)std::min<uint64_t>(tokens, 0xffffffffu)));
}

// ---------------------------------------------------------------- errors, lookup
CU_EXPORT CUresult cuGetErrorName(CUresult e, const char** s) {
  if (!s) return INVALID_VALUE;
  switch (e) {
    case OK: *s = "CUDA_SUCCESS"; break;
    case INVALID_VALUE: *s = "CUDA_ERROR_INVALID_VALUE"; break;
    case OUT_OF_MEMORY: *s = "CUDA_ERROR_OUT_OF_MEMORY"; break;
    case NOT_INITIALIZED: *s = "CUDA_ERROR_NOT_INITIALIZED"; break;
    case NO_DEVICE: *s = "CUDA_ERROR_NO_DEVICE"; break;
    case INVALID_DEVICE: *s = "CUDA_ERROR_INVALID_DEVICE"; break;
    case INVALID_CONTEXT: *s = "CUDA_ERROR_INVALID_CONTEXT"; break;
    case NOT_FOUND: *s = "CUDA_ERROR_NOT_FOUND"; break;
    case NOT_SUPPORTED: *s = "CUDA_ERROR_NOT_SUPPORTED"; break;
    default: *s = "CUDA_ERROR_UNKNOWN"; break;
  }
  return OK;
}
CU_EXPORT CUresult cuGetErrorString(CUresult e, const char** s) { return cuGetErrorName(e, s); }
CU_EXPORT CUresult cuGetProcAddress_v2(const char* name, void** pfn, int, uint64_t, void* status) {
  if (!name || !pfn) return INVALID_VALUE;
  Dl_info me{};
  void* self = dladdr(reinterpret_cast<void*>(&cuInit), &me) ? dlopen(me.dli_fname, RTLD_NOW | RTLD_NOLOAD) : nullptr;
  void* p = nullptr;
  if (self) {
    p = dlsym(self, name);
    if (!p) p = dlsym(self, (std::string(name) + "_v2").c_str());
    dlclose(self);
  }
  *pfn = p;
  if (status) *static_cast<int*>(status) = p ? 0 : 1;  // CU_GET_PROC_ADDRESS_SUCCESS / SYMBOL_NOT_FOUND
  return p ? OK : NOT_FOUND;
}
CU_EXPORT CUresult cuGetProcAddress(const char* name, void** pfn, int v, uint64_t f) { return cuGetProcAddress_v2(name, pfn, v, f, nullptr); }

// leave cleanly: the worker keeps the session open until the client says it is done
__attribute__((destructor)) static void remote_fini() {
  if (g_conn) { tfc_close(g_conn); g_conn = nullptr; }
}
