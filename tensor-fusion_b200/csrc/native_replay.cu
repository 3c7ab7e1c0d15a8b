// native_replay.cu -- the "native CUDA" comparator of BASELINE.md (row B5): the
// identical call stream issued straight to the CUDA runtime by one host thread
// (cudaMallocAsync / cudaMemcpyAsync / cudaMemsetAsync / kernel launch), no
// staging, no batching, no limiter.  Used only to quote the worker's added
// wall-clock ("<= 4 % over native", reference README.md:56).
#include <cuda_runtime.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "tfw_worker.h"

extern "C" TFW_API tfw_status tfw_native_replay(int device, const void* stream, size_t nbytes, uint32_t passes,
                                                 double* seconds_per_pass, uint64_t* payload_bytes,
                                                 uint64_t* calls) {
  if (!stream || !nbytes || !passes || !seconds_per_pass) return TFW_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return TFW_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return TFW_ERR_INVALID;
  if (tfw::preload_kernels() != cudaSuccess) return TFW_ERR_FAILED;
  cudaStream_t st;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return TFW_ERR_FAILED;
  uint8_t* scratch = nullptr;  // pinned landing zone for D2H
  const size_t scratch_bytes = 64u << 20;
  if (cudaHostAlloc(reinterpret_cast<void**>(&scratch), scratch_bytes, cudaHostAllocDefault) != cudaSuccess) { cudaStreamDestroy(st); return TFW_ERR_EXHAUSTED; }
  struct Buf { void* p = nullptr; uint64_t size = 0; };
  std::vector<Buf> bufs;
  const uint8_t* p = static_cast<const uint8_t*>(stream);
  tfw_status rc = TFW_OK;
  double total = 0.0;
  uint64_t pay = 0, ncalls = 0;
  for (uint32_t pass = 0; pass < passes + 1 && rc == TFW_OK; ++pass) {  // pass 0 = warm-up
    cudaStreamSynchronize(st);
    const auto t0 = std::chrono::steady_clock::now();
    size_t pos = 0;
    pay = 0; ncalls = 0;
    while (pos + TFCS_HDR_BYTES <= nbytes) {
      tfcs_frame_hdr h;
      std::memcpy(&h, p + pos, sizeof h);
      if (h.magic != TFCS_MAGIC) { rc = TFW_ERR_PROTOCOL; break; }
      pos += TFCS_HDR_BYTES;
      ++ncalls;
      auto ok = [&](uint32_t hd, uint64_t off, uint64_t len) { return hd < bufs.size() && bufs[hd].p && off <= bufs[hd].size && len <= bufs[hd].size - off; };
      switch (h.opcode) {
        case TFCS_OP_MALLOC:
          if (h.h0 >= bufs.size()) bufs.resize(h.h0 + 1);
          if (!bufs[h.h0].p && h.length && cudaMallocAsync(&bufs[h.h0].p, h.length, st) == cudaSuccess) bufs[h.h0].size = h.length;
          break;
        case TFCS_OP_FREE:
          if (h.h0 < bufs.size() && bufs[h.h0].p) { cudaFreeAsync(bufs[h.h0].p, st); bufs[h.h0] = Buf{}; }
          break;
        case TFCS_OP_MEMCPY_H2D:
          if (ok(h.h0, h.off0, h.length)) { cudaMemcpyAsync(static_cast<uint8_t*>(bufs[h.h0].p) + h.off0, p + pos, h.length, cudaMemcpyHostToDevice, st); pay += h.length; }
          pos += (size_t)tfcs_pad16(h.length);
          break;
        case TFCS_OP_MEMCPY_D2H:
          if (ok(h.h0, h.off0, h.length) && h.length <= scratch_bytes) cudaMemcpyAsync(scratch, static_cast<uint8_t*>(bufs[h.h0].p) + h.off0, h.length, cudaMemcpyDeviceToHost, st);
          break;
        case TFCS_OP_MEMCPY_D2D:
          if (ok(h.h0, h.off0, h.length) && ok(h.h1, h.off1, h.length)) cudaMemcpyAsync(static_cast<uint8_t*>(bufs[h.h0].p) + h.off0, static_cast<uint8_t*>(bufs[h.h1].p) + h.off1, h.length, cudaMemcpyDeviceToDevice, st);
          break;
        case TFCS_OP_MEMSET:
          if (ok(h.h0, h.off0, h.length)) cudaMemsetAsync(static_cast<uint8_t*>(bufs[h.h0].p) + h.off0, (int)(h.arg0 & 0xff), h.length, st);
          break;
        case TFCS_OP_LAUNCH:
          if (h.arg0 <= TFCS_KERNEL_XOR_IDX && (h.length == 0 || ok(h.h0, h.off0, h.length)))
            tfw::launch_client_kernel(h.arg0, h.arg1, h.arg2, h.length ? static_cast<uint8_t*>(bufs[h.h0].p) + h.off0 : nullptr, h.length, h.off1, st);
          break;
        case TFCS_OP_SYNC: cudaStreamSynchronize(st); break;
        default: break;
      }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = TFW_ERR_FAILED;
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (pass > 0) total += dt;
    for (auto& b : bufs) if (b.p) { cudaFreeAsync(b.p, st); b = Buf{}; }  // each pass starts from an empty session
  }
  cudaStreamSynchronize(st);
  cudaGetLastError();
  cudaFreeHost(scratch);
  cudaStreamDestroy(st);
  *seconds_per_pass = total / passes;
  if (payload_bytes) *payload_bytes = pay;
  if (calls) *calls = ncalls;
  return rc;
}

// Bulk-copy comparator for the process-boundary legs of the bench: what a CUDA application does
// without tensor-fusion -- `ncopies` x cudaMemcpyAsync of `each` bytes between `nsrc` host buffers
// (page-locked or pageable, touched beforehand) and `nbuf` device buffers on one stream, then one
// synchronize.  direction 0 = host -> device, 1 = device -> host.
extern "C" TFW_API tfw_status tfw_native_copy(int device, int direction, int pinned, uint64_t each, uint32_t nsrc, uint32_t nbuf,
                                               uint32_t ncopies, uint32_t passes, double* seconds_per_pass) {
  if (!each || !nsrc || !nbuf || !ncopies || !passes || !seconds_per_pass || direction < 0 || direction > 1) return TFW_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return TFW_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return TFW_ERR_INVALID;
  cudaStream_t st;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return TFW_ERR_FAILED;
  std::vector<void*> host(nsrc, nullptr), dev(nbuf, nullptr);
  tfw_status rc = TFW_OK;
  for (auto& h : host) {
    if (pinned ? cudaHostAlloc(&h, each, cudaHostAllocDefault) != cudaSuccess : (h = std::malloc(each)) == nullptr) { rc = TFW_ERR_EXHAUSTED; break; }
    std::memset(h, 0x5A, each);  // first touch outside the timed region
  }
  for (auto& d : dev) if (rc == TFW_OK && cudaMalloc(&d, each) != cudaSuccess) rc = TFW_ERR_EXHAUSTED;
  double total = 0.0;
  for (uint32_t pass = 0; pass < passes + 1 && rc == TFW_OK; ++pass) {  // pass 0 = warm-up
    cudaStreamSynchronize(st);
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t i = 0; i < ncopies; ++i) {
      if (direction == 0) cudaMemcpyAsync(dev[i % nbuf], host[i % nsrc], each, cudaMemcpyHostToDevice, st);
      else cudaMemcpyAsync(host[i % nsrc], dev[i % nbuf], each, cudaMemcpyDeviceToHost, st);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = TFW_ERR_FAILED;
    if (pass > 0) total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  cudaGetLastError();
  for (auto h : host) if (h) { if (pinned) cudaFreeHost(h); else std::free(h); }
  for (auto d : dev) if (d) cudaFree(d);
  cudaStreamDestroy(st);
  *seconds_per_pass = total / passes;
  return rc;
}
