// kernels.h -- internal launch wrappers of libtfw_b200's sm_100a kernels.
// Not part of the C-ABI (see include/tfw_worker.h for that).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tfw_worker.h"

namespace tfw {

// Geometry of the byte mover.  One tile = one CTA iteration.
constexpr int kMoverThreads = 256;
constexpr int kMoverMinCtas = 3;                                     // resident CTAs per SM (register cap 85)
constexpr int kMoverUnroll = 8;                                       // 16-B vectors in flight per thread
constexpr uint32_t kTileBytes = kMoverThreads * 16u * kMoverUnroll;   // 32 KiB

// Number of tiles a descriptor occupies (>= 1 for len > 0).  Tiles partition the
// 16-byte-aligned *destination* body; the first tile also owns the unaligned
// head bytes and the last tile the tail bytes.
// Head bytes copied one by one so that the vector body starts on a 128-byte line of
// the destination (full-line stores; partial-sector writes cost read-modify-write
// in L2).  Short copies only align to the 16-byte vector.
constexpr uint64_t kLineAlignMin = 4096;
static inline __host__ __device__ uint64_t mover_head(uint64_t dst, uint64_t len) {
  const uint64_t a = len >= kLineAlignMin ? 128u : 16u;
  uint64_t head = (a - (dst & (a - 1))) & (a - 1);
  return head > len ? len : head;
}
static inline __host__ __device__ uint32_t mover_tiles(uint64_t dst, uint64_t len) {
  if (len == 0) return 0;
  const uint64_t head = mover_head(dst, len);
  uint64_t body = (len - head) & ~(uint64_t)15u;
  uint64_t t = (body + kTileBytes - 1) / kTileBytes;
  return t ? (uint32_t)t : 1u;
}

enum MoverKind { kMoverLdg = 0, kMoverTma = 1 };
// A batch may go to the TMA mover when every copy in it has source and destination congruent modulo 16 (fills always are).
static inline bool mover_bulk_ok(const tfw_move_desc* d, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i)
    if (d[i].src && ((d[i].src ^ d[i].dst) & 15u)) return false;
  return true;
}

// Force-load every kernel of this library.  CUDA loads kernels lazily on first
// launch and that load synchronises with running work: a first launch issued
// while a blocking gate kernel spins would stall until the gate's fail-open timer.
cudaError_t preload_kernels();

// descs: device pointer to n descriptors with tile0 filled (exclusive scan).
// ctas_per_sm == 0: one tile per CTA (grid = total_tiles, the default: the hardware
// CTA scheduler keeps the active window contiguous in DRAM -- 6.9 TB/s vs 6.0 TB/s
// for a grid-stride loop, profiles/r01_copy_lab.jsonl); > 0: persistent grid.
cudaError_t launch_mover(const tfw_move_desc* d_descs, uint32_t n, uint32_t total_tiles, int sm_count,
                         int ctas_per_sm, MoverKind kind, cudaStream_t stream);
// Same, but <= kInlineDescs descriptors travel in the kernel parameter block
// (no descriptor upload; the small-call fast path).
constexpr uint32_t kInlineDescs = 24;
cudaError_t launch_mover_inline(const tfw_move_desc* h_descs, uint32_t n, uint32_t total_tiles, int sm_count,
                                int ctas_per_sm, cudaStream_t stream);

// digest: *d_out must be zeroed by the caller; finalisation happens on the host.
cudaError_t launch_digest(const void* d_buf, uint64_t bytes, unsigned long long* d_out, int sm_count,
                          cudaStream_t stream);

// deterministic test pattern: 8-byte word i of the range = digest_mix(seed + (i+1)*K1)
cudaError_t launch_pattern(void* d_buf, uint64_t bytes, uint64_t seed, int sm_count, cudaStream_t stream);

// built-in client kernels (TFCS_OP_LAUNCH); clamp_client_launch = the geometry a launch really runs with
void clamp_client_launch(uint32_t kernel_id, uint64_t len, uint32_t* grid, uint32_t* block);
cudaError_t launch_client_kernel(uint32_t kernel_id, uint32_t grid, uint32_t block, uint8_t* range, uint64_t len,
                                 uint64_t scalar, cudaStream_t stream);

// Digest arithmetic shared by host finalisation (and restated by the oracle).
constexpr uint64_t kDigestK1 = 0x9E3779B97F4A7C15ull;
static inline __host__ __device__ uint64_t digest_mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace tfw
