/*
 * tfw_vram.h -- C-ABI of the vGPU VRAM tiering layer (north_star (c)).
 *
 * The reference only *accounts* for expanded VRAM (virtual capacity = VRAM +
 * %host RAM + %disk, internal/gpuallocator/node_capacity.go:143-162; knobs
 * api/v1/gpupool_types.go:64-84) and leaves the mechanism empty
 * (pkg/hypervisor/worker/vram/vram_trap.go:1-3, worker/state/ctx_migration.go:1,
 * handlers/legacy.go:111-139).  This is the mechanism:
 *
 *   one vGPU address space = one CUDA virtual-address reservation cut into
 *   fixed-size regions; each region is backed by
 *     HOME  physical HBM of the vGPU's own GPU,
 *     PEER  physical HBM of another GPU of the box, mapped into the same VA and
 *           reached by the home GPU over NVLink 5 / NVSwitch, or
 *     HOST  pinned host DRAM (region unmapped until it is prefetched again).
 *   Evict / prefetch copy the bytes with the byte-mover kernel (one-sided P2P
 *   put / get, no collective) or the copy engine (host tier) and then re-map
 *   the region, so client-visible device pointers never change.
 *
 * Byte-moving only: a region's contents are bit-identical before and after any
 * sequence of migrations (tests/test_gpu_vram.py).
 */
#ifndef TFW_VRAM_H
#define TFW_VRAM_H

#include <stddef.h>
#include <stdint.h>

#include "tfw_worker.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfw_vspace tfw_vspace;

typedef enum { TFW_TIER_NONE = 0, TFW_TIER_HOME = 1, TFW_TIER_PEER = 2, TFW_TIER_HOST = 3 } tfw_tier;

#define TFW_VRAM_MAX_PEERS 15
#define TFW_VS_COPY_ENGINE 0x1u /* move peer-tier data with cudaMemcpyAsync instead of the mover kernel (library baseline) */
#define TFW_VS_MOVER_TMA 0x2u
/* Evictions are normally PULLED by the destination GPU (SM-initiated NVLink reads reach 790 GB/s,
 * writes 718 GB/s on B200).  When many vGPU processes share the same peer GPUs, foreign pull
 * kernels time-slice on them; this flag keeps every copy kernel on the tenant's own (home) GPU. */
#define TFW_VS_PUSH_EVICT 0x4u
/* Every peer-tier copy is driven by the GPU that holds the SOURCE (evictions pushed by the home GPU, prefetches pushed by
 * the peer): with both directions busy at once a pull costs read-request traffic on the opposite direction (ncu: 0.27 GB of
 * requests per GiB pulled, profiles/r02_peer_ncu.md), a push only acknowledgements. */
#define TFW_VS_SENDER_DRIVEN 0x8u
/* Regions the policy evicts to a peer keep their VA mapped there, so the home GPU can use them in place over NVLink
 * without bringing them home (the worker's tiered buffers do).  Without it such a region's VA stays unmapped until
 * tfw_vspace_access brings it home again: granting the home GPU access to peer-located memory is by far the most
 * expensive VMM call of a migration. */
#define TFW_VS_PEER_IN_PLACE 0x10u
/* Every peer-tier copy is driven by the HOME GPU (it pulls prefetches and pushes evictions, on two streams): no other GPU
 * ever touches the vGPU's memory, so backings are mapped for the home GPU only -- and re-pointing a region's VA never
 * involves another GPU's page tables (VMM calls on allocations that peers have mapped stall for milliseconds while
 * NVLink copies are in flight, profiles/r02_vmm_lab_2gpu.jsonl). */
#define TFW_VS_HOME_DRIVEN 0x20u
/* A prefetched region's VA is re-pointed when its copy has COMPLETED (the access that needs it waits on the host) instead
 * of when the copy is issued (the client stream waits on the GPU).  cuMemSetAccess on an allocation that a copy is
 * writing takes several times longer (profiles/r02_tier_2gpu_variants.jsonl): a sweep that is bound by the VMM calls
 * prefers this, a worker that must not block its host thread does not. */
#define TFW_VS_REMAP_LATE 0x40u
/* Fixed frames: the home budget is cut into frames that are created and mapped when the space is created -- frame f at
 * the VA of EVERY region r with r % frames == f (CUDA virtual aliasing) -- and never re-mapped: a region is resident when
 * its frame holds its bytes, bringing it home evicts whoever lives in that frame, and no migration makes a VMM call
 * (cuMemUnmap / cuMemSetAccess wait for NVLink copies in flight: ~1 ms per miss under load, profiles/r02_tier_c5_*).
 * The price is the replacement policy: direct-mapped instead of LRU.  For a streaming working set (a sweep) the two
 * choose the same victims; a random-access working set sees conflict misses.  Peer-resident regions are not addressable
 * in place (not with TFW_VS_PEER_IN_PLACE); tfw_vspace_migrate brings a region home only into a free or leaving frame. */
#define TFW_VS_FIXED_FRAMES 0x80u

typedef struct {
  uint32_t struct_size;
  int32_t home_device;
  uint64_t va_bytes;           /* size of the vGPU address space */
  uint64_t region_bytes;       /* tiering granule, multiple of 2 MiB */
  uint64_t home_budget_bytes;  /* HBM the vGPU may keep resident on its own GPU */
  uint64_t peer_budget_bytes;  /* per peer GPU */
  uint64_t host_budget_bytes;  /* pinned host DRAM, allocated up front */
  int32_t peer_devices[TFW_VRAM_MAX_PEERS];
  uint32_t n_peers;
  uint32_t flags;
  uint32_t prefetch_ahead;     /* on a sequential sweep start fetching this many following regions early (0 = off, max 8);
                                  the policy keeps prefetch_ahead + 1 regions of the home budget free or being freed */
} tfw_vspace_config;

typedef struct {
  uint64_t regions_home, regions_peer, regions_host;
  uint64_t evict_bytes_peer, prefetch_bytes_peer; /* over NVLink */
  uint64_t evict_bytes_host, prefetch_bytes_host; /* over PCIe */
  uint64_t mover_launches;                        /* P2P copy kernels */
  uint64_t remaps;                                /* cuMemMap/Unmap pairs */
  uint64_t policy_evictions, policy_prefetches, policy_hits;
  uint64_t policy_hits_inflight;  /* accesses that found their region already on its way in (prefetched ahead) */
  uint64_t policy_prefetch_ahead; /* prefetches started before the region was asked for */
  uint64_t stall_ns;              /* host time spent waiting for a migration to complete */
  uint64_t phys_created, phys_destroyed; /* cuMemCreate / cuMemRelease of region backings (the pool should keep both near zero in steady state) */
  uint64_t vmm_ns;                /* host time inside cuMemMap / cuMemUnmap / cuMemSetAccess / cuMemCreate / cuMemRelease */
} tfw_vspace_stats;

typedef struct {
  uint64_t bytes;     /* bytes moved by this call */
  float copy_ms;      /* CUDA-event time of the copy (kernel or DMA) alone */
  float total_ms;     /* host wall-clock including allocation and re-mapping */
  uint32_t launches;  /* kernels launched by this call */
  uint32_t pad;
} tfw_migrate_result;

TFW_API tfw_status tfw_vspace_create(const tfw_vspace_config* cfg, tfw_vspace** out);
TFW_API tfw_status tfw_vspace_destroy(tfw_vspace* vs);
TFW_API const char* tfw_vspace_last_error(const tfw_vspace* vs);
/* base device pointer of the address space; region i lives at base + i*region_bytes for ever */
TFW_API tfw_status tfw_vspace_info(tfw_vspace* vs, uint64_t* base, uint64_t* region_bytes, uint32_t* n_regions);
/* give an unbacked region zero-filled backing at `tier` (peer_slot indexes cfg.peer_devices) */
TFW_API tfw_status tfw_vspace_populate(tfw_vspace* vs, uint32_t region, uint32_t tier, int32_t peer_slot);
/* move `n` regions to new tiers in one batch (one mover launch moves all of them, striped
 * over their targets); tiers[i] in {HOME, PEER, HOST}; peer_slots[i] used when PEER */
TFW_API tfw_status tfw_vspace_migrate(tfw_vspace* vs, const uint32_t* regions, const uint8_t* tiers,
                                      const int32_t* peer_slots, uint32_t n, tfw_migrate_result* res);
TFW_API tfw_status tfw_vspace_residency(tfw_vspace* vs, uint32_t region, uint32_t* tier, int32_t* device);
/* policy entry point: the vGPU is about to touch `region` -- make it HOME-resident, evicting
 * least-recently-used HOME regions to the emptiest peer (else host) as needed.  Migrations are
 * asynchronous: the wanted region's copy and the evictions that keep room for the next misses are only
 * enqueued (prefetch and eviction run in opposite NVLink directions at once), the region's VA is
 * re-pointed at once and the bound client stream waits for the bytes on the GPU.  Without a bound
 * stream the call returns when the wanted region has arrived (evictions still finish in the background). */
TFW_API tfw_status tfw_vspace_access(tfw_vspace* vs, uint32_t region);
/* The vGPU's execution stream (cudaStream_t on the home GPU).  With it the library orders migrations
 * against the client's kernels on the GPU (events on that stream) instead of requiring the caller to
 * drain it: an eviction waits for the kernels enqueued before the access that followed the region's
 * last touch, a prefetched region's first use waits for its copy.  Contract: a region is touched
 * (tfw_vspace_access) before work that uses it is enqueued, and that work goes to this stream. */
TFW_API tfw_status tfw_vspace_bind_stream(tfw_vspace* vs, void* cuda_stream);
/* wait for every migration in flight and finish its book-keeping */
TFW_API tfw_status tfw_vspace_quiesce(tfw_vspace* vs);
/* The policy path as a client sees it, in one native loop (benchmark + verification): for each of `count`
 * regions starting at `first` (wrapping at the end of the space): tfw_vspace_access, then a digest kernel over
 * the whole region on the client stream (the bound one, else an internal one).  digests[i] receives the
 * digest of the i-th region visited (compare with the oracle's); *seconds = wall-clock of the loop
 * including the final synchronise. */
TFW_API tfw_status tfw_vspace_sweep(tfw_vspace* vs, uint32_t first, uint32_t count, uint64_t* digests, double* seconds);
/* drop a region's backing (its bytes are lost; the backing returns to the pool) */
TFW_API tfw_status tfw_vspace_unpopulate(tfw_vspace* vs, uint32_t region);
/* pinned regions are never chosen as eviction victims by tfw_vspace_access; pins count (pinned != 0 adds one, 0 removes one) */
TFW_API tfw_status tfw_vspace_pin(tfw_vspace* vs, uint32_t region, int pinned);
TFW_API tfw_status tfw_vspace_get_stats(tfw_vspace* vs, tfw_vspace_stats* out);
/* verification helpers, executed on the home GPU through the region's VA (a PEER region is
 * read over NVLink; a HOST region answers TFW_ERR_NOT_SUPPORTED until prefetched):
 * word i of the region = mix64(seed + (i+1)*K1)  (tfw_digest64's mixer) */
TFW_API tfw_status tfw_vspace_fill_pattern(tfw_vspace* vs, uint32_t region, uint64_t seed);
TFW_API tfw_status tfw_vspace_digest(tfw_vspace* vs, uint32_t region, uint64_t* digest);
TFW_API tfw_status tfw_vspace_read(tfw_vspace* vs, uint32_t region, uint64_t off, void* dst, uint64_t n);
TFW_API tfw_status tfw_vspace_write(tfw_vspace* vs, uint32_t region, uint64_t off, const void* src, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* TFW_VRAM_H */
