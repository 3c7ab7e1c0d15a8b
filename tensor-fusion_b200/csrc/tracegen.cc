// tracegen.cc -- synthetic TFCS traces, host only.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "tfw_trace.h"

namespace {

struct SplitMix64 {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};

struct Xoshiro256ss {
  uint64_t s[4];
  explicit Xoshiro256ss(uint64_t seed) {
    SplitMix64 sm{seed};
    for (auto& x : s) x = sm.next();
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  uint64_t below(uint64_t n) { return n ? next() % n : 0; }
  // log-uniform integer in [lo, hi]
  uint64_t log_uniform(uint64_t lo, uint64_t hi) {
    const double u = (double)(next() >> 11) * (1.0 / 9007199254740992.0);
    const double v = std::exp(std::log((double)lo) + u * (std::log((double)hi) - std::log((double)lo)));
    uint64_t r = (uint64_t)v;
    return r < lo ? lo : (r > hi ? hi : r);
  }
};

void payload_fill(uint64_t seed, uint32_t call_id, uint8_t* dst, uint64_t n) {
  Xoshiro256ss g(seed + call_id);
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const uint64_t w = g.next();
    std::memcpy(dst + i, &w, 8);
  }
  if (i < n) {
    const uint64_t w = g.next();
    std::memcpy(dst + i, &w, n - i);
  }
}

// Appends frames into a caller buffer; keeps counting when the buffer is full.
struct Writer {
  uint8_t* out;
  size_t cap, pos = 0;
  uint32_t call_id = 0;
  bool fits(size_t n) const { return out && pos + n <= cap; }
  tfcs_frame_hdr mk(uint16_t op) {
    tfcs_frame_hdr h{};
    h.magic = TFCS_MAGIC;
    h.version = TFCS_VERSION;
    h.opcode = op;
    h.call_id = call_id++;
    return h;
  }
  // returns pointer to the payload area (or nullptr when only counting)
  uint8_t* put(const tfcs_frame_hdr& h, uint64_t payload) {
    const size_t total = TFCS_HDR_BYTES + (size_t)tfcs_pad16(payload);
    uint8_t* p = nullptr;
    if (fits(total)) {
      std::memcpy(out + pos, &h, TFCS_HDR_BYTES);
      p = out + pos + TFCS_HDR_BYTES;
      if (tfcs_pad16(payload) > payload) std::memset(p + payload, 0, (size_t)(tfcs_pad16(payload) - payload));
    }
    pos += total;
    return p;
  }
};

struct Live { uint32_t h; uint64_t size; };

}  // namespace

extern "C" {

void tfw_trace_payload(uint64_t seed, uint32_t call_id, void* dst, uint64_t nbytes) {
  if (dst && nbytes) payload_fill(seed, call_id, static_cast<uint8_t*>(dst), nbytes);
}

void tfw_trace_c1_defaults(tfw_trace_c1_params* p) {
  if (!p) return;
  p->seed = TFW_TRACE_SEED_C1;
  p->ncalls = 1000;
  p->max_live = 64;
  p->max_buffer_bytes = 8ull << 20;
  p->max_payload_bytes = 4ull << 20;
  p->unaligned_percent = 25;
  p->error_permille = 10;
  p->launch_cost = 0;
  p->reserved = 0;
}

tfw_status tfw_trace_gen_c1(const tfw_trace_c1_params* p, void* out, size_t cap, size_t* nbytes) {
  if (!p || !nbytes || p->max_live < 2 || p->max_buffer_bytes < 4096 || p->max_payload_bytes < 64) return TFW_ERR_INVALID;
  Writer w{static_cast<uint8_t*>(out), cap};
  Xoshiro256ss rng(p->seed ^ 0xC1C1C1C1ull);
  std::vector<Live> live;
  uint32_t next_handle = 1;
  auto pick = [&]() -> Live& { return live[rng.below(live.size())]; };
  // (off, len) inside a buffer of `size`; `unal` forces both off 16-byte alignment
  auto range = [&](uint64_t size, bool unal, uint64_t* off, uint64_t* len) {
    uint64_t L = rng.log_uniform(64, p->max_payload_bytes);
    if (L > size) L = size;
    uint64_t O = rng.below(size - L + 1);
    if (!unal) { O &= ~15ull; if (L >= 32) L &= ~15ull; }
    else {
      O |= 1 + rng.below(15);
      if ((L & 15) == 0 && L > 16) L -= 1 + rng.below(15);
    }
    if (O >= size) O = size - 1;
    if (O + L > size) L = size - O;
    *off = O; *len = L;
  };
  for (uint32_t i = 0; i < p->ncalls; ++i) {
    uint32_t r = (uint32_t)rng.below(100);
    int op = r < 40 ? TFCS_OP_MEMCPY_H2D : r < 50 ? TFCS_OP_MEMCPY_D2H : r < 60 ? TFCS_OP_MEMCPY_D2D
             : r < 65 ? TFCS_OP_MEMSET : r < 70 ? TFCS_OP_MALLOC : r < 75 ? TFCS_OP_FREE : TFCS_OP_LAUNCH;
    if (live.size() < 2 || i < 16) op = TFCS_OP_MALLOC;  // warm start: 16 buffers
    if (op == TFCS_OP_MALLOC && live.size() >= p->max_live) op = TFCS_OP_FREE;
    if (op == TFCS_OP_FREE && live.size() <= 2) op = TFCS_OP_MALLOC;
    const bool inject_error = p->error_permille && rng.below(1000) < p->error_permille;
    const bool unal = rng.below(100) < p->unaligned_percent;
    switch (op) {
      case TFCS_OP_MALLOC: {
        tfcs_frame_hdr h = w.mk(TFCS_OP_MALLOC);
        h.h0 = inject_error && !live.empty() ? pick().h : next_handle++;  // error: handle already live
        h.length = rng.log_uniform(4096, p->max_buffer_bytes);
        if (unal) h.length |= 1 + rng.below(15);
        w.put(h, 0);
        if (!inject_error || live.empty()) live.push_back(Live{h.h0, h.length});
        break;
      }
      case TFCS_OP_FREE: {
        tfcs_frame_hdr h = w.mk(TFCS_OP_FREE);
        if (inject_error) { h.h0 = next_handle + 1000; w.put(h, 0); break; }
        const size_t k = rng.below(live.size());
        h.h0 = live[k].h;
        w.put(h, 0);
        live[k] = live.back();
        live.pop_back();
        break;
      }
      case TFCS_OP_MEMCPY_H2D: {
        Live& b = pick();
        tfcs_frame_hdr h = w.mk(TFCS_OP_MEMCPY_H2D);
        h.h0 = b.h;
        range(b.size, unal, &h.off0, &h.length);
        if (inject_error) { if (rng.below(2)) h.h0 = next_handle + 1000; else h.off0 = b.size - h.length + 1 + rng.below(64); }
        uint8_t* pay = w.put(h, h.length);
        if (pay) payload_fill(p->seed, h.call_id, pay, h.length);
        break;
      }
      case TFCS_OP_MEMCPY_D2H: {
        Live& b = pick();
        tfcs_frame_hdr h = w.mk(TFCS_OP_MEMCPY_D2H);
        h.h0 = b.h;
        range(b.size, unal, &h.off0, &h.length);
        if (h.length > (256u << 10)) h.length = (256u << 10) - (unal ? 3 : 0);  // keep the response stream small
        if (inject_error) h.off0 = b.size + 1;
        w.put(h, 0);
        break;
      }
      case TFCS_OP_MEMCPY_D2D: {
        Live& d = pick();
        Live& s = pick();
        tfcs_frame_hdr h = w.mk(TFCS_OP_MEMCPY_D2D);
        h.h0 = d.h; h.h1 = s.h;
        uint64_t lo, ln;
        range(d.size < s.size ? d.size : s.size, unal, &lo, &ln);
        h.length = ln;
        h.off0 = rng.below(d.size - ln + 1);
        h.off1 = rng.below(s.size - ln + 1);
        if (!unal) { h.off0 &= ~15ull; h.off1 &= ~15ull; }
        if (d.h == s.h) {  // same buffer: make the ranges disjoint (split in halves) unless an error is wanted
          const uint64_t half = d.size / 2;
          if (!inject_error) {
            if (ln > half) { ln = half; h.length = ln; }
            h.off0 = rng.below(half - ln + 1);
            h.off1 = half + rng.below(d.size - half - ln + 1);
          } else { h.off0 = 0; h.off1 = ln / 2; }
        } else if (inject_error) { h.off1 = s.size - ln + 1; }
        w.put(h, 0);
        break;
      }
      case TFCS_OP_MEMSET: {
        Live& b = pick();
        tfcs_frame_hdr h = w.mk(TFCS_OP_MEMSET);
        h.h0 = b.h;
        range(b.size, unal, &h.off0, &h.length);
        h.arg0 = (uint32_t)rng.below(256);
        if (inject_error) h.h0 = next_handle + 1000;
        w.put(h, 0);
        break;
      }
      default: {
        tfcs_frame_hdr h = w.mk(TFCS_OP_LAUNCH);
        const uint32_t k = (uint32_t)rng.below(100);
        h.arg0 = k < 70 ? TFCS_KERNEL_NOOP : k < 85 ? TFCS_KERNEL_ADD_U8 : TFCS_KERNEL_XOR_IDX;
        h.arg1 = 1 + (uint32_t)rng.below(296);
        h.arg2 = 32u << rng.below(4);
        h.arg3 = p->launch_cost;
        if (h.arg0 != TFCS_KERNEL_NOOP) {
          Live& b = pick();
          h.h0 = b.h;
          range(b.size, unal, &h.off0, &h.length);
          h.off1 = 1 + rng.below(1u << 20);
        }
        if (inject_error) h.arg0 = 99;
        w.put(h, 0);
        break;
      }
    }
  }
  w.put(w.mk(TFCS_OP_SYNC), 0);
  *nbytes = w.pos;
  return (out && w.pos <= cap) || !out ? TFW_OK : TFW_ERR_EXHAUSTED;
}

tfw_status tfw_trace_gen_bulk(uint64_t seed, uint32_t nbuf, uint32_t ncopies, uint64_t bytes_each, uint32_t nthreads,
                              void* out, size_t cap, size_t* nbytes) {
  if (!nbytes || !nbuf || !bytes_each) return TFW_ERR_INVALID;
  Writer w{static_cast<uint8_t*>(out), cap};
  for (uint32_t b = 0; b < nbuf; ++b) {
    tfcs_frame_hdr h = w.mk(TFCS_OP_MALLOC);
    h.h0 = b + 1;
    h.length = bytes_each;
    w.put(h, 0);
  }
  struct Job { uint8_t* p; uint32_t call_id; };
  std::vector<Job> jobs;
  for (uint32_t c = 0; c < ncopies; ++c) {
    tfcs_frame_hdr h = w.mk(TFCS_OP_MEMCPY_H2D);
    h.h0 = (c % nbuf) + 1;
    h.length = bytes_each;
    uint8_t* pay = w.put(h, bytes_each);
    if (pay) jobs.push_back(Job{pay, h.call_id});
  }
  {
    tfcs_frame_hdr h = w.mk(TFCS_OP_LAUNCH);
    h.arg0 = TFCS_KERNEL_NOOP; h.arg1 = 1; h.arg2 = 32;
    w.put(h, 0);
  }
  w.put(w.mk(TFCS_OP_SYNC), 0);
  *nbytes = w.pos;
  if (out && w.pos > cap) return TFW_ERR_EXHAUSTED;
  if (!jobs.empty()) {
    if (!nthreads) nthreads = 1;
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] { for (size_t j = t; j < jobs.size(); j += nthreads) payload_fill(seed, jobs[j].call_id, jobs[j].p, bytes_each); });
    for (auto& x : th) x.join();
  }
  return TFW_OK;
}

tfw_status tfw_trace_gen_small(uint64_t seed, uint32_t ncalls, uint64_t bytes_each, void* out, size_t cap, size_t* nbytes) {
  if (!nbytes || !bytes_each) return TFW_ERR_INVALID;
  Writer w{static_cast<uint8_t*>(out), cap};
  {
    tfcs_frame_hdr h = w.mk(TFCS_OP_MALLOC);
    h.h0 = 1;
    h.length = bytes_each;
    w.put(h, 0);
  }
  for (uint32_t c = 0; c < ncalls; ++c) {
    tfcs_frame_hdr h = w.mk(TFCS_OP_MEMCPY_H2D);
    h.h0 = 1;
    h.length = bytes_each;
    uint8_t* pay = w.put(h, bytes_each);
    if (pay) payload_fill(seed, h.call_id, pay, bytes_each);
    tfcs_frame_hdr l = w.mk(TFCS_OP_LAUNCH);
    l.arg0 = TFCS_KERNEL_NOOP; l.arg1 = 1; l.arg2 = 32;
    w.put(l, 0);
  }
  w.put(w.mk(TFCS_OP_SYNC), 0);
  *nbytes = w.pos;
  return (out && w.pos > cap) ? TFW_ERR_EXHAUSTED : TFW_OK;
}

}  // extern "C"
