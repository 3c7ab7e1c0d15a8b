/*
 * replay_oracle.c -- sequential CPU replay of a TFCS command stream.
 * TEST INFRASTRUCTURE (see tfo_oracle.h).  "PARITY UNPINNED": the reference's
 * worker (image tensorfusion/tensor-fusion-worker:latest,
 * charts/tensor-fusion/values.yaml:166) is closed source; call sites only:
 * internal/utils/compose.go:1304-1325.  Handle<->pointer indirection follows
 * the mock driver precedent provider/example/device_mock/driver_mock.c:306-355
 * (hipMalloc/hipFree keep an allocation table and account VRAM against a cap).
 *
 * Semantics (DESIGN.md "trace semantics"): frames execute strictly in order;
 * MALLOC zero-fills; every rejected frame produces a RESP_ERROR and has no
 * other effect; D2H / SYNC produce responses in issue order.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tfw_wire.h"
#include "tfo_oracle.h"

enum { ST_INVALID = 1, ST_NOT_FOUND = 2, ST_NOT_SUPPORTED = 3, ST_EXHAUSTED = 4, ST_PROTOCOL = 7 };

typedef struct { uint8_t* p; uint64_t size; int live; } obuf;

struct tfo_session {
  obuf* bufs;
  uint32_t nbufs;
  uint8_t* resp;
  size_t resp_len, resp_cap;
  uint64_t frames, payload, live, vram, errors;
};

/* ---- the CPU engine behind big copies and zero-fills --------------------------------------
 * A CPU implementation of this path that is meant to be fast keeps its threads and its memory:
 *   - a persistent pool of worker threads (tfo_set_threads) splits every copy / fill of 4 MiB or
 *     more into page-aligned slices;
 *   - with the buffer cache on (tfo_set_buffer_cache) a freed buffer keeps its pages and the next
 *     MALLOC of that size re-uses them (zero-filled by the pool), the way cudaMallocAsync's pool
 *     does on the GPU -- no fresh mmap + first-touch page faults per session.
 * Both are off by default: the parity tests replay small traces sequentially. */
static int g_threads = 1;
typedef struct { uint8_t* d; const uint8_t* s; uint64_t n; int fill; } copy_job;
enum { MAXT = 256 };
static pthread_t g_pool[MAXT];
static int g_pool_n = 0, g_pool_stop = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_work = PTHREAD_COND_INITIALIZER, g_done = PTHREAD_COND_INITIALIZER;
static copy_job g_jobs[MAXT];
static int g_njobs = 0, g_next = 0, g_pending = 0;

static void run_job(const copy_job* j) {
  if (j->s) memcpy(j->d, j->s, j->n);
  else memset(j->d, j->fill, j->n);
}
static volatile int g_posted = 0; /* bumped (under the lock) whenever jobs are posted: lets idle workers spin without the lock */
static void* pool_thread(void* a) {
  (void)a;
  pthread_mutex_lock(&g_mu);
  for (;;) {
    if (!g_pool_stop && g_next >= g_njobs) {
      /* a bulk stream posts the next copy within microseconds: look again for a short while before sleeping on the
       * condition variable (a futex wake costs about as much as copying a quarter of a megabyte) */
      const int seen = g_posted;
      pthread_mutex_unlock(&g_mu);
      for (int spin = 0; spin < 20000 && g_posted == seen && !g_pool_stop; ++spin) __builtin_ia32_pause();
      pthread_mutex_lock(&g_mu);
    }
    while (!g_pool_stop && g_next >= g_njobs) pthread_cond_wait(&g_work, &g_mu);
    if (g_pool_stop) break;
    const copy_job j = g_jobs[g_next++];
    pthread_mutex_unlock(&g_mu);
    run_job(&j);
    pthread_mutex_lock(&g_mu);
    if (--g_pending == 0) pthread_cond_signal(&g_done);
  }
  pthread_mutex_unlock(&g_mu);
  return NULL;
}
void tfo_set_threads(int n) {
  n = n > 1 ? (n > MAXT ? MAXT : n) : 1;
  if (g_pool_n) {  // resize: stop the old pool
    pthread_mutex_lock(&g_mu);
    g_pool_stop = 1;
    pthread_cond_broadcast(&g_work);
    pthread_mutex_unlock(&g_mu);
    for (int i = 0; i < g_pool_n; ++i) pthread_join(g_pool[i], NULL);
    g_pool_n = 0;
    g_pool_stop = 0;
  }
  g_threads = n;
  for (int i = 0; i < n - 1; ++i)  // the calling thread takes a slice too
    if (pthread_create(&g_pool[g_pool_n], NULL, pool_thread, NULL) == 0) g_pool_n++;
}
/* copy (s != NULL) or fill n bytes, split over the pool */
static void big_op(uint8_t* d, const uint8_t* s, int fill, uint64_t n) {
  if (g_threads <= 1 || g_pool_n == 0 || n < (4u << 20)) { copy_job j = {d, s, n, fill}; run_job(&j); return; }
  const int nt = g_pool_n + 1;
  const uint64_t per = ((n / (uint64_t)nt) + 4095) & ~(uint64_t)4095;
  copy_job mine = {0, 0, 0, 0};
  pthread_mutex_lock(&g_mu);
  g_njobs = 0; g_next = 0;
  for (int t = 0; t < nt; ++t) {
    const uint64_t o = (uint64_t)t * per;
    if (o >= n) break;
    copy_job j = {d + o, s ? s + o : NULL, n - o < per ? n - o : per, fill};
    if (t == 0) mine = j; else g_jobs[g_njobs++] = j;
  }
  g_pending = g_njobs;
  g_posted++;
  pthread_cond_broadcast(&g_work);
  pthread_mutex_unlock(&g_mu);
  run_job(&mine);
  for (int spin = 0; spin < 20000; ++spin) {  /* the helpers finish within microseconds of this thread */
    if (__atomic_load_n(&g_pending, __ATOMIC_ACQUIRE) == 0) break;
    __builtin_ia32_pause();
  }
  pthread_mutex_lock(&g_mu);
  while (g_pending) pthread_cond_wait(&g_done, &g_mu);
  g_njobs = 0; g_next = 0;
  pthread_mutex_unlock(&g_mu);
}
static void big_copy(uint8_t* d, const uint8_t* s, uint64_t n) { big_op(d, s, 0, n); }

/* buffer cache: freed blocks by exact size, bounded */
enum { CACHE_SLOTS = 256 };
static int g_cache_on = 0;
static struct { uint8_t* p; uint64_t size; } g_cache[CACHE_SLOTS];
void tfo_set_buffer_cache(int on) {
  g_cache_on = on;
  if (!on) for (int i = 0; i < CACHE_SLOTS; ++i) { free(g_cache[i].p); g_cache[i].p = NULL; g_cache[i].size = 0; }
}
static uint8_t* buf_alloc(uint64_t n, int zero) {
  if (g_cache_on)
    for (int i = 0; i < CACHE_SLOTS; ++i)
      if (g_cache[i].p && g_cache[i].size == n) {
        uint8_t* p = g_cache[i].p;
        g_cache[i].p = NULL;
        if (zero) big_op(p, NULL, 0, n);  /* MALLOC hands out zeros whatever the block held before */
        return p;
      }
  return (uint8_t*)(zero ? calloc(1, n) : malloc(n));
}
static void buf_free(uint8_t* p, uint64_t n) {
  if (g_cache_on)
    for (int i = 0; i < CACHE_SLOTS; ++i)
      if (!g_cache[i].p) { g_cache[i].p = p; g_cache[i].size = n; return; }
  free(p);
}

static void resp_put(tfo_session* s, const tfcs_frame_hdr* h, const uint8_t* pay, uint64_t len) {
  const size_t need = TFCS_HDR_BYTES + (size_t)tfcs_pad16(len);
  if (s->resp_len + need > s->resp_cap) {
    size_t nc = s->resp_cap ? s->resp_cap * 2 : 1 << 16;
    while (nc < s->resp_len + need) nc *= 2;
    s->resp = (uint8_t*)realloc(s->resp, nc);
    s->resp_cap = nc;
  }
  memcpy(s->resp + s->resp_len, h, TFCS_HDR_BYTES);
  if (len) memcpy(s->resp + s->resp_len + TFCS_HDR_BYTES, pay, len);
  memset(s->resp + s->resp_len + TFCS_HDR_BYTES + len, 0, (size_t)(tfcs_pad16(len) - len));
  s->resp_len += need;
}

static void resp_error(tfo_session* s, const tfcs_frame_hdr* h, uint32_t code) {
  tfcs_frame_hdr r = *h;
  r.opcode = TFCS_OP_RESP_ERROR;
  r.arg0 = code;
  r.arg1 = h->opcode;
  r.length = 0;
  resp_put(s, &r, NULL, 0);
  s->errors++;
}

static obuf* find(tfo_session* s, uint32_t h) {
  return (h < s->nbufs && s->bufs[h].live) ? &s->bufs[h] : NULL;
}

static int in_range(const obuf* b, uint64_t off, uint64_t len) { return off <= b->size && len <= b->size - off; }

int tfo_replay(const void* stream, size_t nbytes, uint64_t vram_limit, uint32_t flags, tfo_session** out) {
  if (!out) return ST_INVALID;
  tfo_session* s = (tfo_session*)calloc(1, sizeof(*s));
  *out = s;
  const uint8_t* p = (const uint8_t*)stream;
  size_t pos = 0;
  while (pos < nbytes) {
    if (nbytes - pos < TFCS_HDR_BYTES) return ST_PROTOCOL;
    tfcs_frame_hdr h;
    memcpy(&h, p + pos, sizeof h);
    if (h.magic != TFCS_MAGIC || h.version != TFCS_VERSION) return ST_PROTOCOL;
    pos += TFCS_HDR_BYTES;
    switch (h.opcode) {
      case TFCS_OP_NOP: break;
      case TFCS_OP_MALLOC: {
        if (h.h0 >= TFCS_MAX_HANDLES || h.length == 0 || h.length > TFCS_MAX_BUFFER_BYTES) { resp_error(s, &h, ST_INVALID); break; }
        if (find(s, h.h0)) { resp_error(s, &h, ST_INVALID); break; }
        if (vram_limit && s->vram + h.length > vram_limit) { resp_error(s, &h, ST_EXHAUSTED); break; }
        if (h.h0 >= s->nbufs) {
          s->bufs = (obuf*)realloc(s->bufs, sizeof(obuf) * (h.h0 + 1));
          memset(s->bufs + s->nbufs, 0, sizeof(obuf) * (h.h0 + 1 - s->nbufs));
          s->nbufs = h.h0 + 1;
        }
        uint8_t* m = buf_alloc(h.length, !(flags & 0x4u));
        if (!m) { resp_error(s, &h, ST_EXHAUSTED); break; }
        s->bufs[h.h0].p = m; s->bufs[h.h0].size = h.length; s->bufs[h.h0].live = 1;
        s->vram += h.length; s->live++;
        break;
      }
      case TFCS_OP_FREE: {
        obuf* b = find(s, h.h0);
        if (!b) { resp_error(s, &h, ST_NOT_FOUND); break; }
        buf_free(b->p, b->size);
        s->vram -= b->size; s->live--;
        memset(b, 0, sizeof *b);
        break;
      }
      case TFCS_OP_MEMCPY_H2D: {
        const uint64_t padded = tfcs_pad16(h.length);
        if (padded > nbytes - pos) return ST_PROTOCOL;
        obuf* b = find(s, h.h0);
        if (!b) resp_error(s, &h, ST_NOT_FOUND);
        else if (!in_range(b, h.off0, h.length)) resp_error(s, &h, ST_INVALID);
        else { big_copy(b->p + h.off0, p + pos, h.length); s->payload += h.length; }
        pos += (size_t)padded;
        break;
      }
      case TFCS_OP_MEMCPY_D2H: {
        obuf* b = find(s, h.h0);
        if (!b) { resp_error(s, &h, ST_NOT_FOUND); break; }
        if (!in_range(b, h.off0, h.length)) { resp_error(s, &h, ST_INVALID); break; }
        tfcs_frame_hdr r = h;
        r.opcode = TFCS_OP_RESP_D2H;
        resp_put(s, &r, b->p + h.off0, h.length);
        break;
      }
      case TFCS_OP_MEMCPY_D2D: {
        obuf* d = find(s, h.h0);
        obuf* sb = find(s, h.h1);
        if (!d || !sb) { resp_error(s, &h, ST_NOT_FOUND); break; }
        if (!in_range(d, h.off0, h.length) || !in_range(sb, h.off1, h.length)) { resp_error(s, &h, ST_INVALID); break; }
        const uint8_t* sp = sb->p + h.off1;
        uint8_t* dp = d->p + h.off0;
        if (h.length && dp < sp + h.length && sp < dp + h.length) { resp_error(s, &h, ST_INVALID); break; }
        big_copy(dp, sp, h.length);
        break;
      }
      case TFCS_OP_MEMSET: {
        obuf* b = find(s, h.h0);
        if (!b) { resp_error(s, &h, ST_NOT_FOUND); break; }
        if (!in_range(b, h.off0, h.length)) { resp_error(s, &h, ST_INVALID); break; }
        big_op(b->p + h.off0, NULL, (int)(h.arg0 & 0xff), h.length);
        break;
      }
      case TFCS_OP_LAUNCH: {
        if (h.arg0 > TFCS_KERNEL_XOR_IDX) { resp_error(s, &h, ST_NOT_SUPPORTED); break; }
        uint8_t* r = NULL;
        if (h.length) {
          obuf* b = find(s, h.h0);
          if (!b) { resp_error(s, &h, ST_NOT_FOUND); break; }
          if (!in_range(b, h.off0, h.length)) { resp_error(s, &h, ST_INVALID); break; }
          r = b->p + h.off0;
        }
        if (h.arg0 == TFCS_KERNEL_ADD_U8) {
          const uint8_t dlt = (uint8_t)(h.off1 & 0xff);
          for (uint64_t i = 0; i < h.length; ++i) r[i] = (uint8_t)(r[i] + dlt);
        } else if (h.arg0 == TFCS_KERNEL_XOR_IDX) {
          for (uint64_t i = 0; i < h.length; ++i) r[i] ^= (uint8_t)((i * h.off1) >> 3);
        }
        break;
      }
      case TFCS_OP_SYNC: {
        tfcs_frame_hdr r = h;
        r.opcode = TFCS_OP_RESP_SYNC;
        r.arg0 = 0;
        r.length = 0;
        resp_put(s, &r, NULL, 0);
        break;
      }
      default:
        /* Frames that carry a payload the oracle does not interpret (user modules: a CPU cannot run a cubin; a
         * response opcode sent by a client): their bytes are passed over, as the worker's parser does, and the
         * frame is refused.  By-reference copies and arenas need memory shared with a client: refused too. */
        if (tfcs_has_payload(h.opcode)) {
          const uint64_t padded = tfcs_pad16(h.length);
          if (padded > nbytes - pos) return ST_PROTOCOL;
          pos += (size_t)padded;
        }
        resp_error(s, &h, ST_NOT_SUPPORTED);
        break;
    }
    s->frames++;
  }
  return 0;
}

size_t tfo_responses(const tfo_session* s, const uint8_t** p) { if (p) *p = s->resp; return s->resp_len; }
int tfo_buffer(const tfo_session* s, uint32_t handle, const uint8_t** p, uint64_t* size) {
  if (!s || handle >= s->nbufs || !s->bufs[handle].live) return ST_NOT_FOUND;
  if (p) *p = s->bufs[handle].p;
  if (size) *size = s->bufs[handle].size;
  return 0;
}
uint64_t tfo_stat(const tfo_session* s, int which) {
  switch (which) { case 0: return s->frames; case 1: return s->payload; case 2: return s->live; case 3: return s->vram; default: return s->errors; }
}
void tfo_free(tfo_session* s) {
  if (!s) return;
  for (uint32_t i = 0; i < s->nbufs; ++i) if (s->bufs[i].live) buf_free(s->bufs[i].p, s->bufs[i].size);
  free(s->bufs);
  free(s->resp);
  free(s);
}

/* ---- digest ------------------------------------------------------------- */
static uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t tfo_digest(const void* p, uint64_t n) {
  const uint8_t* b = (const uint8_t*)p;
  const uint64_t K1 = 0x9E3779B97F4A7C15ull;
  const uint64_t nw = n >> 3;
  uint64_t acc = 0;
  for (uint64_t i = 0; i < nw; ++i) {
    uint64_t w;
    memcpy(&w, b + 8 * i, 8);
    acc += mix64(w ^ ((i + 1) * K1));
  }
  if (n & 7) {
    uint64_t w = 0;
    memcpy(&w, b + 8 * nw, n & 7);
    acc += mix64(w ^ ((nw + 1) * K1));
  }
  return mix64(acc ^ (n * K1));
}

/* ---- payload stream: splitmix64-seeded xoshiro256** ---------------------- */
static uint64_t sm_next(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t tfo_splitmix64_nth(uint64_t seed, uint32_t n) {
  uint64_t v = 0;
  for (uint32_t i = 0; i <= n; ++i) v = sm_next(&seed);
  return v;
}
static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void tfo_payload(uint64_t seed, uint32_t call_id, void* dst, uint64_t n) {
  uint64_t st = seed + call_id, s[4];
  for (int i = 0; i < 4; ++i) s[i] = sm_next(&st);
  uint8_t* d = (uint8_t*)dst;
  for (uint64_t i = 0; i < n; i += 8) {
    const uint64_t r = rotl64(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    memcpy(d + i, &r, n - i < 8 ? (size_t)(n - i) : 8);
  }
}

/* ---- test pattern of the VRAM tiering tests: word i = mix64(seed + (i+1)*K1) ---- */
void tfo_pattern(uint64_t seed, void* dst, uint64_t nbytes) {
  const uint64_t K1 = 0x9E3779B97F4A7C15ull;
  uint64_t* w = (uint64_t*)dst;
  for (uint64_t i = 0; i < (nbytes >> 3); ++i) w[i] = mix64(seed + (i + 1) * K1);
}
