/*
 * shm_oracle.c -- CPU restatement of the soft-limiter quota file.
 * TEST INFRASTRUCTURE (see tfo_oracle.h).  Follows
 *   pkg/hypervisor/worker/state/soft_limiter_shm.go
 * function by function; offsets are written out numerically
 * instead of sharing include/tf_shm_layout.h with the product, so that a
 * layout slip on either side shows up as a diff.  Pinned by the reference's
 * golden vectors in soft_limiter_shm_test.go (tests/test_oracle_golden.py).
 */
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <sched.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "tfo_shm.h"

/* ---- layout constants (soft_limiter_shm.go:19-27,150-161,201-205,330-352) ---- */
enum {
  MAX_PROCESSES = 2048,
  MAX_DEVICES = 16,
  MAX_UUID_LEN = 64,
  OFF_DISCRIMINANT = 0,
  OFF_V2 = 8,            /* test-pinned: soft_limiter_shm_test.go:221 */
  ENTRY_BYTES = 136,
  E_UUID = 0, E_UP_LIMIT = 64, E_MEM_LIMIT = 72, E_TOTAL_CORES = 80, E_POD_MEM = 88,
  E_RATE = 96, E_CAPACITY = 104, E_TOKENS = 112, E_LAST_UPDATE = 120, E_ACTIVE = 128,
  OFF_DEVICE_COUNT = 0x888,
  OFF_HEARTBEAT = 0x890, /* test-pinned: :226 */
  OFF_PIDS = 0x898,      /* test-pinned: :231 */
  P_LOCK = 0, P_VALUES = 8, P_BITMAP = 8 + 8 * MAX_PROCESSES, P_LEN = 8 + 16 * MAX_PROCESSES,
  P_CREATOR = 16 + 16 * MAX_PROCESSES,
  FILE_BYTES = 0x898 + 24 + 16 * MAX_PROCESSES + 512, /* = 35504 */
  LEGACY_BYTES = FILE_BYTES - 8
};

size_t tfo_shm_file_bytes(void) { return FILE_BYTES; }
size_t tfo_shm_legacy_bytes(void) { return LEGACY_BYTES; }
size_t tfo_shm_offset(const char* what) {
  if (!strcmp(what, "v2")) return OFF_V2;
  if (!strcmp(what, "heartbeat")) return OFF_HEARTBEAT;
  if (!strcmp(what, "pids")) return OFF_PIDS;
  if (!strcmp(what, "device_count")) return OFF_DEVICE_COUNT;
  if (!strcmp(what, "entry")) return ENTRY_BYTES;
  if (!strcmp(what, "tokens")) return E_TOKENS;
  return (size_t)-1;
}

static uint8_t* entry(uint8_t* f, uint32_t idx) { return f + OFF_V2 + (size_t)idx * ENTRY_BYTES; }
static uint64_t* w64(uint8_t* p) { return (uint64_t*)(void*)p; }
static uint32_t* w32(uint8_t* p) { return (uint32_t*)(void*)p; }
static uint64_t f2b(double v) { uint64_t b; memcpy(&b, &v, 8); return b; }
static double b2f(uint64_t b) { double v; memcpy(&v, &b, 8); return v; }
static uint64_t ald(uint64_t* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static void ast(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static int cas(uint64_t* p, uint64_t e, uint64_t n) { return __atomic_compare_exchange_n(p, &e, n, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); }

/* Go math.Max / math.Min */
double tfo_go_max(double x, double y) {
  if (isinf(x) && x > 0) return x;
  if (isinf(y) && y > 0) return y;
  if (isnan(x) || isnan(y)) return b2f(0x7FF8000000000001ull);
  if (x == 0 && x == y) return signbit(x) ? y : x;
  return x > y ? x : y;
}
double tfo_go_min(double x, double y) {
  if (isinf(x) && x < 0) return x;
  if (isinf(y) && y < 0) return y;
  if (isnan(x) || isnan(y)) return b2f(0x7FF8000000000001ull);
  if (x == 0 && x == y) return signbit(x) ? x : y;
  return x < y ? x : y;
}

/* ---- SharedDeviceInfoV2 accessors (:654-748) ---- */
double tfo_shm_get(uint8_t* f, uint32_t idx, int field) {
  static const int off[] = {E_RATE, E_CAPACITY, E_TOKENS, E_LAST_UPDATE};
  return b2f(ald(w64(entry(f, idx) + off[field])));
}
void tfo_shm_set(uint8_t* f, uint32_t idx, int field, double v) {
  static const int off[] = {E_RATE, E_CAPACITY, E_TOKENS, E_LAST_UPDATE};
  ast(w64(entry(f, idx) + off[field]), f2b(v));
}
/* FetchSubERLTokens :715-731 */
double tfo_shm_fetch_sub(uint8_t* f, uint32_t idx, double cost) {
  uint64_t* t = w64(entry(f, idx) + E_TOKENS);
  for (;;) {
    uint64_t cb = ald(t);
    double cur = b2f(cb);
    if (cur < cost) return cur;
    double nv = tfo_go_max(0.0, cur - cost);
    if (cas(t, cb, f2b(nv))) return cur;
  }
}
/* FetchAddERLTokens :734-748 */
double tfo_shm_fetch_add(uint8_t* f, uint32_t idx, double amount) {
  double cap = tfo_shm_get(f, idx, 1);
  uint64_t* t = w64(entry(f, idx) + E_TOKENS);
  for (;;) {
    uint64_t cb = ald(t);
    double cur = b2f(cb);
    double nv = tfo_go_max(0.0, tfo_go_min(cap, cur + amount));
    if (cas(t, cb, f2b(nv))) return cur;
  }
}

/* ---- NewSharedDeviceStateV2 (:404-433) into a zeroed 35 504-byte image ---- */
int tfo_shm_init_image(uint8_t* f, const tfo_dev_cfg* cfgs, size_t n, uint64_t now, uint64_t creator_pid) {
  memset(f, 0, FILE_BYTES);
  for (size_t i = 0; i < n; ++i) if (cfgs[i].device_idx >= MAX_DEVICES) return 1;
  *w32(f + OFF_DISCRIMINANT) = 1; /* rustSharedDeviceStateV2Discriminant :26,949 */
  *w32(f + OFF_DEVICE_COUNT) = (uint32_t)n;
  *w64(f + OFF_HEARTBEAT) = now;
  *w64(f + OFF_PIDS + P_CREATOR) = creator_pid;
  for (size_t i = 0; i < n; ++i) {
    uint8_t* e = entry(f, cfgs[i].device_idx);
    size_t L = strlen(cfgs[i].uuid);
    if (L > MAX_UUID_LEN - 1) L = MAX_UUID_LEN - 1; /* SetUUID :262-275 */
    memset(e + E_UUID, 0, MAX_UUID_LEN);
    memcpy(e + E_UUID, cfgs[i].uuid, L);
    *w32(e + E_TOTAL_CORES) = cfgs[i].total_cuda_cores;
    *w32(e + E_UP_LIMIT) = cfgs[i].up_limit;
    *w64(e + E_MEM_LIMIT) = cfgs[i].mem_limit;
    *w64(e + E_CAPACITY) = f2b(100.0);
    *w64(e + E_RATE) = f2b(10.0);
    *w64(e + E_TOKENS) = f2b(100.0);
    *w64(e + E_LAST_UPDATE) = f2b((double)now);
    *w32(e + E_ACTIVE) = 1;
  }
  return 0;
}

int tfo_shm_has_device(uint8_t* f, uint32_t idx) { return idx < MAX_DEVICES && *w32(entry(f, idx) + E_ACTIVE) != 0; }
/* SetPodMemoryUsed :639-652 */
int tfo_shm_set_pod_memory_used(uint8_t* f, uint32_t idx, uint64_t v) {
  if (idx >= MAX_DEVICES || !tfo_shm_has_device(f, idx)) return 0;
  ast(w64(entry(f, idx) + E_POD_MEM), v);
  return 1;
}
uint64_t tfo_shm_pod_memory_used(uint8_t* f, uint32_t idx) { return ald(w64(entry(f, idx) + E_POD_MEM)); }
/* IsHealthy :521-534 */
int tfo_shm_is_healthy(uint8_t* f, uint64_t timeout_secs, uint64_t now) {
  uint64_t hb = ald(w64(f + OFF_HEARTBEAT));
  if (hb == 0) return 0;
  if (hb > now) return 0;
  return now - hb <= timeout_secs;
}

/* ---- PIDBitmap / PIDSet (:751-840), ShmMutex (:842-878) ---- */
static int bm_test(uint8_t* f, uint32_t off) {
  uint64_t* bm = w64(f + OFF_PIDS + P_BITMAP);
  uint64_t bit = ((uint64_t)1 << 63) >> (off & 63);
  return (bm[off / 64] & bit) == bit;
}
static void bm_set(uint8_t* f, uint32_t off, int v) {
  uint64_t* bm = w64(f + OFF_PIDS + P_BITMAP);
  uint64_t bit = ((uint64_t)1 << 63) >> (off & 63);
  if (v) bm[off / 64] |= bit; else bm[off / 64] &= ~bit;
}
static int proc_alive(uint64_t pid) { return pid != 0 && kill((pid_t)pid, 0) == 0; }
static void mu_lock(uint8_t* f) { /* Lock :851-863: the token is the creator PID stored in the file */
  uint64_t* l = w64(f + OFF_PIDS + P_LOCK);
  uint64_t me = *w64(f + OFF_PIDS + P_CREATOR);
  for (;;) {
    if (cas(l, 0, me)) return;
    uint64_t holder = ald(l);
    if (holder != 0 && holder != me && !proc_alive(holder)) { cas(l, holder, 0); continue; }
    sched_yield();
  }
}
static void mu_unlock(uint8_t* f) { cas(w64(f + OFF_PIDS + P_LOCK), *w64(f + OFF_PIDS + P_CREATOR), 0); }

int tfo_shm_pid_insert(uint8_t* f, uint64_t pid) { /* AddPID + InsertIfAbsent :537-541,791-812 */
  uint64_t* vals = w64(f + OFF_PIDS + P_VALUES);
  uint64_t* len = w64(f + OFF_PIDS + P_LEN);
  int rc = 0;
  mu_lock(f);
  for (uint32_t i = 0; i < MAX_PROCESSES; ++i) if (bm_test(f, i) && vals[i] == pid) goto done;
  if (*len >= MAX_PROCESSES) goto done;
  for (uint32_t i = 0; i < MAX_PROCESSES; ++i) {
    if (bm_test(f, i)) continue;
    vals[i] = pid; bm_set(f, i, 1); (*len)++; rc = 1;
    break;
  }
done:
  mu_unlock(f);
  return rc;
}
int tfo_shm_pid_remove(uint8_t* f, uint64_t pid) { /* RemoveValue :815-829 */
  uint64_t* vals = w64(f + OFF_PIDS + P_VALUES);
  uint64_t* len = w64(f + OFF_PIDS + P_LEN);
  int rc = 0;
  mu_lock(f);
  for (uint32_t i = 0; i < MAX_PROCESSES; ++i) {
    if (bm_test(f, i) && vals[i] == pid) {
      bm_set(f, i, 0); vals[i] = 0;
      if (*len > 0) (*len)--;
      rc = 1;
      break;
    }
  }
  mu_unlock(f);
  return rc;
}
size_t tfo_shm_pid_values(uint8_t* f, uint64_t* out, size_t cap) { /* Values :832-840 */
  uint64_t* vals = w64(f + OFF_PIDS + P_VALUES);
  size_t n = 0;
  mu_lock(f);
  for (uint32_t i = 0; i < MAX_PROCESSES; ++i) if (bm_test(f, i)) { if (n < cap) out[n] = vals[i]; n++; }
  mu_unlock(f);
  return n;
}

/* ---- paths (:63-108, :900-909) ---- */
int tfo_valid_component(const char* s) { return s && *s && !strchr(s, '/') && !strchr(s, '\\') && !strstr(s, ".."); }

int tfo_from_shm_path(const char* path, char* ns, char* name, size_t cap) {
  /* filepath.Clean, split on '/', drop empties; need >= 3 comps and last == "shm" */
  char* tmp = strdup(path);
  char* comps[256];
  int n = 0;
  const int rooted = path[0] == '/';
  for (char* tok = strtok(tmp, "/"); tok; tok = strtok(NULL, "/")) {
    if (!strcmp(tok, ".")) continue;
    if (!strcmp(tok, "..")) { if (n > 0 && strcmp(comps[n - 1], "..")) n--; else if (!rooted && n < 256) comps[n++] = tok; continue; }
    if (n < 256) comps[n++] = tok;
  }
  int rc = 1;
  if (n >= 3 && !strcmp(comps[n - 1], "shm")) {
    snprintf(ns, cap, "%s", comps[n - 3]);
    snprintf(name, cap, "%s", comps[n - 2]);
    rc = 0;
  }
  free(tmp);
  return rc;
}

/* ---- Create / Open (:891-1034) ---- */
struct tfo_shm { uint8_t* data; int fd; char path[1024]; };

static int mkdirs(const char* dir) {
  char buf[1024];
  snprintf(buf, sizeof buf, "%s", dir);
  for (char* p = buf + 1; *p; ++p) if (*p == '/') { *p = 0; if (mkdir(buf, 0755) && errno != EEXIST) return -1; *p = '/'; }
  return (mkdir(buf, 0755) && errno != EEXIST) ? -1 : 0;
}

int tfo_shm_create(const char* base, const char* ns, const char* pod, const tfo_dev_cfg* cfgs, size_t n, tfo_shm** out) {
  if (!ns || !pod || !*ns || !*pod) return 1;
  if (!tfo_valid_component(ns) || !tfo_valid_component(pod)) return 1;
  char dir[900];
  snprintf(dir, sizeof dir, "%s/%s/%s", base, ns, pod);
  if (mkdirs(dir)) return 5;
  tfo_shm* h = (tfo_shm*)calloc(1, sizeof *h);
  snprintf(h->path, sizeof h->path, "%s/shm", dir);
  h->fd = open(h->path, O_RDWR | O_CREAT | O_TRUNC, 0666);
  if (h->fd < 0) { free(h); return 5; }
  if (ftruncate(h->fd, FILE_BYTES)) { close(h->fd); free(h); return 5; }
  h->data = (uint8_t*)mmap(NULL, FILE_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED, h->fd, 0);
  if (h->data == MAP_FAILED) { close(h->fd); free(h); return 5; }
  if (tfo_shm_init_image(h->data, cfgs, n, (uint64_t)time(NULL), (uint64_t)getpid())) { munmap(h->data, FILE_BYTES); close(h->fd); free(h); return 1; }
  *out = h;
  return 0;
}

int tfo_shm_open(const char* base, const char* ns, const char* pod, tfo_shm** out) {
  if (!ns || !pod || !*ns || !*pod) return 1;
  if (!tfo_valid_component(ns) || !tfo_valid_component(pod)) return 1;
  tfo_shm* h = (tfo_shm*)calloc(1, sizeof *h);
  snprintf(h->path, sizeof h->path, "%s/%s/%s/shm", base, ns, pod);
  h->fd = open(h->path, O_RDWR, 0666);
  if (h->fd < 0) { free(h); return 2; }
  struct stat st;
  if (fstat(h->fd, &st)) { close(h->fd); free(h); return 5; }
  if (st.st_size != FILE_BYTES) {
    close(h->fd); free(h);
    return st.st_size == LEGACY_BYTES ? 100 : 101; /* "legacy shared memory layout detected" / "unexpected shared memory size" */
  }
  h->data = (uint8_t*)mmap(NULL, FILE_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED, h->fd, 0);
  if (h->data == MAP_FAILED) { close(h->fd); free(h); return 5; }
  if (*w32(h->data + OFF_DISCRIMINANT) != 1) { munmap(h->data, FILE_BYTES); close(h->fd); free(h); return 102; }
  *out = h;
  return 0;
}
uint8_t* tfo_shm_data(tfo_shm* h) { return h->data; }
void tfo_shm_close(tfo_shm* h) { if (!h) return; munmap(h->data, FILE_BYTES); close(h->fd); free(h); }

/* ---- gate throughput of the CPU restatement (ops/s under contention) ---- */
#include <pthread.h>
typedef struct { uint8_t* f; uint64_t ops; double cost; uint64_t denied; } gate_job;
static void* gate_thread(void* a) {
  gate_job* j = (gate_job*)a;
  uint64_t d = 0;
  for (uint64_t i = 0; i < j->ops; ++i) if (tfo_shm_fetch_sub(j->f, 0, j->cost) < j->cost) ++d;
  j->denied = d;
  return NULL;
}
typedef struct { uint8_t* f; volatile int stop; } refill_job;
static void* refill_thread(void* a) {
  refill_job* j = (refill_job*)a;
  while (!j->stop) { tfo_shm_fetch_add(j->f, 0, 1e9); struct timespec ts = {0, 500000000}; nanosleep(&ts, NULL); }
  return NULL;
}
/* returns ops/s; *deny_ratio out */
double tfo_gate_bench(int nthreads, uint64_t ops_per_thread, double cost, double* deny_ratio) {
  uint8_t* img = (uint8_t*)calloc(1, FILE_BYTES);
  tfo_dev_cfg cfg;
  memset(&cfg, 0, sizeof cfg);
  snprintf(cfg.uuid, sizeof cfg.uuid, "GPU-bench");
  tfo_shm_init_image(img, &cfg, 1, 1, 1);
  tfo_shm_set(img, 0, 1, 1e18);   /* capacity */
  tfo_shm_set(img, 0, 2, 1e9);    /* tokens   */
  pthread_t th[64], rt;
  gate_job jobs[64];
  refill_job rj = {img, 0};
  if (nthreads > 64) nthreads = 64;
  pthread_create(&rt, NULL, refill_thread, &rj);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; ++t) { jobs[t].f = img; jobs[t].ops = ops_per_thread; jobs[t].cost = cost; jobs[t].denied = 0; pthread_create(&th[t], NULL, gate_thread, &jobs[t]); }
  uint64_t denied = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); denied += jobs[t].denied; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  rj.stop = 1;
  pthread_join(rt, NULL);
  free(img);
  const double dt = (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-9;
  if (deny_ratio) *deny_ratio = (double)denied / (double)(ops_per_thread * (uint64_t)nthreads);
  return (double)(ops_per_thread * (uint64_t)nthreads) / dt;
}
