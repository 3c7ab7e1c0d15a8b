/* null_worker.c -- CPU stand-in for the worker's half of the shared-memory transport
 * (include/tfw_shm_ring.h): consumes the client->worker ring as fast as it can (headers parsed,
 * payloads skipped), answers SYNC and D2H (payload = whatever is in the ring).  For measuring the
 * client library's side of the transport without a GPU; never shipped.
 *   null_worker <file> <MiB> [sessions] */
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include "tfw_shm_ring.h"
#include "tfw_wire.h"

static void nap(void) { struct timespec ts = {0, 20000}; nanosleep(&ts, NULL); }

static tfsr_header* H;
static uint8_t *C2W, *W2C;

static void send_bytes(const void* p, uint64_t n, int zeros) {
  const uint8_t* b = (const uint8_t*)p;
  uint64_t head = H->w2c_head;
  while (n) {
    const uint64_t tail = __atomic_load_n(&H->w2c_tail, __ATOMIC_ACQUIRE);
    const uint64_t free_b = H->w2c_size - (head - tail);
    if (!free_b) { __builtin_ia32_pause(); continue; }
    const uint64_t pos = head % H->w2c_size;
    uint64_t k = n < free_b ? n : free_b;
    if (k > H->w2c_size - pos) k = H->w2c_size - pos;
    if (!zeros) { memcpy(W2C + pos, b, k); b += k; }
    head += k;
    __atomic_store_n(&H->w2c_head, head, __ATOMIC_RELEASE);
    n -= k;
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const uint64_t total = (uint64_t)atol(argv[2]) << 20;
  int sessions = argc > 3 ? atoi(argv[3]) : 1;
  int fd = open(argv[1], O_RDWR | O_CREAT, 0666);
  if (fd < 0 || ftruncate(fd, (off_t)total) != 0) return 3;
  void* m = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, 0);
  if (m == MAP_FAILED) return 3;
  H = (tfsr_header*)m;
  memset(H, 0, sizeof *H);
  H->version = TFSR_VERSION;
  H->total_bytes = total;
  tfsr_layout(total, &H->c2w_off, &H->c2w_size, &H->w2c_off, &H->w2c_size);
  C2W = (uint8_t*)m + H->c2w_off;
  W2C = (uint8_t*)m + H->w2c_off;
  H->worker_pid = (uint32_t)getpid();
  H->session = 1;
  __atomic_store_n(&H->magic, TFSR_MAGIC, __ATOMIC_RELEASE);
  __atomic_store_n(&H->worker_ready, 1u, __ATOMIC_RELEASE);
  printf("null worker serving %s\n", argv[1]);
  fflush(stdout);
  while (sessions-- > 0) {
    while (__atomic_load_n(&H->client_pid, __ATOMIC_ACQUIRE) == 0) nap();
    const uint32_t session = H->session;
    uint64_t rd = H->c2w_tail, skip = 0;
    uint8_t hdr_buf[TFCS_HDR_BYTES];
    uint32_t have = 0;
    for (;;) {
      const uint64_t head = __atomic_load_n(&H->c2w_head, __ATOMIC_ACQUIRE);
      uint64_t avail = head - rd;
      if (!avail) {
        if (__atomic_load_n(&H->client_closed, __ATOMIC_ACQUIRE) >= session) break;
        __builtin_ia32_pause();
        continue;
      }
      if (skip) {  // payload bytes: not even read
        const uint64_t k = avail < skip ? avail : skip;
        rd += k; skip -= k;
        __atomic_store_n(&H->c2w_tail, rd, __ATOMIC_RELEASE);
        continue;
      }
      while (avail && have < TFCS_HDR_BYTES) {
        const uint64_t pos = rd % H->c2w_size;
        uint64_t k = TFCS_HDR_BYTES - have;
        if (k > avail) k = avail;
        if (k > H->c2w_size - pos) k = H->c2w_size - pos;
        memcpy(hdr_buf + have, C2W + pos, k);
        have += (uint32_t)k; rd += k; avail -= k;
      }
      if (have < TFCS_HDR_BYTES || ((tfcs_frame_hdr*)hdr_buf)->opcode != TFCS_OP_MEMCPY_H2D)
        __atomic_store_n(&H->c2w_tail, rd, __ATOMIC_RELEASE);  /* an H2D header is released together with its payload */
      if (have < TFCS_HDR_BYTES) continue;
      have = 0;
      tfcs_frame_hdr h;
      memcpy(&h, hdr_buf, sizeof h);
      if (h.magic != TFCS_MAGIC) { fprintf(stderr, "null worker: bad magic\n"); return 4; }
      if (h.opcode == TFCS_OP_MEMCPY_H2D) skip = tfcs_pad16(h.length);
      else if (h.opcode == TFCS_OP_SYNC || h.opcode == TFCS_OP_MEMCPY_D2H) {
        tfcs_frame_hdr r = h;
        r.opcode = h.opcode == TFCS_OP_SYNC ? TFCS_OP_RESP_SYNC : TFCS_OP_RESP_D2H;
        if (h.opcode == TFCS_OP_SYNC) r.length = 0;
        send_bytes(&r, sizeof r, 0);
        if (h.opcode == TFCS_OP_MEMCPY_D2H) send_bytes(NULL, tfcs_pad16(h.length), 1);
      }
    }
    __atomic_store_n(&H->worker_closed, session, __ATOMIC_RELEASE);
    H->c2w_tail = __atomic_load_n(&H->c2w_head, __ATOMIC_ACQUIRE);
    H->w2c_tail = H->w2c_head;
    H->session = session + 1;
    __atomic_store_n(&H->client_pid, 0u, __ATOMIC_RELEASE);
  }
  return 0;
}
