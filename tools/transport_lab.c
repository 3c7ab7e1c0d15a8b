/* transport_lab.c -- per-call cost of the TFCS client in C (no interpreter in the loop):
 *   transport_lab <url> [calls] [bytes]      N x tfc_memcpy_h2d(bytes) + one tfc_sync, then N x tfc_sync */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "tfc_client.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const long calls = argc > 2 ? atol(argv[2]) : 200000;
  const size_t bytes = argc > 3 ? (size_t)atol(argv[3]) : 4096;
  tfc_conn* c = NULL;
  if (tfc_connect(argv[1], &c) != 0) { fprintf(stderr, "connect failed\n"); return 3; }
  uint32_t h = 0;
  tfc_malloc(c, 64u << 20, &h);
  if (tfc_sync(c) != 0) return 4;
  uint8_t* src = malloc(bytes);
  memset(src, 7, bytes);
  double t0 = now_us();
  for (long i = 0; i < calls; ++i) tfc_memcpy_h2d(c, h, ((uint64_t)i * bytes) % ((64u << 20) - bytes), src, bytes);
  if (tfc_sync(c) != 0) return 5;
  const double per_call = (now_us() - t0) / calls;
  const long syncs = calls < 20000 ? calls : 20000;
  t0 = now_us();
  for (long i = 0; i < syncs; ++i) tfc_sync(c);
  const double per_sync = (now_us() - t0) / syncs;
  printf("{\"url\": \"%s\", \"calls\": %ld, \"bytes\": %zu, \"h2d_us_per_call\": %.3f, \"h2d_GBps\": %.2f, \"sync_round_trip_us\": %.2f}\n", argv[1], calls, bytes,
         per_call, bytes / per_call / 1e3, per_sync);
  tfc_close(c);
  return 0;
}
