/* ring_lock_probe.c -- what the worker's liveness probe sees for a ring file: prints "held" or "free".
 * (the same inline functions of include/tfw_shm_ring.h the worker and the client use)
 *   ring_lock_probe <file>            probe once
 *   ring_lock_probe <file> hold       take the client's lock and sleep (to be killed by the test) */
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include "tfw_shm_ring.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  int fd = open(argv[1], O_RDWR);
  if (fd < 0) return 3;
#ifdef TFSR_HAVE_LIVENESS
  if (argc > 2 && !strcmp(argv[2], "hold")) {
    if (tfsr_client_lock(fd, 1) != 0) return 4;
    printf("holding\n");
    fflush(stdout);
    pause();
    return 0;
  }
  printf("%s\n", tfsr_client_alive(fd) ? "held" : "free");
  return 0;
#else
  printf("unsupported\n");
  return 0;
#endif
}
