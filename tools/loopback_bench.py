"""TCP loopback throughput of `tensor-fusion-worker -p <port>`: stream a bulk TFCS trace over one
connection and wait for the SYNC response.  GB/s of payload, end to end (socket -> pinned ring ->
DMA -> unpack)."""
import json
import os
import socket
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensor_fusion_b200 import trace  # noqa: E402

EXE = os.path.join(ROOT, "tensor-fusion_b200", "lib", "tensor-fusion-worker")
nbuf, ncopies, each = 16, 64, 64 << 20
raw = trace.gen_bulk(nbuf, ncopies, each, nthreads=16)
p = subprocess.Popen([EXE, "-p", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                     env=dict(os.environ, TFW_ONESHOT="1", TFW_BIND="127.0.0.1"))
port = int(p.stdout.readline().split()[-1])
s = socket.create_connection(("127.0.0.1", port))
s.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, 8 << 20)
time.sleep(1.5)                      # CUDA context creation in the worker is not part of the number
mv = memoryview(raw)
t0 = time.perf_counter()
s.sendall(mv)
resp = b""
while len(resp) < 64:                # the RESP_SYNC frame
    resp += s.recv(4096)
dt = time.perf_counter() - t0
s.close()
p.wait(timeout=60)
print(json.dumps({"leg": "tcp_loopback_one_connection", "payload_gib": ncopies * each / 2**30, "seconds": round(dt, 3),
                  "payload_GBps": round(ncopies * each / dt / 1e9, 2)}))
