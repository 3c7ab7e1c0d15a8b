import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor_fusion_b200 import wire
from tensor_fusion_b200.gate import Gate
from tensor_fusion_b200.worker import Worker
from tensor_fusion_b200._native import lib
with Worker() as w:
    g = Gate()
    g.set_tokens(0.0)
    print("init", g.state())
    s = lib.tfw_exec_stream(w.h)
    print("stream", s)
    b = wire.Builder(); b.malloc(1, 4096).memset(1, 0, 4096, 1)
    w.run(bytes(b))
    g.enqueue(10.0, s)
    w.submit(bytes(wire.Builder().launch(wire.K_ADD_U8, h=1, n=4096, scalar=1)))
    time.sleep(0.2)
    print("waiting", g.state())
    print("refill before=", g.refill(25.0))
    print("after refill", g.state())
    print(w.read(1)[:4])
    print("final", g.state())
    g.close()
