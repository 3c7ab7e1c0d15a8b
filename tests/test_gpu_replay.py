"""GPU parity: the CUDA worker (through the C-ABI) vs the CPU oracle, same seeded traces.

Bit-exact bar: every client-visible buffer, and the whole response stream
(D2H payloads, SYNC acks, per-frame errors), must equal the oracle's.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(worker, rep, resp_gpu):
    from tensor_fusion_b200 import wire
    resp_cpu = rep.responses()
    assert len(resp_gpu) == len(resp_cpu)
    if resp_gpu != resp_cpu:
        for (hg, pg), (hc, pc) in zip(wire.parse_frames(resp_gpu), wire.parse_frames(resp_cpu)):
            assert hg == hc, f"response header mismatch call {hc['call_id']}"
            assert pg == pc, f"D2H payload mismatch call {hc['call_id']}"
    handles = rep.live_handles()
    assert handles
    for h in handles:
        want = rep.buffer(h)
        got = worker.read(h)
        assert got.nbytes == want.nbytes
        assert np.array_equal(got, want), f"buffer {h} differs at {np.flatnonzero(got != want)[:8]}"


def _mk(mode, **kw):
    from tensor_fusion_b200 import _native as N
    from tensor_fusion_b200.worker import Worker
    flags = {"ldg": N.TFW_F_MOVER_LDG, "tma": N.TFW_F_MOVER_TMA}[mode]
    return Worker(flags=flags, **kw)


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_c1_trace_pageable_input(mode):
    """BASELINE config 1 trace (1k calls, seed 0x7F5EED) fed from ordinary host memory."""
    import oracle
    from tensor_fusion_b200 import trace
    t = trace.gen_c1()
    rep = oracle.Replay(t)
    assert rep.rc == 0
    with _mk(mode) as w:
        n, resp = w.run(t)
        assert n == t.nbytes
        _compare(w, rep, resp)
        st = w.stats()
        assert st["frames"] == rep.stat(0)
        assert st["payload_bytes"] == rep.stat(1)
        assert st["mover_launches"] > 0


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_c1_trace_pinned_input_small_chunks(mode):
    """Same trace, DMA'd in place from pinned memory, tiny staging slots (many batches, wrap-around)."""
    import oracle
    from tensor_fusion_b200 import trace
    from tensor_fusion_b200.worker import PinnedBuffer
    raw = trace.gen_c1(seed=0x1234, ncalls=600)
    pin = PinnedBuffer(raw.nbytes + 64)
    for mis in (0, 16, 5):  # 16-byte aligned stream, and a stream at an odd host address
        view = pin.array[mis: mis + raw.nbytes]
        view[:] = raw
        rep = oracle.Replay(raw)
        with _mk(mode, chunk_bytes=256 << 10, num_slots=3) as w:
            n = w.submit(view)
            assert n == raw.nbytes
            w.flush()
            _compare(w, rep, w.poll())
    pin.free()


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_streaming_arbitrary_cuts(mode):
    """The stream arrives in arbitrary pieces (mid-header, mid-payload): the
    resumable deserializer must give the same result as one big submit."""
    import oracle
    from tensor_fusion_b200 import trace
    raw = trace.gen_c1(seed=77, ncalls=300, max_payload_bytes=256 << 10)
    rep = oracle.Replay(raw)
    rng = np.random.default_rng(5)
    with _mk(mode, chunk_bytes=128 << 10) as w:
        pos, carry, resp = 0, b"", b""
        while pos < raw.nbytes or carry:
            step = int(rng.integers(1, 200_000))
            piece = carry + raw[pos: pos + step].tobytes()
            pos = min(raw.nbytes, pos + step)
            used = w.submit(piece)
            carry = piece[used:]
            resp += w.poll()
            if pos >= raw.nbytes and len(carry) == 0:
                break
            assert len(carry) < 64 or pos < raw.nbytes
        w.flush()
        resp += w.poll()
        _compare(w, rep, resp)


def test_hazards_are_serialised():
    """WAW / RAW / WAR inside what would be one batch must behave sequentially."""
    import oracle
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    rng = np.random.default_rng(1)
    b = wire.Builder()
    b.malloc(1, 1 << 20).malloc(2, 1 << 20)
    a = rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes()
    c = rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes()
    b.h2d(1, 100, a)            # write
    b.h2d(1, 50_000, c)         # WAW overlap with the previous payload
    b.d2d(2, 7, 1, 40_000, 100_000)   # RAW: reads what the two copies wrote
    b.memset(1, 45_000, 30_000, 0xAB)  # WAR: overwrites the range D2D just read
    b.d2d(1, 500_000, 2, 0, 120_000)   # RAW on buffer 2
    b.launch(wire.K_ADD_U8, grid=64, block=128, h=1, off=3, n=700_001, scalar=7)
    b.h2d(1, 0, a[:1000])
    b.launch(wire.K_XOR_IDX, grid=3, block=64, h=2, off=1, n=999_999, scalar=12345)
    b.d2h(1, 1, 65_537).d2h(2, 0, 1 << 20).sync()
    raw = bytes(b)
    rep = oracle.Replay(raw)
    with Worker() as w:
        n, resp = w.run(raw)
        assert n == len(raw)
        _compare(w, rep, resp)
        assert w.stats()["batches_hazard"] >= 3


def test_error_frames_match_oracle():
    import oracle
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    b = wire.Builder()
    b.malloc(1, 4096).malloc(1, 64)                 # duplicate handle
    b.malloc(2, 0)                                   # zero size
    b.h2d(9, 0, b"x" * 100)                           # unknown handle, payload must be skipped
    b.h2d(1, 4090, b"y" * 100)                        # out of range
    b.h2d(1, 4000, b"z" * 96)                         # exactly fits
    b.d2h(1, 4096, 1).d2h(1, 0, 4096)
    b.memset(1, 0, 5000, 1).memset(3, 0, 1, 1)
    b.d2d(1, 0, 1, 8, 100)                            # overlapping D2D
    b.launch(42).launch(wire.K_ADD_U8, h=7, n=5)
    b.free(5).free(1).free(1)
    b.raw(wire.frame(99, call_id=1000))               # unknown opcode
    b.sync()
    raw = bytes(b)
    rep = oracle.Replay(raw, vram_limit=1 << 20)
    with Worker(vram_limit=1 << 20) as w:
        n, resp = w.run(raw)
        assert n == len(raw)
        assert resp == rep.responses()
        codes = [h["arg0"] for h, _ in wire.parse_frames(resp) if h["opcode"] == wire.OP_RESP_ERROR]
        assert len(codes) >= 10


def test_vram_quota_enforced():
    import oracle
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    b = wire.Builder()
    b.malloc(1, 600_000).malloc(2, 600_000).free(1).malloc(3, 500_000).sync()
    raw = bytes(b)
    rep = oracle.Replay(raw, vram_limit=1_000_000)
    with Worker(vram_limit=1_000_000) as w:
        _, resp = w.run(raw)
        assert resp == rep.responses()
        st = w.stats()
        assert st["live_buffers"] == 1 and st["vram_bytes"] == 500_000


def test_bad_magic_is_protocol_error():
    from tensor_fusion_b200 import _native as N
    from tensor_fusion_b200.worker import Worker
    with Worker() as w:
        with pytest.raises(N.TfwError) as e:
            w.submit(b"\0" * 64)
        assert e.value.status == N.TFW_ERR_PROTOCOL


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_resident_trace_replay_matches_oracle(mode):
    """tfw_trace_load/replay (trace resident in HBM) gives the same buffers, replay after replay."""
    import oracle
    from tensor_fusion_b200 import trace
    raw = trace.gen_c1(seed=99, ncalls=500, error_permille=0)
    rep = oracle.Replay(raw)
    with _mk(mode, chunk_bytes=1 << 20) as w:
        t = w.load_trace(raw)
        info = t.info()
        assert info["payload_bytes"] == rep.stat(1)
        for _ in range(2):
            t.replay()
            w.flush()
            assert w.poll() == rep.responses()
            for h in rep.live_handles():
                size, ptr = t.buffer_info(h)
                want = rep.buffer(h)
                assert size == want.nbytes
                assert np.array_equal(w.dev_read(ptr, size), want)
                assert w.dev_digest(ptr, size) == oracle.digest(want)
        t.free()


def test_full_size_bulk_digests():
    """BASELINE config 2 shape at full size (64 MiB payloads): position-sensitive
    digests of every buffer computed on the GPU equal the oracle's digests of the
    payload streams -- a checksum-of-checksums that needs no 4 GiB read-back."""
    import oracle
    from tensor_fusion_b200 import trace, wire
    from tensor_fusion_b200.worker import Worker, PinnedBuffer
    nbuf, ncopies, each = 4, 8, 64 << 20
    size = trace.bulk_size(nbuf, ncopies, each)
    pin = PinnedBuffer(size)
    raw = trace.gen_bulk(nbuf, ncopies, each, into=pin)
    # last writer of buffer b is copy (ncopies - nbuf + b); MALLOCs take call ids 0..nbuf-1
    want = {}
    for b in range(nbuf):
        call_id = nbuf + (ncopies - nbuf + b)
        want[b + 1] = oracle.digest(oracle.payload(trace.SEED_C1, call_id, each))
    with Worker() as w:
        assert w.submit(raw) == size
        w.flush()
        for h, d in want.items():
            assert w.digest(h) == d
        st = w.stats()
        assert st["payload_bytes"] == ncopies * each
        # in-place DMA: exactly the payload span went over PCIe, nothing was copied through the ring
        assert st["h2d_dma_bytes"] <= ncopies * (each + 64)
    pin.free()


def test_double_buffered_rings_with_fences():
    """The receive loop of the worker binary: two pinned rings reused alternately; tfw_fence /
    tfw_fence_wait tell when a ring may be overwritten while the other one is still being consumed."""
    import ctypes as C
    import oracle
    from tensor_fusion_b200 import trace
    from tensor_fusion_b200._native import lib, check
    from tensor_fusion_b200.worker import PinnedBuffer, Worker
    raw = trace.gen_c1(seed=31337, ncalls=500, error_permille=0)
    rep = oracle.Replay(raw)
    ring = 1 << 20
    rings = [PinnedBuffer(ring), PinnedBuffer(ring)]
    tickets = [0, 0]
    with Worker(chunk_bytes=256 << 10) as w:
        pos, cur, carry, resp = 0, 0, b"", b""
        while pos < raw.nbytes or carry:
            if tickets[cur]:
                check(lib.tfw_fence_wait(w.h, tickets[cur]), "tfw_fence_wait", w.h)   # ring `cur` is free again
            n = min(ring - len(carry), raw.nbytes - pos)
            buf = rings[cur].array
            buf[: len(carry)] = np.frombuffer(carry, dtype=np.uint8)
            buf[len(carry): len(carry) + n] = raw[pos: pos + n]
            have = len(carry) + n
            used = w.submit(buf[:have])
            t = C.c_uint64()
            check(lib.tfw_fence(w.h, C.byref(t)), "tfw_fence", w.h)
            tickets[cur] = t.value
            carry = bytes(buf[used:have])
            pos += n
            cur ^= 1
            resp += w.poll()
            assert len(carry) < 64
        w.flush()
        resp += w.poll()
        _compare(w, rep, resp)
    for r in rings:
        r.free()


def test_empty_and_boundary_frames():
    """Zero-length operations, the largest legal handle, one past it, a buffer of one byte, NOP frames."""
    import oracle
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    b = wire.Builder()
    b.raw(wire.frame(wire.OP_NOP, call_id=9000))
    b.malloc(65535, 1).malloc(65536, 16)                    # last legal handle; first illegal one
    b.h2d(65535, 0, b"\x7f").h2d(65535, 1, b"")              # fills the buffer; empty payload at the very end
    b.h2d(65535, 2, b"")                                     # offset past the end, even with length 0
    b.malloc(1, 4096)
    b.h2d(1, 0, b"").memset(1, 4096, 0, 9).d2d(1, 0, 1, 0, 0).d2h(1, 4096, 0).d2h(1, 0, 0)
    b.launch(wire.K_ADD_U8, h=1, n=0, scalar=3).launch(wire.K_NOOP, grid=0, block=0)
    b.d2h(65535, 0, 1).d2h(1, 0, 4096).free(65535).sync()
    raw = bytes(b)
    rep = oracle.Replay(raw)
    with Worker() as w:
        n, resp = w.run(raw)
        assert n == len(raw)
        assert resp == rep.responses()
        frames = list(wire.parse_frames(resp))
        errs = [(h["call_id"], h["arg0"]) for h, _ in frames if h["opcode"] == wire.OP_RESP_ERROR]
        assert len(errs) == 2                                 # handle 65536, and the offset-2 write
        assert np.array_equal(w.read(1), rep.buffer(1))
    # an empty submit and a lone partial header are not errors
    with Worker() as w:
        assert w.submit(b"") == 0
        assert w.submit(raw[:40]) == 0
        assert w.submit(raw[:64 + 30]) == 64


def test_responses_leave_in_pieces_through_a_small_buffer():
    """tfw_poll_responses is a byte stream: a D2H response larger than the caller's buffer (the worker
    executable drains through 16 MiB) must come out in pieces, byte-identical to the oracle."""
    import oracle
    from tensor_fusion_b200 import wire
    from tensor_fusion_b200.worker import Worker
    rng = np.random.default_rng(21)
    data = rng.integers(0, 256, 300_001, dtype=np.uint8).tobytes()
    b = wire.Builder()
    b.malloc(1, 300_001).h2d(1, 0, data).d2h(1, 0, 300_001).sync().d2h(1, 7, 1000).d2h(1, 0, 0).sync()
    raw = bytes(b)
    want = oracle.Replay(raw).responses()
    for cap in (1, 63, 64, 65, 4096, 300_000, 1 << 20):
        with Worker() as w:
            assert w.submit(raw) == len(raw)
            w.flush()
            got = bytearray()
            for _ in range(len(want) + 8):
                piece = w.poll(cap)
                assert len(piece) <= cap
                if not piece:
                    break
                got += piece
            assert bytes(got) == want, cap
