"""Kernel-level parity of the byte mover (tfw_move_batch) against numpy semantics."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(mode):
    from tensor_fusion_b200 import _native as N
    from tensor_fusion_b200.worker import Worker
    return Worker(flags=N.TFW_F_MOVER_TMA if mode == "tma" else N.TFW_F_MOVER_LDG)


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_all_alignment_pairs(mode):
    """Every (src mod 16, dst mod 16) pair, lengths around the vector/tile edges."""
    rng = np.random.default_rng(0)
    span = 200_000
    with _worker(mode) as w:
        src_h = rng.integers(0, 256, span + 64, dtype=np.uint8)
        s = w.dev_alloc(span + 64)
        d = w.dev_alloc(span + 64)
        w.dev_write(s, src_h)
        for length in (1, 15, 16, 17, 31, 33, 255, 4096, 32768, 32769, 65536 + 5, 131072 + 17):
            ref = np.full(span + 64, 0xEE, dtype=np.uint8)
            w.dev_write(d, ref)
            descs = []
            # 256 copies in ONE launch, laid out back to back in the destination
            pos = 0
            for sa in range(16):
                for da in range(16):
                    if pos + 32 + length > span:
                        break
                    doff = ((pos + 15) & ~15) + da
                    soff = int(rng.integers(0, span - length)) & ~15
                    soff += sa
                    descs.append((d + doff, s + soff, length, 0))
                    ref[doff: doff + length] = src_h[soff: soff + length]
                    pos = doff + length
            w.move_batch(descs)
            got = w.dev_read(d, span + 64)
            assert np.array_equal(got, ref), f"len={length}: first diff at {np.flatnonzero(got != ref)[:5]}"
        w.dev_free(s)
        w.dev_free(d)


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_fill_and_mixed_batch(mode):
    rng = np.random.default_rng(3)
    n = 3 << 20
    with _worker(mode) as w:
        src_h = rng.integers(0, 256, n, dtype=np.uint8)
        s, d = w.dev_alloc(n), w.dev_alloc(n)
        w.dev_write(s, src_h)
        ref = np.zeros(n, dtype=np.uint8)
        w.dev_write(d, ref)
        descs = [
            (d + 0, 0, 1_000_003, 0x5A),               # big unaligned-length fill
            (d + 1_000_003, s + 11, 777_777, 0),       # misaligned copy, multi-tile
            (d + 1_800_000, s + 1_048_576, 1 << 20, 0),  # aligned copy, 32 full tiles
            (d + 2_900_001, 0, 7, 0xFF),               # tiny fill
            (d + 2_950_000, s + 5, 3, 0),              # tiny copy
        ]
        ref[0:1_000_003] = 0x5A
        ref[1_000_003:1_000_003 + 777_777] = src_h[11:11 + 777_777]
        ref[1_800_000:1_800_000 + (1 << 20)] = src_h[1_048_576:1_048_576 + (1 << 20)]
        ref[2_900_001:2_900_008] = 0xFF
        ref[2_950_000:2_950_003] = src_h[5:8]
        w.move_batch(descs)
        assert np.array_equal(w.dev_read(d, n), ref)
        w.dev_free(s)
        w.dev_free(d)


@pytest.mark.parametrize("mode", ["ldg", "tma"])
def test_many_small_descriptors(mode):
    """4096 descriptors of 64 B..4 KiB in one launch (binary search over the tile prefix)."""
    rng = np.random.default_rng(9)
    n = 16 << 20
    with _worker(mode) as w:
        src_h = rng.integers(0, 256, n, dtype=np.uint8)
        s, d = w.dev_alloc(n), w.dev_alloc(n)
        w.dev_write(s, src_h)
        ref = np.zeros(n, dtype=np.uint8)
        w.dev_write(d, ref)
        descs, pos = [], 0
        for _ in range(4096):
            ln = int(rng.integers(64, 4096))
            soff = int(rng.integers(0, n - ln))
            doff = pos + int(rng.integers(0, 16))
            descs.append((d + doff, s + soff, ln, 0))
            ref[doff: doff + ln] = src_h[soff: soff + ln]
            pos = doff + ln
        w.move_batch(descs)
        assert np.array_equal(w.dev_read(d, n), ref)
        w.dev_free(s)
        w.dev_free(d)


def test_digest_matches_oracle():
    import oracle
    rng = np.random.default_rng(4)
    with _worker("ldg") as w:
        for n in (8, 9, 15, 4096, 1_000_003, 8 << 20):
            h = rng.integers(0, 256, n, dtype=np.uint8)
            p = w.dev_alloc(n)
            w.dev_write(p, h)
            assert w.dev_digest(p, n) == oracle.digest(h)
            w.dev_free(p)
