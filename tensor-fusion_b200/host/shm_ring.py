"""ctypes mirror of include/tfw_shm_ring.h (shared-memory transport between a client and the worker)."""
import ctypes as C

TFSR_MAGIC, TFSR_VERSION, TFSR_HDR_BYTES = 0x52534654, 1, 4096


class TfsrHeader(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint32), ("total_bytes", C.c_uint64),
                ("c2w_off", C.c_uint64), ("c2w_size", C.c_uint64), ("w2c_off", C.c_uint64), ("w2c_size", C.c_uint64),
                ("worker_pid", C.c_uint32), ("worker_ready", C.c_uint32), ("client_pid", C.c_uint32),
                ("client_closed", C.c_uint32), ("worker_closed", C.c_uint32), ("session", C.c_uint32),
                ("client_lock_session", C.c_uint32), ("reserved0", C.c_uint32), ("pad0", C.c_uint8 * 48),
                ("c2w_head", C.c_uint64), ("pad1", C.c_uint8 * 56), ("c2w_tail", C.c_uint64), ("pad2", C.c_uint8 * 56),
                ("w2c_head", C.c_uint64), ("pad3", C.c_uint8 * 56), ("w2c_tail", C.c_uint64), ("pad4", C.c_uint8 * 56)]


assert C.sizeof(TfsrHeader) == 384 and TfsrHeader.c2w_head.offset == 128 and TfsrHeader.w2c_tail.offset == 320


def layout(total):
    """tfsr_layout(): (c2w_off, c2w_size, w2c_off, w2c_size) for a file of `total` bytes."""
    space = (total - TFSR_HDR_BYTES) & ~4095
    up = (space // 4 * 3) & ~4095
    return TFSR_HDR_BYTES, up, TFSR_HDR_BYTES + up, space - up
