"""Child process of tests/test_rtp_parse.py::test_parser_agrees_with_rfc3550_on_fuzzed_input: every packet is parsed where
it ENDS at a page boundary with an inaccessible page behind it, so a read past the packet kills this process instead of going
unnoticed; verdict and fields are compared with a parser written straight from RFC 3550 §5.1 / §5.3.1."""
import ctypes as C
import mmap
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import rtp_packets as RP  # noqa: E402
from mediastreamer2_b200 import _lib  # noqa: E402


def rfc3550(pkt: bytes):
    """None when malformed, else (marker, pt, seq, ts, ssrc, payload_offset, payload_len)"""
    n = len(pkt)
    if n < 12 or pkt[0] >> 6 != 2:
        return None
    off = 12 + 4 * (pkt[0] & 15)
    if off > n:
        return None
    if pkt[0] & 0x10:
        if off + 4 > n:
            return None
        off += 4 + 4 * int.from_bytes(pkt[off + 2:off + 4], "big")
        if off > n:
            return None
    end = n
    if pkt[0] & 0x20:
        pad = pkt[-1]
        if pad == 0 or off + pad > n:
            return None
        end -= pad
    return (pkt[1] >> 7, pkt[1] & 127, int.from_bytes(pkt[2:4], "big"), int.from_bytes(pkt[4:8], "big"),
            int.from_bytes(pkt[8:12], "big"), off, end - off)


def main(rounds: int) -> int:
    lib = _lib.load()
    page = mmap.PAGESIZE
    libc = C.CDLL(None, use_errno=True)
    libc.mmap.restype = C.c_void_p
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    base = libc.mmap(None, 2 * page, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)
    assert base not in (None, C.c_void_p(-1).value)
    assert libc.mprotect(C.c_void_p(base + page), page, 0) == 0  # PROT_NONE: the guard page
    lib.msb200_rtp_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(2026)
    m, off = _lib.RtpMeta(), C.c_size_t()
    accepted = 0
    for k in range(rounds):
        kind = k % 4
        if kind == 0:    # pure noise with a plausible first byte
            pkt = bytes([int(rng.integers(0x80, 0xC0))]) + rng.integers(0, 256, int(rng.integers(0, 96))).astype(np.uint8).tobytes()
        elif kind == 1:  # a valid packet, truncated anywhere
            good, _, _ = RP.random_packet(rng, int(rng.integers(0, 64)), int(rng.integers(0, 128)))
            pkt = good[:int(rng.integers(0, len(good) + 1))]
        elif kind == 2:  # a valid packet with one byte of its header area flipped
            good, _, _ = RP.random_packet(rng, int(rng.integers(0, 64)), int(rng.integers(0, 128)))
            b = bytearray(good)
            b[int(rng.integers(0, min(len(b), 24)))] ^= 1 << int(rng.integers(0, 8))
            pkt = bytes(b)
        else:            # a valid packet
            pkt, _, _ = RP.random_packet(rng, int(rng.integers(0, 200)), int(rng.integers(0, 128)))
        addr = base + page - len(pkt)
        C.memmove(addr, pkt, len(pkt))
        rc = lib.msb200_rtp_parse(C.c_void_p(addr), len(pkt), C.byref(m), C.byref(off)) if len(pkt) else lib.msb200_rtp_parse(C.c_void_p(base), 0, C.byref(m), C.byref(off))
        exp = rfc3550(pkt)
        if exp is None:
            assert rc == _lib.EINVAL, (k, pkt.hex(), rc)
        else:
            accepted += 1
            got = (m.marker, m.payload_type, m.seq, m.timestamp, m.ssrc, off.value, m.payload_len)
            assert rc == 0 and got == exp, (k, pkt.hex(), rc, got, exp)
    print(f"OK {rounds} {accepted}")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 20000))
