"""CPU: msb200_rtp_parse (host code of libmsb200dsp.so, no device needed) against packets built by hand from RFC 3550."""
import ctypes as C

import numpy as np

import rtp_packets as RP
from mediastreamer2_b200 import _lib


def parse(pkt: bytes):
    lib = _lib.load()
    m, off = _lib.RtpMeta(), C.c_size_t()
    rc = lib.msb200_rtp_parse(pkt, len(pkt), C.byref(m), C.byref(off))
    return rc, m, off.value


def test_fields_and_payload_position_on_random_packets():
    rng = np.random.default_rng(7)
    for _ in range(300):
        pkt, payload, kw = RP.random_packet(rng, int(rng.integers(0, 200)), int(rng.integers(0, 128)))
        rc, m, off = parse(pkt)
        assert rc == 0
        assert (m.timestamp, m.seq, m.ssrc, m.marker, m.payload_type) == (kw["ts"], kw["seq"], kw["ssrc"], int(kw["marker"]), kw["pt"])
        assert m.payload_len == len(payload) and pkt[off:off + m.payload_len] == payload


def test_malformed_packets_are_refused():
    good = RP.build(b"\x55" * 20, 8, 1, 160, 0xABCD)
    assert parse(good)[0] == 0
    assert parse(good[:11])[0] == _lib.EINVAL                       # shorter than a header
    assert parse(bytes([0x40]) + good[1:])[0] == _lib.EINVAL         # version 1
    assert parse(bytes([0x8F]) + good[1:13])[0] == _lib.EINVAL       # 15 CSRCs announced, none there
    assert parse(bytes([0x90]) + good[1:14])[0] == _lib.EINVAL       # extension bit, truncated extension header
    assert parse(bytes([0xA0]) + good[1:-1] + b"\x00")[0] == _lib.EINVAL  # padding bit with a zero count
    assert parse(bytes([0xA0]) + good[1:-1] + b"\xFF")[0] == _lib.EINVAL  # padding longer than the packet


def test_parser_agrees_with_rfc3550_on_fuzzed_input():
    """20 000 inputs — noise, truncated packets, packets with a flipped header bit, valid packets — each parsed with an
    inaccessible page right behind its last byte (a read past the packet would kill the child process) and compared with a
    parser written from the RFC: same verdict, same fields"""
    import subprocess
    import sys
    from pathlib import Path

    r = subprocess.run([sys.executable, str(Path(__file__).resolve().parent / "rtp_fuzz_child.py"), "20000"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-1500:])
    words = r.stdout.split()
    assert words[0] == "OK" and int(words[2]) > 4000  # a good share of the inputs were well-formed packets
