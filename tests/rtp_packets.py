"""RTP packets built by hand from RFC 3550 §5.1 / §5.3.1 (independent of csrc/rtp_batch.cu's parser): fixed header, CSRC
list, optional header extension, optional padding."""
from __future__ import annotations

import struct

import numpy as np


def build(payload: bytes, pt: int, seq: int, ts: int, ssrc: int, marker: bool = False, csrcs=(), ext_words=None, pad: int = 0) -> bytes:
    b0 = 0x80 | (0x20 if pad else 0) | (0x10 if ext_words is not None else 0) | len(csrcs)
    b1 = (0x80 if marker else 0) | (pt & 0x7F)
    out = struct.pack("!BBHII", b0, b1, seq & 0xFFFF, ts & 0xFFFFFFFF, ssrc & 0xFFFFFFFF)
    for c in csrcs:
        out += struct.pack("!I", c)
    if ext_words is not None:
        out += struct.pack("!HH", 0xBEDE, len(ext_words)) + b"".join(struct.pack("!I", w) for w in ext_words)
    out += payload
    if pad:
        out += bytes(pad - 1) + bytes([pad])
    return out


def random_packet(rng: np.random.Generator, n_payload: int, pt: int):
    payload = rng.integers(0, 256, n_payload).astype(np.uint8).tobytes()
    kw = dict(pt=pt, seq=int(rng.integers(0, 65536)), ts=int(rng.integers(0, 2**32)), ssrc=int(rng.integers(0, 2**32)),
              marker=bool(rng.integers(2)), csrcs=tuple(int(x) for x in rng.integers(0, 2**32, int(rng.integers(0, 4)))),
              ext_words=None if rng.integers(2) else [int(x) for x in rng.integers(0, 2**32, int(rng.integers(0, 3)))],
              pad=0 if rng.integers(2) else int(rng.integers(1, 9)))
    return build(payload, **kw), payload, kw
