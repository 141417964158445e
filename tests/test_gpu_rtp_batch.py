"""GPU: the batched RTP payload hand-off (csrc/rtp_batch.cu). Receive: hand-built RTP packets (CSRC lists,
header extensions, padding) of many streams -> one decode launch -> PCM equal to the oracle's G.711 decode of each payload,
meta data equal to the header fields MSRtpRecv copies into the mblk_t (msrtp.c:1078-1080). Send: PCM -> packets whose
payload is the oracle's encoding and whose headers count sequence numbers and timestamps as RFC 3550 says."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import rtp_packets as RP
from mediastreamer2_b200 import _lib

pytestmark = pytest.mark.gpu


def _orc_decode(L, law, codes: np.ndarray) -> np.ndarray:
    out = np.zeros(codes.size, np.int16)
    L.orc_g711_decode(law, O.ptr(np.ascontiguousarray(codes)), O.ptr(out), codes.size)
    return out


def _orc_encode(L, law, pcm: np.ndarray) -> np.ndarray:
    out = np.zeros(pcm.size, np.uint8)
    L.orc_g711_encode(law, O.ptr(np.ascontiguousarray(pcm)), O.ptr(out), pcm.size)
    return out


@pytest.mark.parametrize("law", [0, 1])  # MSB200_G711_ALAW, MSB200_G711_ULAW
def test_rx_packets_to_pcm(ctx, law):
    L = O.oracle()
    lib = ctx.lib
    n, maxp = 97, 240
    pt = 8 if law == 0 else 0
    h = C.c_void_p()
    _lib.check(lib.msb200_rtp_rx_create(ctx.h, n, law, maxp, C.byref(h)))
    row = lib.msb200_rtp_rx_row_samples(h)
    rng = np.random.default_rng(law)
    for tick in range(3):
        _lib.check(lib.msb200_rtp_rx_begin_tick(h))
        sent = {}
        for s in range(n):
            if (s + tick) % 5 == 0:
                continue  # nothing from this stream in this tick
            pkt, payload, kw = RP.random_packet(rng, int(rng.choice([80, 160, 240, 33])), pt)
            assert lib.msb200_rtp_rx_push(h, s, pkt, len(pkt), pt) == len(payload)
            sent[s] = (payload, kw)
        # a packet of another payload type is ignored, a malformed one refused; neither disturbs the stream's row
        other = RP.build(b"\x00" * 80, 96, 1, 1, 1)
        assert lib.msb200_rtp_rx_push(h, 0, other, len(other), pt) == 0
        assert lib.msb200_rtp_rx_push(h, 0, other[:5], 5, pt) == _lib.EINVAL
        pcm = np.zeros((n, row), np.int16)
        meta = (_lib.RtpMeta * n)()
        _lib.check(lib.msb200_rtp_rx_decode(h, O.ptr(pcm), meta))
        for s in range(n):
            if s not in sent:
                assert meta[s].payload_len == 0
                continue
            payload, kw = sent[s]
            m = meta[s]
            assert (m.payload_len, m.timestamp, m.seq, m.marker, m.payload_type, m.ssrc) == (
                len(payload), kw["ts"], kw["seq"], int(kw["marker"]), pt, kw["ssrc"])
            exp = _orc_decode(L, law, np.frombuffer(payload, np.uint8))
            assert np.array_equal(pcm[s, :len(payload)], exp), (tick, s)
    lib.msb200_rtp_rx_destroy(h)


@pytest.mark.parametrize("law", [0, 1])
def test_tx_pcm_to_packets(ctx, law):
    L = O.oracle()
    lib = ctx.lib
    n, spp = 64, 160
    h = C.c_void_p()
    _lib.check(lib.msb200_rtp_tx_create(ctx.h, n, law, spp, C.byref(h)))
    pkt_bytes = lib.msb200_rtp_tx_packet_bytes(h)
    assert pkt_bytes == 12 + spp
    for s in range(n):
        _lib.check(lib.msb200_rtp_tx_set_stream(h, s, 0x1000 + s, 8 if law == 0 else 0, 65530 + s, 2**32 - 200 + s))
    rng = np.random.default_rng(5 + law)
    sent_count = np.zeros(n, np.int64)
    for tick in range(4):
        pcm = rng.integers(-32768, 32768, size=(n, spp)).astype(np.int16)
        marker = (rng.integers(0, 2, n)).astype(np.uint8)
        send = np.ones(n, np.uint8)
        send[tick::7] = 0
        out = C.c_void_p()
        _lib.check(lib.msb200_rtp_tx_encode(h, O.ptr(pcm), O.ptr(marker), O.ptr(send), C.byref(out)))
        arena = np.ctypeslib.as_array((C.c_uint8 * (n * pkt_bytes)).from_address(out.value)).reshape(n, pkt_bytes).copy()
        for s in range(n):
            if not send[s]:
                continue
            m, off = _lib.RtpMeta(), C.c_size_t()
            assert lib.msb200_rtp_parse(arena[s].tobytes(), pkt_bytes, C.byref(m), C.byref(off)) == 0
            k = int(sent_count[s])
            assert off.value == 12 and m.payload_len == spp
            assert m.seq == (65530 + s + k) % 65536 and m.timestamp == (2**32 - 200 + s + k * spp) % 2**32  # both wrap
            assert m.ssrc == 0x1000 + s and m.marker == marker[s] and m.payload_type == (8 if law == 0 else 0)
            assert np.array_equal(arena[s, 12:], _orc_encode(L, law, pcm[s])), (tick, s)
            sent_count[s] += 1
    lib.msb200_rtp_tx_destroy(h)
