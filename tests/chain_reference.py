"""CPU composition of the oracle pieces into the BASELINE cfg2 graph (test infrastructure):
2x resampler -> bufferizer -> (echo canceller + preprocessor per frame) -> volume per block [-> mixer per 10 ms].
Mirrors what an unmodified MSTicker graph of the reference filters does for lockstep streams."""
from __future__ import annotations

import ctypes as C

import numpy as np

import _oracle as O
from _oracle import OrcVolumeState, ptr


class OracleChain:
    def __init__(self, n_streams, in_rate=16000, rate=48000, tail_ms=250, volume_gain=0.8, mixer_pins=0):
        self.L = L = O.oracle()
        self.n, self.in_rate, self.rate, self.pins = n_streams, in_rate, rate, mixer_pins
        self.tick_in, self.tick = in_rate // 100, rate // 100
        self.rs_ref = [L.orc_resampler_new(1, in_rate, rate, 3) for _ in range(n_streams)]
        self.rs_mic = [L.orc_resampler_new(1, in_rate, rate, 3) for _ in range(n_streams)]
        self.aec = [L.orc_aec_new(rate, tail_ms, 64) for _ in range(n_streams)]
        self.F = L.orc_aec_frame_size(self.aec[0])
        self.vol = []
        for _ in range(n_streams):
            st = OrcVolumeState()
            L.orc_volume_init(C.byref(st), rate)
            st.gain = st.target_gain = st.static_gain = volume_gain
            self.vol.append(st)
        self.buf_ref = [np.zeros(0, np.int16) for _ in range(n_streams)]
        self.buf_mic = [np.zeros(0, np.int16) for _ in range(n_streams)]
        self.buf_mix = [np.zeros(0, np.int16) for _ in range(n_streams)]

    def tick_stream(self, s, ref_in, mic_in):
        """returns the EC+volume output samples produced by stream s this tick"""
        L = self.L
        o1 = np.zeros(self.tick + 8, np.int16)
        o2 = np.zeros(self.tick + 8, np.int16)
        n1 = L.orc_msresample_block(self.rs_ref[s], ptr(np.ascontiguousarray(ref_in)), self.tick_in, ptr(o1))
        n2 = L.orc_msresample_block(self.rs_mic[s], ptr(np.ascontiguousarray(mic_in)), self.tick_in, ptr(o2))
        self.buf_ref[s] = np.concatenate([self.buf_ref[s], o1[:n1]])
        self.buf_mic[s] = np.concatenate([self.buf_mic[s], o2[:n2]])
        outs = []
        F = self.F
        while len(self.buf_mic[s]) >= F:
            mic, ref = np.ascontiguousarray(self.buf_mic[s][:F]), np.ascontiguousarray(self.buf_ref[s][:F])
            self.buf_mic[s], self.buf_ref[s] = self.buf_mic[s][F:], self.buf_ref[s][F:]
            out = np.zeros(F, np.int16)
            L.orc_aec_process_frame(self.aec[s], ptr(mic), ptr(ref), ptr(out))
            L.orc_volume_process(C.byref(self.vol[s]), ptr(out), F)
            outs.append(out)
        return np.concatenate(outs) if outs else np.zeros(0, np.int16)

    def tick_all(self, ref_in, mic_in):
        """ref_in/mic_in [n][tick_in] -> without mixer: [n][k*F]; with mixer: [n][tick] (conference outputs)"""
        outs = [self.tick_stream(s, ref_in[s], mic_in[s]) for s in range(self.n)]
        if not self.pins:
            return np.stack(outs)
        L, P, T = self.L, self.pins, self.tick
        res = np.zeros((self.n, T), np.int16)
        for s in range(self.n):
            self.buf_mix[s] = np.concatenate([self.buf_mix[s], outs[s]])
        for room in range(self.n // P):
            blk = np.zeros((P, T), np.int16)
            present = np.zeros(P, np.uint8)
            for p in range(P):
                s = room * P + p
                if len(self.buf_mix[s]) >= T:
                    blk[p] = self.buf_mix[s][:T]
                    self.buf_mix[s] = self.buf_mix[s][T:]
                    present[p] = 1
            gain, active = np.ones(P, np.float32), np.ones(P, np.uint8)
            out = np.zeros((P, T), np.int16)
            L.orc_mixer_process(1, P, T, 1, ptr(gain), ptr(active), ptr(blk), ptr(present), ptr(out))
            res[room * P:(room + 1) * P] = out
        return res

    def close(self):
        L = self.L
        for r in self.rs_ref + self.rs_mic:
            L.orc_resampler_free(r)
        for a in self.aec:
            L.orc_aec_free(a)


def smoke_chain(ctx, n_streams=4, ticks=12):
    """used by __graft_entry__.smoke(): a few ticks of the resident device chain vs the CPU composition"""
    from mediastreamer2_b200 import filters as F
    from synth import cfg2_stream

    data = [cfg2_stream(s, 160 * ticks, 16000) for s in range(n_streams)]
    ref = np.stack([d[0] for d in data]).reshape(n_streams, ticks, 160)
    mic = np.stack([d[1] for d in data]).reshape(n_streams, ticks, 160)
    ch = F.AudioChain(ctx, n_streams)
    oc = OracleChain(n_streams)
    worst = 0
    for t in range(ticks):
        out, n = ch.tick(np.ascontiguousarray(ref[:, t]), np.ascontiguousarray(mic[:, t]))
        exp = oc.tick_all(ref[:, t], mic[:, t])
        assert exp.shape[1] == n, (t, exp.shape, n)
        if n:
            worst = max(worst, int(np.abs(out[:, :n].astype(np.int32) - exp.astype(np.int32)).max()))
    assert worst <= 2, f"chain deviates from the oracle by {worst} LSB"
    ch.close()
    oc.close()
    return worst
