"""BASELINE cfg5 graph in the UNMODIFIED reference MSTicker, built from the plugin's B200 filters:

    per stream:  far-end source -> MSResample(16k->48k) -> MSSpeexEC.in0 ; MSSpeexEC.out0 -> speaker sink
                 mic source     -> MSResample(16k->48k) -> MSSpeexEC.in1 ; MSSpeexEC.out1 -> MSVolume(0.8)
                                -> MSAudioMixer (conference mode, rooms of P) pin p ; mixer pin p -> encoder-stub sink

Run as a script (the plugin's execution mode is fixed per process by MSB200_BATCH, so tests and the bench spawn it):

    MSB200_BATCH=0|<slots> python tests/graph_runner.py --streams 8 --pins 4 --ticks 60 --dump out.npz [--timing]
                                                        [--tickers K]   rooms are dealt round-robin to K MSTickers
                                                        (K threads; SURVEY.md cfg5: one ticker per 256 streams)

--dump  writes every sink's sample stream (parity: batch mode == synchronous mode, one ticker interval later per stage)
--timing prints one JSON line: wall time per tick of the free-running (gated, never sleeping) ticker, p50/p99, launches.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import _oracle as O  # noqa: E402
from _oracle import RefGraph  # noqa: E402
from synth import cfg2_stream  # noqa: E402

IN_RATE, RATE, TAIL_MS, GAIN = 16000, 48000, 250, 0.8


def _alaw(pcm: np.ndarray) -> np.ndarray:
    L = O.oracle()
    pcm = np.ascontiguousarray(pcm, np.int16)
    code = np.zeros(pcm.size, np.uint8)
    L.orc_g711_encode(0, O.ptr(pcm), O.ptr(code), pcm.size)
    return code


def build(g: RefGraph, n_streams: int, pins: int, ticks: int, pool: int = 8, codec: str = "none"):
    """codec == "alaw": the decode / encode stubs of cfg5 made real — the sources emit 8 kHz G.711 A-law payloads (80 bytes
    per tick) into MSAlawDec, and every mixer output goes through MSResample(48k->8k) -> MSAlawEnc (20 ms packets)"""
    in_rate = 8000 if codec == "alaw" else IN_RATE
    ti = in_rate // 100
    base = [cfg2_stream(s, ti * ticks, in_rate) for s in range(min(pool, n_streams))]
    sources, spk_sinks, out_sinks, mixers = [], [], [], []  # sources: ONE per room (attaching it schedules the whole room)
    for s in range(n_streams):
        if s % pins == 0:
            mix = g.new("MSAudioMixer")
            g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", RATE)
            g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", 1)
            mixers.append(mix)
        far, mic = base[s % len(base)][0], base[s % len(base)][1]
        if s >= len(base):  # distinct streams from a small pool: rotate so that rooms do not mix identical signals
            far, mic = np.roll(far, 37 * s), np.roll(mic, 37 * s)
        if codec == "alaw":
            s_far, s_mic = g.source(_alaw(far), ti), g.source(_alaw(mic), ti)
            d_far, d_mic = g.new("MSAlawDec"), g.new("MSAlawDec")
        else:
            s_far, s_mic = g.source(far, ti * 2), g.source(mic, ti * 2)
        rs_far, rs_mic = g.new("MSResample"), g.new("MSResample")
        for r in (rs_far, rs_mic):
            g.call_int(r, "MS_FILTER_SET_SAMPLE_RATE", in_rate)
            g.call_int(r, "MS_FILTER_SET_OUTPUT_SAMPLE_RATE", RATE)
        ec = g.new("MSSpeexEC")
        g.call_int(ec, "MS_FILTER_SET_SAMPLE_RATE", RATE)
        g.call_int(ec, "MS_ECHO_CANCELLER_SET_TAIL_LENGTH", TAIL_MS)
        vol = g.new("MSVolume")
        g.call_int(vol, "MS_FILTER_SET_SAMPLE_RATE", RATE)
        g.call_float(vol, "MS_VOLUME_SET_GAIN", GAIN)
        spk, out = g.sink(), g.sink()
        if codec == "alaw":
            g.link(s_far, 0, d_far, 0)
            g.link(d_far, 0, rs_far, 0)
            g.link(s_mic, 0, d_mic, 0)
            g.link(d_mic, 0, rs_mic, 0)
        else:
            g.link(s_far, 0, rs_far, 0)
            g.link(s_mic, 0, rs_mic, 0)
        g.link(rs_far, 0, ec, 0)
        g.link(ec, 0, spk, 0)
        g.link(rs_mic, 0, ec, 1)
        g.link(ec, 1, vol, 0)
        g.link(vol, 0, mixers[-1], s % pins)
        if codec == "alaw":
            rs_out, enc = g.new("MSResample"), g.new("MSAlawEnc")
            g.call_int(rs_out, "MS_FILTER_SET_SAMPLE_RATE", RATE)
            g.call_int(rs_out, "MS_FILTER_SET_OUTPUT_SAMPLE_RATE", in_rate)
            g.link(mixers[-1], s % pins, rs_out, 0)
            g.link(rs_out, 0, enc, 0)
            g.link(enc, 0, out, 0)
        else:
            g.link(mixers[-1], s % pins, out, 0)
        if s % pins == 0:
            sources.append(s_far)
        spk_sinks.append(spk)
        out_sinks.append(out)
    return sources, spk_sinks, out_sinks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--pins", type=int, default=4)
    ap.add_argument("--ticks", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--dump", default="")
    ap.add_argument("--timing", action="store_true")
    ap.add_argument("--tickers", type=int, default=1)
    ap.add_argument("--codec", choices=["none", "alaw"], default="none")
    ap.add_argument("--churn", default="", help="T1,T2: detach the LAST room's graph after T1 ticks and attach it again after T2 ticks "
                    "(its filters leave and re-join their batch groups; the other rooms must not notice)")
    a = ap.parse_args()
    g = RefGraph(plugins_dir=str(O.PLUGIN_DIR))
    sources, spk_sinks, out_sinks = build(g, a.streams, a.pins, a.ticks, codec=a.codec)
    if a.timing and not a.dump:  # timing runs: the recording sinks only count (no growing buffers inside the timed ticks)
        for k in spk_sinks + out_sinks:
            g.L.ref_sink_set_discard(k, 1)
    # a room's streams are one connected graph through its mixer: whole rooms are dealt to the tickers
    tickers = [g.L.ref_ticker_new() for _ in range(max(1, a.tickers))]
    for r, src in enumerate(sources):
        assert g.L.ref_ticker_attach(tickers[r % len(tickers)], src) == 0

    def run(n):  # every ticker advances n ticks, concurrently
        for t in tickers:
            g.L.ref_ticker_release(t, n)
        for t in tickers:
            g.L.ref_ticker_wait(t)

    per_tick = []
    if a.churn:
        t1, t2 = (int(v) for v in a.churn.split(","))
        last = len(sources) - 1
        tk = tickers[last % len(tickers)]
        run(t1)
        assert g.L.ref_ticker_detach(tk, sources[last]) == 0   # postprocess: the room's filters leave their groups
        run(t2 - t1)
        assert g.L.ref_ticker_attach(tk, sources[last]) == 0   # preprocess: they join again (fresh slots, reset state)
        run(a.ticks - t2)
    elif a.timing:
        run(a.warmup)
        if int(os.environ.get("MSB200_PROFILE", "0") or 0) > 0:  # the profile covers the timed ticks only (no joins, no set-up)
            C.CDLL(str(O.PLUGIN_DIR / "libmsb200filters.so")).msb200_filters_host_profile_reset()
        for _ in range(a.ticks - a.warmup):
            t0 = time.perf_counter()
            run(1)
            per_tick.append(time.perf_counter() - t0)
    else:
        run(a.ticks)
    if a.dump:
        np.savez(a.dump, **{f"out{i}": g.read(k, np.uint8 if a.codec == "alaw" else np.int16)[0] for i, k in enumerate(out_sinks)},
                 **{f"spk{i}": g.read(k)[0] for i, k in enumerate(spk_sinks)})
    if a.timing:
        ms = np.array(per_tick) * 1000.0
        stats = {"mode": "batch" if int(os.environ.get("MSB200_BATCH", "0") or 0) > 0 else "sync",
                 "batch_slots": int(os.environ.get("MSB200_BATCH", "0") or 0), "streams": a.streams, "pins": a.pins,
                 "tickers": len(tickers), "codec": a.codec,
                 "ticks_timed": len(ms), "tick_ms_mean": float(ms.mean()), "tick_ms_p50": float(np.percentile(ms, 50)),
                 "tick_ms_p99": float(np.percentile(ms, 99)), "tick_ms_max": float(ms.max()),
                 "late_ticks_10ms": int((ms > 10.0).sum()),
                 "stream_ticks_per_s": float(a.streams / (ms.mean() / 1000.0))}
        try:
            plug = C.CDLL(str(O.PLUGIN_DIR / "libmsb200filters.so"))
            gr, fl, un = C.c_int(), C.c_ulonglong(), C.c_ulonglong()
            plug.msb200_filters_batch_stats(C.byref(gr), C.byref(fl), C.byref(un))
            stats.update({"batch_groups": gr.value, "batch_launches": fl.value})
            if int(os.environ.get("MSB200_PROFILE", "0") or 0) > 0:  # thread-milliseconds inside our process() calls, per kind,
                buf = C.create_string_buffer(4096)                    # over the timed ticks; the rest of a ticker thread's
                plug.msb200_filters_host_profile(buf, 4096)           # time is the reference's own runtime (and the harness)
                stats["host_profile_ms"] = {k: v["ms"] for k, v in json.loads(buf.value.decode()).items()}
                stats["thread_ms_total"] = float(ms.sum() * len(tickers))
        except (OSError, AttributeError):
            pass
        print(json.dumps(stats), flush=True)
    for r, src in enumerate(sources):
        g.L.ref_ticker_detach(tickers[r % len(tickers)], src)
    for t in tickers:
        g.L.ref_ticker_destroy(t)
    g.close()


if __name__ == "__main__":
    main()
