#!/usr/bin/env python
"""Cuts the reference's own multi-rate speech corpus into a small fixture for the resampler anchors
(tests/test_oracle_resample.py, tests/test_gpu_resample_anchor.py). Run HERE (needs /root/reference); the GPU box only sees
the committed tests/golden/resample_voice.npz.

  tester/sounds/test_silence_voice_{8000,16000,32000,44100,48000}.wav — the SAME 25.7 s recording at five rates (the
      reference's audio-stream tests play them, tester/mediastreamer2_basic_audio_tester.c:58-62): seconds 8.0 - 10.5
      (speech) of each, sample-aligned (start = 8 s x rate)
  tester/sounds/hello8000.wav — BASELINE cfg1's input (header length field is bogus: payload = everything after the 44-byte
      header): the first 2.5 s
"""
import wave
from pathlib import Path

import numpy as np

REF = Path("/root/reference/tester/sounds")
OUT = Path(__file__).resolve().parent / "resample_voice.npz"


def main():
    data = {}
    for rate in (8000, 16000, 32000, 44100, 48000):
        w = wave.open(str(REF / f"test_silence_voice_{rate}.wav"))
        assert w.getframerate() == rate and w.getnchannels() == 1 and w.getsampwidth() == 2
        pcm = np.frombuffer(w.readframes(w.getnframes()), np.int16)
        data[f"voice_{rate}"] = pcm[8 * rate:8 * rate + (5 * rate) // 2].copy()
    raw = (REF / "hello8000.wav").read_bytes()[44:]
    data["hello_8000"] = np.frombuffer(raw[:2 * (len(raw) // 2)], np.int16)[:20000].copy()
    np.savez_compressed(OUT, **data)
    print(OUT, OUT.stat().st_size, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
