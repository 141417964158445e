"""Independent anchors for an MSResample implementation (the oracle on the CPU, the device bank on the GPU).

speexdsp is not in the reference tree, so neither implementation can be pinned bit for bit against the library. These
checks do not depend on our restatement of it:

  sine_fit()      a windowed-sinc resampler is linear-phase: an in-band sine must come out with unit gain, delayed by exactly
                  filt_len / 2 INPUT samples, and nothing else but rounding noise. A wrong table constant, cut-off, phase
                  step or history length shows up as gain, delay or residual.
  similarity()    normalised cross-correlation at the nominal delay against scipy.signal.resample_poly (an unrelated
                  polyphase design) and against the reference's OWN recording of the same speech at the other rate
                  (tester/sounds/test_silence_voice_*.wav, cut into tests/golden/resample_voice.npz).
  ms_audio_diff   the reference's own similarity measure (src/utils/audiodiff.c:578-651), run from the unmodified source in
                  oracle/_ref when it is available.
"""
from __future__ import annotations

import wave
from math import gcd
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "resample_voice.npz"

# the 8 ratios tests/test_gpu_audio.py runs + the two 44.1 kHz directions (interpolated mode both ways)
RATIOS = [(8000, 48000), (16000, 48000), (48000, 16000), (48000, 8000), (44100, 48000), (16000, 8000), (8000, 16000),
          (32000, 48000), (48000, 44100), (32000, 16000)]


def nominal_delay_out(filt_len: int, in_rate: int, out_rate: int) -> float:
    """group delay in OUTPUT samples: filt_len / 2 input samples (no zero-skipping: msresample.c never calls skip_zeros)"""
    return filt_len / 2 * out_rate / in_rate


def sine_fit(resample, in_rate: int, out_rate: int, filt_len: int, frac_of_nyquist: float, amp: float = 10000.0):
    """-> (gain_db, delay_error_in_output_samples, max_residual_lsb). `resample(x) -> y` processes 10 ms blocks."""
    f = frac_of_nyquist * min(in_rate, out_rate) / 2
    n = np.arange(in_rate)  # 1 s
    x = np.round(amp * np.sin(2 * np.pi * f * n / in_rate)).astype(np.int16)
    y = resample(x).astype(np.float64)
    s0, s1 = 2000, len(y) - 2000
    t = (np.arange(s0, s1) - nominal_delay_out(filt_len, in_rate, out_rate)) / out_rate
    A = np.stack([np.sin(2 * np.pi * f * t), np.cos(2 * np.pi * f * t)], 1)
    coef = np.linalg.lstsq(A, y[s0:s1], rcond=None)[0]
    gain_db = 20 * np.log10(np.hypot(*coef) / amp)
    delay_err = -np.arctan2(coef[1], coef[0]) / (2 * np.pi * f) * out_rate
    resid = np.abs(y[s0:s1] - A @ coef).max()
    return float(gain_db), float(delay_err), float(resid)


def ncorr(a: np.ndarray, b: np.ndarray) -> float:
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.dot(a, b) / np.sqrt(np.dot(a, a) * np.dot(b, b)))


def scipy_resample(x: np.ndarray, in_rate: int, out_rate: int) -> np.ndarray:
    from scipy import signal

    g = gcd(in_rate, out_rate)
    return signal.resample_poly(x.astype(np.float64), out_rate // g, in_rate // g)


def aligned(y: np.ndarray, other: np.ndarray, filt_len: int, in_rate: int, out_rate: int, guard: int = 300):
    """y with its nominal delay removed, cut to the common support with `other` (a zero-delay signal at out_rate)"""
    d = int(round(nominal_delay_out(filt_len, in_rate, out_rate)))
    n = min(len(y) - d, len(other)) - guard
    return y[d + guard:d + n], other[guard:n]


def write_wav(path, pcm: np.ndarray, rate: int):
    w = wave.open(str(path), "wb")
    w.setnchannels(1)
    w.setsampwidth(2)
    w.setframerate(rate)
    w.writeframes(np.ascontiguousarray(pcm, np.int16).tobytes())
    w.close()
