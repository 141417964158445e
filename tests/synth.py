"""Synthetic stream generators shared by tests and bench.py: deterministic per stream id."""
from __future__ import annotations

import numpy as np


def cfg2_stream(stream: int, n_samples: int, rate: int = 16000, echo_delay_ms: float = 20.0, echo_tail_ms: float = 40.0,
                echo_gain_db: float = -12.0):
    """far-end x (amplitude-modulated tone mix), mic = near-end (active in the envelope gaps) + echo(x).

    Returns (far_end s16, mic s16, echo_only float, near_only float)."""
    seed = 0xB2000000 + stream
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / rate
    f = 300 + (seed * 2654435761 % 3100)
    f2 = 300 + (seed * 40503 % 3100)
    env = (np.sin(2 * np.pi * 2.0 * t + 0.3 * (stream % 7)) > -0.2).astype(np.float64)  # ~2 Hz on/off
    # smooth the envelope edges (5 ms) to avoid clicks
    k = max(1, int(0.005 * rate))
    env = np.convolve(env, np.ones(k) / k, mode="same")
    # speech-like far end: a voiced tone pair plus low-passed noise (broadband excitation lets the canceller converge)
    w = rng.standard_normal(n_samples)
    lp = np.empty(n_samples)
    acc = 0.0
    a = 0.55
    for i in range(0, n_samples, 4096):  # one-pole low-pass, blockwise with scipy-free recursion
        seg = w[i:i + 4096]
        y = np.empty(len(seg))
        for k, v in enumerate(seg):
            acc = a * acc + (1 - a) * v
            y[k] = acc
        lp[i:i + 4096] = y
    x = 2500 * np.sin(2 * np.pi * f * t) + 1200 * np.sin(2 * np.pi * (f * 1.7 + 50) * t) + 9000 * lp
    x *= env
    # echo path: exponentially decaying random FIR, delayed
    L = int(echo_tail_ms * rate / 1000)
    h = rng.standard_normal(L) * np.exp(-np.arange(L) / (L / 4.0))
    h *= 10 ** (echo_gain_db / 20) / np.sqrt(np.sum(h * h))
    d = int(echo_delay_ms * rate / 1000)
    echo = np.concatenate([np.zeros(d), np.convolve(x, h)[: n_samples - d]])
    near = 3000 * np.sin(2 * np.pi * f2 * t) * (1.0 - env)
    mic = echo + near + 20 * rng.standard_normal(n_samples)
    to16 = lambda v: np.clip(np.round(v), -32768, 32767).astype(np.int16)
    return to16(x), to16(mic), echo, near
