"""MSSpeexEC has no test of its own in the reference and speexdsp is not in the tree (parity of oracle/oracle_aec.c is
"unpinned"). What CAN be anchored on the reference: its own echo-canceller material and its own metric
(tests/aec_fixture.py). The oracle runs every scenario of the reference's AEC suite that the committed material covers."""
import numpy as np
import pytest

import _oracle as O
import aec_fixture as A
from _oracle import ptr


def oracle_run(mic: np.ndarray, far: np.ndarray) -> np.ndarray:
    L = O.oracle()
    a = L.orc_aec_new(A.RATE, 250, 64)
    Fs = L.orc_aec_frame_size(a)
    assert Fs == 128
    out = np.zeros_like(mic)
    for k in range(len(mic) // Fs):
        s = slice(k * Fs, (k + 1) * Fs)
        L.orc_aec_process_frame(a, ptr(mic[s]), ptr(far[s]), ptr(out[s]))
    L.orc_aec_free(a)
    return out


def test_oracle_aec_on_the_reference_testers_simple_talk_material():
    far, mic, near = A.load()
    out = oracle_run(mic, far)
    erle, keep, corr = A.check_behaviour(out, mic, near, min_erle_db=25.0)
    print("ERLE dB", erle, "near-end level dB", keep, "near-end correlation", corr)


@pytest.mark.parametrize("name", list(A.SCENARIOS))
def test_oracle_aec_reference_suite_metric(name, tmp_path):
    """ms_audio_compare_silence_and_speech (unmodified audiodiff.c) with the suite's windows on the oracle's output"""
    R = O.ref()
    g = A.load_all()
    far, mic, near = A.scenario_signals(g, name)
    out = oracle_run(mic, far)
    sim, energy = A.silence_and_speech(R, tmp_path, near, out, name)
    min_sim, max_energy = A.SCENARIOS[name][6]
    print(name, "similarity", sim, "energy in silence", energy, "suite bounds", A.SCENARIOS[name][5])
    assert sim >= min_sim and sim < 1.0, sim
    assert energy <= max_energy, energy
    if far.any():
        # the canceller alone: against the same chain fed the near-end talker only (notch + preprocessor cancel out)
        ideal = oracle_run(near.copy(), np.zeros_like(near))
        sim_i, _ = A.silence_and_speech(R, tmp_path, ideal, out, name)
        print(name, "similarity with the echo-free run", sim_i)
        assert sim_i >= A.ISOLATED_MIN[name], sim_i  # measured 0.9998 / 0.9943 / 0.923
