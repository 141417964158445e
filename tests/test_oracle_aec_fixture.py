"""MSSpeexEC has no test of its own in the reference and speexdsp is not in the tree (parity of oracle/oracle_aec.c is
"unpinned"). What CAN be anchored on the reference: its own echo-canceller test material. The oracle
must behave as an echo canceller on the reference tester's "simple talk" scenario
(tester/mediastreamer2_aec3_tester.c:45-47, 710-724; fixture cut by tests/golden/make_aec_fixture.py)."""
import numpy as np

import _oracle as O
import aec_fixture as A
from _oracle import ptr


def test_oracle_aec_on_the_reference_testers_simple_talk_material():
    L = O.oracle()
    far, mic, near = A.load()
    a = L.orc_aec_new(A.RATE, 250, 64)
    Fs = L.orc_aec_frame_size(a)
    assert Fs == 128
    out = np.zeros_like(mic)
    for k in range(len(mic) // Fs):
        s = slice(k * Fs, (k + 1) * Fs)
        L.orc_aec_process_frame(a, ptr(mic[s]), ptr(far[s]), ptr(out[s]))
    L.orc_aec_free(a)
    erle, keep, corr = A.check_behaviour(out, mic, near, min_erle_db=25.0)
    print("ERLE dB", erle, "near-end level dB", keep, "near-end correlation", corr)
