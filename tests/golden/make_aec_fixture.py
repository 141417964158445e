"""Cuts the first 11 s of the reference's own echo-canceller test material (tester/sounds/{farend,echo,nearend}_simple_talk.wav,
used by tester/mediastreamer2_aec3_tester.c:45-47, 710-724 "Simple talk") into tests/golden/aec_simple_talk.npz so that
the behavioural AEC tests can run where /root/reference does not exist (the GPU box). Run in the build container:

    python tests/golden/make_aec_fixture.py
"""
import wave
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SOUNDS = Path("/root/reference/tester/sounds")


def rd(name):
    w = wave.open(str(SOUNDS / f"{name}.wav"))
    assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
    return np.frombuffer(w.readframes(w.getnframes()), np.int16).copy()


n = 11 * 16000
np.savez_compressed(HERE / "aec_simple_talk.npz", rate=np.array([16000]), farend=rd("farend_simple_talk")[:n],
                    echo=rd("echo_simple_talk")[:n], nearend=rd("nearend_simple_talk")[:n])
print((HERE / "aec_simple_talk.npz").stat().st_size, "bytes")
