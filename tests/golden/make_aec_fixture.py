"""Cuts the reference's own echo-canceller test material (tester/sounds/*.wav, 16 kHz mono, used by
tester/mediastreamer2_aec3_tester.c:40-50 and its scenarios :601-812) into tests/golden/aec_talk_16k.npz, so that the AEC
anchors can run where /root/reference does not exist (the GPU box). Run in the build container:

    python tests/golden/make_aec_fixture.py

  simple_*   "Simple talk" (far-end and near-end talkers alternate), first 15 s — the suite analyses 11.0 - 14.5 s
  double_*   "Double talk", first 14 s — the suite analyses 9.5 - 13.5 s
  delay_echo "Simple talk with delay change": the simple-talk echo with a 50 ms jump around 9 s, first 15 s
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


ns, nd = 15 * 16000, 14 * 16000
np.savez_compressed(HERE / "aec_talk_16k.npz", rate=np.array([16000]),
                    simple_far=rd("farend_simple_talk")[:ns], simple_echo=rd("echo_simple_talk")[:ns],
                    simple_near=rd("nearend_simple_talk")[:ns],
                    double_far=rd("farend_double_talk")[:nd], double_echo=rd("echo_double_talk")[:nd],
                    double_near=rd("nearend_double_talk")[:nd],
                    delay_echo=rd("echo_delay_change")[:ns])
print((HERE / "aec_talk_16k.npz").stat().st_size, "bytes")
