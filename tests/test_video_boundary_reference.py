"""CPU: the reference's OWN MSPixConv / MSSizeConv (unmodified, in oracle/_ref) over a scaler that calls the oracle. Pins
the test harness the GPU legs compare against (tests/test_gpu_video_plugin.py) and the filters' host semantics our
replacements must show: pass-through, header, timestamps, frame pacing, the aspect-ratio correction with its event."""
import numpy as np
import pytest

import _oracle as O
import video_graph as V
from _oracle import RefGraph


@pytest.fixture()
def oracle_scaler():
    R = O.ref()
    d = V.OracleScalerDesc()
    d.install(R)
    yield d
    R.ref_set_scaler_desc(None)


def test_reference_pixconv_then_sizeconv_over_the_oracle_scaler(oracle_scaler):
    w, h, tw, th = 64, 48, 32, 24
    frames = [V.synth_frame(V.MS_YUY2, w, h, t) for t in range(4)]
    data, tri, dims = V.run_pixconv_sizeconv(RefGraph(), frames, V.MS_YUY2, w, h, target=(tw, th), want_b200=False)
    assert len(tri) == 4 and (dims == (tw, th)).all()
    assert (tri[:, 1] == tw * th * 3 // 2).all()
    assert tri[:, 2].tolist() == [1000 + 90 * k for k in range(4)]  # timestamps travel through both filters
    assert oracle_scaler.calls == 8
    L = O.oracle()
    a = L.orc_scaler_new(w, h, 6, w, h, 0)
    b = L.orc_scaler_new(w, h, 0, tw, th, 0)
    for k, fr in enumerate(frames):
        mid = np.zeros(w * h * 3 // 2, np.uint8)
        exp = np.zeros(tw * th * 3 // 2, np.uint8)
        L.orc_scaler_process(a, O.ptr(fr), O.ptr(mid))
        L.orc_scaler_process(b, O.ptr(mid), O.ptr(exp))
        assert np.array_equal(data[k * exp.size:(k + 1) * exp.size], exp), k
    L.orc_scaler_free(a)
    L.orc_scaler_free(b)


def test_reference_sizeconv_frame_pacing_and_same_size_pass_through(oracle_scaler):
    w, h = 32, 24
    frames = [V.synth_frame(V.MS_YUV420P, w, h, t) for t in range(12)]
    # 25 fps asked of a 100 fps source (one frame per 10 ms tick): one frame in four goes on, untouched (same size)
    data, tri, dims = V.run_pixconv_sizeconv(RefGraph(), frames, V.MS_YUV420P, w, h, target=(w, h), fps=25.0)
    # (the frame kept back at tick 11 is never sent: the filter is not a pump and runs only while frames arrive)
    assert len(tri) == 2 and (dims == (w, h)).all()
    sent = [int((ts - 1000) // 90) for ts in tri[:, 2]]
    assert sent == [4, 8]
    for k, idx in enumerate(sent):
        assert np.array_equal(data[k * frames[0].size:(k + 1) * frames[0].size], frames[idx])
    assert oracle_scaler.calls == 0
