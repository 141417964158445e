"""GPU: the video drop-in boundary (SURVEY §8b "second boundary", rows a7 / a8 / a11) through the reference's own loader
and ticker. Expected frames come from the reference's unmodified MSPixConv / MSSizeConv over the oracle scaler
(tests/video_graph.py, pinned on the CPU by tests/test_video_boundary_reference.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import _oracle as O
import video_graph as V
from _oracle import RefGraph

pytestmark = pytest.mark.gpu


def _expected(frames, fmt, w, h, target=None, fps=None, ticks_per_frame=1):
    R = O.ref()
    d = V.OracleScalerDesc()
    d.install(R)
    try:
        return V.run_pixconv_sizeconv(RefGraph(), frames, fmt, w, h, target=target, fps=fps, ticks_per_frame=ticks_per_frame,
                                      want_b200=False)
    finally:
        R.ref_set_scaler_desc(None)


def _plugin_graph():
    assert (O.PLUGIN_DIR / "libmsb200filters.so").exists(), "plugin/lib/libmsb200filters.so must travel with the snapshot"
    return RefGraph(plugins_dir=str(O.PLUGIN_DIR))


def _same(a, b):
    (da, ta, ma), (db, tb, mb) = a, b
    assert len(ta) == len(tb) and len(ta) > 0
    assert np.array_equal(ma, mb), (ma, mb)                       # the {w, h} headers below b_rptr
    assert np.array_equal(ta[:, 1:], tb[:, 1:]), (ta, tb)         # sizes and timestamps
    assert np.array_equal(da, db)


# targets keep the input's aspect ratio (MSSizeConv corrects any other target and then waits for the host: the dedicated
# test below covers that path)
CASES = [(V.MS_YUY2, 64, 48, (32, 24)), (V.MS_UYVY, 80, 48, (40, 24)), (V.MS_RGB24, 64, 48, (96, 72)), (V.MS_RGB24_REV, 64, 48, None),
         (V.MS_RGBA32, 48, 32, (24, 16)), (V.MS_YUV420P, 96, 64, (48, 32)), (V.MS_YUYV, 640, 480, (320, 240))]


@pytest.mark.parametrize("fmt,w,h,target", CASES)
def test_b200_pixconv_sizeconv_descs_equal_the_reference_filters(fmt, w, h, target):
    """the plugin's own MSPixConv / MSSizeConv (ids 29 / 31 by name through the reference factory), synchronous lanes"""
    frames = [V.synth_frame(fmt, w, h, t) for t in range(5)]
    exp = _expected(frames, fmt, w, h, target)
    got = V.run_pixconv_sizeconv(_plugin_graph(), frames, fmt, w, h, target=target, want_b200=True)
    _same(exp, got)
    assert np.array_equal(exp[1][:, 0], got[1][:, 0])  # same tick: the synchronous lane adds no latency


@pytest.mark.parametrize("fmt,w,h,target", CASES[:4])
def test_reference_filters_over_the_gpu_scaler_desc(fmt, w, h, target):
    """row a11: ms_video_set_scaler_impl(msb200_ms_scaler_desc()), then the reference's OWN filters"""
    frames = [V.synth_frame(fmt, w, h, t) for t in range(3)]
    exp = _expected(frames, fmt, w, h, target)
    R = O.ref()
    g = _plugin_graph()  # loads the plugin (the desc lives there); the filters below are created from the REFERENCE descs
    plug = C.CDLL(str(O.PLUGIN_DIR / "libmsb200filters.so"))
    plug.msb200_ms_scaler_desc.restype = C.c_void_p
    R.ref_set_scaler_desc(C.c_void_p(plug.msb200_ms_scaler_desc()))
    try:
        # the plugin's descs shadow the built-ins by name, so ask the factory without plugins for the filters
        got = V.run_pixconv_sizeconv(RefGraph(), frames, fmt, w, h, target=target, want_b200=False)
    finally:
        R.ref_set_scaler_desc(None)
        g.close()
    _same(exp, got)


def test_b200_sizeconv_frame_pacing_orientation_and_aspect_events():
    w, h = 64, 48
    frames = [V.synth_frame(V.MS_YUV420P, w, h, t) for t in range(12)]
    for kw in (dict(target=(w, h), fps=25.0), dict(target=(32, 24), fps=50.0), dict(target=(24, 32)), dict(target=(32, 32))):
        exp = _expected(frames, V.MS_YUV420P, w, h, **kw)
        got = V.run_pixconv_sizeconv(_plugin_graph(), frames, V.MS_YUV420P, w, h, want_b200=True, **kw)
        assert len(exp[1]) == len(got[1]), kw
        if len(exp[1]):
            _same(exp, got)


def test_b200_video_batch_lane_is_the_synchronous_mode_one_tick_later(monkeypatch):
    """MSB200_BATCH: the filters of one ticker share a lane; frames leave one tick later, byte-identical"""
    fmt, w, h, target = V.MS_YUY2, 64, 48, (32, 24)
    frames = [V.synth_frame(fmt, w, h, t) for t in range(6)]
    exp = _expected(frames, fmt, w, h, target)
    # a fresh process: the plugin reads MSB200_BATCH once
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, 'tests'); import _oracle as O, video_graph as V; from _oracle import RefGraph\\n"
        f"frames=[V.synth_frame({fmt},{w},{h},t) for t in range(6)]\\n"
        f"d,t,m=V.run_pixconv_sizeconv(RefGraph(plugins_dir=str(O.PLUGIN_DIR)),frames,{fmt},{w},{h},target={target},want_b200=True,extra_ticks=5)\\n"
        "np.savez(sys.argv[1], d=d, t=t, m=m)\\n")
    out = str(O.ROOT / "gpurun_out" / "_vbatch.npz") if (O.ROOT / "gpurun_out").exists() else "/tmp/_vbatch.npz"
    env = dict(os.environ, MSB200_BATCH="8")
    subprocess.run([sys.executable, "-c", code.replace("\\n", "\n"), out], check=True, env=env, cwd=str(O.ROOT))
    z = np.load(out)
    _same(exp, (z["d"], z["t"], z["m"]))


def test_b200_pixconv_takes_rgb565_which_the_reference_filter_drops():
    """MS_RGB565 is a member of MSPixFmt and libswscale reads it (msvideo.c:610-611), but the reference's MSPixConv cannot
    wrap such a frame (ms_picture_init_from_mblk_with_size has no case for it, msvideo.c:120-156) and drops it. The plugin's
    MSPixConv converts it; expected = the oracle's RGB565 reader, itself pinned against the live libswscale."""
    w, h = 64, 48
    frames = [V.synth_frame(V.MS_RGB565, w, h, t) for t in range(3)]
    ref = _expected(frames, V.MS_RGB565, w, h)
    assert len(ref[1]) == 0
    data, tri, dims = V.run_pixconv_sizeconv(_plugin_graph(), frames, V.MS_RGB565, w, h, want_b200=True)
    assert len(tri) == 3 and (dims == (w, h)).all()
    L = O.oracle()
    o = L.orc_scaler_new(w, h, 8, w, h, 0)
    for k, fr in enumerate(frames):
        exp = np.zeros(w * h * 3 // 2, np.uint8)
        L.orc_scaler_process(o, O.ptr(fr), O.ptr(exp))
        assert np.array_equal(data[k * exp.size:(k + 1) * exp.size], exp), k
    L.orc_scaler_free(o)
