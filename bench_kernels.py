"""Per-kernel measurements for every SURVEY.md §8 row that bench.py's two headline blocks do not cover: each bank's
`*_process_dev` entry point on device-resident buffers LARGER than the 126 MB L2 (so the bytes really come from HBM),
timed with CUDA events on the launching stream after warm-up. Reports ms per launch, achieved GB/s on the row's
ALGORITHMIC bytes (SURVEY §8d / DESIGN §5) and the fraction of the measured HBM peak. Imported by bench.py ("kernels"
block); runnable alone:

    python bench_kernels.py
"""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _time(ctx, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(iters):
        fn()
    return ctx.timer_stop_ms() / iters


def _fill(ctx, dev, nbytes, seed, lo=0, hi=256, dtype=np.uint8):
    rng = np.random.default_rng(seed)
    chunk = rng.integers(lo, hi, (4 << 20) // np.dtype(dtype).itemsize).astype(dtype)
    for off in range(0, nbytes, chunk.nbytes):
        n = min(chunk.nbytes, nbytes - off)
        ctx.h2d(dev + off, chunk[: n // chunk.itemsize])


def kernels_bench(ctx, hbm_peak_gbs: float) -> dict:
    from mediastreamer2_b200 import _lib
    from mediastreamer2_b200 import filters as F

    lib = ctx.lib
    P = C.c_void_p
    rows = {}

    def row(name, what, algo_bytes, ms, units, unit_name):
        gbs = algo_bytes / (ms / 1000.0) / 1e9
        rows[name] = {"what": what, "ms_per_launch": ms, "algorithmic_bytes": int(algo_bytes), "achieved_gbs": gbs,
                      "frac_of_hbm_peak": gbs / hbm_peak_gbs, unit_name + "_per_s": units / (ms / 1000.0)}

    # ---------------------------------------------------------------- audio banks, 48 kHz mono 10 ms blocks
    n, w = 262144, 480  # 252 MB of s16 per buffer: twice the L2
    d_a, d_b = ctx.dev_alloc(n * w * 2 * 2), ctx.dev_alloc(n * w * 2 * 2)
    _fill(ctx, d_a, n * w * 2 * 2, 1, -9000, 9000, np.int16)
    # a3 MSAudioMixer: 16-party conference rooms
    rooms, pins = n // 16, 16
    mix = F.AudioMixer(ctx, rooms, pins, w, True)
    d_present = ctx.dev_alloc(rooms * pins)
    ctx.h2d(d_present, np.ones(rooms * pins, np.uint8))
    ms = _time(ctx, lambda: _lib.check(lib.msb200_mixer_process_dev(mix.h, P(d_a), P(d_present), P(d_b))))
    row("mixer_kernel", f"a3 MSAudioMixer conference mode, {rooms} rooms x {pins} pins x {w} samples", rooms * pins * w * 2 * 2, ms, rooms, "room_ticks")
    mix.close()
    ctx.dev_free(d_present)
    # a4 MSVolume light path (in place)
    vol = F.Volume(ctx, n, 48000, w)
    vol.set_gain(-1, 0.8)
    ms = _time(ctx, lambda: _lib.check(lib.msb200_volume_process_dev(vol.h, P(d_a), w, w)))
    row("volume_kernel", f"a4 MSVolume light path, {n} streams x {w} samples in place", n * w * 2 * 2, ms, n, "stream_ticks")
    vol.close()
    # a5 MSEqualizer (512-tap FIR at 48 kHz): compute-bound by design, reported against the same yardstick
    ne = 16384
    eq = F.Equalizer(ctx, ne, 48000, w)
    eq.set_gain(0, 1000.0, 2.0, 300.0)  # every stream runs the full 512-tap FIR whatever its taps are
    ms = _time(ctx, lambda: _lib.check(lib.msb200_equalizer_process_dev(eq.h, P(d_a), w, w)), iters=5)
    row("eq_fir_kernel", f"a5 MSEqualizer 512-tap FIR, {ne} streams x {w} samples in place (compute-bound: 2*512 flop/sample)", ne * w * 2 * 2, ms, ne, "stream_ticks")
    rows["eq_fir_kernel"]["tflops"] = ne * w * 2 * 512 / (ms / 1000.0) / 1e12
    eq.close()
    # a6 MSChannelAdapter mono -> stereo
    ms = _time(ctx, lambda: _lib.check(lib.msb200_chanadapt_process_dev(ctx.h, 0, n, w, P(d_a), None, P(d_b))))
    row("chanadapt_kernel", f"a6 MSChannelAdapter mono->stereo, {n} streams x {w} frames", n * w * 2 * 3, ms, n, "stream_ticks")
    # a1 MSResample 16 kHz -> 48 kHz
    rs = F.Resample(ctx, n, 16000, 48000, 1, 160)
    got = C.c_int()
    ms = _time(ctx, lambda: _lib.check(lib.msb200_resample_process_dev(rs.h, P(d_a), 160, 160, P(d_b), 481, C.byref(got))))
    row("resample_up_kernel<3>", f"a1 MSResample 16k->48k, {n} streams x 160 frames", n * (160 + 480) * 2, ms, n, "stream_ticks")
    rs.close()
    # a1 general kernel: 48 kHz -> 16 kHz (144-tap down-sampling filter)
    nd = 65536
    rs = F.Resample(ctx, nd, 48000, 16000, 1, 480)
    ms = _time(ctx, lambda: _lib.check(lib.msb200_resample_process_dev(rs.h, P(d_a), 480, 480, P(d_b), 161, C.byref(got))), iters=5)
    row("resample_kernel", f"a1 MSResample 48k->16k (general kernel), {nd} streams x 480 frames", nd * (480 + 160) * 2, ms, nd, "stream_ticks")
    rs.close()
    # f3 MSGenericPLC, 16 kHz: (i) every stream receives its block (history + 5 ms continuity delay), (ii) every stream
    # lost its block right after a received one, i.e. the worst case: one 800-point + one 1600-point transform per stream
    try:
        ns, wn = 16384, 160
        plc = F.GenericPLC(ctx, ns, 16000, wn)
        N = plc.history_samples
        d_mode_pkt, d_mode_lost = ctx.dev_alloc(ns), ctx.dev_alloc(ns)
        ctx.h2d(d_mode_pkt, np.full(ns, 1, np.uint8))
        ctx.h2d(d_mode_lost, np.full(ns, 2, np.uint8))
        pkt = lambda: _lib.check(lib.msb200_plc_process_dev(plc.h, P(d_a), wn, wn, P(d_mode_pkt)))  # noqa: E731
        ms_pkt = _time(ctx, pkt)

        def pkt_then_lost():
            pkt()
            _lib.check(lib.msb200_plc_process_dev(plc.h, P(d_b), wn, wn, P(d_mode_lost)))
        ms_pair = _time(ctx, pkt_then_lost, iters=5)
        row("plc_kernel (received)", f"f3 MSGenericPLC received block, {ns} streams x {wn} samples @16 kHz", ns * (2 * wn + 2 * N + 80) * 2, ms_pkt, ns, "stream_ticks")
        row("plc_kernel (concealed)", f"f3 MSGenericPLC first concealed block (800 + 1600-point float FFTs per stream), {ns} streams",
            ns * (wn + 3 * N + 4 * N + 320) * 2, max(ms_pair - ms_pkt, 1e-6), ns, "concealments")
        plc.close()
        ctx.dev_free(d_mode_pkt)
        ctx.dev_free(d_mode_lost)
    except Exception as e:  # noqa: BLE001 - a side measurement must not take the headline line down with it
        rows["plc_kernel"] = {"error": repr(e)}
    ctx.dev_free(d_a)
    ctx.dev_free(d_b)
    # ---------------------------------------------------------------- video, 1080p, 128 frames (>= 400 MB per buffer)
    nf, sw, sh = 128, 1920, 1080
    src_b = sw * sh * 3  # largest source format here (RGB24)
    d_src, d_dst = ctx.dev_alloc(nf * src_b), ctx.dev_alloc(nf * src_b)
    _fill(ctx, d_src, nf * src_b, 2)
    i420 = sw * sh * 3 // 2
    ms = _time(ctx, lambda: _lib.check(lib.msb200_nv12_to_i420_dev(ctx.h, nf, P(d_src), i420, sw * sh, 0, sw, sh, sw, sw, 1, 0, P(d_dst))))
    row("nv12_fast_kernel", f"a9 NV12->I420, {nf} frames {sw}x{sh}", nf * i420 * 2, ms, nf, "frames")
    ms = _time(ctx, lambda: _lib.check(lib.msb200_nv12_to_i420_dev(ctx.h, nf, P(d_src), i420, sw * sh, 90, sh, sw, sw, sw, 1, 0, P(d_dst))))
    row("nv12_rot_kernel<90>", f"a9 NV12->I420 rotated 90 degrees, {nf} frames {sw}x{sh}", nf * i420 * 2, ms, nf, "frames")
    for name, sf, df, dw, dh, sbytes, dbytes, what in (
            ("scale_plane_strip_kernel", _lib.PIX_YUV420P, _lib.PIX_YUV420P, 1280, 720, i420, 1280 * 720 * 3 // 2, "a8 MSSizeConv I420 1080p -> I420 720p"),
            ("scale_plane_strip_kernel<INTER>", _lib.PIX_NV12, _lib.PIX_YUV420P, 1280, 720, i420, 1280 * 720 * 3 // 2,
             "cfg4 reference-shaped two-step as ONE pass: NV12 1080p -> I420 720p, CbCr plane read in place"),
            ("rgb_to_i420_march_kernel<0>", _lib.PIX_RGB24, _lib.PIX_YUV420P, sw, sh, sw * sh * 3, i420, "a7 MSPixConv RGB24 -> I420 1080p"),
            ("rgb24_to_i420_kernel<1>", _lib.PIX_RGB24_REV, _lib.PIX_YUV420P, sw, sh, sw * sh * 3, i420, "a7 MSPixConv BGR24 -> I420 1080p"),
            ("packed422_to_i420_kernel", _lib.PIX_YUY2, _lib.PIX_YUV420P, sw, sh, sw * sh * 2, i420, "a7 MSPixConv YUY2 -> I420 1080p"),
            ("scale_down_kernel<64,2,rgb>", _lib.PIX_NV12, _lib.PIX_RGB24, 640, 360, i420, 640 * 360 * 3, "a8/a11 NV12 1080p -> RGB24 360p (3:1 down-scale tiles)"),
            ("scale_down_kernel<64,1,planar>", _lib.PIX_NV12, _lib.PIX_YUV420P, 960, 540, i420, 960 * 540 * 3 // 2, "a8 NV12 1080p -> I420 540p (2:1 down-scale tiles)"),
            ("scale_down_kernel<32,2,planar>", _lib.PIX_YUV420P, _lib.PIX_YUV420P, 480, 270, i420, 480 * 270 * 3 // 2, "a8 MSSizeConv I420 1080p -> I420 270p thumbnail (4:1 down-scale tiles)")):
        sc = F.Scaler(ctx, sw, sh, sf, dw, dh, df)
        ms = _time(ctx, lambda: sc.process_dev(nf, d_src, d_dst), iters=5)
        row(name, f"{what}, {nf} frames", nf * (sbytes + dbytes), ms, nf, "frames")
        sc.close()
    # f4: 16-tile mosaic — sixteen 480x270 participants scaled to 320x180 and composed into 1280x720 canvases by the
    # scaler's own launches (set_canvas); 64 canvases per launch = 1024 source frames (199 MB in, 88 MB out)
    try:
        tw, th, cw, ch, tiles, ncanv = 320, 180, 1280, 720, 16, 64
        msw, msh = 480, 270
        sc = F.Scaler(ctx, msw, msh, _lib.PIX_YUV420P, tw, th, _lib.PIX_YUV420P)
        rects = (_lib.Rect * tiles)(*[_lib.Rect((k % 4) * tw, (k // 4) * th, tw, th) for k in range(tiles)])
        _lib.check(lib.msb200_scaler_set_canvas(sc.h, cw, ch, tiles, rects))
        n_src = tiles * ncanv
        sbytes, cbytes = msw * msh * 3 // 2, cw * ch * 3 // 2
        ms = _time(ctx, lambda: sc.process_dev(n_src, d_src, d_dst), iters=5)
        row("scale_plane_strip_kernel (mosaic)", f"f4 16-tile 720p mosaic: 16 x I420 {msw}x{msh} -> 320x180 tiles of one 1280x720 canvas, {ncanv} canvases",
            n_src * sbytes + ncanv * cbytes, ms, ncanv, "canvases")
        sc.close()
        # the same mosaic from sixteen 720p participants (4:1 thumbnails): 16 canvases per launch = 256 source frames (354 MB in)
        msw, msh, ncanv = 1280, 720, 16
        sc = F.Scaler(ctx, msw, msh, _lib.PIX_YUV420P, tw, th, _lib.PIX_YUV420P)
        _lib.check(lib.msb200_scaler_set_canvas(sc.h, cw, ch, tiles, rects))
        n_src, sbytes = tiles * ncanv, msw * msh * 3 // 2
        ms = _time(ctx, lambda: sc.process_dev(n_src, d_src, d_dst), iters=5)
        row("scale_down_kernel (mosaic)", f"f4 16-tile 720p mosaic from 720p participants: 16 x I420 {msw}x{msh} -> 320x180 tiles of one 1280x720 canvas, {ncanv} canvases",
            n_src * sbytes + ncanv * cbytes, ms, ncanv, "canvases")
        sc.close()
        # f4: ms_yuv_buf_copy_with_pix_strides, 1080p I420 -> NV12 (planar to semi-planar), 128 frames
        lay_p = _lib.YuvLayout((C.c_size_t * 3)(0, sw * sh, sw * sh * 5 // 4), (C.c_int32 * 3)(sw, sw // 2, sw // 2), (C.c_int32 * 3)(1, 1, 1), i420)
        lay_s = _lib.YuvLayout((C.c_size_t * 3)(0, sw * sh, sw * sh + 1), (C.c_int32 * 3)(sw, sw, sw), (C.c_int32 * 3)(1, 2, 2), i420)
        roi = _lib.Rect(0, 0, sw, sh)
        ms = _time(ctx, lambda: _lib.check(lib.msb200_yuv_copy_strided_dev(ctx.h, nf, P(d_src), C.byref(lay_p), roi, P(d_dst), C.byref(lay_s), roi)))
        row("yuv_copy_rows+chroma_kernel", f"f4 ms_yuv_buf_copy_with_pix_strides I420 -> NV12, {nf} frames {sw}x{sh}", nf * i420 * 2, ms, nf, "frames")
    except Exception as e:  # noqa: BLE001
        rows["mosaic"] = {"error": repr(e)}
    ctx.dev_free(d_src)
    ctx.dev_free(d_dst)
    return rows


if __name__ == "__main__":
    from mediastreamer2_b200 import filters as F

    c = F.Context(0)
    out = kernels_bench(c, 6455.9)
    for k, v in out.items():
        if "error" in v:
            print(k, v)
            continue
        print(f"{k:32s} {v['ms_per_launch']:9.4f} ms  {v['achieved_gbs']:8.1f} GB/s  frac {v['frac_of_hbm_peak']:.3f}   {v['what']}")
    print(json.dumps(out))
    c.close()
