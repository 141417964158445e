"""CPU-side check of the drop-in boundary: the UNMODIFIED reference plugin loader (ms_factory_load_plugins,
src/base/msfactory.c:531-759) dlopens libmsb200filters.so, calls libmsb200filters_init(), and from then on
ms_factory_create_filter_from_name() hands out the B200 descs for the built-in names. No compute here."""
import pytest

import _oracle as O
from _oracle import RefGraph

NAMES = ["MSAudioMixer", "MSVolume", "MSChannelAdapter", "MSEqualizer", "MSResample", "MSSpeexEC",
         "MSAlawEnc", "MSAlawDec", "MSUlawEnc", "MSUlawDec", "MSAudioFlowControl", "MSGenericPLC"]


def _need_plugin():
    if not (O.PLUGIN_DIR / "libmsb200filters.so").exists():
        pytest.skip("plugin not built (needs the mediastreamer2 headers)")


def test_plugin_overrides_builtin_filters_by_name():
    _need_plugin()
    g = RefGraph(plugins_dir=str(O.PLUGIN_DIR))
    for name in NAMES:
        f = g.new(name)
        assert g.text(f).startswith("B200:"), (name, g.text(f))
    g.close()


def test_reference_factory_without_plugin_keeps_builtin_filters():
    g = RefGraph()
    for name in ["MSAudioMixer", "MSVolume", "MSChannelAdapter", "MSEqualizer"]:
        assert not g.text(g.new(name)).startswith("B200:")
    g.close()


def test_plugin_method_tables_accept_the_reference_method_ids():
    """every method id the reference's filters export is accepted (returns 0) by the replacement, before any GPU use"""
    import ctypes as C

    from _oracle import EqualizerGain, MixerCtl

    _need_plugin()
    g = RefGraph(plugins_dir=str(O.PLUGIN_DIR))
    mix = g.new("MSAudioMixer")
    assert g.call_int(mix, "MS_FILTER_SET_SAMPLE_RATE", 48000) == 0
    assert g.call_int(mix, "MS_FILTER_SET_NCHANNELS", 1) == 0
    assert g.call_int(mix, "MS_AUDIO_MIXER_ENABLE_CONFERENCE_MODE", 1) == 0
    ctl = MixerCtl(pin=3)
    ctl.param.gain = 0.5
    assert g.call(mix, "MS_AUDIO_MIXER_SET_INPUT_GAIN", ctl) == 0
    bad = MixerCtl(pin=77)
    assert g.call(mix, "MS_AUDIO_MIXER_SET_INPUT_GAIN", bad) == -1  # same error convention as audiomixer.c:375-378
    rate = C.c_int(0)
    assert g.call(mix, "MS_FILTER_GET_SAMPLE_RATE", rate) == 0 and rate.value == 48000
    vol = g.new("MSVolume")
    assert g.call_float(vol, "MS_VOLUME_SET_GAIN", 0.8) == 0
    got = C.c_float(0)
    assert g.call(vol, "MS_VOLUME_GET_GAIN", got) == 0 and abs(got.value - 0.8) < 1e-7
    eq = g.new("MSEqualizer")
    assert g.call_int(eq, "MS_FILTER_SET_SAMPLE_RATE", 16000) == 0
    assert g.call(eq, "MS_EQUALIZER_SET_GAIN", EqualizerGain(1000, 2.0, 200)) == 0
    n = C.c_int(0)
    assert g.call(eq, "MS_EQUALIZER_GET_NUM_FREQUENCIES", n) == 0 and n.value == 128
    rs = g.new("MSResample")
    assert g.call_int(rs, "MS_FILTER_SET_SAMPLE_RATE", 8000) == 0
    assert g.call_int(rs, "MS_FILTER_SET_OUTPUT_SAMPLE_RATE", 48000) == 0
    ec = g.new("MSSpeexEC")
    assert g.call_int(ec, "MS_FILTER_SET_SAMPLE_RATE", 48000) == 0
    assert g.call_int(ec, "MS_ECHO_CANCELLER_SET_TAIL_LENGTH", 250) == 0
    assert g.call_int(ec, "MS_ECHO_CANCELLER_SET_DELAY", 0) == 0
    # G.711 codecs: the same answers as alaw.c:140-160 / ulaw.c, incl. the first-match quirk of the "ptime:" attribute
    for nm in ("MSAlawEnc", "MSUlawEnc"):
        enc = g.new(nm)
        v = C.c_int(-1)
        assert g.call(enc, "MS_FILTER_GET_SAMPLE_RATE", v) == 0 and v.value == 8000
        assert g.call(enc, "MS_FILTER_GET_NCHANNELS", v) == 0 and v.value == 1
        assert g.call(enc, "MS_AUDIO_ENCODER_GET_PTIME", v) == 0 and v.value == 0
        assert g.call_ptr(enc, "MS_FILTER_ADD_FMTP", C.c_char_p(b"maxptime=60; ptime=80")) == 0
        assert g.call(enc, "MS_AUDIO_ENCODER_GET_PTIME", v) == 0 and v.value == 60
        assert g.call_ptr(enc, "MS_FILTER_ADD_ATTR", C.c_char_p(b"ptime:30")) == 0
        assert g.call(enc, "MS_AUDIO_ENCODER_GET_PTIME", v) == 0 and v.value == 30
        assert g.call_ptr(enc, "MS_FILTER_ADD_ATTR", C.c_char_p(b"ptime:100")) == 0
        assert g.call(enc, "MS_AUDIO_ENCODER_GET_PTIME", v) == 0 and v.value == 10  # strstr("ptime:10") hits first
    for nm in ("MSAlawDec", "MSUlawDec"):
        dec = g.new(nm)
        v = C.c_int(-1)
        assert g.call(dec, "MS_DECODER_HAVE_PLC", v) == 0 and v.value == 0
        assert g.call(dec, "MS_FILTER_GET_SAMPLE_RATE", v) == 0 and v.value == 8000
    plc = g.new("MSGenericPLC")  # msgenericplc.c:185-189
    v = C.c_int(-1)
    assert g.call_int(plc, "MS_FILTER_SET_SAMPLE_RATE", 16000) == 0
    assert g.call(plc, "MS_FILTER_GET_SAMPLE_RATE", v) == 0 and v.value == 16000
    assert g.call_int(plc, "MS_FILTER_SET_NCHANNELS", 1) == 0
    assert g.call(plc, "MS_GENERIC_PLC_SET_CN", (C.c_uint8 * 36)()) == 0
    g.close()


def test_g711_reference_filters_answer_the_same():
    """the unmodified reference encoders give the answers asserted above (the quirk included)"""
    import ctypes as C

    g = RefGraph()
    for nm in ("MSAlawEnc", "MSUlawEnc"):
        enc = g.new(nm)
        v = C.c_int(-1)
        assert g.call_ptr(enc, "MS_FILTER_ADD_FMTP", C.c_char_p(b"maxptime=60; ptime=80")) == 0
        assert g.call(enc, "MS_AUDIO_ENCODER_GET_PTIME", v) == 0 and v.value == 60
        assert g.call_ptr(enc, "MS_FILTER_ADD_ATTR", C.c_char_p(b"ptime:100")) == 0
        assert g.call(enc, "MS_AUDIO_ENCODER_GET_PTIME", v) == 0 and v.value == 10
    g.close()
