"""CUDA Encodec decode (through the C ABI) vs the oracle (1e-3 relative to the peak; the oracle itself is pinned against
transformers' EncodecModel in test_oracle_encodec.py), chunked decode with linearOverlapAdd, batched == serial, errors."""
import numpy as np
import pytest

from conftest import max_rel_to_peak, rel_err
from oracle import encodec as oe

pytestmark = pytest.mark.gpu
TOL = 1e-3


def make(b2a, cfg, W):
    return b2a.Encodec(b2a.EncodecConfig(**cfg.__dict__), weights=W)


def test_decode_24khz_geometry_vs_oracle(b2a):
    cfg = oe.EncodecConfig()                                      # the 24 kHz model: 32 filters, 2 x LSTM(512), ratios 8,5,4,2
    W = oe.init_weights(cfg, 7, n_codebooks=8)
    m = make(b2a, cfg, W)
    assert m.num_codebooks == 8
    codes = np.random.default_rng(1).integers(0, 1024, size=(1, 3, 8, 41))
    y = m.decode(codes)
    ref = oe.decode(cfg, W, codes)
    assert y.shape == ref.shape == (3, 41 * 320, 1)
    assert max_rel_to_peak(y, ref) < TOL, max_rel_to_peak(y, ref)
    assert rel_err(y, ref) < TOL
    # fewer codebooks (lower bandwidth), scales, batched == serial, decodeAudio
    y2 = m.decode(codes[:, :, :2, :], [np.array([0.5, 2.0, 1.0])])
    ref2 = oe.decode(cfg, W, codes[:, :, :2, :], [np.array([0.5, 2.0, 1.0])])
    assert max_rel_to_peak(y2, ref2) < TOL
    assert np.abs(m.decode(codes[:, 1:2]) - y[1:2]).max() < 1e-5 * np.abs(y).max()
    assert np.array_equal(m.decode_audio(b2a.EncodecEncodedAudio(codes, [None])), y)
    # padding-mask truncation (Encodec.swift:397-399)
    assert m.decode(codes, padding_mask=np.ones((3, 1000), bool)).shape == (3, 1000, 1)


def test_more_than_eight_rows_and_single_frame(b2a):
    cfg = oe.EncodecConfig(num_filters=8, hidden_size=16, codebook_dim=16, codebook_size=64)
    W = oe.init_weights(cfg, 3, n_codebooks=4)
    m = make(b2a, cfg, W)
    codes = np.random.default_rng(2).integers(0, 64, size=(1, 11, 4, 1))           # T = 1: reflect padding clamps (pad > length)
    y = m.decode(codes)
    ref = oe.decode(cfg, W, codes)
    assert y.shape == (11, 320, 1) and max_rel_to_peak(y, ref) < TOL
    codes = np.random.default_rng(3).integers(0, 64, size=(1, 9, 4, 70))           # crosses the 64-token tile, 2 LSTM launches
    assert max_rel_to_peak(m.decode(codes), oe.decode(cfg, W, codes)) < TOL


@pytest.mark.parametrize("kw", [dict(use_causal_conv=False), dict(pad_mode="constant"), dict(use_conv_shortcut=False),
                                dict(num_lstm_layers=1), dict(num_lstm_layers=0), dict(audio_channels=2),
                                dict(trim_right_ratio=0.5), dict(upsampling_ratios=[3, 2], compress=1)])
def test_config_variants(b2a, kw):
    cfg = oe.EncodecConfig(num_filters=8, hidden_size=16, codebook_dim=16, codebook_size=64, **kw)
    W = oe.init_weights(cfg, 5, n_codebooks=3)
    m = make(b2a, cfg, W)
    codes = np.random.default_rng(4).integers(0, 64, size=(1, 2, 3, 19))
    y, ref = m.decode(codes), oe.decode(cfg, W, codes)
    assert y.shape == ref.shape
    assert max_rel_to_peak(y, ref) < TOL, max_rel_to_peak(y, ref)


def test_chunked_decode_overlap_add(b2a):
    cfg = oe.EncodecConfig(chunk_length_s=0.04, overlap=0.5, num_filters=8, hidden_size=16, codebook_dim=16, codebook_size=64)
    W = oe.init_weights(cfg, 1, n_codebooks=2)
    m = make(b2a, cfg, W)
    assert m.chunk_length == 960 and m.chunk_stride == 480
    codes = np.random.default_rng(1).integers(0, 64, size=(3, 2, 2, 3))
    scales = [None, np.array([2.0, 0.5]), None]
    y, ref = m.decode(codes, scales), oe.decode(cfg, W, codes, scales)
    assert y.shape == ref.shape == (2, 2 * 480 + 960, 1)
    assert max_rel_to_peak(y, ref) < TOL


def test_errors(b2a):
    cfg = oe.EncodecConfig(num_filters=8, hidden_size=16, codebook_dim=16, codebook_size=64)
    W = oe.init_weights(cfg, 3, n_codebooks=2)
    m = make(b2a, cfg, W)
    E = b2a.AudioGenerationError
    with pytest.raises(E) as e:                                    # "Expected one frame" (Encodec.swift:375-377)
        m.decode(np.zeros((2, 1, 2, 4), np.int32))
    assert e.value.case == "audioDecodingFailed"
    with pytest.raises(E) as e:                                    # more codebooks than the checkpoint holds
        m.decode(np.zeros((1, 1, 3, 4), np.int32))
    assert e.value.case == "invalidInput"
    with pytest.raises(E) as e:
        m.decode(np.zeros((1, 1, 2, 0), np.int32))
    assert e.value.case == "audioDecodingFailed"
    with pytest.raises(E) as e:
        make(b2a, oe.EncodecConfig(norm_type="time_group_norm", num_filters=8, hidden_size=16, codebook_dim=16, codebook_size=64), W)
    assert e.value.case == "invalidInput"
    W2 = dict(W); W2.pop("decoder.layers.1.lstm.1.Wh")
    with pytest.raises(E) as e:
        make(b2a, cfg, W2)
    assert e.value.case == "modelNotInitialized"


def test_decode_vs_committed_golden(b2a):
    """tests/golden/codecs.npz: first / last 64 samples and the (mean, |mean|, min, max) of the 24 kHz-geometry decode."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "codecs.npz")
    cfg = oe.EncodecConfig()
    W = oe.init_weights(cfg, 7, n_codebooks=8)
    codes = np.random.default_rng(1).integers(0, 1024, size=(1, 3, 8, 41))
    y = make(b2a, cfg, W).decode(codes)
    peak = max(abs(g["encodec_stats"][2]), abs(g["encodec_stats"][3]))
    assert y.shape == tuple(g["encodec_shape"])
    assert np.abs(y[:, :64, 0] - g["encodec_first"]).max() < TOL * peak and np.abs(y[:, -64:, 0] - g["encodec_last"]).max() < TOL * peak
    yy = y.astype(np.float64).reshape(-1)
    assert np.abs(np.array([yy.mean(), np.abs(yy).mean(), yy.min(), yy.max()]) - g["encodec_stats"]).max() < TOL * peak
