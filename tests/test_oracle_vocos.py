"""Oracle Vocos restatement: iSTFT head against torch.istft-free closed forms, shapes of the reference's tests.  CPU only."""
import numpy as np
import torch

from oracle import vocos as ov


def test_output_length_matches_reference_shape_tests():
    # Tests/MLXAudioCodecsTests.swift:419-431,473-479: dim 512, 8 layers, n_fft 1024, hop 256 -> (L-1)*hop samples
    cfg = ov.VocosConfig(dim=64, intermediate_dim=128, num_layers=1)
    W = ov.init_weights(cfg, 0)
    y = ov.decode(cfg, W, np.zeros((1, 10, 100), np.float32))
    assert y.shape == (1, 9 * 256)


def test_istft_head_is_window_sum_normalised_ola():
    # a spectrum whose irfft is the constant 1 frame: every output sample = sum(w)/sum(w) = 1 (window-SUM, not squared)
    cfg = ov.VocosConfig(dim=4, n_fft=16, hop_length=4)
    w = {"head.out.weight": np.zeros((18, 4), np.float32), "head.out.bias": np.zeros(18, np.float32)}
    w["head.out.bias"][0] = np.log(16.0)            # DC magnitude 16, phase 0 -> irfft = 1 everywhere
    w["head.out.bias"][1:9] = -50.0                  # other magnitudes ~ 0
    x = torch.zeros(1, 6, 4, dtype=torch.float64)
    y = ov.istft_head(cfg, w, x)
    assert y.shape == (1, 20) and np.abs(y - 1.0).max() < 1e-6


def test_hann_is_symmetric():
    h = ov.hann_symmetric(1024).numpy()
    assert h[0] == 0 and abs(h[-1]) < 1e-12 and abs(h[511] - h[512]) < 1e-12


def test_oracle_reproduces_committed_golden():
    """tests/golden/codecs.npz (tests/golden/make_golden.py --only codecs): the fixture the GPU tests also compare with."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "codecs.npz")
    cfg = ov.VocosConfig(num_layers=2)
    W = ov.init_weights(cfg, 7)
    f = np.random.default_rng(1).standard_normal((2, 37, cfg.input_channels)).astype(np.float32)
    y = ov.decode(cfg, W, f)
    assert tuple(g["vocos_shape"]) == y.shape and np.abs(y[:, :64] - g["vocos_first"]).max() < 1e-6
    yy = np.asarray(y, dtype=np.float64).reshape(-1)
    assert np.allclose([yy.mean(), np.abs(yy).mean(), yy.min(), yy.max()], g["vocos_stats"], rtol=1e-6, atol=1e-9)
