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


def test_istft_head_overlap_add_matches_torch_istft_up_to_its_normalisation():
    """torch.istft(center=True) = OLA(irfft * w) / OLA(w^2), trimmed by n_fft/2; the reference divides by OLA(w) instead (Vocos.swift:123-160,
    window-SUM normalisation).  So oracle * OLA(w) / OLA(w^2) must equal torch.istft on the same spectrum: an independent check of the
    inverse FFT, the windowing, the overlap-add indexing and the centre trim."""
    cfg = ov.VocosConfig(dim=8, n_fft=64, hop_length=16)
    rng = np.random.default_rng(0)
    W = {"head.out.weight": rng.standard_normal((cfg.n_fft + 2, cfg.dim)) * 0.3, "head.out.bias": rng.standard_normal(cfg.n_fft + 2) * 0.1}
    L = 23
    x = torch.from_numpy(rng.standard_normal((2, L, cfg.dim)))
    y = ov.istft_head(cfg, W, x)
    h = x @ torch.from_numpy(W["head.out.weight"]).T + torch.from_numpy(W["head.out.bias"])
    half = cfg.n_fft // 2 + 1
    mag = torch.clamp(torch.exp(h[..., :half]), max=1e2)
    spec = torch.complex(mag * torch.cos(h[..., half:]), mag * torch.sin(h[..., half:]))
    # irfft ignores the imaginary part of the DC and Nyquist bins; torch.istft insists on a Hermitian-consistent input, so zero them
    spec[..., 0] = torch.complex(spec[..., 0].real, torch.zeros_like(spec[..., 0].real))
    spec[..., -1] = torch.complex(spec[..., -1].real, torch.zeros_like(spec[..., -1].real))
    win = ov.hann_symmetric(cfg.n_fft)
    ref = torch.istft(spec.transpose(1, 2), cfg.n_fft, cfg.hop_length, window=win, center=True, length=(L - 1) * cfg.hop_length).numpy()
    out_len = (L - 1) * cfg.hop_length + cfg.n_fft
    w1, w2 = np.zeros(out_len), np.zeros(out_len)
    for i in range(L):
        w1[i * cfg.hop_length: i * cfg.hop_length + cfg.n_fft] += win.numpy()
        w2[i * cfg.hop_length: i * cfg.hop_length + cfg.n_fft] += win.numpy() ** 2
    a = cfg.n_fft // 2
    ratio = (w1 / np.maximum(w2, 1e-300))[a: a + (L - 1) * cfg.hop_length]
    assert y.shape == ref.shape == (2, (L - 1) * cfg.hop_length)
    assert np.abs(y * ratio - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_convnext_block_matches_transformers_convnext_layer():
    """Vocos's ConvNeXtBlock (VocosBackbone.swift:18-100: depthwise k7 "same" -> LayerNorm 1e-6 -> Linear -> exact GELU -> Linear -> gamma ->
    residual) against transformers' 2-D ConvNextLayer on a height-1 image: with padding 3 only the middle row of its 7x7 depthwise
    kernel meets data, so it is the 1-D block whose taps are that row."""
    from transformers import ConvNextConfig
    from transformers.models.convnext.modeling_convnext import ConvNextLayer
    cfg = ov.VocosConfig(input_channels=12, dim=12, intermediate_dim=48, num_layers=1, input_kernel_size=1)
    W = ov.init_weights(cfg, 5)
    layer = ConvNextLayer(ConvNextConfig(hidden_act="gelu", layer_scale_init_value=1.0), dim=12).double()
    p = "backbone.convnext.0."
    with torch.no_grad():
        layer.dwconv.weight.zero_()
        layer.dwconv.weight[:, 0, 3, :] = torch.as_tensor(np.asarray(W[p + "dwconv.weight"]), dtype=torch.float64)[:, :, 0]      # MLX [C, k, 1]
        layer.dwconv.bias.copy_(torch.as_tensor(np.asarray(W[p + "dwconv.bias"])))
        layer.layernorm.weight.copy_(torch.as_tensor(np.asarray(W[p + "norm.weight"]))); layer.layernorm.bias.copy_(torch.as_tensor(np.asarray(W[p + "norm.bias"])))
        layer.pwconv1.weight.copy_(torch.as_tensor(np.asarray(W[p + "pwconv1.weight"]))); layer.pwconv1.bias.copy_(torch.as_tensor(np.asarray(W[p + "pwconv1.bias"])))
        layer.pwconv2.weight.copy_(torch.as_tensor(np.asarray(W[p + "pwconv2.weight"]))); layer.pwconv2.bias.copy_(torch.as_tensor(np.asarray(W[p + "pwconv2.bias"])))
        layer.layer_scale_parameter.copy_(torch.as_tensor(np.asarray(W[p + "gamma"])))
    h = torch.randn(2, 31, 12, dtype=torch.float64)                       # [B, L, C], the state between blocks
    with torch.no_grad():
        ref = layer(h.transpose(1, 2)[:, :, None, :])[:, :, 0, :].transpose(1, 2)
    assert (ov.convnext_layer(cfg, W, 0, h) - ref).abs().max() < 1e-12
