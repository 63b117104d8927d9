"""CUDA Vocos decode (through the C ABI) vs the oracle (1e-3 relative to the peak), shapes of the reference's own tests,
batched == serial."""
import numpy as np
import pytest

from conftest import max_rel_to_peak, rel_err
from oracle import vocos as ov

pytestmark = pytest.mark.gpu
TOL = 1e-3


def make(b2a, cfg, W):
    return b2a.Vocos(cfg.input_channels, cfg.dim, cfg.intermediate_dim, cfg.num_layers, cfg.n_fft, cfg.hop_length,
                     cfg.input_kernel_size, cfg.dw_kernel_size, weights=W)


@pytest.mark.parametrize("cfg", [ov.VocosConfig(dim=128, intermediate_dim=256, num_layers=2),
                                 ov.VocosConfig(),                                               # reference test geometry
                                 ov.VocosConfig(input_channels=512, dim=768, intermediate_dim=2304, num_layers=3, n_fft=2048,
                                                hop_length=512, input_kernel_size=1, dw_kernel_size=3)])   # Soprano's
def test_decode_vs_oracle(b2a, cfg):
    W = ov.init_weights(cfg, 7)
    m = make(b2a, cfg, W)
    f = np.random.default_rng(1).standard_normal((2, 37, cfg.input_channels)).astype(np.float32)
    y = m.decode(f)
    ref = ov.decode(cfg, W, f)
    assert y.shape == ref.shape == (2, 36 * cfg.hop_length)
    assert max_rel_to_peak(y, ref) < TOL, max_rel_to_peak(y, ref)
    assert rel_err(y, ref) < TOL
    # channel-first input is transposed like VocosBackbone does; batched == serial
    assert np.array_equal(m.decode(f.transpose(0, 2, 1).copy()), y) or cfg.input_channels == 37
    assert np.abs(m.decode(f[1:2]) - y[1:2]).max() < 1e-6


def test_shapes_and_errors(b2a):
    cfg = ov.VocosConfig(dim=128, intermediate_dim=256, num_layers=1)
    W = ov.init_weights(cfg, 3)
    m = make(b2a, cfg, W)
    assert m.decode(np.zeros((1, 2, 100), np.float32)).shape == (1, 256)          # (L-1)*hop
    assert m.decode(np.zeros((3, 130, 100), np.float32)).shape == (3, 129 * 256)  # crosses a 64-token tile boundary
    with pytest.raises(b2a.AudioGenerationError) as e:
        m.decode(np.zeros((1, 1, 100), np.float32))
    assert e.value.case == "audioDecodingFailed"
    with pytest.raises(b2a.AudioGenerationError):
        b2a.Vocos(100, 100, 256, 1, 1024, 256, weights=W)                           # dim not a multiple of 64
    W2 = dict(W); W2.pop("head.out.bias")
    with pytest.raises(b2a.AudioGenerationError) as e:
        make(b2a, cfg, W2)
    assert e.value.case == "modelNotInitialized"


def test_decode_vs_committed_golden(b2a):
    """tests/golden/codecs.npz: first 64 samples per row and the (mean, |mean|, min, max) of the reference-geometry decode."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "codecs.npz")
    cfg = ov.VocosConfig(num_layers=2)
    W = ov.init_weights(cfg, 7)
    f = np.random.default_rng(1).standard_normal((2, 37, cfg.input_channels)).astype(np.float32)
    y = make(b2a, cfg, W).decode(f)
    peak = max(abs(g["vocos_stats"][2]), abs(g["vocos_stats"][3]))
    assert y.shape == tuple(g["vocos_shape"]) and np.abs(y[:, :64] - g["vocos_first"]).max() < TOL * peak
    yy = y.astype(np.float64).reshape(-1)
    assert np.abs(np.array([yy.mean(), np.abs(yy).mean(), yy.min(), yy.max()]) - g["vocos_stats"]).max() < TOL * peak


def test_adalayernorm_model_vs_oracle(b2a):
    """AdaLayerNorm (Vocos.swift:17-47, VocosBackbone.swift:44-79,178-193): scale / shift are Linears of the bandwidth conditioning row;
    different rows of a batch may carry different conditioning; no conditioning -> error where the reference fatalErrors."""
    cfg = ov.VocosConfig(dim=128, intermediate_dim=256, num_layers=2, adanorm_num_embeddings=4)
    W = ov.init_weights(cfg, 11)
    assert "backbone.norm.scale.weight" in W and "backbone.norm.weight" not in W and "backbone.final_layer_norm.weight" in W
    m = b2a.Vocos(cfg.input_channels, cfg.dim, cfg.intermediate_dim, cfg.num_layers, cfg.n_fft, cfg.hop_length, cfg.input_kernel_size,
                  cfg.dw_kernel_size, adanorm_num_embeddings=4, weights=W)
    f = np.random.default_rng(2).standard_normal((3, 70, cfg.input_channels)).astype(np.float32)
    cond = np.eye(4, dtype=np.float32)[[2, 0, 3]]                          # one-hot bandwidth ids, one per utterance
    y, ref = m.decode(f, bandwidth_id=cond), ov.decode(cfg, W, f, cond)
    assert y.shape == ref.shape and max_rel_to_peak(y, ref) < TOL, max_rel_to_peak(y, ref)
    soft = np.random.default_rng(3).random((3, 4)).astype(np.float32)      # any conditioning row, not only one-hot
    assert max_rel_to_peak(m.decode(f, bandwidth_id=soft), ov.decode(cfg, W, f, soft)) < TOL
    assert max_rel_to_peak(m.decode(f, bandwidth_id=cond[::-1].copy()), y) > 1e-2      # the conditioning matters
    with pytest.raises(b2a.AudioGenerationError) as e:
        m.decode(f)
    assert e.value.case == "invalidInput"
    plain = make(b2a, ov.VocosConfig(dim=128, intermediate_dim=256, num_layers=1), ov.init_weights(ov.VocosConfig(dim=128, intermediate_dim=256, num_layers=1), 3))
    with pytest.raises(b2a.AudioGenerationError):
        plain.decode(f[:1], bandwidth_id=cond[:1])                         # a LayerNorm model takes no conditioning
