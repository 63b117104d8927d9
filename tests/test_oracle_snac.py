"""Oracle SNAC restatement vs independent formulations (explicit scatter loop for the transposed conv,
torch weight_norm, float64 code search) and the committed goldens.  CPU only."""
import numpy as np
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err
from oracle import snac


def test_conv_transpose_scatter_semantics():
    # y[t*s + k - pad, o] += x[t, i] * w[o, k, i]  (SURVEY 8c trap 7; EncodecLayers.swift:422-438 spells it out)
    rng = np.random.default_rng(0)
    cin, cout, s, T = 3, 2, 4, 5
    k, pad = 2 * s, 2
    w = {"p.weight_v": rng.standard_normal((cin, k, cout)).astype(np.float32),
         "p.weight_g": rng.uniform(0.5, 1.5, (cin, 1, 1)).astype(np.float32),
         "p.bias": rng.standard_normal(cout).astype(np.float32)}
    x = rng.standard_normal((1, cin, T))
    y = snac.wn_conv_transpose1d(w, "p", torch.as_tensor(x), stride=s, padding=pad).numpy()[0]
    v = w["p.weight_v"].astype(np.float64)
    wn = w["p.weight_g"].astype(np.float64) * v / np.sqrt((v ** 2).sum(axis=(1, 2), keepdims=True))   # [in,k,out]
    ref = np.zeros((cout, T * s))
    for t in range(T):
        for kk in range(k):
            to = t * s + kk - pad
            if 0 <= to < T * s:
                for o in range(cout):
                    ref[o, to] += (x[0, :, t] * wn[:, kk, o]).sum()
    ref += w["p.bias"][:, None]
    assert y.shape == ref.shape and rel_err(y, ref) < 1e-12


def test_weight_norm_matches_torch_parametrization():
    rng = np.random.default_rng(1)
    v = rng.standard_normal((6, 7, 4)).astype(np.float32)          # MLX [out, k, in]
    g = rng.uniform(0.5, 1.5, (6, 1, 1)).astype(np.float32)
    ours = snac.wn_conv_weight({"c.weight_v": v, "c.weight_g": g}, "c")
    conv = torch.nn.utils.parametrizations.weight_norm(torch.nn.Conv1d(4, 6, 7).double())
    with torch.no_grad():
        conv.parametrizations.weight.original1.copy_(torch.as_tensor(v).permute(0, 2, 1))
        conv.parametrizations.weight.original0.copy_(torch.as_tensor(g))
    assert rel_err(ours.numpy(), conv.weight.detach().numpy()) < 1e-6


def test_from_codes_repeat_interleave_and_shapes():
    cfg = snac.SNACConfig()
    W = snac.init_weights(cfg, 1234)
    codes = snac.synth_codes(cfg, 2, 8, seed=2)
    assert [c.shape for c in codes] == [(2, 2), (2, 4), (2, 8)]
    z = snac.from_codes(cfg, W, codes).numpy()
    assert z.shape == (2, 768, 8)
    # level 0 contributes a value constant over each group of 4 latent steps
    z0 = snac.from_codes(cfg, W, [codes[0], np.zeros_like(codes[1]), np.zeros_like(codes[2])]).numpy()
    zb = snac.from_codes(cfg, W, [np.zeros_like(codes[0]), np.zeros_like(codes[1]), np.zeros_like(codes[2])]).numpy()
    d = z0 - zb
    assert np.abs(d[:, :, 0:4] - d[:, :, 0:1]).max() < 1e-12 and np.abs(d[:, :, 4:8] - d[:, :, 4:5]).max() < 1e-12


def test_decode_shape_and_batch_equals_serial():
    cfg = snac.SNACConfig()
    W = snac.init_weights(cfg, 1234)
    codes = snac.synth_codes(cfg, 2, 8, seed=5)
    y = snac.decode(cfg, W, codes)
    assert y.shape == (2, 1, 8 * 512) and np.abs(y).max() <= 1.0
    y1 = snac.decode(cfg, W, [c[1:2] for c in codes])
    assert np.abs(y[1:2] - y1).max() < 1e-12          # Parakeet-style batched == serial


def test_fp32_code_search_agrees_with_float64():
    cfg = snac.SNACConfig()
    W = snac.init_weights(cfg, 1234)
    z = (np.random.default_rng(7).standard_normal((2, cfg.latent, 16)) * 0.5).astype(np.float32)
    _, c32 = snac.quantize(cfg, W, z, fp32_search=True)
    _, c64 = snac.quantize(cfg, W, z, fp32_search=False)
    assert all(np.array_equal(a, b) for a, b in zip(c32, c64))
    assert all(a.min() >= 0 and a.max() < 4096 for a in c32)


def test_goldens():
    g = np.load(GOLDEN / "snac.npz")
    cfg = snac.SNACConfig()
    W = snac.init_weights(cfg, 1234)
    codes = snac.synth_codes(cfg, 2, 16, seed=2)
    rng = np.random.default_rng(7)
    noise = [rng.standard_normal(s).astype(np.float32) for s in snac.noise_shapes(cfg, 2, 16)]
    y = snac.decode(cfg, W, codes, noise)
    assert np.abs(y - g["wave"]).max() < 1e-6
    z = (rng.standard_normal((2, cfg.latent, 16)) * 0.5).astype(np.float32)
    _, qc = snac.quantize(cfg, W, z)
    assert np.array_equal(qc[0], g["q_codes0"]) and np.array_equal(qc[2], g["q_codes2"])


def test_snake_matches_transformers_dac_snake1d():
    """snake (Layers.swift:44-50) against the Snake1d of transformers' DAC (the codec SNAC's layers derive from)."""
    from transformers.models.dac.modeling_dac import Snake1d
    m = Snake1d(6).double()
    with torch.no_grad():
        m.alpha.copy_(torch.rand(1, 6, 1) * 2 + 0.1)
    x = torch.randn(2, 6, 33, dtype=torch.float64)
    assert (snac.snake(x, m.alpha.detach()) - m(x)).abs().max() < 1e-12


def test_code_search_matches_transformers_dac_vector_quantize():
    """VectorQuantize.decodeLatents (VQ.swift:96-120: L2-normalise encodings and codebook, argmax of -(|e|^2 - 2 e.c + |c|^2)) against
    DacVectorQuantize.decode_latents of transformers (SNAC's quantiser is DAC's).  Float64 on the HF side; the oracle's ordered-fp32
    search may only differ on near-ties."""
    from transformers import DacConfig
    from transformers.models.dac.modeling_dac import DacVectorQuantize
    rng = np.random.default_rng(4)
    vq = DacVectorQuantize(DacConfig(codebook_size=4096, codebook_dim=8, hidden_size=16)).double()
    cb = rng.standard_normal((4096, 8)).astype(np.float32)
    with torch.no_grad():
        vq.codebook.weight.copy_(torch.from_numpy(cb).double())
    enc = rng.standard_normal((3, 8, 700)).astype(np.float32)                  # [B, D, T]
    _, ref = vq.decode_latents(torch.from_numpy(enc).double())
    ours = snac.nearest_code_fp32(enc.transpose(0, 2, 1).reshape(-1, 8), cb).reshape(3, 700)
    diff = ours != ref.numpy()
    assert diff.mean() < 2e-3
    if diff.any():                                                             # any disagreement must be a near-tie in cosine similarity
        e = enc.transpose(0, 2, 1).reshape(-1, 8).astype(np.float64)
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        c = cb.astype(np.float64) / np.linalg.norm(cb.astype(np.float64), axis=1, keepdims=True)
        sim = e @ c.T
        rows = np.flatnonzero(diff.reshape(-1))
        assert np.abs(sim[rows, ours.reshape(-1)[rows]] - sim[rows, ref.numpy().reshape(-1)[rows]]).max() < 1e-6


def _load_dac_decoder(cfg, w):
    """transformers' DacDecoder carrying the oracle's (weight-norm folded) SNAC weights; only valid for depthwise=False, noise=False."""
    from transformers import DacConfig
    from transformers.models.dac.modeling_dac import DacDecoder
    dec = DacDecoder(DacConfig(hidden_size=cfg.latent, decoder_hidden_size=cfg.decoder_dim, upsampling_ratios=list(cfg.decoder_rates))).double()
    p = "decoder.model.layers"
    with torch.no_grad():
        dec.conv1.weight.copy_(snac.wn_conv_weight(w, f"{p}.0")); dec.conv1.bias.copy_(snac._t(w[f"{p}.0.bias"]))
        li = 1
        for blk in dec.block:
            b = f"{p}.{li}.block.layers"
            blk.snake1.alpha.copy_(snac._t(w[f"{b}.0.alpha"]))
            blk.conv_t1.weight.copy_(snac.wn_convT_weight(w, f"{b}.1")); blk.conv_t1.bias.copy_(snac._t(w[f"{b}.1.bias"]))
            for j, ru in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
                r = f"{b}.{2 + j}.block.layers"
                ru.snake1.alpha.copy_(snac._t(w[f"{r}.0.alpha"])); ru.snake2.alpha.copy_(snac._t(w[f"{r}.2.alpha"]))
                ru.conv1.weight.copy_(snac.wn_conv_weight(w, f"{r}.1")); ru.conv1.bias.copy_(snac._t(w[f"{r}.1.bias"]))
                ru.conv2.weight.copy_(snac.wn_conv_weight(w, f"{r}.3")); ru.conv2.bias.copy_(snac._t(w[f"{r}.3.bias"]))
            li += 1
        dec.snake1.alpha.copy_(snac._t(w[f"{p}.{li}.alpha"]))
        dec.conv2.weight.copy_(snac.wn_conv_weight(w, f"{p}.{li + 1}")); dec.conv2.bias.copy_(snac._t(w[f"{p}.{li + 1}.bias"]))
    return dec


def test_dense_decoder_matches_transformers_dac_decoder():
    """With depthwise = false and noise = false the SNAC decoder (Layers.swift:364-421) IS the DAC decoder: k7 conv -> 4 x (Snake ->
    transposed conv k = 2s, padding ceil(s/2) -> three dilated residual units) -> Snake -> k7 conv -> tanh.  The oracle in that mode,
    weight norm included, against transformers' DacDecoder with the same weights: an independent check of the layer order, every
    padding, the transposed-convolution semantics (SURVEY 8c trap 7) and Snake."""
    cfg = snac.SNACConfig(encoder_dim=4, encoder_rates=(2, 2), decoder_dim=64, decoder_rates=(8, 8, 4, 2), noise=False, depthwise=False)
    w = snac.init_weights(cfg, 7)
    z = torch.randn(2, cfg.latent, 9, dtype=torch.float64)
    ours = snac.decoder(cfg, w, z, None)
    with torch.no_grad():
        ref = _load_dac_decoder(cfg, w)(z)
    assert ours.shape == ref.shape == (2, 1, 9 * 512) and (ours - ref).abs().max() < 1e-10


def test_depthwise_residual_unit_is_the_dense_unit_with_a_diagonal_kernel():
    """SNAC's depthwise ResidualUnit (Layers.swift:202-232, groups = dim) against transformers' dense DacResidualUnit whose k7 kernel is
    the depthwise one placed on the channel diagonal."""
    from transformers.models.dac.modeling_dac import DacResidualUnit
    cfg = snac.SNACConfig(encoder_dim=4, encoder_rates=(2, 2), decoder_dim=32, decoder_rates=(2,), noise=False, depthwise=True)
    w = snac.init_weights(cfg, 3)
    C, dil = 16, 3
    r = "decoder.model.layers.2.block.layers.3"                 # second residual unit of the only block (layers 0, 1 = dw + pw stem)
    ru = DacResidualUnit(C, dilation=dil).double()
    with torch.no_grad():
        dw = snac.wn_conv_weight(w, r + ".block.layers.1")       # [C, 1, 7]
        ru.conv1.weight.zero_()
        for c in range(C):
            ru.conv1.weight[c, c] = dw[c, 0]
        ru.conv1.bias.copy_(snac._t(w[r + ".block.layers.1.bias"]))
        ru.conv2.weight.copy_(snac.wn_conv_weight(w, r + ".block.layers.3")); ru.conv2.bias.copy_(snac._t(w[r + ".block.layers.3.bias"]))
        ru.snake1.alpha.copy_(snac._t(w[r + ".block.layers.0.alpha"])); ru.snake2.alpha.copy_(snac._t(w[r + ".block.layers.2.alpha"]))
    x = torch.randn(2, C, 40, dtype=torch.float64)
    with torch.no_grad():
        ref = ru(x)
    assert (snac.residual_unit(w, r, x, dil, C) - ref).abs().max() < 1e-12
