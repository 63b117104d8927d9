"""Oracle Encodec restatement pinned against an independent implementation available offline (transformers' EncodecModel,
random init, float64) + the closed forms of the reference's own helpers.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import encodec as oe


def hf_to_mlx(m) -> dict:
    """transformers EncodecModel -> the reference's keys / MLX layouts (what mlx-community/encodec-24khz-float32 ships):
    Conv1d [out,in,k] -> [out,k,in]; ConvTranspose1d [in,out,k] -> [out,k,in]; LSTM weight_ih/hh -> Wx/Wh, bias_ih+bias_hh."""
    W = {}
    for q, layer in enumerate(m.quantizer.layers):
        W[f"quantizer.layers.{q}.codebook.embed"] = layer.codebook.embed.detach().numpy().astype(np.float32)

    def conv(pre, mod, transposed=False):
        w = mod.conv.weight.detach()
        w = w.permute(1, 2, 0) if transposed else w.permute(0, 2, 1)
        W[pre + "conv.weight"] = w.contiguous().numpy().astype(np.float32)
        W[pre + "conv.bias"] = mod.conv.bias.detach().numpy().astype(np.float32)

    for i, layer in enumerate(m.decoder.layers):
        pre = f"decoder.layers.{i}."
        name = type(layer).__name__
        if name == "EncodecConv1d":
            conv(pre, layer)
        elif name == "EncodecConvTranspose1d":
            conv(pre, layer, transposed=True)
        elif name == "EncodecLSTM":
            for l in range(layer.lstm.num_layers):
                W[pre + f"lstm.{l}.Wx"] = getattr(layer.lstm, f"weight_ih_l{l}").detach().numpy().astype(np.float32)
                W[pre + f"lstm.{l}.Wh"] = getattr(layer.lstm, f"weight_hh_l{l}").detach().numpy().astype(np.float32)
                W[pre + f"lstm.{l}.bias"] = (getattr(layer.lstm, f"bias_ih_l{l}") + getattr(layer.lstm, f"bias_hh_l{l}")).detach().numpy().astype(np.float32)
        elif name == "EncodecResnetBlock":
            for bi, sub in enumerate(layer.block):
                if type(sub).__name__ == "EncodecConv1d":
                    conv(pre + f"block.{bi}.", sub)
            conv(pre + "shortcut.", layer.shortcut)
    return W


@pytest.fixture(scope="module")
def hf():
    from transformers import EncodecConfig as HC, EncodecModel
    torch.manual_seed(0)
    m = EncodecModel(HC()).eval().float()
    for layer in m.quantizer.layers:                       # codebooks are zero-initialised buffers
        layer.codebook.embed.normal_()
    return m


def test_decoder_matches_transformers(hf):
    cfg = oe.EncodecConfig()
    W = hf_to_mlx(hf)
    codes = np.random.default_rng(0).integers(0, 1024, size=(2, 8, 23))
    with torch.no_grad():
        md = hf.double()
        emb = md.quantizer.decode(torch.from_numpy(codes).transpose(0, 1))           # [B, dim, T]
        ref = md.decoder(emb).numpy()                                                 # [B, 1, T*320]
        hf.float()
    mine_emb = oe.quantizer_decode(W, codes)
    assert np.abs(mine_emb - emb.numpy().transpose(0, 2, 1)).max() < 1e-5
    y = oe.decode(cfg, W, codes[None])
    assert y.shape == (2, 23 * 320, 1)
    err = np.abs(y[:, :, 0] - ref[:, 0, :]).max() / np.abs(ref).max()
    assert err < 1e-5, err                                                            # fp32-rounded weights vs HF's own: ~1e-7


def test_layout_and_keys_follow_the_module_array():
    cfg = oe.EncodecConfig()
    kinds = [k for _, k, _ in oe.decoder_layout(cfg)]
    assert kinds == ["conv", "lstm"] + ["elu", "convt", "resnet"] * 4 + ["elu", "conv"]
    assert cfg.hop_length == 320 and cfg.frame_rate == 75 and cfg.num_quantizers == 32
    W = oe.init_weights(cfg, 0, n_codebooks=2)
    assert W["decoder.layers.3.conv.weight"].shape == (256, 16, 512)
    assert W["decoder.layers.4.block.1.conv.weight"].shape == (128, 3, 256)
    assert W["decoder.layers.4.block.3.conv.weight"].shape == (256, 1, 128)
    assert W["decoder.layers.1.lstm.1.Wh"].shape == (2048, 512)


def test_reflect_pad_clamps_short_inputs():
    # EncodecLayers.swift:160-186: left index min(p - i, L - 1), right index max(L - 2 - i, 0)
    x = np.arange(3, dtype=np.float64).reshape(1, 3, 1)
    y = oe.pad1d(x, 6, 2, "reflect")[0, :, 0]
    assert y.tolist() == [2, 2, 2, 2, 2, 1, 0, 1, 2, 1, 0]
    assert oe.pad1d(x, 2, 1, "constant")[0, :, 0].tolist() == [0, 0, 0, 1, 2, 0]


def test_conv_transpose_trims_stride_samples_on_the_right():
    cfg = oe.EncodecConfig()
    w = np.ones((1, 4, 1), np.float32); b = np.zeros(1, np.float32)
    y = oe.conv_transpose1d(cfg, np.ones((1, 3, 1)), w, b, stride=2)           # (3-1)*2+4 = 8 -> trim 2 right -> 6 = T*s
    assert y.shape == (1, 6, 1) and y[0, :, 0].tolist() == [1, 1, 2, 2, 2, 2]


def test_linear_overlap_add_weights():
    # Encodec.swift:315-317: w[t] = 0.5 - |(t+1)/(L+1) - 0.5|; constant frames stay constant where any weight is non-zero
    f = [np.ones((1, 8, 1)), np.ones((1, 8, 1)), np.ones((1, 8, 1))]
    y = oe.linear_overlap_add(f, 4)
    assert y.shape == (1, 16, 1) and np.allclose(y, 1.0)
    g = [np.zeros((1, 4, 1)), np.ones((1, 4, 1))]
    z = oe.linear_overlap_add(g, 2)[0, :, 0]
    wv = 0.5 - np.abs((np.arange(4) + 1) / 5 - 0.5)
    assert np.allclose(z[2:4], wv[:2] / (wv[2:] + wv[:2])) and np.allclose(z[4:], 1.0) and np.allclose(z[:2], 0.0)


def test_chunked_decode_uses_overlap_add():
    cfg = oe.EncodecConfig(chunk_length_s=0.04, overlap=0.5, num_filters=4, hidden_size=8, codebook_dim=8, codebook_size=16)
    W = oe.init_weights(cfg, 1, n_codebooks=2)
    codes = np.random.default_rng(1).integers(0, 16, size=(3, 1, 2, 3))
    y = oe.decode(cfg, W, codes, [None, np.array([2.0]), None])
    assert cfg.chunk_length == 960 and cfg.chunk_stride == 480 and y.shape == (1, 2 * 480 + 960, 1)


def test_oracle_reproduces_committed_golden():
    """tests/golden/codecs.npz (tests/golden/make_golden.py --only codecs): the fixture the GPU tests also compare with."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "codecs.npz")
    cfg = oe.EncodecConfig()
    W = oe.init_weights(cfg, 7, n_codebooks=8)
    codes = np.random.default_rng(1).integers(0, 1024, size=(1, 3, 8, 41))
    z = oe.decode(cfg, W, codes)
    assert tuple(g["encodec_shape"]) == z.shape
    assert np.abs(z[:, :64, 0] - g["encodec_first"]).max() < 1e-6 and np.abs(z[:, -64:, 0] - g["encodec_last"]).max() < 1e-6
    zz = z.reshape(-1)
    assert np.allclose([zz.mean(), np.abs(zz).mean(), zz.min(), zz.max()], g["encodec_stats"], rtol=1e-6, atol=1e-9)


def test_linear_overlap_add_matches_transformers():
    """linearOverlapAdd (Encodec.swift:304-356) against transformers' EncodecModel._linear_overlap_add on ragged last frames."""
    import torch
    from transformers import EncodecModel
    rng = np.random.default_rng(2)
    for n_frames, L, last, hop in ((3, 8, 8, 4), (4, 12, 7, 6), (2, 10, 10, 9)):
        frames = [rng.standard_normal((2, L if i < n_frames - 1 else last, 1)) for i in range(n_frames)]
        ours = oe.linear_overlap_add(frames, hop)
        ref = EncodecModel._linear_overlap_add([torch.from_numpy(f).permute(0, 2, 1) for f in frames], hop).permute(0, 2, 1).numpy()
        assert ours.shape == ref.shape and np.abs(ours - ref).max() < 1e-12
