"""CPU model of the (experimental) CUDA speech-tokenizer decoder's DATA FLOW, checked against the oracle.

csrc/speech_tokenizer.cu has not run on a GPU yet; what can be checked without one is its design: the planar activation
buffers with a history prefix, the history carry with ping-pong state, the tap -> frame mapping of the implicit convolution,
the phase-major transposed-convolution weight matrix (taken from the library itself through the host-only
b2a_speech_tokenizer_debug_layout), the fused epilogue order and the reference's bias-twice behaviour at chunk boundaries.
``Machine`` below follows b2a_speech_tokenizer::step_dev statement by statement with numpy standing in for each kernel
(implicit_conv() implements exactly the Args contract documented in csrc/implicit_conv.cuh).  If this model matches the oracle
and the CUDA kernels match their contracts, the CUDA path matches the oracle.  CPU only."""
import ctypes as C
import math

import numpy as np
import pytest

import implicit_conv_model
from implicit_conv_model import implicit_conv
from oracle import qwen3_tts_codec as oc


def lib_layout(b2a, w, stride=0):
    """[out, k, in] MLX weight -> the library's GEMM matrix [rows, taps, kpad] (host-only C entry)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    out, k, cin = w.shape
    rows, taps, kpad = C.c_int32(), C.c_int32(), C.c_int32()
    cap = max(stride, 1) * out * k * (cin + 64)
    buf = np.zeros(cap, np.float32)
    b2a._ffi.check(b2a._ffi.lib().b2a_speech_tokenizer_debug_layout(b2a._ffi.ptr(w), out, k, cin, stride, b2a._ffi.ptr(buf), cap,
                                                                    C.byref(rows), C.byref(taps), C.byref(kpad)))
    return buf[: rows.value * taps.value * kpad.value].reshape(rows.value, taps.value, kpad.value).astype(np.float64), cin


class PlaneState:
    def __init__(self, B, H, Cc):
        self.s, self.H = [np.zeros((B, H, Cc)), np.zeros((B, H, Cc))], H


class Machine:
    """b2a_speech_tokenizer: constructor = weight preparation, step() = step_dev."""

    def __init__(self, b2a, cfg, W, B):
        self.cfg, self.B = cfg, B
        f64 = lambda k: W[k].double().numpy()
        lay = lambda k, stride=0: lib_layout(b2a, W[k].numpy(), stride)
        snake = lambda p: (np.exp(f64(p + ".alpha")), 1.0 / (np.exp(f64(p + ".beta")) + 1e-9))
        nq, ns = cfg.num_quantizers, cfg.num_semantic_quantizers
        self.D2 = cfg.codebook_dim // 2
        emb = []
        for qi in range(nq):
            p = f"quantizer.rvq_first.vq.layers.{qi}" if qi < ns else f"quantizer.rvq_rest.vq.layers.{qi - ns}"
            emb.append(f64(p + ".codebook.embedding_sum") / np.maximum(f64(p + ".codebook.cluster_usage"), 1e-5)[:, None])
        self.emb = emb
        w = np.concatenate([f64("quantizer.rvq_first.output_proj.weight")[:, 0, :], f64("quantizer.rvq_rest.output_proj.weight")[:, 0, :]], axis=1)
        self.rvq_proj = (w[:, None, :], 2 * self.D2)
        self.pre_conv, self.pre_b = lay("pre_conv.conv.weight"), f64("pre_conv.conv.bias")
        lin = lambda k: (f64(k)[:, None, :], W[k].shape[1])
        self.in_proj, self.in_b = lin("pre_transformer.input_proj.weight"), f64("pre_transformer.input_proj.bias")
        self.out_proj, self.out_b = lin("pre_transformer.output_proj.weight"), f64("pre_transformer.output_proj.bias")
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"pre_transformer.layers.{i}."
            qkv = np.concatenate([f64(p + f"self_attn.{n}_proj.weight") for n in "qkv"], axis=0)
            gu = np.concatenate([f64(p + "mlp.gate_proj.weight"), f64(p + "mlp.up_proj.weight")], axis=0)
            self.layers.append(dict(qkv=(qkv[:, None, :], cfg.hidden_size), o=lin(p + "self_attn.o_proj.weight"), gu=(gu[:, None, :], cfg.hidden_size),
                                    down=lin(p + "mlp.down_proj.weight"), ln1=f64(p + "input_layernorm.weight"), ln2=f64(p + "post_attention_layernorm.weight"),
                                    sa=f64(p + "self_attn_layer_scale.scale"), sm=f64(p + "mlp_layer_scale.scale"), K=None, V=None))
        self.ups = []
        for i, f in enumerate(cfg.upsampling_ratios):
            p = f"upsample.{i}.layers."
            self.ups.append(dict(f=f, ct=lay(p + "0.conv.weight", f), ct_b=f64(p + "0.conv.bias"), dw=f64(p + "1.dwconv.conv.weight")[:, :, 0], dw_b=f64(p + "1.dwconv.conv.bias"),
                                 ln_w=f64(p + "1.norm.weight"), ln_b=f64(p + "1.norm.bias"), pw1=lin(p + "1.pwconv1.weight"), pw1_b=f64(p + "1.pwconv1.bias"),
                                 pw2=lin(p + "1.pwconv2.weight"), pw2_b=f64(p + "1.pwconv2.bias"), gamma=f64(p + "1.gamma"),
                                 st=PlaneState(B, 6, cfg.latent_dim)))
        self.dec0, self.dec0_b = lay("decoder.0.conv.weight"), f64("decoder.0.conv.bias")
        self.blocks = []
        for b, r in enumerate(cfg.upsample_rates):
            p = f"decoder.{1 + b}.block."
            cin, cout = cfg.decoder_dim >> b, cfg.decoder_dim >> (b + 1)
            rus = []
            for j, d in enumerate((1, 3, 9)):
                q = p + f"{2 + j}."
                rus.append(dict(dil=d, a1=snake(q + "act1"), a2=snake(q + "act2"), c1=lay(q + "conv1.conv.weight"), c1_b=f64(q + "conv1.conv.bias"),
                                c2=lay(q + "conv2.conv.weight"), c2_b=f64(q + "conv2.conv.bias"), st=PlaneState(B, 6 * d, cout)))
            self.blocks.append(dict(rate=r, cin=cin, cout=cout, sn=snake(p + "0"), ct=lay(p + "1.conv.weight", r), ct_b=f64(p + "1.conv.bias"),
                                    st=PlaneState(B, 1, cin), ru=rus))
        n = len(cfg.upsample_rates)
        self.out_snake, self.out_w, self.out_b_ = snake(f"decoder.{n + 1}"), f64(f"decoder.{n + 2}.conv.weight")[0], float(f64(f"decoder.{n + 2}.conv.bias")[0])
        self.st_pre, self.st_dec0 = PlaneState(B, 2, cfg.codebook_dim), PlaneState(B, 6, cfg.latent_dim)
        self.st_out = PlaneState(B, self.out_w.shape[0] - 1, cfg.output_dim)
        self.final_norm = f64("pre_transformer.norm.weight")
        self.parity = self.chunk_idx = self.cache_len = 0

    def carry(self, X, st, T):                                    # carry_planes_kernel
        old, new, H = st.s[self.parity], st.s[self.parity ^ 1], st.H
        X[:, :H] = old
        for f in range(H):
            src = T + f
            new[:, f] = old[:, src] if src < H else X[:, src]

    def update_f32(self, x, st, T):                               # state_update_f32_kernel
        old, new, H = st.s[self.parity], st.s[self.parity ^ 1], st.H
        for f in range(H):
            src = T + f
            new[:, f] = old[:, src] if src < H else x[:, src - H]

    def front_half(self, codes):
        """step_dev sections 1-3: gathers, pre_conv, transformer layers.  Returns the residual stream [B, T, hidden]."""
        c, B = self.cfg, self.B
        codes = np.asarray(codes)
        _, nq, T = codes.shape
        cbd, L, Hd, I, nh, nkv, hd = c.codebook_dim, c.latent_dim, c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim
        # 1. gathers + projections
        P0 = np.zeros((B, T, 2 * self.D2))
        for qi in range(nq):
            off = 0 if qi < c.num_semantic_quantizers else self.D2
            P0[:, :, off:off + self.D2] += self.emb[qi][codes[:, qi]]
        P1 = np.zeros((B, T + 2, cbd))
        implicit_conv(*self.rvq_proj, P0, T, hl=P1, Hout=2)
        # 2. pre_conv -> input_proj
        self.carry(P1, self.st_pre, T)
        P0 = np.zeros((B, T, L))
        implicit_conv(*self.pre_conv, P1, T, bias=self.pre_b, hl=P0)
        Xh = np.zeros((B, T, Hd))
        implicit_conv(*self.in_proj, P0, T, bias=self.in_b, xo=Xh)
        # 3. transformer
        rms = lambda x, w: w * (x / np.sqrt((x * x).mean(-1, keepdims=True) + c.rms_norm_eps))
        inv = 1.0 / (c.rope_theta ** (np.arange(0, hd, 2) / hd))
        pos = self.cache_len + np.arange(T)
        cs, sn = np.cos(pos[:, None] * inv), np.sin(pos[:, None] * inv)
        half = hd // 2
        for Ly in self.layers:
            P0 = rms(Xh, Ly["ln1"])
            Q = np.zeros((B, T, (nh + 2 * nkv) * hd))
            implicit_conv(*Ly["qkv"], P0, T, xo=Q)
            heads = Q.reshape(B, T, nh + 2 * nkv, hd)
            rot = heads[:, :, : nh + nkv].copy()                                      # rope_cache_kernel
            x1, x2 = rot[..., :half].copy(), rot[..., half:].copy()
            rot[..., :half] = x1 * cs[None, :, None, :] - x2 * sn[None, :, None, :]
            rot[..., half:] = x2 * cs[None, :, None, :] + x1 * sn[None, :, None, :]
            q, k, v = rot[:, :, :nh], rot[:, :, nh:], heads[:, :, nh + nkv:]
            Ly["K"] = k if Ly["K"] is None else np.concatenate([Ly["K"], k], axis=1)
            Ly["V"] = v if Ly["V"] is None else np.concatenate([Ly["V"], v], axis=1)
            P1 = np.zeros((B, T, nh * hd))
            for h in range(nh):                                                       # attn_kernel
                kvh = h // (nh // nkv)
                for t in range(T):
                    n = self.cache_len + t + 1
                    s = np.einsum("bd,bpd->bp", q[:, t, h], Ly["K"][:, :n, kvh]) / math.sqrt(hd)
                    p = np.exp(s - s.max(-1, keepdims=True))
                    P1[:, t, h * hd:(h + 1) * hd] = np.einsum("bp,bpd->bd", p / p.sum(-1, keepdims=True), Ly["V"][:, :n, kvh])
            implicit_conv(*Ly["o"], P1, T, xo=Xh, add=True, gamma=Ly["sa"])
            P0 = rms(Xh, Ly["ln2"])
            Q = np.zeros((B, T, 2 * I))
            implicit_conv(*Ly["gu"], P0, T, xo=Q)
            g, u = Q[..., :I], Q[..., I:]
            P1 = g / (1.0 + np.exp(-g)) * u
            implicit_conv(*Ly["down"], P1, T, xo=Xh, add=True, gamma=Ly["sm"])
        return Xh, T


def machine_step(m, codes):
    """step_dev: front half, then sections 4-7 (final norm / output projection, upsample layers, decoder, output conv)."""
    c, B = m.cfg, m.B
    Xh, T = m.front_half(codes)
    first = m.chunk_idx == 0
    L = c.latent_dim
    rms = lambda x, w: w * (x / np.sqrt((x * x).mean(-1, keepdims=True) + c.rms_norm_eps))
    P0 = rms(Xh, m.final_norm)
    cur = np.zeros((B, T, L))
    implicit_conv(*m.out_proj, P0, T, bias=m.out_b, hl=cur)
    # 4. upsample layers
    Tc, Hcur = T, 0
    for i, U in enumerate(m.ups):
        Xc = np.zeros((B, Tc * U["f"], L))
        implicit_conv(*U["ct"], cur, Tc, up=U["f"], bias=U["ct_b"], xo=Xc)
        Tc *= U["f"]
        old = U["st"].s[m.parity]                                                     # dw_ln_kernel
        ext = np.concatenate([old, Xc], axis=1)
        dwv = sum(U["dw"][None, None, :, kk] * ext[:, kk:kk + Tc] for kk in range(7)) + U["dw_b"]
        mu, var = dwv.mean(-1, keepdims=True), dwv.var(-1, keepdims=True)
        cur = (dwv - mu) / np.sqrt(var + 1e-6) * U["ln_w"] + U["ln_b"]
        m.update_f32(Xc, U["st"], Tc)
        other = np.zeros((B, Tc, 4 * L))
        implicit_conv(*U["pw1"], cur, Tc, bias=U["pw1_b"], gelu=True, hl=other)
        Hcur = 0 if i + 1 < len(m.ups) else m.st_dec0.H
        cur = np.zeros((B, Hcur + Tc, L))
        implicit_conv(*U["pw2"], other, Tc, bias=U["pw2_b"], xo=Xc, add=True, gamma=U["gamma"], hl=cur, Hout=Hcur)
    # 5. decoder.0
    m.carry(cur, m.st_dec0, Tc)
    other = np.zeros((B, 1 + Tc, c.decoder_dim))
    implicit_conv(*m.dec0, cur, Tc, bias=m.dec0_b, hl=other, Hout=1, sa=m.blocks[0]["sn"][0], sb=m.blocks[0]["sn"][1])
    cur = other
    # 6. decoder blocks
    for b, Bk in enumerate(m.blocks):
        m.carry(cur, Bk["st"], Tc)
        r, cout = Bk["rate"], Bk["cout"]
        Xc = np.zeros((B, Tc * r, cout))
        H0 = Bk["ru"][0]["st"].H
        other = np.zeros((B, H0 + Tc * r, cout))
        implicit_conv(*Bk["ct"], cur, Tc, up=r, bias=Bk["ct_b"], xo=Xc, hl=other, Hout=H0, sa=Bk["ru"][0]["a1"][0], sb=Bk["ru"][0]["a1"][1],
                      bias_twice_t0=not first)
        Tc *= r
        cur = other
        for j, R in enumerate(Bk["ru"]):
            m.carry(cur, R["st"], Tc)
            other = np.zeros((B, Tc, cout))
            implicit_conv(*R["c1"], cur, Tc, dil=R["dil"], bias=R["c1_b"], hl=other, sa=R["a2"][0], sb=R["a2"][1])
            if j < 2:
                nxt, Hn = Bk["ru"][j + 1]["a1"], Bk["ru"][j + 1]["st"].H
            elif b + 1 < len(m.blocks):
                nxt, Hn = m.blocks[b + 1]["sn"], 1
            else:
                nxt, Hn = None, 0
            cur = np.zeros((B, Hn + Tc, cout)) if nxt is not None else None
            implicit_conv(*R["c2"], other, Tc, bias=R["c2_b"], xo=Xc, add=True, hl=cur, Hout=Hn,
                          sa=None if nxt is None else nxt[0], sb=None if nxt is None else nxt[1])
    # 7. output snake + conv + clip (final_conv_kernel)
    H = m.st_out.H
    ext = np.concatenate([m.st_out.s[m.parity], Xc], axis=1)
    act = ext + m.out_snake[1] * np.sin(m.out_snake[0] * ext) ** 2
    y = sum((act[:, kk:kk + Tc] * m.out_w[kk]).sum(-1) for kk in range(H + 1)) + m.out_b_
    m.update_f32(Xc, m.st_out, Tc)
    m.parity ^= 1
    m.chunk_idx += 1
    m.cache_len += T
    return np.clip(y, -1.0, 1.0)


def design_config():
    """codebook_dim / latent multiples of 8 with a 96-style non-multiple-of-64 channel count in the decoder (48, 24)."""
    return oc.tiny_config(latent_dim=40, codebook_dim=16, decoder_dim=192, hidden_size=24, head_dim=8, upsample_rates=[4, 3, 2], upsampling_ratios=[2, 2])


@pytest.mark.parametrize("chunks", [[6], [1, 1, 1, 1], [3, 1, 4], [2, 7]])
def test_data_flow_model_matches_oracle_streaming(b2a, chunks):
    cfg = design_config()
    W = oc.init_weights(cfg, 21)
    B = 2
    m = Machine(b2a, cfg, W, B)
    d = oc.SpeechTokenizerDecoder(cfg, W)
    d.reset_streaming_state()
    codes = np.random.default_rng(4).integers(0, cfg.codebook_size, (B, cfg.num_quantizers, sum(chunks)))
    s = 0
    for n in chunks:
        y = machine_step(m, codes[:, :, s:s + n])
        ref = d.streaming_step(codes[:, :, s:s + n])[:, 0].numpy()
        assert y.shape == ref.shape and np.abs(ref).max() > 1e-3
        assert np.abs(y - ref).max() < 1e-6, (s, np.abs(y - ref).max())      # library layouts are float32-rounded weights
        s += n


@pytest.mark.parametrize("operand,lo,hi", [("bf16", 1e-6, 6e-4), ("fp16", 1e-7, 1.5e-4)])
def test_predicted_numerics_of_the_hilo_arithmetic_stay_inside_the_parity_tolerance(b2a, monkeypatch, operand, lo, hi):
    """The same data-flow model with the kernel's arithmetic (bf16 hi/lo operands, three products, fp32 accumulation): the error
    it predicts against the float64 oracle must leave room under the 1e-3 parity bar of the GPU tests.  (Measured here: about
    3e-4 of the peak at this geometry and at the shipped one with bf16 pairs, 6e-5 with fp16 pairs -- Args::f16, B2A_ST_FP16=1 --;
    plain fp32 arithmetic gives 3e-6, so the operand split is the whole budget.)"""
    import sys
    monkeypatch.setattr(sys.modules[__name__], "implicit_conv", implicit_conv_model.implicit_conv_hilo)
    monkeypatch.setattr(implicit_conv_model, "OPERAND", operand)
    cfg = oc.mid_config()
    W = oc.init_weights(cfg, 5)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (1, cfg.num_quantizers, 12))
    y = machine_step(Machine(b2a, cfg, W, 1), codes)
    ref = oc.SpeechTokenizerDecoder(cfg, W)(codes)[:, 0].numpy()
    err = np.abs(y - ref).max() / np.abs(ref).max()
    assert lo < err < hi, err


def test_library_weight_layouts(b2a):
    rng = np.random.default_rng(0)
    w = rng.standard_normal((5, 6, 70)).astype(np.float32)                     # [out, k, in], in = 70 -> kpad 128
    g, cin = lib_layout(b2a, w)
    assert g.shape == (5, 6, 128) and cin == 70
    assert np.array_equal(g[:, :, :70], w) and not g[:, :, 70:].any()
    g, _ = lib_layout(b2a, w, stride=3)                                        # k = 2 * 3: rows = 3 * 5 phase-major, 2 taps
    assert g.shape == (15, 2, 128)
    for rho in range(3):
        assert np.array_equal(g[rho * 5:(rho + 1) * 5, 1, :70], w[:, rho, :])          # tap 1 <-> input frame q     <-> kernel index rho
        assert np.array_equal(g[rho * 5:(rho + 1) * 5, 0, :70], w[:, rho + 3, :])      # tap 0 <-> input frame q - 1 <-> kernel index rho + stride
    g, _ = lib_layout(b2a, w[:, :3], stride=3)                                 # k = stride: one tap, no history
    assert g.shape == (15, 1, 128) and np.array_equal(g[5:10, 0, :70], w[:, 1, :])
    with pytest.raises(b2a.AudioGenerationError):
        lib_layout(b2a, w[:, :5], stride=3)                                    # kernel not a multiple of the stride
