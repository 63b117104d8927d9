"""Thread-level emulation of the SIMT kernels of csrc/speech_tokenizer.cu (experimental row-N1 path, not yet run on a GPU): the index
arithmetic of each kernel is replayed in Python exactly as written -- block / thread ids, lane-owned dimensions, batches of four keys,
tile rows, the 16-byte vector decomposition -- and compared with the kernel's contract.  The data-flow test
(test_speech_tokenizer_design.py) checks what the kernels are asked to do; this one checks how they index.  CPU only."""
import math

import numpy as np
import pytest


# --------------------------------------------------------------------------------------------- carry_planes_kernel
@pytest.mark.parametrize("B,T,H,C", [(2, 5, 6, 16), (1, 3, 54, 8), (3, 1, 1, 24), (2, 70, 18, 32)])
def test_carry_planes_kernel_indexing(B, T, H, C):
    rng = np.random.default_rng(B * 100 + T)
    C8 = C // 8
    X = rng.standard_normal((2 * B * (H + T) * C8, 8))             # "uint4" elements of X = [2][B][H + T][C8]
    X.reshape(2, B, H + T, C8, 8)[:, :, :H] = np.nan               # the history prefix is garbage before the kernel runs
    old = rng.standard_normal((2 * B * H * C8, 8))
    new = np.full_like(old, np.nan)
    Xk = X.copy()
    per_plane = B * H * C8
    for i in range(2 * per_plane):                                 # one thread per i, as written
        c = i % C8
        r = i // C8
        f = r % H; r //= H
        b = r % B; p = r // B
        xrow = (p * B + b) * (H + T)
        ov = old[i]
        Xk[(xrow + f) * C8 + c] = ov
        src = T + f
        new[i] = old[((p * B + b) * H + src) * C8 + c] if src < H else X[(xrow + src) * C8 + c]     # reads only frames >= H of X: never a slot this kernel writes
    # contract: X[:, :, :H] = old ; new = last H frames of [old | fresh frames]
    Xr = X.reshape(2, B, H + T, C8, 8).copy()
    oldr = old.reshape(2, B, H, C8, 8)
    Xr[:, :, :H] = oldr
    full = np.concatenate([oldr, X.reshape(2, B, H + T, C8, 8)[:, :, H:]], axis=2)
    assert np.array_equal(Xk.reshape(2, B, H + T, C8, 8), Xr)
    assert np.array_equal(new.reshape(2, B, H, C8, 8), full[:, :, -H:])


# --------------------------------------------------------------------------------------------- state_update_f32_kernel
def test_state_update_f32_kernel_indexing():
    B, T, H, C = 2, 4, 6, 5
    rng = np.random.default_rng(0)
    x, old = rng.standard_normal((B * T * C,)), rng.standard_normal((B * H * C,))
    new = np.empty_like(old)
    for i in range(B * H * C):
        c = i % C
        bf = i // C
        f, b = bf % H, bf // H
        src = T + f
        new[i] = old[(b * H + src) * C + c] if src < H else x[(b * T + (src - H)) * C + c]
    full = np.concatenate([old.reshape(B, H, C), x.reshape(B, T, C)], axis=1)
    assert np.array_equal(new.reshape(B, H, C), full[:, -H:])


# --------------------------------------------------------------------------------------------- attn_kernel
@pytest.mark.parametrize("DPL,T,pos0,nh,nkv", [(2, 5, 0, 4, 2), (1, 3, 6, 2, 2), (4, 2, 9, 2, 1), (2, 1, 11, 4, 4)])
def test_attn_kernel_online_softmax_in_batches_of_four(DPL, T, pos0, nh, nkv):
    HD, B, cap = DPL * 32, 2, 16
    rng = np.random.default_rng(DPL + T)
    ld = (nh + 2 * nkv) * HD
    qkv = rng.standard_normal((B * T, ld))
    Kc, Vc = rng.standard_normal((B, nkv, cap, HD)), rng.standard_normal((B, nkv, cap, HD))
    scale = 1.0 / math.sqrt(HD)
    out = np.zeros((B * T, nh * HD))
    for b in range(B):
        for h in range(nh):
            for t in range(T):                                     # one warp
                n = b * T + t
                kvh = h // (nh // nkv)
                q = [qkv[n, h * HD + lane * DPL: h * HD + (lane + 1) * DPL] * scale for lane in range(32)]     # lane-owned dims
                acc = [np.zeros(DPL) for _ in range(32)]
                m, l = -np.inf, 0.0
                nkeys = pos0 + t + 1
                for p0 in range(0, nkeys, 4):
                    s = []
                    for u in range(4):
                        d0 = [float(q[lane] @ Kc[b, kvh, p0 + u, lane * DPL:(lane + 1) * DPL]) if p0 + u < nkeys else 0.0 for lane in range(32)]
                        s.append(sum(d0))                          # wsum: the butterfly leaves the total in every lane
                    mx = m
                    for u in range(4):
                        if p0 + u < nkeys:
                            mx = max(mx, s[u])
                    corr = math.exp(m - mx) if m != -np.inf else 0.0
                    l *= corr
                    acc = [a * corr for a in acc]
                    for u in range(4):
                        if p0 + u < nkeys:
                            e = math.exp(s[u] - mx)
                            l += e
                            for lane in range(32):
                                acc[lane] = acc[lane] + e * Vc[b, kvh, p0 + u, lane * DPL:(lane + 1) * DPL]
                    m = mx
                for lane in range(32):
                    out[n, h * HD + lane * DPL: h * HD + (lane + 1) * DPL] = acc[lane] / l
    # contract: causal softmax attention of query t over cache positions [0, pos0 + t]
    for b in range(B):
        for h in range(nh):
            kvh = h // (nh // nkv)
            for t in range(T):
                n = b * T + t
                k = Kc[b, kvh, : pos0 + t + 1]
                sc = (k @ qkv[n, h * HD:(h + 1) * HD]) * scale
                p = np.exp(sc - sc.max()); p /= p.sum()
                assert np.abs(out[n, h * HD:(h + 1) * HD] - p @ Vc[b, kvh, : pos0 + t + 1]).max() < 1e-12


# --------------------------------------------------------------------------------------------- rope_cache_kernel
def test_rope_cache_kernel_indexing():
    B, T, nh, nkv, hd, cap, pos0 = 2, 3, 4, 2, 8, 10, 4
    half, ld = hd // 2, (nh + 2 * nkv) * hd
    rng = np.random.default_rng(2)
    qkv = rng.standard_normal((B * T, ld))
    q0 = qkv.copy()
    inv = 1.0 / (10000.0 ** (np.arange(0, hd, 2) / hd))
    Kc, Vc = np.zeros((B, nkv, cap, hd)), np.zeros((B, nkv, cap, hd))
    for n in range(B * T):                                         # one block per token
        b, t = n // T, n % T
        pos = pos0 + t
        row = qkv[n]
        for idx in range((nh + nkv) * half):
            head, i = idx // half, idx % half
            sn, cs = math.sin(pos * inv[i]), math.cos(pos * inv[i])
            x1, x2 = row[head * hd + i], row[head * hd + i + half]
            o1, o2 = x1 * cs - x2 * sn, x2 * cs + x1 * sn
            if head < nh:
                row[head * hd + i], row[head * hd + i + half] = o1, o2
            else:
                Kc[b, head - nh, pos, i], Kc[b, head - nh, pos, i + half] = o1, o2
        for idx in range(nkv * hd):
            kvh, d = idx // hd, idx % hd
            Vc[b, kvh, pos, d] = row[(nh + nkv) * hd + idx]
    # contract: x * cos + rotate_half(x) * sin on q (in place) and k (into the cache at pos0 + t); v copied
    for n in range(B * T):
        b, t = n // T, n % T
        ang = (pos0 + t) * inv
        cos, sin = np.concatenate([np.cos(ang)] * 2), np.concatenate([np.sin(ang)] * 2)
        rot = lambda x: np.concatenate([-x[half:], x[:half]])
        for h in range(nh):
            x = q0[n, h * hd:(h + 1) * hd]
            assert np.abs(qkv[n, h * hd:(h + 1) * hd] - (x * cos + rot(x) * sin)).max() < 1e-12
        for kv in range(nkv):
            x = q0[n, (nh + kv) * hd:(nh + kv + 1) * hd]
            assert np.abs(Kc[b, kv, pos0 + t] - (x * cos + rot(x) * sin)).max() < 1e-12
            assert np.array_equal(Vc[b, kv, pos0 + t], q0[n, (nh + nkv + kv) * hd:(nh + nkv + kv + 1) * hd])


# --------------------------------------------------------------------------------------------- final_conv_kernel
@pytest.mark.parametrize("T,C,k", [(150, 96, 7), (64, 16, 7), (1, 24, 3)])
def test_final_conv_kernel_tiles(T, C, k):
    FC_TILE, FC_THREADS, B = 64, 128, 2
    H, rows, ldc = k - 1, 64 + k - 1, C + 1
    rng = np.random.default_rng(T)
    x, st = rng.standard_normal((B, T, C)), rng.standard_normal((B, H, C))
    sa, sb, w, bias = np.exp(rng.standard_normal(C) * 0.3), np.exp(rng.standard_normal(C) * 0.3), rng.standard_normal((k, C)), 0.3
    snake = lambda v, a, ib: v + ib * np.sin(a * v) ** 2
    wave = np.full((B, T), np.nan)
    for by in range(B):
        for bx in range(-(-T // FC_TILE)):
            t0 = bx * FC_TILE
            tile = np.zeros(rows * ldc)
            for i in range(rows * C):                              # staged by all threads (strided loop): every (row, channel) once
                rr, c = i // C, i % C
                ti = t0 - H + rr
                if ti >= 0:
                    v = x[by, ti, c] if ti < T else 0.0
                else:
                    v = st[by, H + ti, c]
                tile[rr * ldc + c] = snake(v, sa[c], sb[c])
            red = np.zeros(FC_THREADS)
            for tid in range(FC_THREADS):
                o, part = tid & (FC_TILE - 1), tid // FC_TILE
                cbeg = part * ((C + 1) // 2)
                cend = min(C, cbeg + (C + 1) // 2)
                acc = 0.0
                for kk in range(k):
                    for c in range(cbeg, cend):
                        acc += w[kk, c] * tile[(o + kk) * ldc + c]
                red[tid] = acc
            for tid in range(FC_THREADS):
                o, part = tid & (FC_TILE - 1), tid // FC_TILE
                if part == 0 and t0 + o < T:
                    wave[by, t0 + o] = min(1.0, max(-1.0, red[o] + red[o + FC_TILE] + bias))
    ext = np.concatenate([st, x], axis=1)
    act = snake(ext, sa, sb)
    ref = np.clip(sum((act[:, kk:kk + T] * w[kk]).sum(-1) for kk in range(k)) + bias, -1, 1)
    assert np.abs(wave - ref).max() < 1e-12


# --------------------------------------------------------------------------------------------- dw_ln_kernel
def test_dw_ln_kernel_indexing():
    DL_THREADS, DL_MAXV, B, T, C, k = 256, 4, 2, 5, 300, 7
    H = k - 1
    rng = np.random.default_rng(3)
    x, st = rng.standard_normal((B * T, C)), rng.standard_normal((B * H, C))
    dw_w, dw_b, ln_w, ln_b = rng.standard_normal((C, k)).reshape(-1), rng.standard_normal(C), rng.standard_normal(C), rng.standard_normal(C)
    out = np.zeros((B * T, C))
    for n in range(B * T):
        b, t = n // T, n % T
        v = np.zeros((DL_THREADS, DL_MAXV))
        for tid in range(DL_THREADS):
            for j in range(DL_MAXV):
                c = tid + j * DL_THREADS
                if c < C:
                    val = dw_b[c]
                    for kk in range(k):
                        ti = t - H + kk
                        xin = x[b * T + ti, c] if ti >= 0 else st[b * H + (H + ti), c]
                        val += dw_w[c * k + kk] * xin
                    v[tid, j] = val
        mean = v.sum() / C
        q = sum((v[tid, j] - mean) ** 2 for tid in range(DL_THREADS) for j in range(DL_MAXV) if tid + j * DL_THREADS < C)
        r = 1.0 / math.sqrt(q / C + 1e-6)
        for tid in range(DL_THREADS):
            for j in range(DL_MAXV):
                c = tid + j * DL_THREADS
                if c < C:
                    out[n, c] = (v[tid, j] - mean) * r * ln_w[c] + ln_b[c]
    ext = np.concatenate([st.reshape(B, H, C), x.reshape(B, T, C)], axis=1)
    W = dw_w.reshape(C, k)
    dwv = sum(W[None, None, :, kk] * ext[:, kk:kk + T] for kk in range(k)) + dw_b
    ref = (dwv - dwv.mean(-1, keepdims=True)) / np.sqrt(dwv.var(-1, keepdims=True) + 1e-6) * ln_w + ln_b
    assert np.abs(out.reshape(B, T, C) - ref).max() < 1e-10


# --------------------------------------------------------------------------------------------- ic::implicit_conv_kernel (coordinates and epilogue mapping)
@pytest.mark.parametrize("M,taps,cin,B,T,dil,up,grid", [(200, 3, 72, 2, 70, 2, 1, 5), (96, 2, 96, 1, 67, 1, 2, 3), (256, 1, 64, 3, 5, 1, 1, 148)])
def test_implicit_conv_kernel_tiles_and_epilogue_mapping(M, taps, cin, B, T, dil, up, grid):
    """Producer coordinates (tile -> (b, frame tile, m tile); k-block -> (tap, channel block) -> TMA box origin), the accumulator layout the
    MMAs leave in TMEM (lane = weight row, column = frame, + 64 for the lo half) and the epilogue's (warp, lane, j) -> (m, frame, phase) mapping,
    replayed CTA by CTA and compared with the kernel's contract.  The tensor-core products themselves are taken as exact."""
    from implicit_conv_model import implicit_conv
    BM, BK, HALF = 128, 64, 64
    rng = np.random.default_rng(M + T)
    cblocks = -(-cin // BK)
    Wg = np.zeros((M, taps, cblocks * BK)); Wg[:, :, :cin] = rng.standard_normal((M, taps, cin))
    H = (taps - 1) * dil
    Ttot = H + T
    X = rng.standard_normal((B, Ttot, cin))
    Cout, Hout = M // up, 3
    bias = rng.standard_normal(Cout)
    To = T * up
    xo = np.full((B, To, Cout), np.nan)
    hl = np.zeros((B, Hout + To, Cout))
    m_tiles, t_tiles = -(-M // BM), -(-T // HALF)
    tiles = B * t_tiles * m_tiles
    Wflat = Wg.reshape(M, taps * cblocks * BK)                     # the K-major matrix the A tensor map reads
    done = np.zeros((B, To, Cout), dtype=int)
    for cta in range(min(grid, tiles)):
        for t in range(cta, tiles, min(grid, tiles)):
            nt, mt = t // m_tiles, t % m_tiles
            b, tt = nt // t_tiles, nt % t_tiles
            D = np.zeros((BM, HALF))                               # hi + lo columns already summed (the epilogue adds v[j] + w[j])
            kb = 0
            for j in range(taps):
                frame = tt * HALF + 0 + j * dil                    # shift0 = 0: the input carries exactly H history frames
                for cb in range(cblocks):
                    A = np.zeros((BM, BK))
                    rows = slice(mt * BM, min(M, (mt + 1) * BM))
                    A[: rows.stop - rows.start] = Wflat[rows, kb * BK:(kb + 1) * BK]       # TMA box {kb * 64, mt * 128}, rows past M zero-filled
                    Bt = np.zeros((HALF, BK))                                               # TMA box {cb * 64, frame, b, 0}: frames / channels out of range zero-filled
                    f1, c1 = min(Ttot, frame + HALF), min(cin, (cb + 1) * BK)
                    if f1 > frame and c1 > cb * BK:
                        Bt[: f1 - frame, : c1 - cb * BK] = X[b, frame:f1, cb * BK:c1]
                    D += A @ Bt.T
                    kb += 1
            for warp in range(2, 18):                              # epilogue warps
                q, c0 = warp & 3, ((warp - 2) >> 2) * 16
                for lane in range(32):
                    m = mt * BM + q * 32 + lane
                    if m >= M:
                        continue
                    rho, co = (m // Cout, m % Cout) if up > 1 else (0, m)
                    for jj in range(16):
                        tf = tt * HALF + c0 + jj
                        if tf >= T:
                            continue
                        val = D[q * 32 + lane, c0 + jj] + bias[co]
                        fo = tf * up + rho
                        xo[b, fo, co] = val
                        hl[b, Hout + fo, co] = val
                        done[b, fo, co] += 1
    assert (done == 1).all()                                       # every output written exactly once
    rx, rh = np.zeros((B, To, Cout)), np.zeros((B, Hout + To, Cout))
    implicit_conv(Wg, cin, X, T, dil=dil, up=up, bias=bias, xo=rx, hl=rh, Hout=Hout)
    assert np.abs(xo - rx).max() < 1e-10 and np.abs(hl - rh).max() < 1e-10
