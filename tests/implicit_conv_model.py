"""numpy statement of the Args contract of ic::implicit_conv_kernel (csrc/implicit_conv.cuh).  Shared by the CPU data-flow
model (test_speech_tokenizer_design.py) and the gated GPU kernel test (test_gpu_implicit_conv.py).  Test infrastructure."""
import math

import numpy as np


def implicit_conv(Wg, cin, X, T, *, dil=1, shift0=0, up=1, bias=None, gamma=None, gelu=False, add=False, bias_twice_t0=False,
                  xo=None, hl=None, Hout=0, sa=None, sb=None):
    """The Args contract of ic::implicit_conv_kernel.  X: planes as one float64 array [B, Ttot, cin]; xo [B, T*up, Cout] and
    hl [B, Hout + T*up, Cout] are written in place."""
    M, taps, _ = Wg.shape
    B, Ttot, _ = X.shape
    Cout = M // up
    acc = np.zeros((B, T, M))
    for j in range(taps):
        for t in range(T):
            f = t + shift0 + j * dil
            if 0 <= f < Ttot:                                      # out-of-range frames are TMA zero fill
                acc[:, t, :] += X[:, f, :] @ Wg[:, j, :cin].T
    for rho in range(up):
        val = acc[:, :, rho * Cout:(rho + 1) * Cout].copy()
        if bias is not None:
            val += bias
            if bias_twice_t0:
                val[:, 0, :] += bias
        if gelu:
            val = 0.5 * val * (1.0 + np.vectorize(math.erf)(val / math.sqrt(2.0)))
        if gamma is not None:
            val = val * gamma
        fo = np.arange(T) * up + rho
        if add:
            val = val + xo[:, fo, :]
        if xo is not None:
            xo[:, fo, :] = val
        if hl is not None:
            hv = val + sb * np.sin(sa * val) ** 2 if sa is not None else val
            hl[:, Hout + fo, :] = hv
