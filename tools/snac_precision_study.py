"""CPU study: which operand-splitting scheme do SNAC's tensor-core GEMMs need to stay inside the 1e-3 parity bar?

The float64 oracle decode (oracle/snac.py, SNAC-24 kHz geometry, random-init weights, explicit noise) is re-run with every DENSE
convolution / transposed convolution replaced by an emulation of the tensor-core arithmetic: operands rounded to bf16 or fp16 (optionally
as hi + lo pairs), the listed products summed, fp32 result.  Depthwise convolutions stay exact (they run on the CUDA cores in fp32).

    python tools/snac_precision_study.py MODE [T]      MODE: exact | 3 | 2x | 2w | 1   (bf16)  |  3h | 2x16 | 1h   (fp16)

Results (T = 16 latent steps, max |error| / peak):  see profiles/r01_snac_precision_study.md."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, torch.nn.functional as F
from oracle import snac
def bf16(x): return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)
def split(x):
    hi=bf16(x); lo=bf16(x-hi); return hi,lo
MODE=sys.argv[1]
orig_c, orig_t = snac.wn_conv1d, snac.wn_conv_transpose1d
def f16(x): return x.to(torch.float32).to(torch.float16).to(torch.float64)
def products(conv, W, x):
    if MODE in ("2x16","1h","3h"):
        Wh=f16(W); Wl=f16(W-Wh); xh=f16(x); xl=f16(x-xh)
        if MODE=="2x16": return conv(Wh,xh)+conv(Wh,xl)
        if MODE=="3h": return conv(Wh,xh)+conv(Wh,xl)+conv(Wl,xh)
        return conv(Wh,xh)
    Wh,Wl=split(W); xh,xl=split(x)
    if MODE=="exact": return conv(W,x)
    if MODE=="3": return conv(Wh,xh)+conv(Wh,xl)+conv(Wl,xh)
    if MODE=="2x": return conv(Wh,xh)+conv(Wh,xl)          # bf16 weights, 16-bit activations
    if MODE=="2w": return conv(Wh,xh)+conv(Wl,xh)          # 16-bit weights, bf16 activations
    if MODE=="1": return conv(Wh,xh)
def wn_conv1d(w, prefix, x, *, padding=0, dilation=1, groups=1, stride=1):
    if groups!=1: return orig_c(w,prefix,x,padding=padding,dilation=dilation,groups=groups,stride=stride)
    b=w.get(prefix+".bias"); W=snac.wn_conv_weight(w,prefix)
    y=products(lambda W_,x_: F.conv1d(x_,W_,None,stride=stride,padding=padding,dilation=dilation), W, x)
    y=y.to(torch.float32).to(torch.float64)
    return y if b is None else y+snac._t(b)[None,:,None]
def wn_conv_transpose1d(w, prefix, x, *, stride, padding):
    b=w.get(prefix+".bias"); W=snac.wn_convT_weight(w,prefix)
    y=products(lambda W_,x_: F.conv_transpose1d(x_,W_,None,stride=stride,padding=padding,output_padding=0), W, x)
    y=y.to(torch.float32).to(torch.float64)
    return y if b is None else y+snac._t(b)[None,:,None]
cfg=snac.SNACConfig(); W=snac.init_weights(cfg,1234)
T=int(sys.argv[2]) if len(sys.argv)>2 else 8
codes=snac.synth_codes(cfg,1,T,seed=2)
rng=np.random.default_rng(0)
noise=[rng.standard_normal(s).astype(np.float32) for s in snac.noise_shapes(cfg,1,T)]
ref=snac.decode(cfg,W,codes,noise)
snac.wn_conv1d, snac.wn_conv_transpose1d = wn_conv1d, wn_conv_transpose1d
y=snac.decode(cfg,W,codes,noise)
ref=np.asarray(ref); y=np.asarray(y)
print(MODE, "max err/peak", np.abs(y-ref).max()/np.abs(ref).max(), "rel L2", np.linalg.norm(y-ref)/np.linalg.norm(ref), ref.shape)
