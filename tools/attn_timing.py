"""Debug: per-phase SM-clock timestamps of attn_decode_kernel (build with B2A_ATTN_TIMING)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m
from bench import ORPHEUS
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 320
tts = m.LlamaTTSModel.random_init(ORPHEUS, max_batch=8, max_context=640)
print("ms/step", tts.time_steps(8, ctx, 10))
lib = C.CDLL(str(m._ffi.LIB_PATH))
n = 8 * 8 * 8 * 10
buf = np.zeros(n, dtype=np.int64)
lib.b2a_debug_attn_ts(buf.ctypes.data_as(C.c_void_p), n)
ts = buf.reshape(10, 8, 8, 8)     # [z][y][x][stamp]
act = ts[: (ctx // 64 + 1)]
d = lambda a, b: (act[..., b] - act[..., a]).astype(np.float64)
names = ["start->after pdl_wait", "->staging done", "->bulk landed", "->QK/softmax/PV/partials", "->fence+atomic"]
for i, nm in enumerate(names):
    x = d(i, i + 1)
    print(f"{nm:32s} mean {x.mean():8.0f} cyc  min {x.min():8.0f} max {x.max():8.0f}")
last = act[..., 6] > act[..., 5]
x = (act[..., 6] - act[..., 5])[last]
print(f"{'merge (last CTAs)':32s} mean {x.mean():8.0f} cyc  n={last.sum()}")
x = (act[..., 5] - act[..., 1])
print(f"{'after wait -> done (non-merge)':32s} mean {x.mean():8.0f} max {x.max():8.0f}")
t0 = act[..., 0].min()
print("span first start -> last end (cycles):", max(act[..., 5].max(), act[..., 6].max()) - t0, "  last start:", act[..., 0].max() - t0)
