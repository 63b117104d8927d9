"""Debug: per-phase globaltimer stamps of the last GEMM-chain launch of a decode step (B2A_CHAIN_DEBUG=1)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

os.environ["B2A_CHAIN_DEBUG"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mlx_audio_swift_b200 as m  # noqa: E402
from bench import ORPHEUS  # noqa: E402

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 320
tts = m.LlamaTTSModel.random_init(ORPHEUS, max_batch=8, max_context=640)
print("ms/step", tts.time_steps(8, ctx, 10))
lib = C.CDLL(str(m._ffi.LIB_PATH))
buf = np.zeros(148 * 64, dtype=np.uint64)
rc = lib.b2a_debug_chain_ts(tts._h, buf.ctypes.data_as(C.c_void_p), buf.size)
ts = buf.reshape(148, 64).astype(np.float64)
t0 = ts[:, 0][ts[:, 0] > 0].min()
names = {0: "epi start (after dep wait)", 40: "producer start", 41: "producer end"}
for gi in range(4):
    names.update({1 + 6 * gi: f"g{gi} norm begin", 2 + 6 * gi: f"g{gi} norm done", 3 + 6 * gi: f"g{gi} norm arrived",
                  4 + 6 * gi: f"g{gi} first acc ready", 5 + 6 * gi: f"g{gi} epilogue done", 6 + 6 * gi: f"g{gi} arrived",
                  32 + 2 * gi: f"g{gi} producer flush begin", 33 + 2 * gi: f"g{gi} producer flush end"})
for slot in sorted(names, key=lambda s: (np.nanmedian(np.where(ts[:, s] > 0, ts[:, s], np.nan)) if (ts[:, s] > 0).any() else 1e30)):
    v = ts[:, slot]
    ok = v > 0
    if not ok.any():
        continue
    d = (v[ok] - t0) / 1e3
    print(f"{names[slot]:32s} n={ok.sum():4d}  min {d.min():8.2f}  median {np.median(d):8.2f}  max {d.max():8.2f} us")
