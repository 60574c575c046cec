"""Timeline of one CTA of the tensor-core LSTM kernel (experiment build with -DIC3_TC_EXP_TRACE only):

    IC3_NVCC_EXTRA=-DIC3_TC_EXP_TRACE python -m ic3net_b200.build --force
    python profiles/microbench/trace_lstm.py
    python -m ic3net_b200.build --force

Prints, for CTA 0 and each of its work items, when the producer / MMA issuer / epilogue (warp 0) reached their
hand-over points (microseconds from the first stamp), i.e. how far the three stages actually overlap.
"""
import argparse
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from ic3net_b200 import _lib  # noqa: E402
from ic3net_b200.comm import CommNetMLP  # noqa: E402

B, N, H, O = 8192, 10, 128, 64
args = argparse.Namespace(nagents=N, hid_size=H, comm_passes=1, recurrent=True, continuous=False, rnn_type='LSTM',
                          share_weights=False, naction_heads=[5, 2], comm_init='uniform', hard_attn=True,
                          comm_mode='avg', comm_mask_zero=False, seed=0)
torch.manual_seed(0)
net = CommNetMLP(args, O)
dev = torch.device('cuda')
x = torch.randn(B, N, O, device=dev)
h, c = torch.randn(B * N, H, device=dev) * 0.1, torch.randn(B * N, H, device=dev) * 0.1
info = {'comm_action': torch.ones(B, N, dtype=torch.uint8, device=dev)}
lib = _lib.load()
for _ in range(5):
    net([x, (h, c)], info)
torch.cuda.synchronize()
lib.ic3_debug_tc_trace_clear()
net([x, (h, c)], info)
buf = (C.c_ulonglong * (4 * 512))()
lib.ic3_debug_tc_trace.argtypes = [C.c_void_p]
assert lib.ic3_debug_tc_trace(buf) == 0
t = np.frombuffer(buf, dtype=np.uint64).reshape(4, 512).astype(np.int64)
t0 = min(int(v) for v in t.ravel() if v > 0)
us = lambda v: (int(v) - t0) / 1e3 if v > 0 else float('nan')
print("item | producer: chunk0 slot free, chunk11 slot free | issuer: acc free, chunk0 landed, chunk11 landed |"
      " epilogue w0: enter, acc full, math done")
for i in range(10):
    p = t[0, 2 * i:2 * i + 2]
    m = t[1, 3 * i:3 * i + 3]
    e = t[2, 4 * i:4 * i + 3]
    if not m.any():
        break
    print("%2d | %7.2f %7.2f | %7.2f %7.2f %7.2f | %7.2f %7.2f %7.2f" % ((i,) + tuple(us(v) for v in list(p) + list(m) + list(e))),
          "| own stage landed (pair kernel): %7.2f %7.2f" % tuple(us(v) for v in t[3, 2 * i:2 * i + 2]) if t[3].any() else "")
