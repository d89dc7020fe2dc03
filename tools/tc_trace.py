"""Dump the per-role clock64 trace of CTA 0 of the tcgen05 kernel for one layer shape.
   MIGAN_TC_TRACE="512,64,64" python tools/tc_trace.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import migan_b200
from migan_b200 import synthetic, _abi

R, N = 512, int(os.environ.get("N", "16"))
dev = torch.device("cuda:0")
m = migan_b200.Generator(R, path="tc")
m.load_state_dict(synthetic.export_style_state_dict(R))
m = m.to(dev).eval()
x = synthetic.synthetic_input(R, N).to(dev)
for _ in range(3):
    m(x)
torch.cuda.synchronize()
buf = np.zeros(4096, dtype=np.uint64)
_abi.check(_abi.load().migan_debug_read_tc_trace(buf.ctypes.data_as(ctypes.c_void_p)))
t = buf.reshape(4, 64, 16).astype(np.int64)
t0 = t[t > 0].min()
names = {0: ["tile", "g0wait", "g0got", "g1wait", "g1got"], 1: ["wacc", "gotacc", "gotA", "gotB", "commit"],
         2: ["wfull", "gotfull", "freed", "end", "ld0", "cmp0", "wg0", "bar0a", "sts0", "bar0b", "st0", "rgb0", "rgb1", "-", "top", "-"], 3: ["win", "gotin", "gotA", "done", "arrived"]}
roles = ["producer", "mma", "epilogue", "prologue"]
print("trace for", os.environ.get("MIGAN_TC_TRACE"), "ablate", os.environ.get("MIGAN_TC_ABLATE", "0"))
for it in list(range(0, 6)) + list(range(20, 26)):
    for r in range(4):
        ev = " ".join("%s=%d" % (names[r][e], t[r, it, e] - t0) for e in range(len(names[r])) if t[r, it, e] > 0)
        print("tile %2d %-9s %s" % (it, roles[r], ev))
for r in range(4):
    col = t[r, :, 1 if r != 0 else 0]
    v = col[col > 0]
    if len(v) > 8:
        d = np.diff(v[4:])
        print("%-9s per-tile period (clk): median %d  mean %d" % (roles[r], np.median(d), d.mean()))
