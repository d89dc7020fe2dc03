"""Timing experiments on the tcgen05 kernel: run migan-512 forwards with pipeline stages skipped
(MIGAN_TC_ABLATE bitmask; results are numerically wrong) and print per-launch times."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import migan_b200
from migan_b200 import synthetic

R, N = 512, int(os.environ.get("N", "16"))
dev = torch.device("cuda:0")
m = migan_b200.Generator(R, path="tc")
m.load_state_dict(synthetic.export_style_state_dict(R))
m = m.to(dev).eval()
x = synthetic.synthetic_input(R, N).to(dev)
for _ in range(3):
    m(x)
m.set_profiling(True)
acc = {}
for _ in range(5):
    m(x); torch.cuda.synchronize()
    for lb, ms, nb, fl in m.profile_steps():
        a = acc.setdefault(lb, [0.0, nb]); a[0] += ms / 5
tag = "ablate=%s prefetch=%s" % (os.environ.get("MIGAN_TC_ABLATE", "0"), os.environ.get("MIGAN_TC_PREFETCH", "-"))
keys = ["encoder.b512.conv1.sepconv_tc", "encoder.b512.conv2.sepconv_tc", "encoder.b256.conv1.sepconv_tc", "encoder.b128.conv1.sepconv_tc",
        "encoder.b64.conv1.sepconv_tc", "synthesis.b512.conv1.sepconv_tc", "synthesis.b512.conv2.sepconv_tc", "synthesis.b256.conv2.sepconv_tc"]
print(tag, " ".join("%s=%.3f" % (k.replace("encoder.", "e.").replace("synthesis.", "s.").replace(".sepconv_tc", ""), acc[k][0]) for k in keys),
      "total_tc=%.3f" % sum(v[0] for k, v in acc.items() if "sepconv" in k), flush=True)
