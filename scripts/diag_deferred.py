"""Logit error of the deferred-LayerNorm bf16 path vs the LayerNorm-kernel bf16 path, both against the fp32 CUDA path
(B = 64, L = 20, 'rich' weights with the x10 head of the tests; errors are divided by 10 = reference logit scale)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn, _lib as L
from rgqa_b200.engine import Engine, EngineConfig
lib = L.load()
sd = syn.lxmert_state_dict(variant="rich", head_sigma=0.2)
def eng(prec):
    e = Engine(EngineConfig(arch="lxmert", precision=prec)); e.load_weights(sd.items()); return e
e32, e16 = eng("fp32"), eng("bf16")
for B in (64, 256):
    inp = syn.make_inputs(B, 20, seed=5 + B)
    args = [inp[k].to(e32.device) for k in ("feat", "pos", "ids", "seg", "mask")]
    _, l32 = e32.forward(*args)
    out = {}
    for name, v in (("layernorm kernels", 0), ("deferred", 1)):
        lib.rgqa_set_option(b"defer_ln", v)
        _, l = e16.forward(*args)
        d = (l - l32).abs() / 10.0
        out[name] = (float(d.pow(2).mean().sqrt()), float(d.max()), float((d <= 2e-2).float().mean()))
    lib.rgqa_set_option(b"defer_ln", 1)
    for k, (rms, mx, fr) in out.items():
        print(f"B={B:4d} {k:18s}: rms {rms:.5f}  max {mx:.4f}  within 2e-2: {fr:.5f}")
