"""UNITER-base (BASELINE config 3) forward + MSP at B = 1024 (and 32): device-resident inputs, CUDA-event timing."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn, _lib as L
from rgqa_b200.engine import Engine, EngineConfig
lib = L.load()
cfg = EngineConfig(arch="uniter", precision="bf16")
cfg.l_layers, cfg.x_layers, cfg.r_layers, cfg.vocab_size, cfg.visual_pos_dim = 12, 0, 0, 28996, 7
eng = Engine(cfg); eng.load_weights(syn.uniter_state_dict(variant="init").items())
GF_PER_Q = 9.7511
for B in (32, 1024):
    ins = []
    for j in range(2):
        inp = syn.make_inputs(B, 20, seed=3 + j, vocab=28996)
        ins.append([inp[k].to(eng.device) for k in ("feat", "pos7", "ids", "seg", "mask")])
    for opt in (0, 1):
        lib.rgqa_set_option(b"defer_ln", opt)
        def step(i):
            pooled, logits = eng.forward(*ins[i & 1])
            return eng.score("msp", logits=logits)
        for i in range(6): step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record()
        for i in range(K): step(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"UNITER B={B:5d} defer_ln={opt}: {ms:7.2f} ms/step  {B/ms*1e3:8.0f} q/s  {B/ms*1e3*GF_PER_Q/1e3:6.1f} TFLOP/s", flush=True)
lib.rgqa_set_option(b"defer_ln", 1)
