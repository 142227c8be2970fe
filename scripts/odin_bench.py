"""ODIN scoring (forward + input-gradient backward + forward, fp32 path) at the reference scripts' batch size."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn
from rgqa_b200.lxrt.input_grad import LXMERTInputGrad
from rgqa_b200.tasks.gqa_odin import GraphedInputGradScorer, odin_scores
sd = syn.lxmert_state_dict(variant="init")
ig = LXMERTInputGrad(sd, gemm=os.environ.get("GEMM", "fp32"))
for B in (32, 128):
    inp = syn.make_inputs(B, 20, seed=1)
    args = [inp[k].to(ig.dev) for k in ("feat", "pos", "ids", "seg", "mask")]
    for _ in range(2): odin_scores(ig, *args, 1.0, 1e-4)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    K = 5
    for _ in range(K): odin_scores(ig, *args, 1.0, 1e-4)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"ODIN B={B}: {dt*1e3:.1f} ms/batch  {B/dt:.0f} q/s ({ig.gemm} matrix products, launches issued from Python)")
    gs = GraphedInputGradScorer(ig, "odin", 1.0, 1e-4)
    want = odin_scores(ig, *args, 1.0, 1e-4)[0].clone()
    for _ in range(3): got = gs(*args)[0]
    torch.cuda.synchronize()
    err = (got - want).abs().max().item()
    t0 = time.perf_counter()
    for _ in range(4 * K): gs(*args)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / (4 * K)
    print(f"ODIN B={B}: {dt*1e3:.1f} ms/batch  {B/dt:.0f} q/s (one CUDA-graph replay per batch; max |score - eager| = {err:.2e})")
