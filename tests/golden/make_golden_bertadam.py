"""Generates tests/golden/bertadam_steps.npz with the UNMODIFIED reference optimiser (src/lxrt/optimization.py) and the
reference trainer's clipping (nn.utils.clip_grad_norm_(params, 5.), src/tasks/gqa_conf.py:197): three tensors, four
steps, warmup_linear schedule (warmup 0.1, t_total 10: steps 0-1 inside the warm-up), weight decay 0.01.
Run in the build container:  python tests/golden/make_golden_bertadam.py"""
import os, sys
import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from lxrt.optimization import BertAdam  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
g = torch.Generator().manual_seed(31)
shapes = [(7, 33), (129,), (64, 48)]
params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.05) for s in shapes]
out = {f"p0_{i}": p.detach().numpy().copy() for i, p in enumerate(params)}
opt = BertAdam(params, lr=1e-3, warmup=0.1, t_total=10)
STEPS = 4
for k in range(STEPS):
    scale = [3.0, 0.02, 1.0, 40.0][k]                 # global norms on both sides of the 5.0 clip
    for i, p in enumerate(params):
        p.grad = torch.randn(p.shape, generator=g) * scale
        out[f"g{k}_{i}"] = p.grad.numpy().copy()
    out[f"norm{k}"] = np.array(float(torch.nn.utils.clip_grad_norm_(params, 5.)))
    out[f"lr{k}"] = np.array(opt.get_lr()[0])
    opt.step()
    for i, p in enumerate(params):
        out[f"p{k + 1}_{i}"] = p.detach().numpy().copy()
for i, p in enumerate(params):
    out[f"m_{i}"] = opt.state[p]["next_m"].numpy().copy()
    out[f"v_{i}"] = opt.state[p]["next_v"].numpy().copy()
# BCE loss * A and its gradient on a small batch (gqa_conf.py:193-194)
logit = torch.randn(6, 41, generator=g, requires_grad=True) * 3
logit.retain_grad()
target = (torch.rand(6, 41, generator=g) < 0.05).float()
z = logit.clone().detach().requires_grad_(True)
loss = torch.nn.BCEWithLogitsLoss()(z, target) * z.size(1)
loss.backward()
out.update(bce_logit=z.detach().numpy(), bce_target=target.numpy(), bce_loss=np.array(loss.item()), bce_dlogit=z.grad.numpy())
np.savez_compressed(os.path.join(HERE, "bertadam_steps.npz"), **out)
print("norms", [float(out[f"norm{k}"]) for k in range(STEPS)], "lrs", [float(out[f"lr{k}"]) for k in range(STEPS)])
