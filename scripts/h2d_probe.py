"""Pinned host -> device bandwidth for one step's feature block (B=1024: 302 MB fp32)."""
import torch
dev = torch.device("cuda")
x = torch.empty(1024, 36, 2048, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device=dev)
for n in (1, 2, 3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"H2D {x.numel()*4/1e6:.0f} MB in {ms:.2f} ms = {x.numel()*4/ms/1e6:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
