"""The four hot kernels of a visual self-attention + FFN layer in the round-2b form (single-word fp16 state), three launches each at
M = 36 864 (B = 1024): target of ONE `ncu --set full -k regex:"gemm_bf16_pair|qkv_attn_pair"` capture (the last launch of each
kind is the warm one).  Q|K|V projection + attention (fp16 x fp16), attention-output producer (bf16 x bf16, fp16 state in / out),
FFN1 consumer (fp16 x fp16, erf-GELU), FFN2 producer (K = 3072)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import ops, _lib as L
from scripts.r02_kernel_bench import head_major
lib = L.load()
dev = torch.device("cuda")
B, S, H = 1024, 36, 768
M = B * S
gen = torch.Generator().manual_seed(0)
g = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
h = (torch.randn(M, H, generator=gen) * 1.5).to(torch.float16).to(dev)
v = h.float()
st = torch.stack([v.view(M, 6, 128).sum(-1), (v * v).view(M, 6, 128).sum(-1)], -1).contiguous()
def folded(N):
    w = (torch.randn(N, H, generator=gen) * 0.03).to(dev); bias = torch.randn(N, generator=gen).to(dev)
    return ops.fold_layernorm_weight(w, bias, g, b, dtype=torch.float16)
wq, cq, bq = folded(2304)
w1, c1, b1 = folded(3072)
w_hm, b_hm = head_major(wq, bq)
wo = (torch.randn(H, H, generator=gen) * 0.03).to(torch.bfloat16).to(dev); bo = torch.randn(H, generator=gen).to(dev)
w2 = (torch.randn(H, 3072, generator=gen) * 0.03).to(torch.bfloat16).to(dev)
ctx = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    L.check(lib.rgqa_qkv_attention_fused_a16(L.current_stream_ptr(), h.data_ptr(), H, w_hm.data_ptr(), b_hm.data_ptr(), st.data_ptr(),
                                             1e-12, None, ctx.data_ptr(), H, B, S, 12), "fused")              # Q|K|V + self-attention
for _ in range(3):
    h2, s2 = ops.linear_residual_h16(ctx, wo, bo, h, st, g, None)                                            # attention-output producer
for _ in range(3):
    inter = ops.linear_a16_folded(h2, s2, w1, b1, epilogue=L.EPI_GELU)                                       # FFN1 consumer (erf-GELU)
for _ in range(3):
    ops.linear_residual_h16(inter, w2, bo, h2, s2, g, None)                                                  # FFN2 producer
torch.cuda.synchronize()
print("ok")
