"""CUDA-event timings of the round-2 fused kernels at the B = 1024 shapes, next to the kernels they replace.
  python scripts/r02_kernel_bench.py [--out gpurun_out/r02_kernels.json]
Each kernel: 3 warm-up launches, then 20 timed launches alternating between two operand sets (> L2 together where the
operands are large), median.  Not a bench.py number: per-kernel evidence for DESIGN.md / profiles/README.md."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import _lib as L, ops  # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn(0)
        fn(1)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i & 1)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def head_major(w, b):
    idx = torch.cat([torch.arange(part * 768 + h * 64, part * 768 + h * 64 + 64) for h in range(12) for part in range(3)]).to(w.device)
    return w[idx].contiguous(), b[idx].contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = L.load()
    res = {}
    g = torch.Generator().manual_seed(0)
    B = 1024

    # ---- visual feature encoder: one kernel (fp32 in / bf16 in) vs cast + GEMM + combine ----
    M, K = B * 36, 2048
    feats = [torch.rand(M, K, generator=g).to(dev) for _ in range(2)]
    featsb = [f.to(torch.bfloat16) for f in feats]
    W1 = (torch.randn(768, K, generator=g) * 0.02).to(dev)
    W1b = W1.to(torch.bfloat16)
    vec = lambda s=0.1, o=0.0: (o + s * torch.randn(768, generator=g)).to(dev)
    b1, g1, be1, b2, g2, be2 = vec(), vec(0.1, 1.0), vec(), vec(), vec(0.1, 1.0), vec()
    box = torch.rand(M, 4, generator=g).to(dev)
    W2 = (torch.randn(768, 4, generator=g) * 0.02).to(dev)
    hi = torch.empty(M, 768, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    st = torch.empty(M, 12, device=dev)
    # the entry point builds its tables per call (and syncs): time the kernel inside through the engine instead -> use events
    # around the ABI call minus the table work is not possible; so call the internal launcher via the engine path below.

    def visn(i, f32):
        src = feats[i] if f32 else featsb[i]
        w = W1 if f32 else W1b
        L.check(lib.rgqa_visn_embed(L.current_stream_ptr(), src.data_ptr(), L.DT_F32 if f32 else L.DT_BF16, w.data_ptr(), b1.data_ptr(),
                                    g1.data_ptr(), be1.data_ptr(), box.data_ptr(), 4, W2.data_ptr(), b2.data_ptr(), g2.data_ptr(),
                                    be2.data_ptr(), 1e-12, hi.data_ptr(), lo.data_ptr(), st.data_ptr(), M, K), "visn")
    # rgqa_visn_embed (a test entry point) builds its box tables per call and synchronises: the kernel's own duration is taken
    # from the library's event pair around the launch (rgqa_prof_enable / rgqa_prof_collect, kind 0)
    import ctypes as C

    def prof_time(fn, n=10):
        for i in range(4):
            fn(i & 1)
        ms, fl, cnt = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_longlong * 4)()
        lib.rgqa_prof_enable(1)
        lib.rgqa_prof_collect(ms, fl, cnt)
        for i in range(n):
            fn(i & 1)
        torch.cuda.synchronize()
        lib.rgqa_prof_collect(ms, fl, cnt)
        lib.rgqa_prof_enable(0)
        return ms[0] * 1e3 / max(cnt[0], 1)
    res["visn_embed_f32_tf32_us"] = prof_time(lambda i: visn(i, True))
    res["visn_embed_bf16_us"] = prof_time(lambda i: visn(i, False))
    pre = torch.empty(M, 768, device=dev)

    def old_visn(i):
        fb = ops.cast_bf16(feats[i])
        x = ops.linear_bf16(fb, W1b, b1, out_dtype=torch.float32)
        ops.visn_combine(x, box, W2, b2, g1, be1, g2, be2, out_dtype=torch.bfloat16)
    res["visn_old_cast_gemm_combine_us"] = timeit(old_visn, n=10)

    # ---- self-attention block front half: fused vs projection + attention ----
    for name, S in (("visual_S36", 36), ("language_S20", 20), ("uniter_S56", 56)):
        Mr = B * S
        his = [torch.randn(Mr, 768, generator=g).to(torch.bfloat16).to(dev) for _ in range(2)]
        stats = torch.zeros(Mr, 12, device=dev)
        stats[:, 1] = 768.0
        w = (torch.randn(2304, 768, generator=g) * 0.03).to(torch.bfloat16).to(dev)
        bias = (torch.randn(2304, generator=g) * 0.1).to(dev)
        w_hm, b_hm = head_major(w, bias)
        lens = torch.randint(1, S + 1, (B,), generator=g)
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(dev)
        ctx = torch.empty(Mr, 768, dtype=torch.bfloat16, device=dev)

        def fused(i):
            L.check(lib.rgqa_qkv_attention_fused(L.current_stream_ptr(), his[i].data_ptr(), 768, w_hm.data_ptr(), b_hm.data_ptr(),
                                                 stats.data_ptr(), 1e-12, mask.data_ptr(), ctx.data_ptr(), 768, B, S, 12), "fused")

        def split(i):
            qkv = ops.linear_bf16_folded(his[i], stats.view(Mr, 6, 2), w, bias, out_dtype=torch.bfloat16)
            ops.attention(qkv, qkv, B, S, S, key_mask=mask)
        res[f"qkv_attn_fused_{name}_us"] = timeit(fused)
        res[f"qkv_then_attention_{name}_us"] = timeit(split)
        fl = 2.0 * Mr * 2304 * 768
        res[f"qkv_attn_fused_{name}_tflops"] = fl / (res[f"qkv_attn_fused_{name}_us"] * 1e-6) / 1e12

    # ---- answer head with scorer epilogue vs GEMM + scorer kernel ----
    h1 = torch.randn(B, 1536, generator=g).to(torch.bfloat16).to(dev)
    w3 = (torch.randn(1842, 1536, generator=g) * 0.02).to(torch.bfloat16).to(dev)
    b3 = vec()[:1].repeat(1842).contiguous()
    work = torch.empty(int(lib.rgqa_reduce_work_bytes(B, 1842)), dtype=torch.uint8, device=dev)
    sc = torch.empty(4, B, device=dev)
    lb = torch.empty(B, dtype=torch.int32, device=dev)

    def head_fused(i):
        L.check(lib.rgqa_gemm_bf16_reduce_logits(L.current_stream_ptr(), h1.data_ptr(), 1536, w3.data_ptr(), 1536, b3.data_ptr(), None, B,
                                                 1842, 1536, 0, work.data_ptr(), work.numel(), sc.data_ptr(), lb.data_ptr()), "head")

    def head_split(i):
        lg = ops.linear_bf16(h1, w3, b3, out_dtype=torch.float32)
        ops.score_logits(lg, "energy")
    res["head_fc3_scorer_epilogue_us"] = timeit(head_fused)
    res["head_fc3_then_score_kernel_us"] = timeit(head_split)

    # ---- Mahalanobis: tensor cores (split bf16) vs fp32 CUDA cores ----
    from rgqa_b200 import synthetic as syn
    from rgqa_b200.engine import Engine, EngineConfig
    eng = Engine(EngineConfig(l_layers=1, x_layers=0, r_layers=0))
    eng.load_weights(syn.lxmert_state_dict(l_layers=1, x_layers=0, r_layers=0).items())
    eng.set_maha(*syn.maha_stats())
    x = torch.tanh(torch.randn(B, 768, generator=g)).to(dev)
    res["mahalanobis_tensor_core_us"] = timeit(lambda i: eng.score("maha", pooled=x))
    lib.rgqa_set_option(b"maha_tc", 0)
    res["mahalanobis_fp32_cuda_core_us"] = timeit(lambda i: eng.score("maha", pooled=x))
    lib.rgqa_set_option(b"maha_tc", 1)
    for k, v in res.items():
        print(f"{k:50s} {v:10.1f}")
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
