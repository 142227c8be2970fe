"""BASELINE north_star: "argmax answer agreement >= 99.9 %" for the bf16 path — MEASURED on 4 096 questions of the full-depth
(9/5/5) model against the fp32 CUDA path of the same weights (which the other GPU tests pin to the reference goldens / the
oracle at <= 1e-4), reported overall and per margin bucket, and ASSERTED where the statement is well-posed: random-init logits
have top-2 margins down to 1e-4, far inside any bf16-operand implementation's 2e-2 logit noise, so a swap there says
nothing; every question whose fp32 top-2 margin exceeds 4x the logit tolerance must keep its answer (100 %), and at least 99.9 %
of those above 2x.  Two weight sets: the reference's init (sigma 0.02 everywhere) and the same with the answer head's last layer
at sigma 0.2, which emulates a trained model's margins."""
import json

import pytest
import torch

from rgqa_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.mark.parametrize("head_sigma", [None, 0.2])
def test_bf16_answer_agreement_table_on_4096_questions(cuda_device, head_sigma):
    from rgqa_b200.engine import Engine, EngineConfig
    sd = syn.lxmert_state_dict(variant="rich", head_sigma=head_sigma)
    engs = {}
    for prec in ("fp32", "bf16"):
        engs[prec] = Engine(EngineConfig(precision=prec))
        engs[prec].load_weights(sd.items())
    N, B = 4096, 1024
    l32, l16, lab16 = [], [], []
    for c in range(N // B):
        inp = syn.make_inputs(B, 20, seed=100 + c)
        t = [inp["feat"].to(cuda_device), inp["pos"].to(cuda_device)] + [inp[k].to(torch.int32).to(cuda_device) for k in ("ids", "seg", "mask")]
        l32.append(engs["fp32"].forward(*t, want_pooled=False)[1])
        out, _, lg = engs["bf16"].forward_scored(*t, ["msp"], want_logits=True)   # the fused scorer path the sweep uses
        l16.append(lg)
        lab16.append(out["msp"][1])
    torch.cuda.synchronize()
    for e in engs.values():
        e.close()
    l32, l16, lab16 = torch.cat(l32), torch.cat(l16), torch.cat(lab16).long()
    d = (l16 - l32).abs()
    t2 = l32.topk(2, dim=1).values
    margin = t2[:, 0] - t2[:, 1]
    # the answer = arg-max of the logits.  (The MSP label is the arg-max of sigmoid(logit), which is the same class until the
    # sigmoid saturates to 1.0f — the x10 head reaches |logit| ~ 40 — where the reference's own label degenerates to "first
    # saturated class"; that tie semantics is pinned bit for bit in test_fused_scorers_gpu.py, not here.)
    same = l16.argmax(1) == l32.argmax(1)
    if head_sigma is None:
        assert torch.equal(lab16, l16.argmax(1))   # no saturation at init scale: the fused scorer's label is the arg-max
    table = {"questions": N, "head_sigma": head_sigma, "logit_std": float(l32.std()), "max_dlogit": float(d.max()),
             "rms_dlogit": float(d.pow(2).mean().sqrt()), "overall": float(same.float().mean())}
    for k in (1, 2, 4):
        sel = margin > k * TOL
        table[f"margin>{k}tol"] = {"n": int(sel.sum()), "agree": float(same[sel].float().mean())}
    print("AGREEMENT " + json.dumps(table))
    scale = max(1.0, float(l32.std()) / 0.79)   # the 2e-2 bound is stated for logits of std ~0.8; a x10 head scales the noise with it
    assert table["rms_dlogit"] <= 6e-3 * scale and table["max_dlogit"] <= 4e-2 * scale, table
    sel4, sel2 = margin > 4 * TOL * scale, margin > 2 * TOL * scale
    assert int(sel4.sum()) >= N // 20, table                     # not vacuous
    assert bool(same[sel4].all()), table                          # 100 % on the well-conditioned bucket
    assert float(same[sel2].float().mean()) >= 0.999, table
