"""Generates tests/golden/train_steps.npz: three optimisation steps of the UNMODIFIED reference LXMERT modules (2 language /
1 cross / 1 relational layer, the real vocabulary size) with the reference trainer's recipe — BCEWithLogitsLoss × A,
clip_grad_norm_(5.), the reference BertAdam (warmup_linear) — and BOTH dropout probabilities set to 0 (SURVEY §8(d) config 5:
"loss parity vs reference for 3 steps with dropout off").  Stored: per-step loss and pre-clip gradient norm, the first
step's gradient of a few tensors, and a sample of the parameters after step 3.
Run in the build container:  python tests/golden/make_golden_train.py"""
import os, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
for name in ("boto3", "botocore", "botocore.exceptions"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
sys.path[:0] = [os.path.join(REF, "src"), REF]
sys.argv = ["x", "--fromScratch"]
import lxrt.modeling as M                      # noqa: E402
from lxrt.optimization import BertAdam         # noqa: E402
from rgqa_b200 import synthetic as syn         # noqa: E402

LAYERS = dict(l_layers=2, x_layers=1, r_layers=1)
M.VISUAL_CONFIG.l_layers, M.VISUAL_CONFIG.x_layers, M.VISUAL_CONFIG.r_layers = 2, 1, 1
torch.manual_seed(0)
cfg = M.BertConfig(30522, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
enc = M.LXRTFeatureExtraction(cfg, mode='x')
A = 1842
head = torch.nn.Sequential(torch.nn.Linear(768, 1536), M.GeLU(), M.BertLayerNorm(1536, eps=1e-12), torch.nn.Linear(1536, A))
sd = syn.lxmert_state_dict(variant="rich", **LAYERS)
enc.load_state_dict({k[len(syn.LXMERT_PREFIX):]: v for k, v in sd.items() if k.startswith(syn.LXMERT_PREFIX)}, strict=True)
head.load_state_dict({k[len("logit_fc."):]: v for k, v in sd.items() if k.startswith("logit_fc.")}, strict=True)
enc.train(); head.train()
params = list(enc.parameters()) + list(head.parameters())
named = {syn.LXMERT_PREFIX + k: v for k, v in enc.named_parameters()}
named.update({"logit_fc." + k: v for k, v in head.named_parameters()})
assert set(named) == set(sd)
opt = BertAdam(params, lr=1e-4, warmup=0.1, t_total=20)
bce = torch.nn.BCEWithLogitsLoss()
B, L = 6, 20
g = torch.Generator().manual_seed(404)
out = {}
WATCH = ["logit_fc.3.bias", "logit_fc.2.weight", "lxrt_encoder.model.bert.pooler.dense.weight",
         "lxrt_encoder.model.bert.encoder.x_layers.0.visual_attention.att.key.weight",
         "lxrt_encoder.model.bert.encoder.x_layers.0.visual_attention.output.LayerNorm.weight",
         "lxrt_encoder.model.bert.encoder.layer.0.attention.self.query.bias",
         "lxrt_encoder.model.bert.encoder.r_layers.0.output.dense.weight",
         "lxrt_encoder.model.bert.encoder.visn_fc.box_fc.weight", "lxrt_encoder.model.bert.encoder.visn_fc.visn_fc.bias",
         "lxrt_encoder.model.bert.embeddings.position_embeddings.weight", "lxrt_encoder.model.bert.embeddings.LayerNorm.bias",
         "lxrt_encoder.model.bert.embeddings.token_type_embeddings.weight"]
def cut(t):          # big matrices: the first 4 rows are enough to pin them (rows are full reductions over the batch)
    return t[:4] if t.numel() > 20000 else t


for step in range(3):
    inp = syn.make_inputs(B, L, seed=900 + step)
    target = torch.zeros(B, A)
    target[torch.arange(B // 2), torch.randint(0, A, (B // 2,), generator=g)] = 1.0      # the other half: RP negatives (all-zero rows)
    opt.zero_grad()
    pooled = enc(inp["ids"], inp["seg"], inp["mask"], visual_feats=(inp["feat"], inp["pos"]))
    logit = head(pooled)
    loss = bce(logit, target) * logit.size(1)
    loss.backward()
    if step == 0:
        for n in WATCH:
            out["grad0:" + n] = cut(named[n].grad).numpy().copy()
        out["grad0_word_rows"] = named[syn.LXMERT_PREFIX + "bert.embeddings.word_embeddings.weight"].grad[[101, 102, 0]].numpy().copy()
    norm = torch.nn.utils.clip_grad_norm_(params, 5.)
    opt.step()
    out[f"loss{step}"], out[f"norm{step}"] = np.array(loss.item()), np.array(float(norm))
    out[f"target{step}"] = target.numpy().copy()
    print(step, loss.item(), float(norm))
for n in WATCH:
    out["final:" + n] = cut(named[n].detach()).numpy().copy()
np.savez_compressed(os.path.join(HERE, "train_steps.npz"), **out)
print("bytes", os.path.getsize(os.path.join(HERE, "train_steps.npz")))
