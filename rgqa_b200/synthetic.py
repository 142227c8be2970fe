"""Seeded synthetic weights and inputs (there are no checkpoints, features or vocab files offline).

The tensors carry the REFERENCE state-dict key names, so the same dict loads
into the reference ``GQAModel`` (golden-vector generation, ``tests/golden/make_golden.py``), into the
CPU oracle and into the CUDA handle.  Everything is drawn from ``torch.Generator`` on the CPU in a
fixed key order, so it is reproducible on any box with the same torch build.

Variants
  ``init``  what ``BertPreTrainedModel.init_bert_weights`` produces (src/lxrt/modeling.py:688-699):
            Linear/Embedding ~ N(0, 0.02), biases 0, LayerNorm (1, 0).
  ``rich``  same weights plus non-trivial biases ~ N(0, 0.02) and LayerNorm gamma ~ 1 + N(0, 0.1),
            beta ~ N(0, 0.1): exercises every bias / gamma / beta path that ``init`` would hide.
``head_sigma`` rescales ``logit_fc.3.weight`` (e.g. 0.2) to emulate a trained head's logit margins.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

H, I, A_DEFAULT = 768, 3072, 1842
LXMERT_PREFIX = "lxrt_encoder.model."
UNITER_PREFIX = "encoder.model."


def _linear(p, n_out, n_in):
    return [(p + ".weight", (n_out, n_in), "w"), (p + ".bias", (n_out,), "b")]


def _ln(p, n):
    return [(p + ".weight", (n,), "g"), (p + ".bias", (n,), "beta")]


def _att(p, self_name):
    out = []
    for n in ("query", "key", "value"):
        out += _linear(f"{p}.{self_name}.{n}", H, H)
    out += _linear(p + ".output.dense", H, H) + _ln(p + ".output.LayerNorm", H)
    return out


def _bert_layer(p):
    return (_att(p + ".attention", "self") + _linear(p + ".intermediate.dense", I, H)
            + _linear(p + ".output.dense", H, I) + _ln(p + ".output.LayerNorm", H))


def _head(num_answers):
    return _linear("logit_fc.0", 2 * H, H) + _ln("logit_fc.2", 2 * H) + _linear("logit_fc.3", num_answers, 2 * H)


def lxmert_spec(l_layers=9, x_layers=5, r_layers=5, num_answers=A_DEFAULT, vocab=30522, max_pos=512):
    """[(reference key, shape, kind)] for GQAModel (src/tasks/gqa_model.py:14-28)."""
    b = LXMERT_PREFIX + "bert."
    s = [(b + "embeddings.word_embeddings.weight", (vocab, H), "w"),
         (b + "embeddings.position_embeddings.weight", (max_pos, H), "w"),
         (b + "embeddings.token_type_embeddings.weight", (2, H), "w")]
    s += _ln(b + "embeddings.LayerNorm", H)
    s += _linear(b + "encoder.visn_fc.visn_fc", H, 2048) + _ln(b + "encoder.visn_fc.visn_layer_norm", H)
    s += _linear(b + "encoder.visn_fc.box_fc", H, 4) + _ln(b + "encoder.visn_fc.box_layer_norm", H)
    for i in range(l_layers):
        s += _bert_layer(f"{b}encoder.layer.{i}")
    for i in range(x_layers):
        x = f"{b}encoder.x_layers.{i}"
        s += _att(x + ".visual_attention", "att") + _att(x + ".lang_self_att", "self") + _att(x + ".visn_self_att", "self")
        s += _linear(x + ".lang_inter.dense", I, H) + _linear(x + ".lang_output.dense", H, I) + _ln(x + ".lang_output.LayerNorm", H)
        s += _linear(x + ".visn_inter.dense", I, H) + _linear(x + ".visn_output.dense", H, I) + _ln(x + ".visn_output.LayerNorm", H)
    for i in range(r_layers):
        s += _bert_layer(f"{b}encoder.r_layers.{i}")
    s += _linear(b + "pooler.dense", H, H)
    if num_answers:
        s += _head(num_answers)
    return s


def uniter_spec(layers=12, num_answers=A_DEFAULT, vocab=28996, max_pos=512):
    """[(reference key, shape, kind)] for GQAUNITER (src/uniter/uniter.py:15-44)."""
    u = UNITER_PREFIX + "uniter."
    s = [(u + "embeddings.word_embeddings.weight", (vocab, H), "w"),
         (u + "embeddings.position_embeddings.weight", (max_pos, H), "w"),
         (u + "embeddings.token_type_embeddings.weight", (2, H), "w")]
    s += _ln(u + "embeddings.LayerNorm", H)
    s += _linear(u + "img_embeddings.img_linear", H, 2048) + _ln(u + "img_embeddings.img_layer_norm", H)
    s += _ln(u + "img_embeddings.pos_layer_norm", H) + _linear(u + "img_embeddings.pos_linear", H, 7)
    s += _ln(u + "img_embeddings.LayerNorm", H)
    for i in range(layers):
        s += _bert_layer(f"{u}encoder.layer.{i}")
    s += _linear(u + "pooler.dense", H, H)
    if num_answers:
        s += _head(num_answers)
    return s


def make_state_dict(spec, seed=9595, variant="init", head_sigma=None):
    """Draw every tensor of ``spec`` in order from one CPU generator (seed 9595 = reference default, param.py:92)."""
    if variant not in ("init", "rich"):
        raise ValueError(f"unknown variant {variant!r}")
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = OrderedDict()
    for key, shape, kind in spec:
        if kind == "w":
            t = torch.randn(shape, generator=g) * 0.02
            if head_sigma is not None and key == "logit_fc.3.weight":
                t = t * (head_sigma / 0.02)
        elif kind == "b":
            t = torch.randn(shape, generator=g) * 0.02 if variant == "rich" else torch.zeros(shape)
        elif kind == "g":
            t = 1.0 + torch.randn(shape, generator=g) * 0.1 if variant == "rich" else torch.ones(shape)
        else:
            t = torch.randn(shape, generator=g) * 0.1 if variant == "rich" else torch.zeros(shape)
        sd[key] = t
    return sd


def lxmert_state_dict(seed=9595, variant="init", head_sigma=None, **kw):
    return make_state_dict(lxmert_spec(**kw), seed, variant, head_sigma)


def uniter_state_dict(seed=9595, variant="init", head_sigma=None, **kw):
    return make_state_dict(uniter_spec(**kw), seed, variant, head_sigma)


def make_inputs(B, L=20, seed=1234, vocab=30522, num_rois=36, feat_dim=2048):
    """Synthetic batch per SURVEY.md §8(d): non-negative RoI features, boxes in [0,1], [CLS] .. [SEP] ids.

    Returns dict(feat f32 [B,36,2048], pos f32 [B,36,4], pos7 f32 [B,36,7], ids/seg/mask int64 [B,L]).
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    feat = torch.rand(B, num_rois, feat_dim, generator=g)
    x1 = torch.rand(B, num_rois, generator=g) * 0.5
    y1 = torch.rand(B, num_rois, generator=g) * 0.5
    x2 = x1 + torch.rand(B, num_rois, generator=g) * 0.5
    y2 = y1 + torch.rand(B, num_rois, generator=g) * 0.5
    pos = torch.stack([x1, y1, x2, y2], dim=-1)
    w, h = x2 - x1, y2 - y1
    pos7 = torch.stack([y1, x1, y2, x2, w, h, w * h], dim=-1)  # gqa_data.py:240-250 ordering
    lens = torch.randint(5, L + 1, (B,), generator=g)
    ids = torch.randint(1000, vocab, (B, L), generator=g)
    ar = torch.arange(L)[None, :]
    mask = (ar < lens[:, None]).long()
    ids = ids * mask
    ids[:, 0] = 101
    ids[torch.arange(B), lens - 1] = 102
    seg = torch.zeros(B, L, dtype=torch.long)
    return {"feat": feat, "pos": pos, "pos7": pos7, "ids": ids, "seg": seg, "mask": mask, "lens": lens}


def maha_stats(seed=4321, num_answers=A_DEFAULT, hidden=H, n=4096):
    """Class means and a well-conditioned precision matrix; fp32 like gqa_maha.py:199-201."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    mu = torch.randn(num_answers, hidden, generator=g) * 0.1
    Z = torch.randn(n, hidden, generator=g, dtype=torch.float64)
    P = torch.linalg.inv(Z.t() @ Z / n + 0.1 * torch.eye(hidden, dtype=torch.float64))
    P = 0.5 * (P + P.t())
    return mu, P.float()
