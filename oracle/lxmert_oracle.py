"""ORACLE — test infrastructure, not product code.

CPU (fp32, torch functional ops on the host) restatement of the reference's LXMERT / UNITER
evaluation forward, written from the reference's algorithm with the file:line each function follows
(paths relative to chihhuiho/RGQA).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``rgqa_b200/`` does.

Pinned: ``tests/test_oracle_cpu.py`` checks this restatement against golden vectors produced by the
UNMODIFIED reference modules run in the build container (``tests/golden/make_golden.py`` imports
``/root/reference/src/lxrt/modeling.py`` etc., loads the same seeded state dict, dumps pooled/logits).

The oracle consumes a state dict with the reference's key names (with or without the
``lxrt_encoder.model.`` / ``encoder.model.`` prefixes) so it sees exactly the tensors the CUDA handle does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # BertLayerNorm(..., eps=1e-12) everywhere (modeling.py:275,354,409,500,504; gqa_model.py:25)
HEADS = 12


def _sd(sd, prefix):
    """View of the state dict with ``prefix`` stripped (and a DataParallel ``module.`` tolerated, gqa_conf.py:378-380)."""
    out = {}
    for k, v in sd.items():
        k2 = k
        for p in ("module.", prefix, "module."):
            if k2.startswith(p):
                k2 = k2[len(p):]
        out[k2] = v.float()
    return out


def gelu(x):
    """src/lxrt/modeling.py:112-118 — erf form, not the tanh approximation."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps=LN_EPS)


def embeddings(sd, p, input_ids, token_type_ids):
    """BertEmbeddings.forward, src/lxrt/modeling.py:278-292 (dropout is identity in eval)."""
    L = input_ids.shape[1]
    position_ids = torch.arange(L, dtype=torch.long).unsqueeze(0).expand_as(input_ids)
    e = (sd[p + ".word_embeddings.weight"][input_ids] + sd[p + ".position_embeddings.weight"][position_ids]
         + sd[p + ".token_type_embeddings.weight"][token_type_ids])
    return layer_norm(sd, p + ".LayerNorm", e)


def visual_feat_encoder(sd, p, feats, boxes):
    """VisualFeatEncoder.forward, src/lxrt/modeling.py:507-517."""
    x = layer_norm(sd, p + ".visn_layer_norm", linear(sd, p + ".visn_fc", feats))
    y = layer_norm(sd, p + ".box_layer_norm", linear(sd, p + ".box_fc", boxes))
    return (x + y) / 2


def bert_attention(sd, p, hidden, context, ext_mask):
    """BertAttention.forward, src/lxrt/modeling.py:320-347; ext_mask is additive [B,1,1,Lk] or None."""
    B, Lq, Hd = hidden.shape
    Lk = context.shape[1]
    d = Hd // HEADS

    def split(x, n):  # transpose_for_scores, :315-318
        return x.view(B, n, HEADS, d).permute(0, 2, 1, 3)

    q = split(linear(sd, p + ".query", hidden), Lq)
    k = split(linear(sd, p + ".key", context), Lk)
    v = split(linear(sd, p + ".value", context), Lk)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    if ext_mask is not None:
        scores = scores + ext_mask
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    return ctx.view(B, Lq, Hd)


def att_output(sd, p, hidden, input_tensor):
    """BertAttOutput.forward :357-361 and BertOutput.forward :411-415 (same shape of computation)."""
    return layer_norm(sd, p + ".LayerNorm", linear(sd, p + ".dense", hidden) + input_tensor)


def self_att_layer(sd, p, x, ext_mask, self_name="self"):
    """BertSelfattLayer.forward, :383-386."""
    return att_output(sd, p + ".output", bert_attention(sd, p + "." + self_name, x, x, ext_mask), x)


def cross_att_layer(sd, p, x, ctx, ctx_mask):
    """BertCrossattLayer.forward, :370-373."""
    return att_output(sd, p + ".output", bert_attention(sd, p + ".att", x, ctx, ctx_mask), x)


def ffn(sd, inter_p, out_p, x):
    """BertIntermediate :398-401 + BertOutput :411-415."""
    return att_output(sd, out_p, gelu(linear(sd, inter_p + ".dense", x)), x)


def bert_layer(sd, p, x, ext_mask):
    """BertLayer.forward, :425-429."""
    a = self_att_layer(sd, p + ".attention", x, ext_mask)
    return ffn(sd, p + ".intermediate", p + ".output", a)


def x_layer(sd, p, lang, lang_mask, visn, visn_mask):
    """LXRTXLayer.forward, :477-488: cross (shared module, both directions from the same inputs :455-459),
    then per-stream self attention (:461-465), then per-stream FFN (:467-475)."""
    lang_c = cross_att_layer(sd, p + ".visual_attention", lang, visn, visn_mask)
    visn_c = cross_att_layer(sd, p + ".visual_attention", visn, lang, lang_mask)
    lang_s = self_att_layer(sd, p + ".lang_self_att", lang_c, lang_mask)
    visn_s = self_att_layer(sd, p + ".visn_self_att", visn_c, visn_mask)
    lang_o = ffn(sd, p + ".lang_inter", p + ".lang_output", lang_s)
    visn_o = ffn(sd, p + ".visn_inter", p + ".visn_output", visn_s)
    return lang_o, visn_o


def count_layers(sd, stem):
    n = 0
    while f"{stem}.{n}.attention.self.query.weight" in sd or f"{stem}.{n}.visual_attention.att.query.weight" in sd:
        n += 1
    return n


def lxmert_encode(state_dict, feat, pos, input_ids, token_type_ids, attention_mask, return_trace=False):
    """LXRTFeatureExtraction.forward in mode 'x' (:1020-1026) = LXRTModel.forward (:845-886).

    feat [B,36,2048] f32, pos [B,36,4] f32, ids/seg/mask [B,L] int64.  Returns pooled [B,768]
    (and, with return_trace, a dict of intermediate activations for debugging the CUDA path)."""
    sd = _sd(state_dict, "lxrt_encoder.model.")
    ext = (1.0 - attention_mask[:, None, None, :].float()) * -10000.0  # :857-865
    visn_mask = None                                                    # entry.py:109 passes no visual mask
    trace = {}
    lang = embeddings(sd, "bert.embeddings", input_ids, token_type_ids)  # :876
    visn = visual_feat_encoder(sd, "bert.encoder.visn_fc", feat, pos)    # :551
    trace["emb"], trace["visn_fc"] = lang, visn
    for i in range(count_layers(sd, "bert.encoder.layer")):              # :553-554
        lang = bert_layer(sd, f"bert.encoder.layer.{i}", lang, ext)
    for i in range(count_layers(sd, "bert.encoder.r_layers")):           # :557-558
        visn = bert_layer(sd, f"bert.encoder.r_layers.{i}", visn, visn_mask)
    trace["lang_l"], trace["visn_r"] = lang, visn
    for i in range(count_layers(sd, "bert.encoder.x_layers")):           # :561-563
        lang, visn = x_layer(sd, f"bert.encoder.x_layers.{i}", lang, ext, visn, visn_mask)
        trace[f"lang_x{i}"], trace[f"visn_x{i}"] = lang, visn
    pooled = torch.tanh(linear(sd, "bert.pooler.dense", lang[:, 0]))     # BertPooler :575-581
    return (pooled, trace) if return_trace else pooled


def gqa_head(state_dict, x):
    """GQAModel.logit_fc, src/tasks/gqa_model.py:22-28: Linear -> GeLU -> LayerNorm(1536) -> Linear."""
    sd = _sd(state_dict, "lxrt_encoder.model.")
    h = layer_norm(sd, "logit_fc.2", gelu(linear(sd, "logit_fc.0", x)))
    return linear(sd, "logit_fc.3", h)


def gqa_model_forward(state_dict, feat, pos, input_ids, token_type_ids, attention_mask):
    """GQAModel_maha.forward (gqa_model.py:62-75) on pre-tokenised input: returns (logit, pooled x)."""
    with torch.no_grad():
        x = lxmert_encode(state_dict, feat, pos, input_ids, token_type_ids, attention_mask)
        return gqa_head(state_dict, x), x


def uniter_encode(state_dict, feat, pos7, input_ids, token_type_ids, attention_mask):
    """UniterModel.forward, src/uniter/modeling.py:623-635, driven as UniterEncoder.forward does
    (src/uniter/entry.py:83-99: visual mask all ones, visual token type all ones)."""
    sd = _sd(state_dict, "encoder.model.")
    B, V = feat.shape[0], feat.shape[1]
    mask = torch.cat([attention_mask, torch.ones(B, V, dtype=attention_mask.dtype)], dim=1)
    ext = (1.0 - mask[:, None, None, :].float()) * -10000.0
    txt = embeddings(sd, "uniter.embeddings", input_ids, token_type_ids)
    type_emb = sd["uniter.embeddings.token_type_embeddings.weight"][torch.ones(B, V, dtype=torch.long)]
    p = "uniter.img_embeddings"
    im = layer_norm(sd, p + ".img_layer_norm", linear(sd, p + ".img_linear", feat))    # :607
    ps = layer_norm(sd, p + ".pos_layer_norm", linear(sd, p + ".pos_linear", pos7))    # :608
    img = layer_norm(sd, p + ".LayerNorm", im + ps + type_emb)                          # :609-610
    h = torch.cat([txt, img], dim=1)
    for i in range(count_layers(sd, "uniter.encoder.layer")):
        h = bert_layer(sd, f"uniter.encoder.layer.{i}", h, ext)
    return torch.tanh(linear(sd, "uniter.pooler.dense", h[:, 0]))


def gqa_uniter_forward(state_dict, feat, pos7, input_ids, token_type_ids, attention_mask):
    """GQAUNITER_maha-style forward (src/uniter/uniter.py:32-44): (logit, pooled)."""
    with torch.no_grad():
        x = uniter_encode(state_dict, feat, pos7, input_ids, token_type_ids, attention_mask)
        sd = _sd(state_dict, "encoder.model.")
        h = layer_norm(sd, "logit_fc.2", gelu(linear(sd, "logit_fc.0", x)))
        return linear(sd, "logit_fc.3", h), x
