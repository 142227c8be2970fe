"""d loss / d (feats, boxes) through the LXMERT GQA model — the ``loss.backward()`` of ODIN
(``src/tasks/gqa_odin.py:97-131``) and of the Mahalanobis input pre-processing (``src/tasks/gqa_maha.py:238-268``),
SURVEY §8(f) rank 1.  fp32 path; every operator is a kernel of this package called through the C ABI
(``rgqa_b200.ops``): the forward GEMM / attention / LayerNorm kernels of the precise path for the taped forward, the
dgrad kernels of ``csrc/backward.cu`` plus the same fp32 GEMM against transposed weights for the backward.  PyTorch is
used for tensor storage, views and copies only.

No weight gradients are formed (the reference calls ``backward()`` on the whole graph but only reads ``feats.grad`` and
``boxes.grad``), and the nine language-only layers are not differentiated at all: the question tokens do not depend on
the image.  Structure of the graph: ``src/lxrt/modeling.py`` — VisualFeatEncoder :507-517, BertLayer :425-429,
LXRTXLayer :455-488 (ONE shared cross-attention module serving both directions), BertPooler :575-581, answer head
``src/tasks/gqa_model.py:22-28``.
"""
from __future__ import annotations

import torch

from .. import _lib as L
from .. import ops

P = "lxrt_encoder.model.bert."


class _Lin:
    def __init__(self, w, b):
        self.w, self.b = w.contiguous(), b.contiguous()
        self.wt = w.t().contiguous()          # [K, N]: the dgrad GEMM  dX = dY · W  is  linear_f32(dY, W^T)
        self._b16 = {}

    def b16(self, transposed):
        """bf16 copy of w (or of its transpose) for the tensor-core mode, made on first use."""
        if transposed not in self._b16:
            self._b16[transposed] = ops.cast_bf16(self.wt if transposed else self.w)
        return self._b16[transposed]


class _Att:
    """BertAttention + BertAttOutput parameters; Q|K|V stacked like the engine does."""

    def __init__(self, sd, prefix, self_name):
        s = prefix + "." + self_name
        self.qkv = _Lin(torch.cat([sd[s + ".query.weight"], sd[s + ".key.weight"], sd[s + ".value.weight"]], 0),
                        torch.cat([sd[s + ".query.bias"], sd[s + ".key.bias"], sd[s + ".value.bias"]], 0))
        self.out = _Lin(sd[prefix + ".output.dense.weight"], sd[prefix + ".output.dense.bias"])
        self.g, self.beta = sd[prefix + ".output.LayerNorm.weight"].contiguous(), sd[prefix + ".output.LayerNorm.bias"].contiguous()


class _Ffn:
    def __init__(self, sd, inter, outp):
        self.inter = _Lin(sd[inter + ".dense.weight"], sd[inter + ".dense.bias"])
        self.out = _Lin(sd[outp + ".dense.weight"], sd[outp + ".dense.bias"])
        self.g, self.beta = sd[outp + ".LayerNorm.weight"].contiguous(), sd[outp + ".LayerNorm.bias"].contiguous()


class LXMERTInputGrad:
    """Taped fp32 forward + input-gradient backward.  ``state_dict``: the reference GQAModel's keys (fp32)."""

    def __init__(self, state_dict, l_layers=9, x_layers=5, r_layers=5, heads=12, eps=1e-12, device="cuda:0", gemm="fp32"):
        """``gemm="bf16"``: the encoder's matrix products (forward and dgrad) run on the tcgen05 tensor cores with bf16
        operands cast on the fly, fp32 accumulation and fp32 activations / gradients everywhere else (mixed precision)."""
        assert gemm in ("fp32", "bf16")
        self.gemm = gemm
        dev = torch.device(device)
        sd = {k: v.detach().to(dev, torch.float32) for k, v in state_dict.items()}
        self.dev, self.heads, self.eps = dev, heads, eps
        e = P + "embeddings."
        self.word, self.posemb, self.typeemb = (sd[e + n + ".weight"].contiguous() for n in
                                                ("word_embeddings", "position_embeddings", "token_type_embeddings"))
        self.emb_g, self.emb_b = sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"]
        v = P + "encoder.visn_fc."
        self.visn_fc = _Lin(sd[v + "visn_fc.weight"], sd[v + "visn_fc.bias"])
        self.box_fc = _Lin(sd[v + "box_fc.weight"], sd[v + "box_fc.bias"])
        self.visn_g, self.visn_b = sd[v + "visn_layer_norm.weight"].contiguous(), sd[v + "visn_layer_norm.bias"].contiguous()
        self.box_g, self.box_b = sd[v + "box_layer_norm.weight"].contiguous(), sd[v + "box_layer_norm.bias"].contiguous()
        enc = P + "encoder."
        self.l = [(_Att(sd, f"{enc}layer.{i}.attention", "self"), _Ffn(sd, f"{enc}layer.{i}.intermediate", f"{enc}layer.{i}.output"))
                  for i in range(l_layers)]
        self.r = [(_Att(sd, f"{enc}r_layers.{i}.attention", "self"), _Ffn(sd, f"{enc}r_layers.{i}.intermediate", f"{enc}r_layers.{i}.output"))
                  for i in range(r_layers)]
        self.x = []
        for i in range(x_layers):
            x = f"{enc}x_layers.{i}."
            self.x.append(dict(cross=_Att(sd, x + "visual_attention", "att"), ls=_Att(sd, x + "lang_self_att", "self"),
                               vs=_Att(sd, x + "visn_self_att", "self"), lf=_Ffn(sd, x + "lang_inter", x + "lang_output"),
                               vf=_Ffn(sd, x + "visn_inter", x + "visn_output")))
        self.pooler = _Lin(sd[P + "pooler.dense.weight"], sd[P + "pooler.dense.bias"])
        self.fc0 = _Lin(sd["logit_fc.0.weight"], sd["logit_fc.0.bias"])
        self.fc_g, self.fc_b = sd["logit_fc.2.weight"].contiguous(), sd["logit_fc.2.bias"].contiguous()
        self.fc3 = _Lin(sd["logit_fc.3.weight"], sd["logit_fc.3.bias"])
        self.tape = None

    def _mm(self, x, lin, transposed=False, epilogue=L.EPI_BIAS, residual=None):
        """epilogue(x · Wᵀ + b [+ residual]) (forward) or x · W [+ residual] (``transposed``: the dgrad product)."""
        w = lin.wt if transposed else lin.w
        b = None if transposed else lin.b
        if self.gemm == "bf16" and w.shape[1] % 64 == 0 and x.shape[0] >= 64:
            return ops.linear_bf16(ops.cast_bf16(x.contiguous()), lin.b16(transposed), b, epilogue=epilogue, residual=residual,
                                   out_dtype=torch.float32)
        return ops.linear_f32(x, w, b, epilogue=epilogue, residual=residual)

    # ---- forward blocks (each returns its output and what its backward needs) ----
    def _att_fwd(self, w, x_q, x_kv, B, Lq, Lk, mask, qkv_q=None, qkv_kv=None):
        qkv_q = self._mm(x_q, w.qkv) if qkv_q is None else qkv_q
        qkv_kv = qkv_q if x_kv is x_q and qkv_kv is None else (self._mm(x_kv, w.qkv) if qkv_kv is None else qkv_kv)
        ctx = ops.attention(qkv_q, qkv_kv, B, Lq, Lk, key_mask=mask, heads=self.heads)
        pre = self._mm(ctx, w.out, epilogue=L.EPI_RESID, residual=x_q)
        y = ops.layernorm(pre, w.g, w.beta, self.eps)
        return y, dict(qkv_q=qkv_q, qkv_kv=qkv_kv, pre=pre, B=B, Lq=Lq, Lk=Lk, mask=mask)

    def _ffn_fwd(self, w, x):
        pg = self._mm(x, w.inter)
        inter = ops.gelu_fwd(pg)
        pre = self._mm(inter, w.out, epilogue=L.EPI_RESID, residual=x)
        return ops.layernorm(pre, w.g, w.beta, self.eps), dict(pg=pg, pre=pre)

    def forward(self, feat, pos, ids, seg, mask):
        """feat [B,36,2048] f32, pos [B,36,4] f32, ids / seg / mask [B,L] int -> (logits [B,A], pooled [B,768]); tapes."""
        dev = self.dev
        B, V = feat.shape[0], feat.shape[1]
        Lt = ids.shape[1]
        ids32, mask32 = ids.to(dev, torch.int32).contiguous(), mask.to(dev, torch.int32).contiguous()
        seg32 = torch.zeros_like(ids32) if seg is None else seg.to(dev, torch.int32).contiguous()
        feat2 = feat.to(dev, torch.float32).reshape(B * V, -1).contiguous()
        pos2 = pos.to(dev, torch.float32).reshape(B * V, -1).contiguous()
        t = dict(B=B, V=V, L=Lt)
        lang = ops.embed_ln(ids32, seg32, self.word, self.posemb, self.typeemb, self.emb_g, self.emb_b, self.eps)
        t["pre_v"] = self._mm(feat2, self.visn_fc)
        t["pre_b"] = ops.linear_f32(pos2, self.box_fc.w, self.box_fc.b)
        visn = ops.visn_combine(t["pre_v"], pos2, self.box_fc.w, self.box_fc.b, self.visn_g, self.visn_b, self.box_g, self.box_b, self.eps)
        for a, f in self.l:                                   # language only: no tape needed
            lang, _ = self._att_fwd(a, lang, lang, B, Lt, Lt, mask32)
            lang, _ = self._ffn_fwd(f, lang)
        t["r"] = []
        for a, f in self.r:
            visn, ta = self._att_fwd(a, visn, visn, B, V, V, None)
            visn, tf = self._ffn_fwd(f, visn)
            t["r"].append((ta, tf))
        t["x"] = []
        for w in self.x:
            c = w["cross"]
            qkv_l = self._mm(lang, c.qkv)
            qkv_v = self._mm(visn, c.qkv)
            l1, tcl = self._att_fwd(c, lang, visn, B, Lt, V, None, qkv_q=qkv_l, qkv_kv=qkv_v)        # lang queries, visual keys
            v1, tcv = self._att_fwd(c, visn, lang, B, V, Lt, mask32, qkv_q=qkv_v, qkv_kv=qkv_l)      # visual queries, lang keys
            l2, tls = self._att_fwd(w["ls"], l1, l1, B, Lt, Lt, mask32)
            v2, tvs = self._att_fwd(w["vs"], v1, v1, B, V, V, None)
            lang, tlf = self._ffn_fwd(w["lf"], l2)
            visn, tvf = self._ffn_fwd(w["vf"], v2)
            t["x"].append(dict(cl=tcl, cv=tcv, ls=tls, vs=tvs, lf=tlf, vf=tvf))
        first = lang.view(B, Lt, -1)[:, 0, :].contiguous()
        pooled = ops.linear_f32(first, self.pooler.w, self.pooler.b, epilogue=L.EPI_TANH)
        t["h0"] = ops.linear_f32(pooled, self.fc0.w, self.fc0.b)
        t["g"] = ops.gelu_fwd(t["h0"])
        n = ops.layernorm(t["g"], self.fc_g, self.fc_b, self.eps)
        logits = ops.linear_f32(n, self.fc3.w, self.fc3.b)
        t["pooled"] = pooled
        self.tape = t
        return logits, pooled

    # ---- backward blocks ----
    def _ffn_bwd(self, w, tp, dy):
        d_pre = ops.layernorm_bwd(dy, tp["pre"], w.g, eps=self.eps)
        d_inter = self._mm(d_pre, w.out, transposed=True)
        d_pg = ops.gelu_bwd(d_inter, tp["pg"])
        return self._mm(d_pg, w.inter, transposed=True, epilogue=L.EPI_RESID, residual=d_pre)

    def _att_bwd(self, w, tp, dy, dqkv_q, dqkv_kv):
        """Gradient of one attention block w.r.t. its query-side input through the residual and the output projection;
        dQ / dK / dV land in the given fused gradient buffers (the caller turns those into input gradients)."""
        d_pre = ops.layernorm_bwd(dy, tp["pre"], w.g, eps=self.eps)
        d_ctx = self._mm(d_pre, w.out, transposed=True)
        ops.attention_bwd(tp["qkv_q"], tp["qkv_kv"], d_ctx, dqkv_q, dqkv_kv, tp["B"], tp["Lq"], tp["Lk"], key_mask=tp["mask"], heads=self.heads,
                          bf16=self.gemm == "bf16")      # mixed precision: the attention products on the tensor cores as well
        return d_pre

    def _self_att_bwd(self, w, tp, dy):
        dqkv = torch.empty_like(tp["qkv_q"])
        d_pre = self._att_bwd(w, tp, dy, dqkv, dqkv)
        return self._mm(dqkv, w.qkv, transposed=True, epilogue=L.EPI_RESID, residual=d_pre)

    def backward(self, dlogits=None, dpooled=None):
        """(d loss / d feat [B,36,2048], d loss / d pos [B,36,4]) for the last forward, given d loss / d logits and / or
        d loss / d pooled."""
        t = self.tape
        assert t is not None, "backward() needs a forward() first"
        B, V, Lt = t["B"], t["V"], t["L"]
        H = self.pooler.w.shape[0]
        dp = None if dpooled is None else dpooled.to(self.dev, torch.float32).contiguous()
        if dlogits is not None:
            dn = ops.linear_f32(dlogits.to(self.dev, torch.float32).contiguous(), self.fc3.wt)
            dg = ops.layernorm_bwd(dn, t["g"], self.fc_g, eps=self.eps)
            dh0 = ops.gelu_bwd(dg, t["h0"])
            dp = ops.linear_f32(dh0, self.fc0.wt) if dp is None else ops.linear_f32(dh0, self.fc0.wt, epilogue=L.EPI_RESID, residual=dp)
        assert dp is not None, "nothing to differentiate"
        d_first = ops.linear_f32(ops.tanh_bwd(dp, t["pooled"]), self.pooler.wt)
        dl = torch.zeros((B, Lt, H), dtype=torch.float32, device=self.dev)
        dl[:, 0, :] = d_first                                   # BertPooler reads token 0 only (:578)
        dl = dl.view(B * Lt, H)
        dv = torch.zeros((B * V, H), dtype=torch.float32, device=self.dev)
        for w, tp in zip(reversed(self.x), reversed(t["x"])):
            dl2 = self._ffn_bwd(w["lf"], tp["lf"], dl)
            dv2 = self._ffn_bwd(w["vf"], tp["vf"], dv)
            dl1 = self._self_att_bwd(w["ls"], tp["ls"], dl2)
            dv1 = self._self_att_bwd(w["vs"], tp["vs"], dv2)
            c = w["cross"]
            dqkv_l = torch.empty_like(tp["cl"]["qkv_q"])        # [dQ_l | dK_l | dV_l]: dQ from direction 1, dK/dV from direction 2
            dqkv_v = torch.empty_like(tp["cv"]["qkv_q"])
            d_pre_l = self._att_bwd(c, tp["cl"], dl1, dqkv_l, dqkv_v)
            d_pre_v = self._att_bwd(c, tp["cv"], dv1, dqkv_v, dqkv_l)
            dl = self._mm(dqkv_l, c.qkv, transposed=True, epilogue=L.EPI_RESID, residual=d_pre_l)
            dv = self._mm(dqkv_v, c.qkv, transposed=True, epilogue=L.EPI_RESID, residual=d_pre_v)
        for (a, f), (ta, tf) in zip(reversed(self.r), reversed(t["r"])):
            dv = self._self_att_bwd(a, ta, self._ffn_bwd(f, tf, dv))
        # VisualFeatEncoder: output = (LN(visn_fc(feat)) + LN(box_fc(pos))) / 2
        d_pre_v = ops.layernorm_bwd(dv, t["pre_v"], self.visn_g, eps=self.eps, scale=0.5)
        d_pre_b = ops.layernorm_bwd(dv, t["pre_b"], self.box_g, eps=self.eps, scale=0.5)
        dfeat = self._mm(d_pre_v, self.visn_fc, transposed=True).view(B, V, -1)
        dpos = ops.linear_f32(d_pre_b, self.box_fc.wt).view(B, V, -1)
        return dfeat, dpos
