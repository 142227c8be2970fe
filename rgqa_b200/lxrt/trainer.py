"""One RP-finetune optimisation step of the LXMERT GQA model on the GPU — ``GQA.train``'s loop body
(``src/tasks/gqa_conf.py:148-198``): forward, ``BCEWithLogitsLoss × A``, ``backward()``, ``clip_grad_norm_(…, 5.)``,
``BertAdam.step()`` — SURVEY §8 row a22 / config 5.  fp32 path (master weights, activations, gradients all fp32 on the CUDA
cores); every operator is a kernel of this package behind the C ABI, PyTorch holds the tensors.

* parameters, gradients and both Adam moments live in four FLAT fp32 buffers (the reference's single parameter group,
  ``gqa_conf.py:123-127``): zeroing the gradients is one memset, the global norm one reduction, the optimiser ONE kernel
  launch; data parallelism = one NCCL all-reduce (mean over ranks, like DDP) per finished layer's contiguous slice of the flat
  gradient, issued during the backward and overlapped with it (``buckets``), the optimiser waiting only for the last one;
* backward = the dgrad chain of ``input_grad.py`` extended down through the nine language layers and the embeddings, plus the
  weight gradients (``rgqa_gemm_f32_tn``, ``rgqa_colsum``, ``rgqa_ln_param_grad``, ``rgqa_embedding_scatter``), accumulated
  straight into the flat gradient buffer under the reference's parameter names;
* ``dropout=0`` (default) is the "dropout off" configuration SURVEY §8(d) config 5 asks parity for: the golden trajectory comes
  from the UNMODIFIED reference modules + the reference's own BertAdam with both dropout probabilities set to 0.  ``dropout=0.1``
  is the reference's training configuration with this package's counter-based masks (regenerated in the backward, never stored;
  not torch's random stream, so parity under dropout is statistical).
"""
from __future__ import annotations

import gc

import torch
import torch.distributed as dist

from .. import _lib as L
from .. import ops
from .optimization import SCHEDULES

P = "lxrt_encoder.model.bert."


class _Att:
    def __init__(self, prefix, self_name):
        s = prefix + "." + self_name
        self.q, self.k, self.v = s + ".query", s + ".key", s + ".value"
        self.out, self.ln = prefix + ".output.dense", prefix + ".output.LayerNorm"


class _Ffn:
    def __init__(self, inter, outp):
        self.inter, self.out, self.ln = inter + ".dense", outp + ".dense", outp + ".LayerNorm"


class LXMERTTrainer:
    def __init__(self, state_dict, l_layers=9, x_layers=5, r_layers=5, heads=12, eps=1e-12, device="cuda:0", lr=5e-5, warmup=0.1,
                 t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01, clip_norm=5.0, gemm="fp32",
                 dropout=0.0, seed=9595, bucketed=True, use_graph=False):
        """``gemm="bf16"``: the large matrix products (forward, dgrad and wgrad of every nn.Linear with K a multiple of 64)
        run on the tcgen05 tensor cores with bf16 operands cast on the fly and fp32 accumulation / outputs; master weights,
        activations, gradients and the optimiser stay fp32 (mixed precision; not the reference's arithmetic — parity with
        the fp32 trajectory holds to ~1e-2, see tests/test_train_gpu.py)."""
        assert gemm in ("fp32", "bf16") and 0.0 <= dropout < 1.0
        dev = torch.device(device)
        self.gemm = gemm
        self.bucketed = bucketed   # False: the single blocking all-reduce after the backward (A/B of the overlap)
        self.use_graph = use_graph  # capture the gradient pass of each input shape into a CUDA graph (dropout off only)
        self._graphs = {}
        self._capturing = False
        # dropout > 0: the reference's training-mode dropout sites (embeddings, attention probabilities, the two output
        # projections of every block, the visual feature encoder) with this package's counter-based masks (not torch's)
        self.dropout, self.seed, self._site = dropout, seed, 0
        self.dev, self.heads, self.eps, self.clip_norm = dev, heads, eps, clip_norm
        self.hp = dict(lr=lr, warmup=warmup, t_total=t_total, schedule=schedule, b1=b1, b2=b2, e=e, weight_decay=weight_decay)
        self.step_count = 0
        names = list(state_dict.keys())
        sizes = [state_dict[n].numel() for n in names]
        pad = lambda n: (n + 63) // 64 * 64            # every tensor starts 256-byte aligned (TMA operands / vector stores)
        total = sum(pad(sz) for sz in sizes)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.p, self.g = {}, {}
        self._span = {}                                  # name -> (offset, padded size) inside the flat buffers
        off = 0
        for n, sz in zip(names, sizes):
            self.p[n] = self.flat_p[off:off + sz].view(state_dict[n].shape)
            self.g[n] = self.flat_g[off:off + sz].view(state_dict[n].shape)
            self.p[n].copy_(state_dict[n].detach().to(dev, torch.float32))
            self._span[n] = (off, pad(sz))
            off += pad(sz)
        enc = P + "encoder."
        self.l = [(_Att(f"{enc}layer.{i}.attention", "self"), _Ffn(f"{enc}layer.{i}.intermediate", f"{enc}layer.{i}.output")) for i in range(l_layers)]
        self.r = [(_Att(f"{enc}r_layers.{i}.attention", "self"), _Ffn(f"{enc}r_layers.{i}.intermediate", f"{enc}r_layers.{i}.output"))
                  for i in range(r_layers)]
        self.x = [dict(cross=_Att(f"{enc}x_layers.{i}.visual_attention", "att"), ls=_Att(f"{enc}x_layers.{i}.lang_self_att", "self"),
                       vs=_Att(f"{enc}x_layers.{i}.visn_self_att", "self"),
                       lf=_Ffn(f"{enc}x_layers.{i}.lang_inter", f"{enc}x_layers.{i}.lang_output"),
                       vf=_Ffn(f"{enc}x_layers.{i}.visn_inter", f"{enc}x_layers.{i}.visn_output")) for i in range(x_layers)]
        self._derived = {}
        self._a16 = {}          # per-step bf16 copies of activations / gradients (_act16)
        self.tape = None
        # Gradient buckets in the order the backward FINISHES them (data parallelism): answer head + pooler, the
        # cross-modality layers last to first, the object-relational and language layers last to first, then the visual feature
        # encoder + embeddings.  Each bucket is one contiguous slice of the flat gradient (the state dict groups a layer's tensors
        # together), all-reduced on NCCL's own stream the moment its last weight gradient is enqueued, i.e. overlapped with the rest
        # of the backward; only the last bucket's collective is exposed.
        groups = [["logit_fc.", P + "pooler."]] + [[f"{enc}x_layers.{i}."] for i in reversed(range(x_layers))] + \
                 [[f"{enc}r_layers.{i}."] for i in reversed(range(r_layers))] + [[f"{enc}layer.{i}."] for i in reversed(range(l_layers))] + \
                 [[enc + "visn_fc.", P + "embeddings."]]
        self.buckets = []
        owner = {}
        for gi, prefixes in enumerate(groups):
            mine = [n for n in names if any(n.startswith(px) for px in prefixes)]
            for n in mine:
                owner[n] = gi
            spans = sorted(self._span[n] for n in mine)
            # split into maximal contiguous runs (head and pooler, or visn_fc and the embeddings, need not be neighbours)
            runs, cur = [], None
            for o, ln in spans:
                if cur is not None and o == cur[1]:
                    cur[1] = o + ln
                else:
                    cur = [o, o + ln]
                    runs.append(cur)
            self.buckets.append([(lo, hi) for lo, hi in runs])
        assert set(owner) == set(names), sorted(set(names) - set(owner))[:5]
        assert sum(hi - lo for b in self.buckets for lo, hi in b) == total
        self._works, self._overlap = [], False
        self.comm_stats = {"buckets": len(self.buckets), "bytes": 4 * total}
        # CUDA-graph mode: the gradient pass is captured as THREE graphs cut at bucket boundaries (after the cross-modality layers,
        # and half-way down the language stack), so that the all-reduce of a finished segment's gradients runs on NCCL's stream
        # while the next segment replays; only the last segment's collective (~28 % of the bytes) is exposed.
        self._graph_cuts = sorted({x_layers, x_layers + r_layers + l_layers // 2} - {len(self.buckets) - 1}) if len(self.buckets) > 2 else []
        self._capture = None

    # ---- parameter access: stacked Q|K|V and transposed copies are rebuilt lazily after every optimiser step ----
    def _w(self, name):
        return self.p[name + ".weight"], self.p[name + ".bias"]

    def _wt(self, name):
        key = ("t", name)
        if key not in self._derived:
            self._derived[key] = self.p[name + ".weight"].t().contiguous()
        return self._derived[key]

    def _qkv(self, a):
        key = ("qkv", a.q)
        if key not in self._derived:
            w = torch.cat([self.p[a.q + ".weight"], self.p[a.k + ".weight"], self.p[a.v + ".weight"]], 0)
            b = torch.cat([self.p[a.q + ".bias"], self.p[a.k + ".bias"], self.p[a.v + ".bias"]], 0)
            self._derived[key] = (w, b, w.t().contiguous())
        return self._derived[key]

    # ---- matrix products: fp32 CUDA cores, or bf16 tensor cores with fp32 in / out ----
    def _b16(self, key, make):
        if key not in self._derived:
            self._derived[key] = ops.cast_bf16(make().contiguous())
        return self._derived[key]

    def _ln(self, x, name):
        """LayerNorm(x); in mixed precision the kernel also leaves the bf16 copy the next GEMM wants (no cast kernel for it)."""
        if self.gemm == "bf16" and x.shape[1] in (768, 1536):
            y32, y16 = ops.layernorm_dual(x, *self._w(name), self.eps)
            self._a16[(y32.data_ptr(), tuple(y32.shape), tuple(y32.stride()))] = (y32, y16)
            return y32
        return ops.layernorm(x, *self._w(name), self.eps)

    @staticmethod
    def _castable(t):
        return t.dim() == 2 and t.stride(1) == 1 and t.shape[1] % 8 == 0 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0

    def _act16(self, t, colsum_into=None):
        """bf16 copy of an fp32 activation or gradient, made ONCE per step: the forward's GEMM operand is the weight gradient's, a
        layer's dy feeds both its dgrad and its wgrad.  ``colsum_into``: the column sums of t (the bias gradient) ride on the cast.
        The source tensor is kept alive with its copy, so that a recycled address can never alias a stale entry."""
        if t.dtype == torch.bfloat16:          # already a GEMM operand (bf16 attention context / GELU output of the mixed-precision step)
            assert colsum_into is None
            return t
        key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()))
        hit = self._a16.get(key)
        if hit is not None:
            if colsum_into is not None:
                ops.linear_wgrad_bias_(colsum_into, t)
            return hit[1]
        if self._castable(t):
            out = ops.cast_bf16_colsum(t, colsum_into)
        else:
            out = ops.cast_bf16(t.contiguous())
            if colsum_into is not None:
                ops.linear_wgrad_bias_(colsum_into, t)
        self._a16[key] = (t, out)
        return out

    def _mm(self, x, w, b, epilogue=L.EPI_BIAS, residual=None, wkey=None):
        """epilogue(x · wᵀ + b [+ residual]) with w [N, K] fp32 (``wkey`` names its cached bf16 copy)."""
        if self.gemm == "bf16" and w.shape[1] % 64 == 0 and x.shape[0] >= 64:
            wb = self._b16(("b16", wkey), lambda: w)
            return ops.linear_bf16(self._act16(x), wb, b, epilogue=epilogue, residual=residual, out_dtype=torch.float32)
        return ops.linear_f32(x, w, b, epilogue=epilogue, residual=residual)

    def _tn_ok(self, dy, x):
        """dW += dyᵀ · x straight from the bf16 rows (tensor cores read both operands MN-major): needs 64-row multiples and 16-byte rows."""
        return (self.gemm == "bf16" and dy.shape[0] % 64 == 0 and dy.shape[0] >= 64 and dy.shape[1] % 8 == 0 and dy.shape[1] >= 64 and
                x.shape[1] % 32 == 0 and x.shape[1] >= 64 and self._castable(dy) and self._castable(x))

    def _wgrad(self, wname, dy, x, xt=None):
        """G[w] += dyᵀ · x, G[b] += colsum(dy).  ``xt``: a one-element list caching the bf16 transpose of x across calls that share x."""
        gw, gb = self.g[wname + ".weight"], self.g[wname + ".bias"]
        M = dy.shape[0]
        if x.dtype != torch.float32 and not self._tn_ok(dy, x):
            x = x.float()                          # (a bf16 operand copy met a shape the transposed-operand GEMM does not take)
        if self._tn_ok(dy, x):
            # one pass over dy: its bf16 copy (the dgrad GEMM reuses it) + the bias gradient; x's bf16 copy is the forward's
            ops.linear_bf16_accumulate_tn_(gw, self._act16(dy, colsum_into=gb), self._act16(x))
        elif self.gemm == "bf16" and M % 64 == 0 and M >= 64 and x.shape[1] >= 64 and dy.shape[1] >= 64:
            a = ops.transpose_cast_bf16(dy, colsum_into=gb)  # [N, M]: the product runs over the rows; G[b] rides on the same read of dy
            if xt is not None and xt:
                bt = xt[0]
            else:
                bt = ops.transpose_cast_bf16(x)              # [K, M]   (transposing copy + cast in one pass)
                if xt is not None:
                    xt.append(bt)
            ops.linear_bf16_accumulate_(gw, a, bt)           # G[w] += a · btᵀ, K (= the token rows) split over the SMs
        else:
            ops.linear_wgrad_(gw, gb, dy, x)

    def _next_seed(self):
        """A fresh mask stream per dropout site and optimisation step."""
        self._site += 1
        return (self.seed * 1_000_003 + self.step_count) * 4096 + self._site

    # ---- forward blocks ----
    def _att_fwd(self, a, x_q, x_kv, B, Lq, Lk, mask, qkv_q=None, qkv_kv=None):
        w, b, _ = self._qkv(a)
        qkv_q = self._mm(x_q, w, b, wkey=a.q) if qkv_q is None else qkv_q
        if qkv_kv is None:
            qkv_kv = qkv_q if x_kv is x_q else self._mm(x_kv, w, b, wkey=a.q)
        wo, bo = self._w(a.out)
        s_att = s_out = 0
        if self.dropout > 0:
            s_att, s_out = self._next_seed(), self._next_seed()
            ctx = ops.attention_dropout(qkv_q, qkv_kv, B, Lq, Lk, key_mask=mask, heads=self.heads, p=self.dropout, seed=s_att)
            pre = ops.dropout_add(self._mm(ctx, wo, bo, wkey=a.out), x_q, self.dropout, s_out)       # dense -> dropout -> + input
        elif self.gemm == "bf16" and qkv_q.shape[0] % 64 == 0 and self._castable(qkv_q) and self._castable(qkv_kv):
            # mixed precision: the attention on the tensor cores too (bf16 Q, K, V; bf16 context = the output projection's operand)
            q16 = self._act16(qkv_q)
            ctx = ops.attention(q16, q16 if qkv_kv is qkv_q else self._act16(qkv_kv), B, Lq, Lk, key_mask=mask, heads=self.heads)
            pre = self._mm(ctx, wo, bo, epilogue=L.EPI_RESID, residual=x_q, wkey=a.out)
        else:
            ctx = ops.attention(qkv_q, qkv_kv, B, Lq, Lk, key_mask=mask, heads=self.heads)
            pre = self._mm(ctx, wo, bo, epilogue=L.EPI_RESID, residual=x_q, wkey=a.out)
        y = self._ln(pre, a.ln)
        return y, dict(x_q=x_q, qkv_q=qkv_q, qkv_kv=qkv_kv, ctx=ctx, pre=pre, B=B, Lq=Lq, Lk=Lk, mask=mask, s_att=s_att, s_out=s_out)

    def _ffn_fwd(self, f, x):
        pg = self._mm(x, *self._w(f.inter), wkey=f.inter)
        inter = ops.gelu_fwd(pg, out_dtype=torch.bfloat16 if (self.gemm == "bf16" and pg.shape[0] % 64 == 0 and self.dropout == 0) else torch.float32)
        wo, bo = self._w(f.out)
        s_out = 0
        if self.dropout > 0:
            s_out = self._next_seed()
            pre = ops.dropout_add(self._mm(inter, wo, bo, wkey=f.out), x, self.dropout, s_out)
        else:
            pre = self._mm(inter, wo, bo, epilogue=L.EPI_RESID, residual=x, wkey=f.out)
        return self._ln(pre, f.ln), dict(x=x, pg=pg, inter=inter, pre=pre, s_out=s_out)

    def forward(self, feat, pos, ids, seg, mask):
        dev = self.dev
        B, V, Lt = feat.shape[0], feat.shape[1], ids.shape[1]
        ids32, mask32 = ids.to(dev, torch.int32).contiguous(), mask.to(dev, torch.int32).contiguous()
        seg32 = torch.zeros_like(ids32) if seg is None else seg.to(dev, torch.int32).contiguous()
        feat2 = feat.to(dev, torch.float32).reshape(B * V, -1).contiguous()
        pos2 = pos.to(dev, torch.float32).reshape(B * V, -1).contiguous()
        e = P + "embeddings."
        word, posemb, typeemb = self.p[e + "word_embeddings.weight"], self.p[e + "position_embeddings.weight"], self.p[e + "token_type_embeddings.weight"]
        t = dict(B=B, V=V, L=Lt, ids=ids32.view(-1), seg=seg32.view(-1), feat2=feat2, pos2=pos2)
        # BertEmbeddings (:278-292); the pre-LayerNorm sum is re-formed with tensor indexing for the LayerNorm backward
        t["e_pre"] = (word[ids32.long()] + posemb[:Lt][None, :, :] + typeemb[seg32.long()]).view(B * Lt, -1).contiguous()
        self._site = 0
        lang = ops.embed_ln(ids32, seg32, word, posemb, typeemb, *self._w(e + "LayerNorm"), self.eps)
        if self.dropout > 0:
            t["s_emb"] = self._next_seed()
            lang = ops.dropout_add(lang, None, self.dropout, t["s_emb"])
        v = P + "encoder.visn_fc."
        t["pre_v"] = self._mm(feat2, *self._w(v + "visn_fc"), wkey=v + "visn_fc")
        t["pre_b"] = ops.linear_f32(pos2, *self._w(v + "box_fc"))
        visn = ops.visn_combine(t["pre_v"], pos2, *self._w(v + "box_fc"), *self._w(v + "visn_layer_norm"), *self._w(v + "box_layer_norm"), self.eps)
        if self.dropout > 0:
            t["s_visn"] = self._next_seed()
            visn = ops.dropout_add(visn, None, self.dropout, t["s_visn"])
        t["l"], t["r"], t["x"] = [], [], []
        for a, f in self.l:
            lang, ta = self._att_fwd(a, lang, lang, B, Lt, Lt, mask32)
            lang, tf = self._ffn_fwd(f, lang)
            t["l"].append((ta, tf))
        for a, f in self.r:
            visn, ta = self._att_fwd(a, visn, visn, B, V, V, None)
            visn, tf = self._ffn_fwd(f, visn)
            t["r"].append((ta, tf))
        for w in self.x:
            c = w["cross"]
            wq, bq, _ = self._qkv(c)
            qkv_l, qkv_v = self._mm(lang, wq, bq, wkey=c.q), self._mm(visn, wq, bq, wkey=c.q)
            l1, tcl = self._att_fwd(c, lang, visn, B, Lt, V, None, qkv_q=qkv_l, qkv_kv=qkv_v)
            v1, tcv = self._att_fwd(c, visn, lang, B, V, Lt, mask32, qkv_q=qkv_v, qkv_kv=qkv_l)
            l2, tls = self._att_fwd(w["ls"], l1, l1, B, Lt, Lt, mask32)
            v2, tvs = self._att_fwd(w["vs"], v1, v1, B, V, V, None)
            lang, tlf = self._ffn_fwd(w["lf"], l2)
            visn, tvf = self._ffn_fwd(w["vf"], v2)
            t["x"].append(dict(cl=tcl, cv=tcv, ls=tls, vs=tvs, lf=tlf, vf=tvf))
        t["first"] = lang.view(B, Lt, -1)[:, 0, :].contiguous()
        t["pooled"] = ops.linear_f32(t["first"], *self._w(P + "pooler.dense"), epilogue=L.EPI_TANH)
        t["h0"] = ops.linear_f32(t["pooled"], *self._w("logit_fc.0"))
        t["g"] = ops.gelu_fwd(t["h0"])
        t["n"] = ops.layernorm(t["g"], *self._w("logit_fc.2"), self.eps)
        logits = ops.linear_f32(t["n"], *self._w("logit_fc.3"))
        self.tape = t
        return logits, t["pooled"]

    # ---- backward blocks (input gradient returned, parameter gradients accumulated into self.g) ----
    def _lin_wgrad(self, name, dy, x):
        self._wgrad(name, dy, x)

    def _dgrad(self, dy, name, wt, residual=None):
        """dy · W (+ residual): W given as its transposed fp32 copy wt [K, N] (a callable builds it only if the fp32 path needs it)."""
        epi = L.EPI_RESID if residual is not None else L.EPI_BIAS
        w = self.p.get(name + ".weight")                  # [N, K]; absent for the stacked Q|K|V pseudo-names
        if self.gemm == "bf16" and w is not None and w.shape[0] % 64 == 0 and dy.shape[0] >= 64:
            key = ("b16", ("t", name))
            if key not in self._derived:                  # bf16 W^T straight from the fp32 master weight: one pass, no fp32 transposed copy
                self._derived[key] = ops.transpose_cast_bf16(w)
            return ops.linear_bf16(self._act16(dy), self._derived[key], None, epilogue=epi, residual=residual,
                                   out_dtype=torch.float32)
        return self._mm(dy, wt() if callable(wt) else wt, None, epilogue=epi, residual=residual, wkey=("t", name))

    def _ln_bwd(self, name, dy, x, scale=1.0):
        if x.shape[1] in (768, 1536):     # one pass over (dy, x): dx and both parameter gradients
            return ops.layernorm_bwd_params_(self.g[name + ".weight"], self.g[name + ".bias"], dy, x, self.p[name + ".weight"],
                                             eps=self.eps, scale=scale)
        ops.layernorm_param_grad_(self.g[name + ".weight"], self.g[name + ".bias"], dy, x, self.eps, scale)
        return ops.layernorm_bwd(dy, x, self.p[name + ".weight"], eps=self.eps, scale=scale)

    def _ffn_bwd(self, f, tp, dy):
        d_pre = self._ln_bwd(f.ln, dy, tp["pre"])
        d_h = ops.dropout_add(d_pre, None, self.dropout, tp["s_out"]) if tp["s_out"] else d_pre    # through the output dropout
        self._lin_wgrad(f.out, d_h, tp["inter"])
        d_pg = ops.gelu_bwd(self._dgrad(d_h, f.out, lambda: self._wt(f.out)), tp["pg"])
        self._lin_wgrad(f.inter, d_pg, tp["x"])
        return self._dgrad(d_pg, f.inter, lambda: self._wt(f.inter), residual=d_pre)

    def _att_bwd(self, a, tp, dy, dqkv_q, dqkv_kv):
        d_pre = self._ln_bwd(a.ln, dy, tp["pre"])
        d_h = ops.dropout_add(d_pre, None, self.dropout, tp["s_out"]) if tp["s_out"] else d_pre
        self._lin_wgrad(a.out, d_h, tp["ctx"])
        d_ctx = self._dgrad(d_h, a.out, lambda: self._wt(a.out))
        ops.attention_bwd(tp["qkv_q"], tp["qkv_kv"], d_ctx, dqkv_q, dqkv_kv, tp["B"], tp["Lq"], tp["Lk"], key_mask=tp["mask"], heads=self.heads,
                          p=self.dropout if tp["s_att"] else 0.0, seed=tp["s_att"],
                          bf16=self.gemm == "bf16" and not tp["s_att"])     # mixed precision: the attention products on the tensor cores too
        return d_pre

    def _qkv_wgrad(self, a, dqkv, x):
        H = dqkv.shape[1] // 3
        if self._tn_ok(dqkv[:, :H], x) and self._castable(dqkv):
            # the three column blocks of dqkv are cast (with their bias gradients) into ONE bf16 [rows, 3H] buffer: its blocks are the
            # transposed-operand wgrad GEMMs' dy, the whole of it the stacked dgrad GEMM's
            d16 = torch.empty(dqkv.shape, dtype=torch.bfloat16, device=dqkv.device)
            for j, name in enumerate((a.q, a.k, a.v)):
                ops.cast_bf16_colsum(dqkv[:, j * H:(j + 1) * H], colsum_into=self.g[name + ".bias"], out=d16[:, j * H:(j + 1) * H])
            self._a16[(dqkv.data_ptr(), tuple(dqkv.shape), tuple(dqkv.stride()))] = (dqkv, d16)
            x16 = self._act16(x)
            for j, name in enumerate((a.q, a.k, a.v)):
                ops.linear_bf16_accumulate_tn_(self.g[name + ".weight"], d16[:, j * H:(j + 1) * H], x16)
            return
        xt = []                                               # Q, K and V share their input: transpose it once
        for j, name in enumerate((a.q, a.k, a.v)):
            self._wgrad(name, dqkv[:, j * H:(j + 1) * H], x, xt)

    def _self_att_bwd(self, a, tp, dy):
        dqkv = torch.empty_like(tp["qkv_q"])
        d_pre = self._att_bwd(a, tp, dy, dqkv, dqkv)
        self._qkv_wgrad(a, dqkv, tp["x_q"])
        return self._dgrad(dqkv, "qkv:" + a.q, self._qkv(a)[2], residual=d_pre)

    def _bucket_done(self, k):
        """Every weight gradient of bucket k is enqueued: start its all-reduce (mean over ranks) behind them, on NCCL's stream."""
        cap = self._capture
        if cap is not None and k in self._graph_cuts:      # under capture: close this segment's graph, open the next one
            cap["graphs"][-1].capture_end()
            cap["ends"].append(k)
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=cap["graphs"][0].pool())
            cap["graphs"].append(g)
        if not self._overlap:
            return
        for lo, hi in self.buckets[k]:
            self._works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.AVG, async_op=True))

    def backward(self, dlogits):
        t = self.tape
        B, V, Lt = t["B"], t["V"], t["L"]
        H = t["pooled"].shape[1]
        dlogits = dlogits.to(self.dev, torch.float32).contiguous()
        self._lin_wgrad("logit_fc.3", dlogits, t["n"])
        dn = ops.linear_f32(dlogits, self._wt("logit_fc.3"))
        dg = self._ln_bwd("logit_fc.2", dn, t["g"])
        dh0 = ops.gelu_bwd(dg, t["h0"])
        self._lin_wgrad("logit_fc.0", dh0, t["pooled"])
        d_pp = ops.tanh_bwd(ops.linear_f32(dh0, self._wt("logit_fc.0")), t["pooled"])
        self._lin_wgrad(P + "pooler.dense", d_pp, t["first"])
        nb = 0
        self._bucket_done(nb)
        dl = torch.zeros((B, Lt, H), dtype=torch.float32, device=self.dev)
        dl[:, 0, :] = ops.linear_f32(d_pp, self._wt(P + "pooler.dense"))
        dl = dl.view(B * Lt, H)
        dv = torch.zeros((B * V, H), dtype=torch.float32, device=self.dev)
        for w, tp in zip(reversed(self.x), reversed(t["x"])):
            dl2, dv2 = self._ffn_bwd(w["lf"], tp["lf"], dl), self._ffn_bwd(w["vf"], tp["vf"], dv)
            dl1, dv1 = self._self_att_bwd(w["ls"], tp["ls"], dl2), self._self_att_bwd(w["vs"], tp["vs"], dv2)
            c = w["cross"]
            dqkv_l, dqkv_v = torch.empty_like(tp["cl"]["qkv_q"]), torch.empty_like(tp["cv"]["qkv_q"])
            d_pre_l = self._att_bwd(c, tp["cl"], dl1, dqkv_l, dqkv_v)
            d_pre_v = self._att_bwd(c, tp["cv"], dv1, dqkv_v, dqkv_l)
            self._qkv_wgrad(c, dqkv_l, tp["cl"]["x_q"])                 # ONE module, both modalities: gradients add up
            self._qkv_wgrad(c, dqkv_v, tp["cv"]["x_q"])
            wt = self._qkv(c)[2]
            dl = self._dgrad(dqkv_l, "qkv:" + c.q, wt, residual=d_pre_l)
            dv = self._dgrad(dqkv_v, "qkv:" + c.q, wt, residual=d_pre_v)
            nb += 1
            self._bucket_done(nb)
        for (a, f), (ta, tf) in zip(reversed(self.r), reversed(t["r"])):
            dv = self._self_att_bwd(a, ta, self._ffn_bwd(f, tf, dv))
            nb += 1
            self._bucket_done(nb)
        for (a, f), (ta, tf) in zip(reversed(self.l), reversed(t["l"])):
            dl = self._self_att_bwd(a, ta, self._ffn_bwd(f, tf, dl))
            nb += 1
            self._bucket_done(nb)
        v = P + "encoder.visn_fc."
        if self.dropout > 0:
            dv = ops.dropout_add(dv, None, self.dropout, t["s_visn"])
            dl = ops.dropout_add(dl, None, self.dropout, t["s_emb"])
        d_pre_v = self._ln_bwd(v + "visn_layer_norm", dv, t["pre_v"], scale=0.5)
        d_pre_b = self._ln_bwd(v + "box_layer_norm", dv, t["pre_b"], scale=0.5)
        self._lin_wgrad(v + "visn_fc", d_pre_v, t["feat2"])
        self._lin_wgrad(v + "box_fc", d_pre_b, t["pos2"])
        e = P + "embeddings."
        d_e = self._ln_bwd(e + "LayerNorm", dl, t["e_pre"])
        # all three tables are nn.Embedding(..., padding_idx=0) in the reference (modeling.py:272-274): word 0, position 0
        # and token type 0 get no gradient
        ops.embedding_scatter_(self.g[e + "word_embeddings.weight"], t["ids"], d_e, padding_idx=0)
        ops.embedding_scatter_(self.g[e + "token_type_embeddings.weight"], t["seg"], d_e, padding_idx=0)
        pos_ids = torch.arange(Lt, dtype=torch.int32, device=self.dev).repeat(B)
        ops.embedding_scatter_(self.g[e + "position_embeddings.weight"], pos_ids, d_e, padding_idx=0)
        self._bucket_done(nb + 1)

    # ---- one optimisation step (gqa_conf.py:148-198 without the data-side lines) ----
    def lr_now(self):
        hp = self.hp
        if hp["t_total"] != -1:
            return hp["lr"] * SCHEDULES[hp["schedule"]](self.step_count / hp["t_total"], hp["warmup"])
        return hp["lr"]

    def _grad_pass(self, feat, pos, ids, seg, mask, target):
        """zero grads, forward, loss, backward (+ the bucketed all-reduces): everything of a step that does not depend on the
        step counter, i.e. what a CUDA graph can replay.  Returns the fp64 loss accumulator (device)."""
        lib = L.load()
        self._a16.clear()
        self.flat_g.zero_()
        logits, _ = self.forward(feat, pos, ids, seg, mask)
        target = target.to(self.dev, torch.float32).contiguous()
        B, A = logits.shape
        dlogits = torch.empty_like(logits)
        loss = torch.zeros(1, dtype=torch.float64, device=self.dev)
        L.check(lib.rgqa_bce_loss_grad(L.current_stream_ptr(), L.ptr(logits), L.ptr(target), B, A, L.ptr(dlogits), L.ptr(loss)),
                "rgqa_bce_loss_grad")
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not self._capturing
        # NCCL: one bucket per finished layer, each all-reduce overlapped with the remaining backward (see __init__);
        # other backends (the gloo CPU tests): one blocking all-reduce of the flat gradient after the backward.
        # Under CUDA-graph capture no collective is issued at all (NCCL work handles captured with async_op hung on B200 /
        # NCCL 2.28 in this build): _graphed_grad_pass all-reduces the flat gradient once after the replay.
        self._overlap = multi and dist.get_backend() == "nccl" and self.bucketed
        self._works = []
        self.backward(dlogits)
        if self._overlap:
            for w in self._works:
                w.wait()                                                # the compute stream waits; the host does not
            self._works = []
        elif multi:
            buf = self.flat_g if dist.get_backend() == "nccl" else self.flat_g.cpu()
            dist.all_reduce(buf)                                        # ONE collective per step
            self.flat_g.copy_(buf.to(self.dev) / dist.get_world_size())
        return loss

    def _graphed_grad_pass(self, feat, pos, ids, seg, mask, target):
        """The ~1 000 kernel launches of a gradient pass are issued from Python (~40 ms of host time against ~25 ms of GPU
        work): capture them ONCE per input shape into a CUDA graph over static input buffers and replay.  Only with dropout
        off (the mask seeds depend on the step counter).  The first call of a shape runs eagerly (lazy kernel attributes, NCCL
        channel set-up must not happen under capture)."""
        args = [feat, pos, ids, torch.zeros_like(ids) if seg is None else seg, mask, target]
        key = tuple((tuple(t.shape), t.dtype) for t in args)
        st = self._graphs.get(key)
        if st is None:
            self._graphs[key] = {"seen": 1}
            return self._grad_pass(feat, pos, ids, seg, mask, target)
        if "graph" not in st:
            st["inputs"] = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in args]
            for dst, src in zip(st["inputs"], args):
                dst.copy_(src, non_blocking=True)
            self._derived.clear()
            torch.cuda.synchronize()
            gc.collect()
            side = torch.cuda.Stream(device=self.dev)       # capture needs a non-default stream
            side.wait_stream(torch.cuda.current_stream(self.dev))
            self._capturing = True
            cap = {"graphs": [torch.cuda.CUDAGraph()], "ends": []}
            try:
                with torch.cuda.stream(side):
                    cap["graphs"][0].capture_begin()
                    self._capture = cap
                    st["loss"] = self._grad_pass(*st["inputs"])     # _bucket_done cuts the capture at self._graph_cuts
                    cap["graphs"][-1].capture_end()
                    cap["ends"].append(len(self.buckets) - 1)
            finally:
                self._capture = None
                self._capturing = False
            torch.cuda.current_stream(self.dev).wait_stream(side)
            st["graphs"] = cap["graphs"]
            # the flat-gradient slices each segment finishes, contiguous runs merged (a segment's layers are neighbours)
            st["runs"], first = [], 0
            for end in cap["ends"]:
                spans = sorted(r for k in range(first, end + 1) for r in self.buckets[k])
                runs = []
                for lo, hi in spans:
                    if runs and runs[-1][1] == lo:
                        runs[-1][1] = hi
                    else:
                        runs.append([lo, hi])
                st["runs"].append(runs)
                first = end + 1
            st["graph"] = st["graphs"][0]
            self._derived.clear()          # tensors created under capture live in the graph's pool: never reuse them eagerly
            self._a16.clear()
        for dst, src in zip(st["inputs"], args):
            dst.copy_(src, non_blocking=True)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        works = []
        for g, runs in zip(st["graphs"], st["runs"]):
            g.replay()
            if multi and dist.get_backend() == "nccl":
                # this segment's gradients are final: their all-reduce (mean) goes to NCCL's stream behind the replay just
                # enqueued and overlaps the next segment's replay
                for lo, hi in runs:
                    works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.AVG, async_op=True))
        for w in works:
            w.wait()                                        # the compute stream waits; the host does not
        if multi and dist.get_backend() != "nccl":
            buf = self.flat_g.cpu()
            dist.all_reduce(buf)
            self.flat_g.copy_(buf.to(self.dev) / dist.get_world_size())
        return st["loss"]

    def train_step(self, feat, pos, ids, seg, mask, target):
        """Returns (loss, gradient norm before clipping) as Python floats."""
        lib = L.load()
        if self.use_graph and self.dropout == 0:
            loss = self._graphed_grad_pass(feat, pos, ids, seg, mask, target)
        else:
            loss = self._grad_pass(feat, pos, ids, seg, mask, target)
        sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        stream = L.current_stream_ptr()
        L.check(lib.rgqa_grad_sumsq(stream, L.ptr(self.flat_g), self.flat_g.numel(), L.ptr(sumsq)), "rgqa_grad_sumsq")
        hp = self.hp
        L.check(lib.rgqa_bertadam_step(stream, L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.flat_m), L.ptr(self.flat_v),
                                       self.flat_p.numel(), float(self.lr_now()), hp["b1"], hp["b2"], hp["e"], hp["weight_decay"],
                                       L.ptr(sumsq), float(self.clip_norm)), "rgqa_bertadam_step")
        self.step_count += 1
        self._derived.clear()                                           # stacked / transposed copies are stale now
        self._a16.clear()
        return float(loss), float(sumsq.sqrt())

    def state_dict(self):
        return {n: t.detach().clone() for n, t in self.p.items()}
