"""Python owner of one ``rgqa_handle`` (C ABI): weight upload by reference key, forward, scorers.

This is plumbing only — it allocates outputs with torch, borrows data pointers, passes the current
CUDA stream and raises ``RuntimeError`` with ``rgqa_last_error()`` on any failure.  No arithmetic of the
hot path happens in Python and nothing here falls back to PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib as L


@dataclass
class EngineConfig:
    arch: str = "lxmert"          # "lxmert" | "uniter"
    l_layers: int = 9
    x_layers: int = 5
    r_layers: int = 5
    hidden: int = 768
    heads: int = 12
    intermediate: int = 3072
    vocab_size: int = 30522
    max_position: int = 512
    type_vocab: int = 2
    visual_feat_dim: int = 2048
    visual_pos_dim: int = 4
    num_rois: int = 36
    num_answers: int = 1842
    precision: str = "bf16"       # "bf16" | "fp32"
    ln_eps: float = 1e-12

    def to_c(self) -> L.RgqaConfig:
        if self.arch not in ("lxmert", "uniter"):
            raise ValueError(f"unknown arch {self.arch!r}")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"unknown precision {self.precision!r}")
        return L.RgqaConfig(
            L.ARCH_LXMERT if self.arch == "lxmert" else L.ARCH_UNITER, self.l_layers, self.x_layers, self.r_layers,
            self.hidden, self.heads, self.intermediate, self.vocab_size, self.max_position, self.type_vocab,
            self.visual_feat_dim, self.visual_pos_dim, self.num_rois, self.num_answers,
            L.PREC_BF16 if self.precision == "bf16" else L.PREC_FP32, self.ln_eps)


class Engine:
    def __init__(self, cfg: EngineConfig, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("rgqa_b200.Engine needs a CUDA device (there is no CPU path)")
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = L.load()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            ccfg = cfg.to_c()
            L.check(self._lib.rgqa_create(C.byref(ccfg), C.byref(self._h)), "rgqa_create")
        self.ready = False
        self.has_maha = False

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.rgqa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights --------------------------------------------------------------------------------
    def weight_names(self):
        n = self._lib.rgqa_num_weights(self._h)
        return [self._lib.rgqa_weight_name(self._h, i).decode() for i in range(n)]

    def load_weights(self, named_tensors):
        """named_tensors: iterable of (reference key, tensor); keys may carry the reference's
        ``lxrt_encoder.model.`` / ``encoder.model.`` / ``module.`` prefixes.  Unknown keys raise."""
        with torch.cuda.device(self.device):
            for key, t in named_tensors:
                t = t.detach()
                if t.dtype != torch.float32:
                    t = t.float()
                t = t.contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                L.check(self._lib.rgqa_load_weight(self._h, key.encode(), t.data_ptr(), 1 if t.is_cuda else 0, shape,
                                                   t.dim()), f"rgqa_load_weight({key})")
            L.check(self._lib.rgqa_finalize_weights(self._h), "rgqa_finalize_weights")
        self.ready = True

    # ---- forward --------------------------------------------------------------------------------
    def _prep_common(self, feat, pos, ids, seg, mask):
        if not self.ready:
            raise RuntimeError("rgqa_b200.Engine: weights were not loaded")
        for name, t in (("feat", feat), ("pos", pos), ("ids", ids), ("mask", mask)):
            if not t.is_cuda:
                raise RuntimeError(f"rgqa_b200.Engine: `{name}` must be a CUDA tensor")
        B = feat.shape[0]
        if feat.shape[1:] != (self.cfg.num_rois, self.cfg.visual_feat_dim):
            raise AssertionError(f"expected feat [B,{self.cfg.num_rois},{self.cfg.visual_feat_dim}], got {tuple(feat.shape)}")
        if pos.shape != (B, self.cfg.num_rois, self.cfg.visual_pos_dim):
            raise AssertionError(f"expected pos [B,{self.cfg.num_rois},{self.cfg.visual_pos_dim}], got {tuple(pos.shape)}")
        if feat.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("feat must be float32 or bfloat16")
        feat = feat.contiguous()
        pos = pos.float().contiguous()
        ids = ids.to(torch.int32).contiguous()
        mask = mask.to(torch.int32).contiguous()
        seg = torch.zeros_like(ids) if seg is None else seg.to(torch.int32).contiguous()
        return feat, pos, ids, seg, mask, B, ids.shape[1]

    def forward(self, feat, pos, ids, seg, mask, want_logits=True, want_pooled=True):
        """(pooled f32 [B,768] | None, logits f32 [B,A] | None) — one call into the C ABI."""
        feat, pos, ids, seg, mask, B, Lq = self._prep_common(feat, pos, ids, seg, mask)
        want_logits = want_logits and self.cfg.num_answers > 0
        pooled = torch.empty((B, self.cfg.hidden), dtype=torch.float32, device=feat.device) if want_pooled else None
        logits = torch.empty((B, self.cfg.num_answers), dtype=torch.float32, device=feat.device) if want_logits else None
        fn = self._lib.rgqa_lxmert_forward if self.cfg.arch == "lxmert" else self._lib.rgqa_uniter_forward
        with torch.cuda.device(feat.device):
            L.check(fn(self._h, L.current_stream_ptr(), feat.data_ptr(), L.DT_F32 if feat.dtype == torch.float32 else L.DT_BF16,
                       pos.data_ptr(), ids.data_ptr(), seg.data_ptr(), mask.data_ptr(), B, Lq, L.ptr(pooled), L.ptr(logits)),
                    f"rgqa_{self.cfg.arch}_forward")
        return pooled, logits

    def forward_scored(self, feat, pos, ids, seg, mask, modes, want_logits=False, want_pooled=False):
        """Forward + rejection scorers in ONE C-ABI call (``rgqa_forward_scored``): the scorer reductions run in the
        answer-head GEMM's epilogue / on the tensor cores, the logit and Mahalanobis matrices stay on chip.
        ``modes``: iterable of msp|energy|energy_sum|lse|maha.  Returns ({mode: (score f32 [B], label i32 [B])}, pooled, logits)."""
        feat, pos, ids, seg, mask, B, Lq = self._prep_common(feat, pos, ids, seg, mask)
        codes = [L.SCORE_MODES[m] for m in modes]
        if L.SCORE_MAHA in codes and not self.has_maha:
            raise RuntimeError("rgqa_b200.Engine: Mahalanobis statistics were not set (set_maha)")
        bits = 0
        for c in codes:
            bits |= 1 << c
        dev = feat.device
        scores = torch.empty((5, B), dtype=torch.float32, device=dev)
        labels = torch.empty((2, B), dtype=torch.int32, device=dev)
        pooled = torch.empty((B, self.cfg.hidden), dtype=torch.float32, device=dev) if want_pooled else None
        logits = torch.empty((B, self.cfg.num_answers), dtype=torch.float32, device=dev) if want_logits else None
        with torch.cuda.device(dev):
            L.check(self._lib.rgqa_forward_scored(self._h, L.current_stream_ptr(), feat.data_ptr(),
                                                  L.DT_F32 if feat.dtype == torch.float32 else L.DT_BF16, pos.data_ptr(),
                                                  ids.data_ptr(), seg.data_ptr(), mask.data_ptr(), B, Lq, bits, L.ptr(pooled),
                                                  L.ptr(logits), scores.data_ptr(), labels.data_ptr()), "rgqa_forward_scored")
        out = {m: (scores[c], labels[1 if c == L.SCORE_MAHA else 0]) for m, c in zip(modes, codes)}
        return out, pooled, logits

    def fold_residue(self):
        """(max |c[n]| over the folded consumer weights, number of consumers that kept the exact epilogue) — deferred LayerNorm."""
        n = C.c_int(0)
        return float(self._lib.rgqa_fold_residue(self._h, C.byref(n))), int(n.value)

    def graph_captures(self) -> int:
        return int(self._lib.rgqa_graph_captures(self._h))

    # ---- scorers --------------------------------------------------------------------------------
    def set_maha(self, sample_mean, precision):
        """{mean, precision} from ``sample_estimator`` (src/tasks/gqa_maha.py:120-201); numpy or torch."""
        mu = torch.as_tensor(sample_mean).float().contiguous()
        P = torch.as_tensor(precision).float().contiguous()
        if mu.shape != (self.cfg.num_answers, self.cfg.hidden) or P.shape != (self.cfg.hidden, self.cfg.hidden):
            raise ValueError(f"Mahalanobis statistics have the wrong shape: {tuple(mu.shape)}, {tuple(P.shape)}")
        if mu.is_cuda != P.is_cuda:
            mu, P = mu.cpu(), P.cpu()
        with torch.cuda.device(self.device):
            L.check(self._lib.rgqa_set_maha(self._h, mu.data_ptr(), P.data_ptr(), 1 if mu.is_cuda else 0), "rgqa_set_maha")
        self.has_maha = True

    def score(self, mode, logits=None, pooled=None, return_full=False):
        """(score f32 [B], label int32 [B]) on the device; mode in msp|energy|energy_sum|lse|maha."""
        code = L.SCORE_MODES[mode]
        src = pooled if code == L.SCORE_MAHA else logits
        if src is None:
            raise ValueError(f"scorer {mode!r} needs " + ("pooled features" if code == L.SCORE_MAHA else "logits"))
        B = src.shape[0]
        score = torch.empty(B, dtype=torch.float32, device=src.device)
        label = torch.empty(B, dtype=torch.int32, device=src.device)
        full = None
        if return_full and code == L.SCORE_MAHA:
            full = torch.empty((B, self.cfg.num_answers), dtype=torch.float32, device=src.device)
        lg = logits.contiguous() if logits is not None else None
        pl = pooled.contiguous() if pooled is not None else None
        with torch.cuda.device(src.device):
            L.check(self._lib.rgqa_score(self._h, L.current_stream_ptr(), code, L.ptr(lg), L.ptr(pl), B, score.data_ptr(),
                                         label.data_ptr(), L.ptr(full)), "rgqa_score")
        return (score, label, full) if return_full else (score, label)
