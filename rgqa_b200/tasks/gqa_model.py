"""Host-side mirror of ``src/tasks/gqa_model.py``: ``GQAModel(num_answers).forward(feat, pos, sent) -> logit``
and ``GQAModel_maha`` returning ``(logit, x)``; attributes ``.lxrt_encoder`` and ``.logit_fc`` with the
reference's state-dict keys (``lxrt_encoder.model.bert.…``, ``logit_fc.{0,2,3}.{weight,bias}``).

One forward is: host tokenisation -> one small H2D copy -> ONE C-ABI call (``rgqa_lxmert_forward``) that
runs the whole encoder + pooler + answer head on the caller's stream.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import synthetic as _syn
from ..engine import Engine
from ..lxrt.entry import LXRTEncoder
from ..lxrt.modeling import RefKeyParams
from ..param import args as _default_args

# Max length including <bos> and <eos> (gqa_model.py:11)
MAX_GQA_LENGTH = 30


class _LogitFC(RefKeyParams):
    """Parameters of nn.Sequential(Linear(768,1536), GeLU, LayerNorm(1536), Linear(1536,A)) (gqa_model.py:22-27)."""

    def __init__(self, num_answers):
        super().__init__(_syn._head(num_answers), strip_prefix="logit_fc.")


class GQAModel(nn.Module):
    returns_pooled = False

    def __init__(self, num_answers, args=None, max_seq_length=None):
        super().__init__()
        args = _default_args if args is None else args
        L = max_seq_length or getattr(args, "max_seq_length", None) or MAX_GQA_LENGTH
        self.num_answers = num_answers
        self.lxrt_encoder = LXRTEncoder(args, max_seq_length=L)
        self.logit_fc = _LogitFC(num_answers)
        self._engine = None
        self._engine_fp = None

    # ---- engine (packed weights on the device), rebuilt whenever a parameter changed ----
    def engine(self) -> Engine:
        enc = self.lxrt_encoder.model
        fp = (enc.fingerprint(), self.logit_fc.fingerprint())
        if self._engine is None or self._engine_fp != fp:
            if self._engine is not None:
                self._engine.close()
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("rgqa_b200: call .cuda() first — the model runs only on a CUDA device")
            self._engine = Engine(enc.engine_config(self.num_answers), device=dev)
            named = list(enc.ref_named_parameters()) + [("logit_fc." + k, p) for k, p in self.logit_fc.ref_named_parameters()]
            self._engine.load_weights(named)
            self._engine_fp = fp
        return self._engine

    def forward_ids(self, feat, pos, input_ids, segment_ids, input_mask):
        """Tensor-only form (pre-tokenised questions): returns (logit [B,A] f32, pooled x [B,768] f32)."""
        pooled, logits = self.engine().forward(feat, pos, input_ids, segment_ids, input_mask)
        return logits, pooled

    def forward_scored_ids(self, feat, pos, input_ids, segment_ids, input_mask, methods):
        """Tensor-only forward + scorers in one engine call: [(score f32 [B], label i32 [B]) per method]."""
        out, _, _ = self.engine().forward_scored(feat, pos, input_ids, segment_ids, input_mask, list(methods))
        return [out[m] for m in methods]

    def forward(self, feat, pos, sent):
        """
        :param feat: (b, o, f) float32 (or pre-staged bfloat16) CUDA tensor
        :param pos:  (b, o, 4)
        :param sent: (b,) list of str
        :return: (b, num_answer) logits  [GQAModel_maha: (logits, pooled x)]
        """
        ids, mask, seg = self.lxrt_encoder.tokenize(sent, feat.device)
        logits, pooled = self.forward_ids(feat, pos, ids, seg, mask)
        return (logits, pooled) if self.returns_pooled else logits


class GQAModel_maha(GQAModel):
    """gqa_model.py:46-75: same network, forward returns (logit, x)."""
    returns_pooled = True
