"""Mirror of ``src/uniter/uniter.py``: ``GQAUNITER(num_answers).forward(feat, pos7, sent) -> logit`` with attributes
``.encoder`` (``.model``, ``.tokenizer``) and ``.logit_fc``; state-dict keys ``encoder.model.uniter.*``, ``logit_fc.*``."""
from __future__ import annotations

import torch.nn as nn

from ..engine import Engine
from ..param import args as _default_args
from ..tasks.gqa_model import _LogitFC
from .entry import UniterEncoder

MAX_VQA_LENGTH = 20


class GQAUNITER(nn.Module):
    returns_pooled = False

    def __init__(self, num_answers, args=None):
        super().__init__()
        self.num_answers = num_answers
        self.encoder = UniterEncoder(_default_args if args is None else args)
        self.logit_fc = _LogitFC(num_answers)
        self._engine = None
        self._engine_fp = None

    def engine(self) -> Engine:
        enc = self.encoder.model
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

    def forward_ids(self, feat, pos7, input_ids, segment_ids, input_mask):
        pooled, logits = self.engine().forward(feat, pos7, input_ids, segment_ids, input_mask)
        return logits, pooled

    def forward_scored_ids(self, feat, pos7, input_ids, segment_ids, input_mask, methods):
        """Tensor-only forward + scorers in one engine call: [(score f32 [B], label i32 [B]) per method]."""
        out, _, _ = self.engine().forward_scored(feat, pos7, input_ids, segment_ids, input_mask, list(methods))
        return [out[m] for m in methods]

    def forward(self, feat, pos, sent):
        assert feat.shape[1] == 36, "Not Using 36 ROIs, please change the following 2 lines"
        ids, mask, seg = self.encoder.tokenize(sent, feat.device)
        logits, pooled = self.forward_ids(feat, pos, ids, seg, mask)
        return (logits, pooled) if self.returns_pooled else logits


class GQAUNITER_maha(GQAUNITER):
    returns_pooled = True
