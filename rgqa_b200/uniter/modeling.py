"""Host-side mirror of ``src/uniter/modeling.py``: ``UniterFeatureExtraction`` (:638-656) as a parameter bag with the
reference's state-dict keys (``uniter.embeddings.*``, ``uniter.img_embeddings.*``, ``uniter.encoder.layer.N.*``,
``uniter.pooler.dense.*``) whose forward is one C-ABI call (``rgqa_uniter_forward``)."""
from __future__ import annotations

import torch

from .. import synthetic as _syn
from ..engine import Engine, EngineConfig
from ..lxrt.modeling import BertConfig, RefKeyParams


class UniterFeatureExtraction(RefKeyParams):
    def __init__(self, config=None, precision="bf16"):
        self.config = config if config is not None else BertConfig(28996)
        self.precision = precision
        spec = _syn.uniter_spec(self.config.num_hidden_layers, num_answers=0, vocab=self.config.vocab_size,
                                max_pos=self.config.max_position_embeddings)
        super().__init__(spec, strip_prefix=_syn.UNITER_PREFIX)
        self._engine = None
        self._engine_fp = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, *inputs, **kwargs):
        """No S3 offline (uniter/modeling.py pulls bert-base-cased): random init; load real weights with load_state_dict."""
        kwargs.pop("cache_dir", None)
        kwargs.pop("state_dict", None)
        return cls(BertConfig(28996), *inputs, **kwargs)

    def engine_config(self, num_answers=0):
        return EngineConfig(arch="uniter", l_layers=self.config.num_hidden_layers, x_layers=0, r_layers=0,
                            vocab_size=self.config.vocab_size, max_position=self.config.max_position_embeddings,
                            type_vocab=self.config.type_vocab_size, visual_pos_dim=7, num_answers=num_answers,
                            precision=self.precision)

    def _encoder_engine(self):
        fp = self.fingerprint()
        if self._engine is None or self._engine_fp != fp:
            if self._engine is not None:
                self._engine.close()
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("rgqa_b200: call .cuda() first — the model runs only on a CUDA device")
            self._engine = Engine(self.engine_config(0), device=dev)
            self._engine.load_weights(self.ref_named_parameters())
            self._engine_fp = fp
        return self._engine

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, visual_feats=None,
                visual_token_type_ids=None, visual_attention_mask=None, img_pos_feat=None):
        # the reference's UniterEncoder always passes all-ones visual mask and visual token type 1 (entry.py:92-93)
        if visual_attention_mask is not None and not bool((visual_attention_mask == 1).all()):
            raise NotImplementedError("RoI masking is not on the evaluation path (src/uniter/entry.py:93)")
        if visual_token_type_ids is not None and not bool((visual_token_type_ids == 1).all()):
            raise NotImplementedError("image tokens carry token type 1 (src/uniter/entry.py:92)")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        pooled, _ = self._encoder_engine().forward(visual_feats, img_pos_feat, input_ids, token_type_ids, attention_mask,
                                                   want_logits=False)
        return pooled
