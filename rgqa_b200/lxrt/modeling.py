"""Host-side mirror of the reference's LXMERT model classes (``src/lxrt/modeling.py``).

Only the tensor-only seam survives as a module: ``LXRTFeatureExtraction.forward(input_ids,
token_type_ids, attention_mask, visual_feats=(feat, pos), visual_attention_mask=None)`` (:1020-1030).
The module holds fp32 ``nn.Parameter``s under exactly the reference's state-dict keys
(``bert.encoder.layer.3.attention.self.query.weight`` …) so reference checkpoints load with
``load_state_dict`` and ``state_dict()`` saves them back, but there is no module tree of
BertLayer/BertAttention objects: the arithmetic is one C-ABI call into hand-written CUDA kernels
(``rgqa_lxmert_forward``).  There is no PyTorch fallback — on a box without the extension or without a
GPU, ``forward`` raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import synthetic as _syn
from ..engine import Engine, EngineConfig


class VisualConfig(object):
    """Mirror of modeling.py:VisualConfig (:128-170): only the fields the eval path reads."""

    def __init__(self, l_layers=12, x_layers=5, r_layers=0):
        self.l_layers, self.x_layers, self.r_layers = l_layers, x_layers, r_layers
        self.visual_feat_dim = 2048
        self.visual_pos_dim = 4

    def set_visual_dims(self, feat_dim, pos_dim):
        self.visual_feat_dim, self.visual_pos_dim = feat_dim, pos_dim


VISUAL_CONFIG = VisualConfig()
BertLayerNorm = torch.nn.LayerNorm


class BertConfig(object):
    """BERT-base fields of modeling.py:BertConfig (:173-257)."""

    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02):
        self.vocab_size = vocab_size_or_config_json_file
        self.hidden_size, self.num_hidden_layers = hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size = num_attention_heads, intermediate_size
        self.hidden_act, self.hidden_dropout_prob = hidden_act, hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings, self.type_vocab_size = max_position_embeddings, type_vocab_size
        self.initializer_range = initializer_range


class RefKeyParams(nn.Module):
    """A bag of parameters addressed by reference state-dict key.

    ``nn.Module`` forbids '.' in parameter names, so parameters are registered under an escaped name and
    the (de)serialisation hooks translate: ``state_dict()`` emits ``prefix + reference_key`` and
    ``load_state_dict`` consumes them with the usual missing / unexpected / size-mismatch reporting.
    """

    def __init__(self, spec, strip_prefix=""):
        super().__init__()
        self._ref_keys = []
        self._kinds = {}
        for key, shape, kind in spec:
            if strip_prefix and key.startswith(strip_prefix):
                key = key[len(strip_prefix):]
            elif strip_prefix:
                continue
            self._ref_keys.append(key)
            self._kinds[key] = kind
            self.register_parameter(self._esc(key), nn.Parameter(torch.empty(shape), requires_grad=False))
        self.init_bert_weights()

    @staticmethod
    def _esc(key):
        return key.replace(".", "|")

    def ref_named_parameters(self):
        for key in self._ref_keys:
            yield key, self._parameters[self._esc(key)]

    def init_bert_weights(self, module=None):
        """BertPreTrainedModel.init_bert_weights (modeling.py:688-699): N(0, 0.02) / zeros / LayerNorm (1, 0)."""
        with torch.no_grad():
            for key, p in self.ref_named_parameters():
                kind = self._kinds[key]
                if kind == "w":
                    p.normal_(mean=0.0, std=0.02)
                elif kind == "g":
                    p.fill_(1.0)
                else:
                    p.zero_()

    def fingerprint(self):
        return tuple((p.data_ptr(), p._version) for _, p in self.ref_named_parameters())

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for key, p in self.ref_named_parameters():
            destination[prefix + key] = p if keep_vars else p.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        known = set()
        for key, p in self.ref_named_parameters():
            full = prefix + key
            known.add(full)
            if full not in state_dict:
                missing_keys.append(full)
                continue
            src = state_dict[full]
            if tuple(src.shape) != tuple(p.shape):
                error_msgs.append('size mismatch for {}: copying a param with shape {} from checkpoint, '
                                  'the shape in current model is {}.'.format(full, tuple(src.shape), tuple(p.shape)))
                continue
            with torch.no_grad():
                p.copy_(src)
        if strict:
            for full in state_dict.keys():
                if full.startswith(prefix) and full not in known:
                    unexpected_keys.append(full)


class LXRTFeatureExtraction(RefKeyParams):
    """BERT model for classification (reference: modeling.py:1005-1030).  mode 'x' only (pooled output)."""

    def __init__(self, config=None, mode='lxr', precision="bf16"):
        self.config = config if config is not None else BertConfig()
        if mode != 'x':
            raise NotImplementedError("only mode='x' (pooled cross-modality output) is on the evaluation path "
                                      "(src/lxrt/entry.py:80, src/tasks/gqa_model.py:17)")
        self.mode = mode
        self.precision = precision
        self.visual_config = VisualConfig(VISUAL_CONFIG.l_layers, VISUAL_CONFIG.x_layers, VISUAL_CONFIG.r_layers)
        spec = _syn.lxmert_spec(self.visual_config.l_layers, self.visual_config.x_layers, self.visual_config.r_layers,
                                num_answers=0, vocab=self.config.vocab_size, max_pos=self.config.max_position_embeddings)
        super().__init__(spec, strip_prefix=_syn.LXMERT_PREFIX)
        self._engine = None
        self._engine_fp = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, *inputs, **kwargs):
        """The reference downloads bert-base-uncased from S3 (modeling.py:730-747); offline this builds the
        randomly initialised model — load real weights with load_state_dict / LXRTEncoder.load."""
        kwargs.pop("cache_dir", None)
        kwargs.pop("state_dict", None)
        return cls(BertConfig(), *inputs, **kwargs)

    def engine_config(self, num_answers=0):
        v = self.visual_config
        return EngineConfig(arch="lxmert", l_layers=v.l_layers, x_layers=v.x_layers, r_layers=v.r_layers,
                            vocab_size=self.config.vocab_size, max_position=self.config.max_position_embeddings,
                            type_vocab=self.config.type_vocab_size, visual_feat_dim=v.visual_feat_dim,
                            visual_pos_dim=v.visual_pos_dim, num_answers=num_answers, precision=self.precision)

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
                visual_attention_mask=None):
        if visual_attention_mask is not None:
            raise NotImplementedError("the evaluation path never masks RoIs (src/lxrt/entry.py:109)")
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)      # modeling.py:847-848
        feat, pos = visual_feats
        pooled, _ = self._encoder_engine().forward(feat, pos, input_ids, token_type_ids, attention_mask, want_logits=False)
        return pooled
