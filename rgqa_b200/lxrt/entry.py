"""Host-side mirror of ``src/lxrt/entry.py``: tokenise the questions, build id/mask/segment tensors, call
the model.  Same class, attributes (``.model``, ``.tokenizer``, ``.max_seq_length``, ``.dim``) and methods
(``forward``, ``load``, ``save``, ``multi_gpu``) as the reference's ``LXRTEncoder`` (:80-152)."""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn

from .modeling import VISUAL_CONFIG, LXRTFeatureExtraction
from .tokenization import BertTokenizer, SyntheticTokenizer


class InputFeatures(object):
    def __init__(self, input_ids, input_mask, segment_ids):
        self.input_ids, self.input_mask, self.segment_ids = input_ids, input_mask, segment_ids


def convert_sents_to_features(sents, max_seq_length, tokenizer):
    """entry.py:36-71 — kept for callers that want the per-sentence objects; the model path uses the batched
    ``tokenizer.encode_batch`` which produces the same numbers."""
    ids, mask, seg = tokenizer.encode_batch(list(sents), max_seq_length)
    return [InputFeatures(ids[i].tolist(), mask[i].tolist(), seg[i].tolist()) for i in range(len(sents))]


def set_visual_config(args):
    VISUAL_CONFIG.l_layers = args.llayers
    VISUAL_CONFIG.x_layers = args.xlayers
    VISUAL_CONFIG.r_layers = args.rlayers


def make_tokenizer(args, name="bert-base-uncased", vocab_size=30522):
    """The reference downloads the vocabulary from S3 and fails loudly without it (tokenization.py:27-35).  Here the
    vocabulary comes from ``--vocab`` / ``RGQA_BERT_VOCAB`` (a file or directory) or ``RGQA_BERT_VOCAB_DIR``.  The hash-based
    ``SyntheticTokenizer`` is used only when asked for (``--vocab synthetic``) or, WITH A WARNING, for models that carry no
    checkpoint (random-init parity / benchmark runs); as soon as a checkpoint is involved (``--load*`` flags, ``load()``)
    an unresolved vocabulary is an error: trained embeddings fed hash ids would produce garbage silently."""
    vocab = getattr(args, "vocab", None)
    if vocab == "synthetic":
        return SyntheticTokenizer(vocab_size)
    if vocab:
        return BertTokenizer.from_pretrained(vocab, do_lower_case=True)
    try:
        return BertTokenizer.from_pretrained(name, do_lower_case=True)
    except EnvironmentError as e:
        if any(getattr(args, k, None) for k in ("load", "load_lxmert", "load_lxmert_qa")):
            raise EnvironmentError(
                f"a checkpoint is being loaded but the '{name}' vocabulary cannot be resolved: pass --vocab <vocab.txt> or set "
                "RGQA_BERT_VOCAB_DIR (use --vocab synthetic only for random-init runs)") from e
        warnings.warn(f"no '{name}' vocabulary available offline: using the deterministic SyntheticTokenizer "
                      "(random-init runs only; loading a checkpoint into this model will be refused)", stacklevel=2)
        tok = SyntheticTokenizer(vocab_size)
        tok.is_fallback = True
        return tok


def require_real_vocab(tokenizer, what):
    """Refuse to put trained weights behind the fallback tokenizer (see make_tokenizer)."""
    if getattr(tokenizer, "is_fallback", False):
        raise EnvironmentError(f"{what}: the model was built without a vocabulary (SyntheticTokenizer fallback); pass --vocab "
                               "<vocab.txt> / set RGQA_BERT_VOCAB_DIR, or --vocab synthetic if hash ids are really intended")


class LXRTEncoder(nn.Module):
    def __init__(self, args, max_seq_length, mode='x'):
        super().__init__()
        self.max_seq_length = max_seq_length
        set_visual_config(args)
        self.tokenizer = make_tokenizer(args)
        self.model = LXRTFeatureExtraction.from_pretrained("bert-base-uncased", mode=mode,
                                                           precision=getattr(args, "precision", "bf16"))
        if getattr(args, "from_scratch", False):
            print("initializing all the weights")
            self.model.init_bert_weights()

    def multi_gpu(self):
        """The reference wraps the encoder in single-process nn.DataParallel (entry.py:102-103).  This build
        scales as one process per GPU (torchrun + batch sharding, rgqa_b200.tasks.gqa_eval), so this is a no-op
        kept for call-site compatibility."""
        return None

    @property
    def dim(self):
        return 768

    def tokenize(self, sents, device):
        ids, mask, seg = self.tokenizer.encode_batch(list(sents), self.max_seq_length)
        pin = torch.cuda.is_available()
        to = lambda a: (torch.from_numpy(a).pin_memory() if pin else torch.from_numpy(a)).to(device, non_blocking=True)
        return to(ids), to(mask), to(seg)

    def forward(self, sents, feats, visual_attention_mask=None):
        feat, pos = feats
        input_ids, input_mask, segment_ids = self.tokenize(sents, feat.device)
        return self.model(input_ids, segment_ids, input_mask, visual_feats=feats,
                          visual_attention_mask=visual_attention_mask)

    def save(self, path):
        torch.save(self.model.state_dict(), os.path.join("%s_LXRT.pth" % path))

    def load(self, path):
        """entry.py:126-152: strip a DataParallel ``module.`` prefix, report key differences, load non-strictly."""
        require_real_vocab(self.tokenizer, "LXRTEncoder.load")
        print("Load LXMERT pre-trained model from %s" % path)
        state_dict = torch.load("%s_LXRT.pth" % path, map_location="cpu")
        state_dict = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        load_keys, model_keys = set(state_dict.keys()), set(self.model.state_dict().keys())
        print("\nWeights in loaded but not in model:")
        for key in sorted(load_keys.difference(model_keys)):
            print(key)
        print("\nWeights in model but not in loaded:")
        for key in sorted(model_keys.difference(load_keys)):
            print(key)
        print()
        self.model.load_state_dict(state_dict, strict=False)
