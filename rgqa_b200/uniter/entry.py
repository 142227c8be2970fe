"""Mirror of ``src/uniter/entry.py:63-99``: ``UniterEncoder(args).forward(sents, feats, boxes)`` (L = 20 hard-coded,
36 RoIs asserted, 7-d boxes)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..lxrt.entry import make_tokenizer
from .modeling import UniterFeatureExtraction


class UniterEncoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.max_seq_length = 20
        self.tokenizer = make_tokenizer(args, "bert-base-cased", vocab_size=28996)
        self.model = UniterFeatureExtraction.from_pretrained("bert-base-cased", precision=getattr(args, "precision", "bf16"))
        if getattr(args, "from_scratch", False):
            print("initializing all the weights")
            self.model.init_bert_weights()

    @property
    def dim(self):
        return 768

    def tokenize(self, sents, device):
        ids, mask, seg = self.tokenizer.encode_batch(list(sents), self.max_seq_length)
        pin = torch.cuda.is_available()
        to = lambda a: (torch.from_numpy(a).pin_memory() if pin else torch.from_numpy(a)).to(device, non_blocking=True)
        return to(ids), to(mask), to(seg)

    def forward(self, sents, feats, boxes, visual_attention_mask=None):
        assert feats.shape[1] == 36, "Not Using 36 ROIs, please change the following 2 lines"  # entry.py:91
        ids, mask, seg = self.tokenize(sents, feats.device)
        return self.model(input_ids=ids, token_type_ids=seg, attention_mask=mask, visual_feats=feats, img_pos_feat=boxes)
