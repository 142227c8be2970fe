"""Random-pairing (RP) pseudo-unanswerable batches — ``GQA.train`` with ``--samplePair`` (src/tasks/gqa_conf.py:155-170).

Every question of the batch is paired with the image of the same row and the SENTENCE of a randomly drawn other row
whose image differs; the batch doubles: rows [B, 2B) carry the same features / boxes, the drawn sentences and all-zero
targets (the head is trained with BCE, so an all-zero target is the 'reject' label)."""
from __future__ import annotations

import random

import torch


def sample_pairs(ques_id, img_of, rng=random):
    """Indices of the negative sentence for every row (:157-163): ``rng.choice`` over the batch's question ids, redrawn
    while the drawn question shares the row's image.  ``img_of``: question id -> image id (``dset.id2datum[q]['img_id']``).
    Raises if every question of the batch shares one image (the reference would loop forever)."""
    if len({img_of(q) for q in ques_id}) < 2:
        raise ValueError("random pairing needs at least two different images in the batch")
    out = []
    for qid in ques_id:
        qid_r = rng.choice(ques_id)
        while img_of(qid_r) == img_of(qid):
            qid_r = rng.choice(ques_id)
        out.append(ques_id.index(qid_r))
    return out


def build_rp_batch(ques_id, feats, boxes, sent, target, img_of, rng=random):
    """(ques_id, feats, boxes, sent, target) of size 2B exactly as :155-170 build it (target already without the last,
    'UQ', column: :153)."""
    ind = sample_pairs(ques_id, img_of, rng)
    neg_sent = [sent[i] for i in ind]
    neg_qid = [img_of(q) + "+" + ques_id[i] for q, i in zip(ques_id, ind)]
    return (list(ques_id) + neg_qid, feats.repeat(2, 1, 1), boxes.repeat(2, 1, 1), list(sent) + neg_sent,
            torch.cat([target, torch.zeros_like(target)], 0))
