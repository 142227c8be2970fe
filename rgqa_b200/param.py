"""Flags of the hot path, same names and defaults as the reference's ``src/param.py:33-150``.

The reference parses ``sys.argv`` at import time into a module-level ``args``; task and model files
do ``from param import args``.  This module keeps that contract (``from rgqa_b200.param import args``)
for the flags the evaluation path reads, ignores the training-recipe flags it does not know
(``parse_known_args``), and adds ``--precision`` and ``--maxSeqLength``.
"""
from __future__ import annotations

import argparse
import os
import random
import sys

import numpy as np
import torch


def get_parser():
    p = argparse.ArgumentParser(add_help=False)
    # data / run selection (param.py:84-101)
    p.add_argument("--train", default="train")
    p.add_argument("--valid", default="valid")
    p.add_argument("--test", default=None)
    p.add_argument("--backbone", default="lxmert", choices=["lxmert", "uniter"])
    p.add_argument("--batchSize", dest="batch_size", type=int, default=256)
    p.add_argument("--seed", type=int, default=9595)
    p.add_argument("--output", type=str, default="snap/test")
    p.add_argument("--fast", action="store_const", default=False, const=True)
    p.add_argument("--tiny", action="store_const", default=False, const=True)
    p.add_argument("--load", type=str, default=None)
    p.add_argument("--loadLXMERT", dest="load_lxmert", type=str, default=None)
    p.add_argument("--loadLXMERTQA", dest="load_lxmert_qa", type=str, default=None)
    p.add_argument("--fromScratch", dest="from_scratch", action="store_const", default=False, const=True)
    # model (param.py:118-120)
    p.add_argument("--llayers", default=9, type=int)
    p.add_argument("--xlayers", default=5, type=int)
    p.add_argument("--rlayers", default=5, type=int)
    # scorers (param.py:41-53)
    p.add_argument("--tau", type=float, default=0.5)
    p.add_argument("--noise", type=float, default=0.0)
    p.add_argument("--temperature", type=float, default=1.0)
    # flows beyond the per-batch body: ensemble (param.py:60), RP finetune (:67, :88-91), loss variant (:114)
    p.add_argument("--ensemble_method", default="mean", type=str)
    p.add_argument("--sample_pair", action="store_const", default=False, const=True)
    p.add_argument("--optim", default="bert")
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--mceLoss", dest="mce_loss", action="store_const", default=False, const=True)
    p.add_argument("--tqdm", action="store_const", default=False, const=True)
    # runtime (param.py:133-134)
    p.add_argument("--multiGPU", action="store_const", default=False, const=True)
    p.add_argument("--numWorkers", dest="num_workers", default=0, type=int)
    # additions of this build
    p.add_argument("--precision", default=os.environ.get("RGQA_PRECISION", "bf16"), choices=["bf16", "fp32"])
    p.add_argument("--maxSeqLength", dest="max_seq_length", type=int, default=None)
    p.add_argument("--vocab", type=str, default=os.environ.get("RGQA_BERT_VOCAB", None),
                   help="vocab.txt (or directory); without it a deterministic SyntheticTokenizer is used")
    p.add_argument("--method", default="msp", choices=["msp", "energy", "energy_sum", "lse", "maha", "odin", "ensemble"])
    return p


def parse_args(argv=None):
    args, _unknown = get_parser().parse_known_args(sys.argv[1:] if argv is None else argv)
    # the reference seeds everything at import (param.py:143-145)
    torch.manual_seed(args.seed)
    random.seed(args.seed)
    np.random.seed(args.seed)
    return args


args = parse_args()
