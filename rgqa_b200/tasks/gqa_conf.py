"""Evaluation driver with the reference's command line (``src/tasks/gqa_conf.py:384-453`` test branch, and the
energy / Mahalanobis variants ``gqa_energy.py``, ``gqa_maha.py`` selected with ``--method``):

    python -m rgqa_b200.tasks.gqa_conf --test ClipEasy --load snap/gqa/x/BEST --batchSize 1024 --method msp \
        --output snap/out [--precision bf16] [--syntheticN 11158]
    torchrun --nproc-per-node 8 -m rgqa_b200.tasks.gqa_conf ...        # batch-sharded sweep, one all-gather

Writes ``<output>/<test>_predict.json`` and ``<output>/<test>_result.json`` like the reference (:406-415).
No RGQA data, features or checkpoints exist offline: without ``--load`` the weights are the seeded random init
and the data set is ``SyntheticRGQA`` (seeded features / questions / labels)."""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.distributed as dist

from ..param import get_parser
from .gqa_data import dump_results_json
from .gqa_eval import GQAEval, SyntheticRGQA, shard_range
from .gqa_model import GQAModel_maha


class GQA:
    def __init__(self, args, num_answers=1842):
        """num_answers: the answer table WITHOUT the extra 'UQ' class (gqa_conf.py:80: ``GQAModel(num_answers - 1)``)."""
        self.args = args
        if args.backbone == "uniter":
            from ..uniter.uniter import GQAUNITER_maha
            self.model = GQAUNITER_maha(num_answers, args=args)
        elif args.backbone == "lxmert":
            self.model = GQAModel_maha(num_answers, args=args, max_seq_length=args.max_seq_length)
        else:
            raise ValueError("Backbone is not defined")  # gqa_conf.py:91
        self.output = args.output
        os.makedirs(self.output, exist_ok=True)

    def load(self, path):
        """gqa_conf.py:375-381: strip '.module', non-strict load."""
        from ..lxrt.entry import require_real_vocab
        enc = getattr(self.model, "lxrt_encoder", None) or getattr(self.model, "encoder")
        require_real_vocab(enc.tokenizer, "GQA.load")
        print("Load model from %s" % path)
        state_dict = torch.load("%s.pth" % path, map_location="cpu")
        state_dict = {k.replace(".module", ""): v for k, v in state_dict.items()}
        self.model.load_state_dict(state_dict, strict=False)

    def save(self, name):
        torch.save(self.model.state_dict(), os.path.join(self.output, "%s.pth" % name))

    def train(self, loader, img_of, steps_total=None, epochs=None, gemm=None):
        """``GQA.train`` (gqa_conf.py:140-202) on the GPU: BCE x A, clip 5.0, BertAdam(lr, warmup 0.1, t_total) (:123-127), RP
        negatives with ``--sample_pair`` (:155-170).  ``loader`` yields ``(ques_id, feats, boxes, sent, target)`` host batches
        (``target`` with the trailing 'UQ' column, dropped as :153 does); ``img_of(qid)`` -> image id.  Returns the per-step
        losses; the trained parameters are loaded back into ``self.model``."""
        from ..lxrt.trainer import LXMERTTrainer
        from .rp import build_rp_batch
        args = self.args
        epochs = args.epochs if epochs is None else epochs
        batches_per_epoch = len(loader)
        t_total = int(batches_per_epoch * epochs) if steps_total is None else steps_total        # :122
        if gemm is None:
            gemm = "bf16" if args.precision == "bf16" else "fp32"
        dev = torch.device(f"cuda:{torch.cuda.current_device()}")
        tr = LXMERTTrainer(self.model.state_dict(), l_layers=args.llayers, x_layers=args.xlayers, r_layers=args.rlayers, device=dev,
                           lr=args.lr, warmup=0.1, t_total=t_total, dropout=args.dropout, seed=args.seed, gemm=gemm)
        tok, L = self.model.lxrt_encoder.tokenizer, self.model.lxrt_encoder.max_seq_length
        losses = []
        for _ in range(epochs):
            for ques_id, feats, boxes, sent, target in loader:
                target = target[:, :-1]
                ques_id, sent = list(ques_id), list(sent)
                if args.sample_pair:
                    ques_id, feats, boxes, sent, target = build_rp_batch(ques_id, feats, boxes, sent, target, img_of)
                ids, mask, seg = tok.encode_batch(sent, L)
                loss, _ = tr.train_step(feats, boxes, torch.from_numpy(ids), torch.from_numpy(seg), torch.from_numpy(mask), target)
                losses.append(loss)
        self.model.load_state_dict(tr.state_dict(), strict=True)
        return losses

    def ood_evaluate(self, dataset, batch_size=None, method=None, dump=None, maha=None):
        """Two call forms.  ``ood_evaluate(dataset, batch_size, method, dump=, maha=)``: the batch-sharded sweep over a pre-tokenised
        data set (``GQAEval``).  ``ood_evaluate(eval_tuple, dump=)`` / ``ood_evaluate(eval_tuple, (mean, precision), dump=)``: the
        reference's own signature (gqa_conf.py:296, gqa_energy.py:194, gqa_maha.py:191) — ``eval_tuple = (dset, loader, _)`` whose
        loader yields ``(ques_id, feats, boxes, sent, ...)`` host batches with SENTENCES; the scorer is ``args.method``."""
        if isinstance(dataset, tuple) and len(dataset) == 3:
            pack = batch_size if batch_size is not None else maha
            return self._ood_evaluate_tuple(dataset, self.args.method if method is None else method, dump, pack)
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        lo, hi = shard_range(dataset.n, rank, world)
        if method == "maha":
            if maha is None:
                raise RuntimeError("Mahalanobis needs {mean, precision} (sample_estimator pickle, gqa_maha.py:185-188)")
            self.model.engine().set_maha(*maha)
        ev = GQAEval(self.model, dataset, method=method, tau=self.args.tau)
        return ev.ood_evaluate(dataset.loader(batch_size, lo, hi), dump=dump if rank == 0 else None)


    def _ood_evaluate_tuple(self, eval_tuple, method, dump, pack):
        """The reference loop (gqa_conf.py:301-330): model call on each loader batch, device scorer, ``quesid2ans`` in loader order,
        ``GQAOODEvaluator.evaluate`` + ``dump_result``."""
        from .gqa_data import GQAOODEvaluator
        dset, loader, _ = eval_tuple
        if method == "maha":
            if pack is None:
                raise RuntimeError("Mahalanobis needs {mean, precision} (sample_estimator pickle, gqa_maha.py:185-188)")
            mean, precision = (torch.as_tensor(t).float() for t in pack)     # gqa_maha.py:194-201
            self.model.engine().set_maha(mean, precision)
        dev = next(self.model.parameters()).device
        evaluator = GQAOODEvaluator(dset, self.args.tau)
        quesid2ans = {}
        for datum_tuple in loader:
            ques_id, feats, boxes, sent = datum_tuple[:4]   # avoid handling target
            logit, x = self.model(feats.to(dev), boxes.to(dev), list(sent))
            score, label = self.model.engine().score(method, logits=logit, pooled=x)
            for qid, l, s in zip(ques_id, label.cpu().numpy(), score.cpu().numpy()):
                quesid2ans[qid] = (dset.label2ans[l], s)
        results = evaluator.evaluate(quesid2ans)
        if dump is not None:
            evaluator.dump_result(quesid2ans, dump)
        return results


def get_tuple(args, splits, tokenizer, max_seq_length, rank=0, world=1):
    """``get_tuple`` of the reference (gqa_conf.py:34-60) for the evaluation flows: the GQA-UQ data set (+ the 'UQ' class), this
    rank's contiguous shard of it as a ``SweepLoader`` over the memory-mapped feature store, and the evaluator."""
    from ..ingest import SweepLoader
    from .gqa_data import FAST_IMG_NUM, TINY_IMG_NUM, GQADataset, GQAOODEvaluator, open_feature_store
    dset = GQADataset(splits, root=args.data_root)
    dset.add_uq_class()
    topk = TINY_IMG_NUM if args.tiny else (FAST_IMG_NUM if args.fast else None)
    store = open_feature_store(dset.splits, root=args.data_root, backbone=args.backbone, dtype=args.feature_dtype, topk=topk,
                               store_dir=args.feature_store)
    rows = [d for d in dset.data if d["img_id"] in store]       # GQATorchDataset keeps only questions whose image was loaded
    print("Use %d data in torch dataset" % len(rows))
    lo, hi = shard_range(len(rows), rank, world)
    loader = SweepLoader(store, rows, tokenizer, max_seq_length=max_seq_length, batch_size=args.batch_size, lo=lo, hi=hi)
    return dset, loader, GQAOODEvaluator(dset, args.tau)


def main(argv=None):
    p = argparse.ArgumentParser(parents=[get_parser()])
    p.add_argument("--syntheticN", type=int, default=None, help="evaluate on N synthetic questions instead of a data set")
    p.add_argument("--dataRoot", dest="data_root", default="data", help="where gqa/<split>.json and vg_gqa_imgfeat/ live")
    p.add_argument("--featureStore", dest="feature_store", default=None, help="a converted feature store (rgqa_b200.ingest)")
    p.add_argument("--featureDtype", dest="feature_dtype", default="bf16", choices=["bf16", "f32"])
    p.add_argument("--maha", default=None, help="pickle {mean, precision} written by sample_estimator (gqa_maha.py:185-188)")
    args, _ = p.parse_known_args(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("rgqa_b200 needs a CUDA device; there is no CPU path")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("RGQA_SINGLE_GPU") == "1":   # several ranks on one GPU (tests): every rank drives device 0
        local = 0
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        backend = os.environ.get("RGQA_DIST_BACKEND", "nccl")
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    rank = dist.get_rank() if dist.is_initialized() else 0
    if args.max_seq_length is None:
        args.max_seq_length = 30  # MAX_GQA_LENGTH, gqa_model.py:11
    if args.syntheticN is not None and args.vocab is None:
        args.vocab = "synthetic"     # explicit: a synthetic data set has no real sentences to tokenise
    num_answers = 1842
    if args.syntheticN is None and args.test is not None:
        import json
        with open(os.path.join(args.data_root, "gqa", "trainval_label2ans.json")) as f:
            num_answers = len(json.load(f))
    gqa = GQA(args, num_answers)
    if args.load is not None:
        gqa.load(args.load)
    gqa.model = gqa.model.cuda()
    test = args.test or "synthetic"
    L = 20 if args.backbone == "uniter" else args.max_seq_length   # UniterEncoder hard-codes 20 (uniter/entry.py:66)
    dump = os.path.join(args.output, "%s_predict.json" % test)
    maha = None
    if args.method == "maha":
        if args.maha is not None:
            import pickle
            with open(args.maha, "rb") as f:
                pack = pickle.load(f)
            maha = (pack["mean"], pack["precision"])
        elif args.syntheticN is not None:
            from .. import synthetic as syn
            maha = syn.maha_stats()
        else:
            raise RuntimeError("--method maha needs --maha <sample_estimator pickle> (gqa_maha.py:185-188)")
    if args.syntheticN is not None:
        if args.backbone == "uniter":
            raise SystemExit("the synthetic data set carries 4-d boxes; drive GQAUNITER through a real feature store (7-d boxes)")
        dataset = SyntheticRGQA(args.syntheticN, L=L, seed=args.seed)
        results = gqa.ood_evaluate(dataset, args.batch_size, args.method, dump=dump, maha=maha)
    else:
        if args.test is None:
            raise SystemExit("--test SPLIT (a file <dataRoot>/gqa/SPLIT.json) or --syntheticN N is required")
        enc = getattr(gqa.model, "lxrt_encoder", None) or gqa.model.encoder
        dset, loader, _ = get_tuple(args, args.test, enc.tokenizer, L, rank, world)
        if maha is not None:
            gqa.model.engine().set_maha(*maha)
        ev = GQAEval(gqa.model, dset, method=args.method, tau=args.tau)
        results = ev.ood_evaluate(loader, dump=dump if rank == 0 else None)
    if results is not None:
        print(results)
        dump_results_json(results, os.path.join(args.output, "%s_result.json" % test))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
