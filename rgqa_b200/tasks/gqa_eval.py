"""The evaluation sweep: host loader -> H2D -> model -> rejection scorer -> D2H -> {qid: (answer, score)}.

Mirror of ``GQA.ood_evaluate`` in the reference's task files (``src/tasks/gqa_conf.py:297-333`` MSP,
``src/tasks/gqa_energy.py:193-213`` energy, ``src/tasks/gqa_maha.py:191-279`` Mahalanobis with noise = 0):
same inputs per batch ``(ques_id, feats, boxes, sent)``, same ``quesid2ans`` dictionary, same evaluator and
JSON dump.  What changes is the plumbing around the (unchanged) per-batch contract:

  * the reference does a blocking ``feats.cuda()`` and a ``.cpu().numpy()`` sync per batch (gqa_conf.py:307,327);
    here batches are uploaded from pinned memory on a copy stream into a two-slot device ring while the previous
    batch computes, and (score, label) stay on the device until the end of the sweep;
  * multi-GPU is one process per GPU: each rank sweeps a contiguous shard of the question list and the sweep
    ends with ONE all-gather of (score f32, label i32) — the reference's only multi-GPU mechanism is
    single-process nn.DataParallel (src/lxrt/entry.py:102-103).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .gqa_data import GQAOODEvaluator


def shard_range(n, rank, world):
    """Contiguous index range [lo, hi) of rank's shard of n questions (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SweepRunner:
    """Runs batches through ``model`` + scorer with overlapped H2D, keeps results on the device."""

    def __init__(self, model, method="msp", device=None, fused=True):
        self.model = model
        self.method = method
        self.fused = fused  # False: logits / pooled leave the forward and the stand-alone scorer kernels run (A/B, tests)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]
        self._free = [torch.cuda.Event(), torch.cuda.Event()]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._n = 0
        self.h2d_bytes = 0

    def _upload(self, slot, feats, boxes, ids, mask, seg):
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[slot])  # the compute that last read this slot is done
            bufs = self._slots[slot]
            if bufs is None or bufs[0].shape != feats.shape or bufs[0].dtype != feats.dtype or bufs[2].shape != ids.shape:
                bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in (feats, boxes, ids, mask, seg))
                self._slots[slot] = bufs
            for dst, src in zip(bufs, (feats, boxes, ids, mask, seg)):
                dst.copy_(src, non_blocking=True)
                self.h2d_bytes += src.numel() * src.element_size()
            self._ready[slot].record(self.copy_stream)
        del cur
        return bufs

    def submit(self, feats, boxes, ids, mask, seg=None):
        """Enqueue one batch given HOST tensors (pinned for true overlap); returns device (score, label)."""
        slot = self._n & 1
        self._n += 1
        if seg is None:
            seg = torch.zeros_like(ids)
        f, b, i, m, s = self._upload(slot, feats, boxes, ids, mask, seg)
        stream = torch.cuda.current_stream(self.device)
        stream.wait_event(self._ready[slot])
        out = self.score_device_batch(f, b, i, m, s)
        self._free[slot].record(stream)
        # Bound the host's run-ahead to ONE batch: a loader that recycles two pinned buffer sets (ingest.SweepLoader) refills
        # the set of batch k when it builds batch k+2, so the async H2D of batch k must have executed before submit(k+1)
        # returns.  Waiting for the PREVIOUS upload here keeps this batch's copy overlapped with the previous compute.
        if self._n >= 2:
            self._ready[slot ^ 1].synchronize()
        return out

    def drain(self):
        """Block until every upload issued so far has executed (call before reusing the LAST batch's host buffers)."""
        for ev in self._ready:
            ev.synchronize()

    def score_device_batch(self, feats, boxes, ids, mask, seg=None):
        """One forward + scorer on device tensors: the reference's per-batch body (gqa_conf.py:305-326)."""
        many = isinstance(self.method, (tuple, list))  # several scorers over ONE forward: [(score, label), ...]
        methods = list(self.method) if many else [self.method]
        if self.fused and hasattr(self.model, "forward_scored_ids"):
            # scorer reductions inside the forward (answer-head GEMM epilogue, tensor-core Mahalanobis): one engine call
            res = self.model.forward_scored_ids(feats, boxes, ids, seg, mask, methods)
        else:
            logits, pooled = self.model.forward_ids(feats, boxes, ids, seg, mask)
            eng = self.model.engine()
            res = [eng.score(m, logits=logits, pooled=pooled) for m in methods]
        return res if many else res[0]


class GQAEval:
    """``ood_evaluate`` for a model of this package.  ``dataset`` needs ``label2ans`` (index -> answer string)
    and ``id2datum`` (qid -> {'label': {...}}), like the reference's GQADataset."""

    def __init__(self, model, dataset, method="msp", tau=0.5):
        self.model, self.dataset, self.method, self.tau = model, dataset, method, tau

    def ood_evaluate(self, loader, dump=None, gather=True):
        """loader yields (ques_ids, feats, boxes, ids, mask) host batches of THIS rank's shard, in order.
        Returns the result dict on rank 0 (None elsewhere when gather=True and world_size > 1)."""
        runner = SweepRunner(self.model, self.method)
        qids, scores, labels = [], [], []
        for ques_id, feats, boxes, ids, mask in loader:
            s, l = runner.submit(feats, boxes, ids, mask)
            qids.extend(ques_id)
            scores.append(s)
            labels.append(l)
        score = torch.cat(scores) if scores else torch.empty(0, device=runner.device)
        label = torch.cat(labels) if labels else torch.empty(0, dtype=torch.int32, device=runner.device)
        score, label, qids = gather_results(score, label, qids) if gather else (score.cpu(), label.cpu(), qids)
        if score is None:
            return None
        quesid2ans = build_quesid2ans(qids, label.numpy(), score.numpy(), self.dataset.label2ans)
        evaluator = GQAOODEvaluator(self.dataset, self.tau)
        results = evaluator.evaluate(quesid2ans)
        if dump is not None:
            evaluator.dump_result(quesid2ans, dump)
        return results


def build_quesid2ans(qids, labels, scores, label2ans):
    """{qid: (answer string, score)} in loader order (gqa_conf.py:327-329)."""
    return {qid: (label2ans[int(l)], s) for qid, l, s in zip(qids, labels, scores)}


def gather_results(score, label, qids=None):
    """One all-gather of (score f32, label i32) across ranks (variable shard sizes padded to the max);
    returns CPU tensors (and the concatenated qid list) on rank 0, (None, None, None) elsewhere.
    Works on NCCL (device tensors) and gloo (CPU tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return score.cpu(), label.cpu(), qids
    world, rank = dist.get_world_size(), dist.get_rank()
    backend_dev = score.device if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([score.numel()], dtype=torch.int64, device=backend_dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    # pack (score bits, label) into one int32 [nmax, 2] payload: ONE data all-gather
    payload = torch.zeros((nmax, 2), dtype=torch.int32, device=backend_dev)
    payload[:score.numel(), 0] = score.to(backend_dev).contiguous().view(torch.int32)
    payload[:score.numel(), 1] = label.to(backend_dev)
    out = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(out, payload)
    all_qids = None
    if qids is not None:
        bucket = [None] * world
        dist.all_gather_object(bucket, list(qids))
        all_qids = [q for part in bucket for q in part]
    if rank != 0:
        return None, None, None
    cat = torch.cat([o[:c] for o, c in zip(out, counts)]).cpu()
    return cat[:, 0].contiguous().view(torch.float32), cat[:, 1].contiguous(), all_qids


class SyntheticRGQA:
    """Synthetic stand-in for an RGQA subset (no data offline): N questions with seeded features / ids, a label
    table where a fraction are 'UQ', and a 1842-entry label2ans.  Host tensors are pinned when CUDA exists."""

    def __init__(self, n, L=20, seed=1234, num_answers=1842, uq_frac=0.5, feat_dtype=torch.float32, pool=256):
        from .. import synthetic as syn
        self.n, self.L = n, L
        self.label2ans = [f"ans{i}" for i in range(num_answers)]
        g = np.random.default_rng(seed)
        self.qids = [str(10_000_000 + i) for i in range(n)]
        is_uq = g.random(n) < uq_frac
        gold = g.integers(0, num_answers, n)
        self.id2datum = {q: {"question_id": q, "label": ({"UQ": 1.0} if u else {self.label2ans[a]: 1.0})}
                         for q, u, a in zip(self.qids, is_uq, gold)}
        # a pool of distinct feature rows reused cyclically keeps host memory bounded (302 KB / question in fp32)
        self.pool = min(pool, n)
        inp = syn.make_inputs(self.pool, L, seed=seed)
        pin = torch.cuda.is_available()
        self.feat = inp["feat"].to(feat_dtype)
        self.pos = inp["pos"]
        self.ids = inp["ids"].to(torch.int32)
        self.mask = inp["mask"].to(torch.int32)
        if pin:
            self.feat, self.pos, self.ids, self.mask = (t.pin_memory() for t in (self.feat, self.pos, self.ids, self.mask))

    def loader(self, batch_size, lo=0, hi=None):
        hi = self.n if hi is None else hi
        for s in range(lo, hi, batch_size):
            e = min(s + batch_size, hi)
            idx = torch.arange(s, e) % self.pool
            contiguous = int(idx[0]) + (e - s) <= self.pool
            if contiguous:
                a = int(idx[0])
                sl = slice(a, a + (e - s))
                yield self.qids[s:e], self.feat[sl], self.pos[sl], self.ids[sl], self.mask[sl]
            else:
                yield self.qids[s:e], self.feat[idx], self.pos[idx], self.ids[idx], self.mask[idx]
