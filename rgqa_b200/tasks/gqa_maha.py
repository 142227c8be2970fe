"""Class-conditional Gaussian statistics for the Mahalanobis scorer — ``GQA.sample_estimator`` of the reference
(``src/tasks/gqa_maha.py:120-189``) with the data pass on the GPU.

Reference: pooled feature of every training sample -> host; a Python double loop over (sample, 1841 classes) appends it to
the list of every class it is labelled with (:158-168); class means (:170-173); the class-centred features are stacked
(:175-179) and ``sklearn.covariance.EmpiricalCovariance(assume_centered=False).fit(X).precision_`` is taken (:182-183).

Here each batch only adds to three fp64 device accumulators (``rgqa_estimator_accumulate``): ``n_c``, ``s_c = sum x`` and
``G = sum_i w_i x_i x_i^T``; ranks all-reduce them; the finish is closed-form in those sums:

    mean_c = fp32(s_c / n_c)                         (zero row for an empty class, like :170-173)
    X^T X  = G - sum_c (s_c m_c^T + m_c s_c^T - n_c m_c m_c^T)          with m_c = mean_c
    cov    = X^T X / N - loc loc^T,  loc = sum_c (s_c - n_c m_c) / N,  N = sum_c n_c        (EmpiricalCovariance)
    precision = scipy.linalg.pinvh(cov)                                  (what sklearn's ``precision_`` calls)
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
import torch.distributed as dist

from .. import ops


class MahaEstimator:
    def __init__(self, num_classes, dim=768, device=None):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.C, self.D = num_classes, dim
        self.G = torch.zeros((dim, dim), dtype=torch.float64, device=self.device)
        self.class_sum = torch.zeros((num_classes, dim), dtype=torch.float64, device=self.device)
        self.class_count = torch.zeros(num_classes, dtype=torch.float64, device=self.device)

    def accumulate(self, pooled, target):
        """pooled [B, D] f32 (device); target [B, >= C] multi-hot (the reference drops the last, 'UQ', class: :141)."""
        target = target.to(self.device, dtype=torch.float32, non_blocking=True)
        ops.estimator_accumulate(pooled.contiguous(), target, self.G, self.class_sum, self.class_count)

    def all_reduce(self):
        """Sum the accumulators over the ranks of a sharded pass (one collective over a packed fp64 buffer)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = torch.cat([self.G.flatten(), self.class_sum.flatten(), self.class_count])
            if dist.get_backend() != "nccl":
                packed = packed.cpu()
            dist.all_reduce(packed)
            packed = packed.to(self.device)
            d2, cd = self.D * self.D, self.C * self.D
            self.G = packed[:d2].view(self.D, self.D).clone()
            self.class_sum = packed[d2:d2 + cd].view(self.C, self.D).clone()
            self.class_count = packed[d2 + cd:].clone()

    def finish(self):
        """(sample_class_mean f32 torch [C, D], precision f64 numpy [D, D]) — the pair the reference pickles (:185-188)."""
        return finish_estimator(self.G.cpu().numpy(), self.class_sum.cpu().numpy(), self.class_count.cpu().numpy())


def finish_estimator(G, class_sum, class_count):
    import scipy.linalg
    n = class_count.astype(np.float64)
    seen = n > 0
    mean = np.zeros_like(class_sum, dtype=np.float32)
    mean[seen] = (class_sum[seen] / n[seen, None]).astype(np.float32)
    m = mean.astype(np.float64)
    sm = class_sum.T @ m                                   # sum_c s_c m_c^T
    xtx = G - sm - sm.T + (m * n[:, None]).T @ m
    N = n.sum()
    loc = (class_sum - m * n[:, None]).sum(axis=0) / N
    cov = xtx / N - np.outer(loc, loc)
    cov = 0.5 * (cov + cov.T)
    precision = scipy.linalg.pinvh(cov, check_finite=False)
    return torch.from_numpy(mean), precision


def sample_estimator(model, loader, num_answers, output=None, reduce=True):
    """Drop-in for ``GQA.sample_estimator`` (:120-189).  ``loader`` yields ``(ques_id, feats, boxes, ids, mask, target)``
    host batches of this rank's shard; ``model`` is a ``GQAModel_maha``.  With ``output`` the result is cached in
    ``<output>/sample_estimates.pkl`` exactly like the reference (:122-128, :185-188)."""
    output_file = os.path.join(output, "sample_estimates.pkl") if output else None
    if output_file and os.path.exists(output_file):
        with open(output_file, "rb") as f:
            data = pickle.load(f)
        return data["mean"], data["precision"]
    num_classes = num_answers - 1                                                  # :131
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    est = MahaEstimator(num_classes, model.lxrt_encoder.dim, dev)
    for _, feats, boxes, ids, mask, target in loader:
        f, b, i, m = (t.to(dev, non_blocking=True) for t in (feats, boxes, ids, mask))
        _, pooled = model.forward_ids(f, b, i, None, m)
        est.accumulate(pooled, target[:, :-1] if target.shape[1] == num_answers else target)
    if reduce:
        est.all_reduce()
    mean, precision = est.finish()
    if output_file and (not dist.is_initialized() or dist.get_rank() == 0):
        with open(output_file, "wb") as f:
            pickle.dump({"mean": mean, "precision": precision}, f, protocol=pickle.HIGHEST_PROTOCOL)
    return mean, precision
