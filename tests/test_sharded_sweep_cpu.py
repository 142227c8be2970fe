"""The N > 1 path on CPU (gloo, world_size 2): shard ranges, the single all-gather of (score, label), and the
assembled prediction JSON must equal the 1-rank sweep byte for byte."""
import json
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_scores(n, seed=3):
    g = np.random.default_rng(seed)
    return torch.tensor(np.round(g.random(n), 3), dtype=torch.float32), torch.tensor(g.integers(0, 1842, n), dtype=torch.int32)


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    from rgqa_b200.tasks.gqa_eval import SyntheticRGQA, build_quesid2ans, gather_results, shard_range
    ds = SyntheticRGQA(n, pool=8)
    score, label = _fake_scores(n)
    lo, hi = shard_range(n, rank, world)
    s, l, qids = gather_results(score[lo:hi], label[lo:hi], ds.qids[lo:hi])
    if rank == 0:
        q2a = build_quesid2ans(qids, l.numpy(), s.numpy(), ds.label2ans)
        ev = GQAOODEvaluator(ds, 0.5)
        ev.dump_result(q2a, os.path.join(out_dir, "sharded_predict.json"))
        json.dump({k: float(v) for k, v in ev.evaluate(q2a).items()}, open(os.path.join(out_dir, "sharded_result.json"), "w"))
    else:
        assert s is None and l is None
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from rgqa_b200.tasks.gqa_eval import shard_range
    for n in (0, 1, 7, 11158):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def test_two_rank_sweep_reproduces_single_rank_json(tmp_path):
    sys.path.insert(0, ROOT)
    from rgqa_b200.tasks.gqa_data import GQAOODEvaluator
    from rgqa_b200.tasks.gqa_eval import SyntheticRGQA, build_quesid2ans
    n = 1001  # odd: the shards differ in size, exercising the padded all-gather
    mp.spawn(_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    ds = SyntheticRGQA(n, pool=8)
    score, label = _fake_scores(n)
    q2a = build_quesid2ans(ds.qids, label.numpy(), score.numpy(), ds.label2ans)
    ev = GQAOODEvaluator(ds, 0.5)
    ev.dump_result(q2a, str(tmp_path / "single_predict.json"))
    assert (tmp_path / "single_predict.json").read_bytes() == (tmp_path / "sharded_predict.json").read_bytes()
    want = {k: float(v) for k, v in ev.evaluate(q2a).items()}
    assert json.load(open(tmp_path / "sharded_result.json")) == want  # bit-exact AUAF/FF95/FACC given identical scores


def _estimator_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rgqa_b200.tasks.gqa_eval import shard_range
    from rgqa_b200.tasks.gqa_maha import MahaEstimator
    x, t = _estimator_data()
    lo, hi = shard_range(x.shape[0], rank, world)
    est = MahaEstimator(t.shape[1], x.shape[1], device="cpu")
    xd, td = x[lo:hi].double(), t[lo:hi].double()          # the device kernel's sums, formed on the CPU for this rank's shard
    est.G += (xd * td.sum(1, keepdim=True)).t() @ xd
    est.class_sum += td.t() @ xd
    est.class_count += td.sum(0)
    est.all_reduce()
    mean, prec = est.finish()
    if rank == 0:
        np.savez(os.path.join(out_dir, "est.npz"), mean=mean.numpy(), prec=prec)
    dist.destroy_process_group()


def _estimator_data():
    g = torch.Generator().manual_seed(12)
    N, C, D = 401, 17, 48
    x = torch.tanh(torch.randn(N, D, generator=g))
    t = torch.zeros(N, C)
    t[torch.arange(N), torch.randint(0, C - 1, (N,), generator=g)] = 1
    t[::7, 3] = 1
    return x, t


def test_two_rank_estimator_all_reduce_matches_single_pass(tmp_path):
    """the Mahalanobis estimator shards over ranks with ONE all-reduce of its fp64 sums; the result equals
    the oracle's single-process pass over all samples."""
    sys.path.insert(0, ROOT)
    from oracle import scorers_oracle as S
    mp.spawn(_estimator_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    x, t = _estimator_data()
    ref_mean, ref_prec = S.sample_estimator(x, t)
    got = np.load(tmp_path / "est.npz")
    assert np.abs(got["mean"] - ref_mean.numpy()).max() <= 1e-6
    assert np.abs(got["prec"] - ref_prec).max() <= 1e-6 * np.abs(ref_prec).max()
