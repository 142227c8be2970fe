"""Config 5: RP-finetune steps (forward + backward + clip + BertAdam, fp32 path, dropout off), per-rank B = 64 questions + 64
random-pairing negatives = 128 rows.  Launch with torchrun for data parallelism; prints steps/s and questions/s, and under
torchrun checks that every rank holds identical parameters after the steps.

  python scripts/train_bench.py            |  python -m torch.distributed.run --nproc-per-node 2 scripts/train_bench.py"""
import os, sys, time, random
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn
from rgqa_b200.lxrt.trainer import LXMERTTrainer
from rgqa_b200.tasks.rp import build_rp_batch
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
B, L, A, K = 64, 20, 1842, int(os.environ.get("STEPS", 4))
sd = syn.lxmert_state_dict(variant="init")
tr = LXMERTTrainer(sd, device=f"cuda:{local}", lr=1e-5, warmup=0.1, t_total=1000, gemm=os.environ.get("GEMM", "fp32"))
rng = random.Random(1 + rank)
def batch(step):
    inp = syn.make_inputs(B, L, seed=10_000 * rank + step)
    qids = [f"q{rank}_{step}_{i}" for i in range(B)]
    img = {q: f"img{i}" for i, q in enumerate(qids)}
    target = torch.zeros(B, A); target[torch.arange(B), torch.randint(0, A, (B,))] = 1.0
    # the reference pairs SENTENCES; with pre-tokenised questions that is a row permutation of ids / mask
    sent = list(range(B))
    q2, f2, b2, s2, t2 = build_rp_batch(qids, inp["feat"], inp["pos"], sent, target, lambda q: img[q.split("+")[-1]] if "+" in q else img[q], rng)
    rows = torch.tensor(s2)
    return f2, b2, inp["ids"][rows], inp["seg"][rows], inp["mask"][rows], t2
losses = []
for step in range(2):
    losses.append(tr.train_step(*batch(step))[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for step in range(2, 2 + K):
    losses.append(tr.train_step(*batch(step))[0])
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
if world > 1:
    chk = tr.flat_p.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    assert all(torch.equal(c, allc[0]) for c in allc), "ranks diverged"
if rank == 0:
    print(f"config 5 [gemm={tr.gemm}], {world} rank(s): {dt*1e3:.0f} ms/step, {1/dt:.2f} steps/s, {world*B/dt:.0f} questions/s ({world*2*B/dt:.0f} rows/s); losses {[round(l, 2) for l in losses]}")
if world > 1:
    dist.destroy_process_group()
