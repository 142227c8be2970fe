"""Where does the end-to-end step lose time against the device-resident step?  Variants of the SweepRunner loop."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn
from rgqa_b200.tasks.gqa_model import GQAModel_maha
from rgqa_b200.tasks.gqa_eval import SweepRunner
dev = torch.device("cuda:0")
B, L, K = 1024, 20, 10
class Ns:
    llayers, xlayers, rlayers, from_scratch, precision, vocab, max_seq_length = 9, 5, 5, False, "bf16", None, L
model = GQAModel_maha(1842, args=Ns, max_seq_length=L)
model.load_state_dict(syn.lxmert_state_dict(seed=9595, variant="init"), strict=True)
model = model.to(dev)
mu, P = syn.maha_stats(); model.engine().set_maha(mu, P)
runner = SweepRunner(model, ("energy", "maha"))
host, resident = [], []
for j in range(2):
    inp = syn.make_inputs(B, L, seed=1234 + j)
    h = tuple(t.pin_memory() for t in (inp["feat"], inp["pos"], inp["ids"].to(torch.int32), inp["mask"].to(torch.int32), inp["seg"].to(torch.int32)))
    host.append(h); resident.append(tuple(t.to(dev) for t in h))
out_host = [[(torch.empty(B).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()) for _ in range(2)] for _ in range(2)]
def run(name, step):
    for i in range(6): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(K): step(i)
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{name:46s}: {e0.elapsed_time(e1)/K:6.2f} ms/step (host enqueue {t_host/K*1e3:5.2f} ms/step)", flush=True)
def dev_step(i):
    return runner.score_device_batch(*resident[i & 1])
def e2e(i):
    res = runner.submit(*host[i & 1])
    for (s, l), (hs, hl) in zip(res, out_host[i & 1]):
        hs.copy_(s, non_blocking=True); hl.copy_(l, non_blocking=True)
def e2e_nod2h(i):
    runner.submit(*host[i & 1])
def e2e_small_only(i):  # upload everything except the feature block
    f, p, ids, mask, seg = host[i & 1]
    rf = resident[i & 1][0]
    runner2.submit_with_feat(rf, p, ids, mask, seg) if False else runner.score_device_batch(rf, *(t.to(dev, non_blocking=True) for t in (p, ids, mask, seg)))
cs = torch.cuda.Stream()
stage = [torch.empty_like(resident[0][0]), torch.empty_like(resident[0][0])]
def copy_only_plus_dev(i):  # feature H2D on a side stream with no dependency at all + resident compute
    with torch.cuda.stream(cs):
        stage[i & 1].copy_(host[i & 1][0], non_blocking=True)
    return dev_step(i)
run("device-resident", dev_step)
run("e2e (SweepRunner.submit + D2H)", e2e)
run("e2e without the D2H of results", e2e_nod2h)
run("resident feats, small inputs uploaded", e2e_small_only)
run("resident compute + independent feature H2D", copy_only_plus_dev)
run("device-resident again", dev_step)
