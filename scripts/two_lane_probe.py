"""Does running consecutive batches on two handles / two streams (their small tail kernels overlapping the next batch's
big GEMMs) beat one handle on one stream?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rgqa_b200 import synthetic as syn
from rgqa_b200.tasks.gqa_model import GQAModel_maha
from rgqa_b200.tasks.gqa_eval import SweepRunner
dev = torch.device("cuda:0")
B, L, K = 1024, 20, 16
class Ns:
    llayers, xlayers, rlayers, from_scratch, precision, vocab, max_seq_length = 9, 5, 5, False, "bf16", None, L
sd = syn.lxmert_state_dict(seed=9595, variant="init")
mu, P = syn.maha_stats()
models, runners = [], []
for j in range(2):
    m = GQAModel_maha(1842, args=Ns, max_seq_length=L); m.load_state_dict(sd, strict=True); m = m.to(dev)
    m.engine().set_maha(mu, P); models.append(m); runners.append(SweepRunner(m, ("energy", "maha")))
resident = []
for j in range(2):
    inp = syn.make_inputs(B, L, seed=1234 + j)
    resident.append(tuple(t.to(dev) for t in (inp["feat"], inp["pos"], inp["ids"].to(torch.int32), inp["mask"].to(torch.int32), inp["seg"].to(torch.int32))))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def one_lane(i):
    return runners[0].score_device_batch(*resident[i & 1])
def two_lanes(i):
    with torch.cuda.stream(streams[i & 1]):
        return runners[i & 1].score_device_batch(*resident[i & 1])
def run(name, step):
    for i in range(8): step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K): step(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K * 1e3
    print(f"{name:28s}: {dt:6.2f} ms/step  {B/dt*1e3:8.0f} q/s", flush=True)
for rep in range(2):
    run("one handle, one stream", one_lane)
    run("two handles, two streams", two_lanes)
