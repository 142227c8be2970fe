#!/usr/bin/env python
"""bench.py — the RGQA evaluation hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1]): LXMERT-base (9/5/5) GQA eval forward + energy (top-2) and Mahalanobis
scoring, batch 1024 questions per GPU per step, 36 x 2048 RoI features + 20 tokens, bf16 tensor-core path,
random-init weights (seed 9595), synthetic inputs (seed 1234).  A "step" is one batch through the whole path.

  value   questions/s with inputs (fp32 features, boxes, ids) already resident in HBM; the fp32->bf16 feature
          cast is part of every step.  Two resident batches (2 x 302 MB > 126 MB L2) alternate.
  e2e     the same metric through the public sweep API (rgqa_b200.tasks.gqa_eval.SweepRunner over
          GQAModel.forward_ids + Engine.score) from PINNED HOST buffers: per step one H2D of the batch and one
          D2H of (score, label) for both scorers, uploads double-buffered against compute.
  --impl reference   the reference algorithm on the host cores (the CPU oracle, a port: the reference is
          pure Python/PyTorch with no build), all threads, bounded sample of B=32 per step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 when the variable is unset.  The reference arm is meant to use every host core it can, so
# that default is undone before the OpenMP runtime reads it (torch.set_num_threads below then sizes the pools).
if "reference" in sys.argv and os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
    del os.environ["OMP_NUM_THREADS"]

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GQA questions/sec, LXMERT fwd+OOD score, 1/2/4/8 B200; tensor-pipe % of peak"
GFLOP_PER_Q = {20: 10.5827, 30: 12.8330}  # 2*M*N*K per GEMM + 4*Sq*Sk*H per attention
MAHA_GFLOP_PER_Q = 0.00401


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="questions per GPU per step")
    ap.add_argument("--seq-len", type=int, default=20)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="eval", choices=["eval", "train"],
                    help="eval = the headline (BASELINE.json configs[1]); train = the RP-finetune step (configs[4]), a side measurement")
    ap.add_argument("--no-graph", action="store_true", help="--workload train: issue the ~1 000 launches of a step from Python")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the UNITER (configs[2]) side measurement")
    ap.add_argument("--ref-batch", type=int, default=32)
    return ap.parse_args()


def measured_gemm_traffic():
    """dram__bytes_read + dram__bytes_write per tcgen05 GEMM-class launch (average over the 121 launches of one step: the plain
    GEMMs, the Q|K|V + attention kernels, the visual encoder), from the committed ncu capture of this same command
    (profiles/r02b_launch_summary.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02b_launch_summary.json")) as f:
            return float(json.load(f)["tcgen05_gemm"]["dram_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (port of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(seq_len, batch, steps, warmup, budget_s=None):
    from oracle import lxmert_oracle as O
    from oracle import scorers_oracle as S
    from rgqa_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.lxmert_state_dict()
    inp = syn.make_inputs(batch, seq_len, seed=1234)
    mu, P = syn.maha_stats()

    def step():
        logit, x = O.gqa_model_forward(sd, inp["feat"], inp["pos"], inp["ids"], inp["seg"], inp["mask"])
        S.energy_top2(logit)
        S.mahalanobis(x, mu, P)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": batch * done / dt, "unit": "questions/s", "cores": cores, "kind": "port",
            "sample": f"{done} steps of B={batch}, L={seq_len}: oracle LXMERT forward + energy + literal 1842-class "
                      f"Mahalanobis loop, fp32, torch {torch.__version__} with {cores} threads",
            "ms_per_step": 1e3 * dt / done, "steps": done}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_rate(a.seq_len, a.ref_batch, a.steps, min(a.warmup, 2), budget_s=150)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "questions/s", "n_gpus": a.gpus,
            "steps": r["steps"], "warmup": min(a.warmup, 2), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LXMERT-base 9/5/5 GQA eval forward + energy(top-2) + Mahalanobis, "
                                   f"B={a.ref_batch} per step (bounded sample of the B={a.batch} workload), L={a.seq_len}, "
                                   "36x2048 RoI, random-init seed 9595"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# parity of the TIMED configuration (same process, same handle, same kernels and CUDA graphs as the timed region)
# ------------------------------------------------------------------------------------------------
LOGIT_TOL = 2e-2  # BASELINE north_star: bf16 logits within 2e-2 abs


def parity_block(model, eng, dev, batch_dev, batch_host, mu, P, precision, rows=8):
    """(1) first `rows` questions of the timed batch against the CPU oracle (logits, energy, Mahalanobis score / label);
    (2) the whole timed batch against the fp32 CUDA path of the same weights (itself pinned to the oracle at 1e-4 by the GPU
    tests): answer agreement overall and on the buckets whose fp32 top-2 logit margin exceeds 2x / 4x the logit tolerance.
    Raises when out of tolerance: a bench number without parity is not a number."""
    from oracle import lxmert_oracle as O
    from oracle import scorers_oracle as S
    from rgqa_b200 import synthetic as syn
    from rgqa_b200.engine import Engine
    f, p, ids, mask, seg = batch_dev
    out, pooled, logits = eng.forward_scored(f, p, ids, seg, mask, ["energy", "maha"], want_logits=True, want_pooled=True)
    torch.cuda.synchronize()
    sd = syn.lxmert_state_dict(seed=9595, variant="init")
    hf, hp, hids, hmask, hseg = (t[:rows] for t in batch_host)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_logit, ref_x = O.gqa_model_forward(sd, hf.float(), hp, hids.long(), hseg.long(), hmask.long())
    e_ref, el_ref = S.energy_top2(ref_logit)
    m_ref, ml_ref, m_full = S.mahalanobis(ref_x, mu, P)
    dlogit = float((logits[:rows].cpu() - ref_logit).abs().max())
    e_rel = float(((out["energy"][0][:rows].cpu() - e_ref).abs() / e_ref.abs()).max())
    m_rel = float(((out["maha"][0][:rows].cpu() - m_ref).abs() / m_ref.abs()).max())
    # labels: a swap is a failure only where the reference margin is clear of the tolerance
    top2 = ref_logit.topk(2, dim=1).values
    clear_l = (top2[:, 0] - top2[:, 1]) > 2 * LOGIT_TOL
    mt2 = m_full.topk(2, dim=1).values
    clear_m = (mt2[:, 0] - mt2[:, 1]) > 2e-3 * mt2[:, 0].abs()
    lab_e = out["energy"][1][:rows].cpu().long() == el_ref.long()
    lab_m = out["maha"][1][:rows].cpu().long() == ml_ref.long()
    # fp32 CUDA path on the whole batch
    cfg32 = model.lxrt_encoder.model.engine_config(model.num_answers)
    cfg32.precision = "fp32"
    e32 = Engine(cfg32, device=dev)
    e32.load_weights(sd.items())
    _, l32 = e32.forward(f.float(), p, ids, seg, mask, want_pooled=False)
    torch.cuda.synchronize()
    e32.close()
    d_all = (logits - l32).abs()
    t2 = l32.topk(2, dim=1).values
    margin = t2[:, 0] - t2[:, 1]
    same = out["energy"][1].long() == l32.argmax(1)
    agree = {"n": int(same.numel()), "overall": float(same.float().mean())}
    for k in (2, 4):
        sel = margin > k * LOGIT_TOL
        agree[f"margin_gt_{k}tol"] = {"n": int(sel.sum()), "agree": float(same[sel].float().mean()) if int(sel.sum()) else None}
    par = {"rows_vs_oracle": rows, "max_dlogit": dlogit, "energy_rel": e_rel, "maha_rel": m_rel,
           "labels_equal": {"energy": int(lab_e.sum()), "maha": int(lab_m.sum()), "of": rows},
           "batch_vs_fp32_cuda": {"max_dlogit": float(d_all.max()), "rms_dlogit": float(d_all.pow(2).mean().sqrt()),
                                  "frac_within_2e-2": float((d_all <= LOGIT_TOL).float().mean()), "answer_agreement": agree},
           "tolerances": {"logit_abs_hard_cap": 3e-2 if precision == "bf16" else 1e-4, "score_rel": 2e-2 if precision == "bf16" else 1e-3}}
    cap, srel = par["tolerances"]["logit_abs_hard_cap"], par["tolerances"]["score_rel"]
    ok = (dlogit <= cap and float(d_all.max()) <= (4e-2 if precision == "bf16" else 1e-4) and e_rel <= srel and m_rel <= srel and
          bool(lab_e[clear_l].all()) and bool(lab_m[clear_m].all()) and
          (agree["margin_gt_4tol"]["agree"] in (None, 1.0)))
    par["ok"] = bool(ok)
    return par


DTYPE_DETAIL = {
    "bf16": "16-bit tensor-core operands, fp32 accumulation (tcgen05 kind::f16): weights, attention context, GELU intermediates and "
            "Q|K|V in bf16; the pre-LayerNorm hidden state — A operand of the Q|K|V / FFN1 / pooler GEMMs — and the copies of those "
            "GEMMs' folded weights in fp16 (11 instead of 8 significand bits, saturating; option state_f16, extra.all_bf16_hilo_state "
            "is the all-bf16 variant); LayerNorm / softmax statistics, pooled output, logits and scores fp32",
    "fp32": "fp32 end to end on the CUDA cores",
}


def uniter_extra(dev, steps=6):
    """BASELINE configs[2] next to the headline: UNITER-base (12 layers over the joint 56-token sequence), B = 1024, forward +
    MSP through rgqa_forward_scored, device-resident inputs; 9.7511 GFLOP / question."""
    from rgqa_b200 import synthetic as syn
    from rgqa_b200.engine import Engine, EngineConfig
    cfg = EngineConfig(arch="uniter", precision="bf16", l_layers=12, x_layers=0, r_layers=0, vocab_size=28996, visual_pos_dim=7)
    eng = Engine(cfg, device=dev)
    eng.load_weights(syn.uniter_state_dict(variant="init").items())
    B = 1024
    ins = []
    for j in range(2):
        inp = syn.make_inputs(B, 20, seed=3 + j, vocab=28996)
        ins.append([inp[k].to(dev) for k in ("feat", "pos7")] + [inp[k].to(torch.int32).to(dev) for k in ("ids", "seg", "mask")])
    for i in range(4):
        eng.forward_scored(*ins[i & 1], ["msp"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.forward_scored(*ins[i & 1], ["msp"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.close()
    qps = B / ms * 1e3
    peak_tf = peaks()[0]
    return {"uniter_b1024_qps": qps, "uniter_ms_per_step": ms, "uniter_tflops": qps * 9.7511e9 / 1e12,
            "uniter_frac_of_sustained_peak": qps * 9.7511e9 / 1e12 / peak_tf, "uniter_steps": steps,
            "uniter_workload": "UNITER-base 12 layers, 36 RoI + 20 tokens, forward + MSP, B=1024, bf16 (BASELINE.json configs[2])"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import ctypes as C
    import torch.distributed as dist
    from rgqa_b200 import _lib as L
    from rgqa_b200 import synthetic as syn
    from rgqa_b200.tasks.gqa_eval import SweepRunner, gather_results
    from rgqa_b200.tasks.gqa_model import GQAModel_maha

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    for kv in filter(None, os.environ.get("RGQA_OPTS", "").split(",")):  # e.g. RGQA_OPTS=two_streams=0,gemm_ln=1
        k, v = kv.split("=")
        L.check(lib.rgqa_set_option(k.encode(), int(v)), "rgqa_set_option")

    class Ns:  # the flags GQAModel reads (rgqa_b200.param)
        llayers, xlayers, rlayers, from_scratch, precision, vocab, max_seq_length = 9, 5, 5, False, a.precision, None, a.seq_len
    model = GQAModel_maha(1842, args=Ns, max_seq_length=a.seq_len)
    model.load_state_dict(syn.lxmert_state_dict(seed=9595, variant="init"), strict=True)
    model = model.to(dev)
    eng = model.engine()
    mu, P = syn.maha_stats()
    eng.set_maha(mu, P)

    B, Lq = a.batch, a.seq_len
    methods = ("energy", "maha")
    runner = SweepRunner(model, methods, device=dev)
    # two distinct batches, host (pinned) and device-resident copies
    host = []
    for i in range(2):
        inp = syn.make_inputs(B, Lq, seed=1234 + 17 * i + 1000 * rank)
        host.append(tuple(t.pin_memory() for t in (inp["feat"], inp["pos"], inp["ids"].to(torch.int32),
                                                   inp["mask"].to(torch.int32), inp["seg"].to(torch.int32))))
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    d2h_bytes = len(methods) * B * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident timing ----------------
    def dev_step(i):
        f, p, ids, mask, seg = resident[i & 1]
        return runner.score_device_batch(f, p, ids, mask, seg)

    # priming (not counted as warm-up): the library captures one CUDA graph per (shape, input buffers) on the second
    # call of each of the two alternating batches; capture + instantiation must not land in a timed region
    for i in range(4):
        dev_step(i)
    warm = [dev_step(i) for i in range(a.warmup)]
    # the end-of-sweep all-gather also gets its warm-up call: NCCL connects its channels lazily on the first collective
    # of a shape (tens of ms), which is start-up cost like context or graph creation, not part of a sweep's steady state
    for m in range(len(methods)):
        gather_results(torch.cat([k[m][0] for k in warm] * max(1, a.steps // max(1, a.warmup))),
                       torch.cat([k[m][1] for k in warm] * max(1, a.steps // max(1, a.warmup))))
    del warm
    two_streams_default = 0 if "two_streams=0" in os.environ.get("RGQA_OPTS", "") else 1

    def timed_region(profile):
        """K steps + the end-of-sweep all-gather, device-timed with CUDA events, max over ranks."""
        barrier()
        if profile:  # per-kernel durations are only meaningful when kernels do not share the GPU: one stream
            lib.rgqa_set_option(b"two_streams", 0)
        lib.rgqa_prof_enable(1 if profile else 0)
        n0 = lib.rgqa_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        keep = []
        for i in range(a.steps):
            keep.append(dev_step(i))
        # the sweep ends with ONE all-gather of (score, label) per scorer (N > 1)
        for m in range(len(methods)):
            gather_results(torch.cat([k[m][0] for k in keep]), torch.cat([k[m][1] for k in keep]))
        ev1.record()
        barrier()
        lib.rgqa_prof_enable(0)
        if profile:
            lib.rgqa_set_option(b"two_streams", two_streams_default)
        return max_over_ranks(ev0.elapsed_time(ev1)), (lib.rgqa_launch_count() - n0) // max(a.steps, 1)

    # region 1: the headline number, nothing but the path's own kernels on the stream
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    ms_total, launches = timed_region(False)
    clocks = sampler.stop()
    ms_step = ms_total / a.steps
    value = world * B * a.steps / (ms_total * 1e-3)
    # ---------------- end-to-end from pinned host memory (measured right after the headline region, before the
    # event-instrumented one) ----------------
    e2e = None
    if not a.no_e2e:
        out_host = [[(torch.empty(B, dtype=torch.float32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
                     for _ in methods] for _ in range(2)]

        def e2e_step(i):
            f, p, ids, mask, seg = host[i & 1]
            res = runner.submit(f, p, ids, mask, seg)
            for (s, l), (hs, hl) in zip(res, out_host[i & 1]):
                hs.copy_(s, non_blocking=True)
                hl.copy_(l, non_blocking=True)

        for i in range(4):  # priming: the two staging slots of the runner each get their CUDA graph (see above)
            e2e_step(i)
        for i in range(a.warmup):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            e2e_step(i)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * B * a.steps / (ms_e2e * 1e-3), "unit": "questions/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / a.steps,
               "host_feature_dtype": "f32 (reference loader format, src/tasks/gqa_data.py:173-238)"}
        # the same sweep from the ingest store's format (rgqa_b200/ingest.py: bf16 features, RNE = the device cast bit for bit):
        # half the H2D bytes, which is what eight ranks pulling out of one host DRAM need
        host_b = [(hb[0].to(torch.bfloat16).pin_memory(),) + tuple(hb[1:]) for hb in host]

        def e2e_step_b(i):
            f, p, ids, mask, seg = host_b[i & 1]
            res = runner_b.submit(f, p, ids, mask, seg)
            for (s, l), (hs, hl) in zip(res, out_host[i & 1]):
                hs.copy_(s, non_blocking=True)
                hl.copy_(l, non_blocking=True)
        runner_b = SweepRunner(model, methods, device=dev)
        for i in range(4 + a.warmup):
            e2e_step_b(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            e2e_step_b(i)
        e1.record()
        barrier()
        ms_b = max_over_ranks(e0.elapsed_time(e1))
        e2e["bf16_host_features"] = {"value": world * B * a.steps / (ms_b * 1e-3), "ms_per_step": ms_b / a.steps,
                                     "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host_b[0])}

    # region 2: the same K steps again with CUDA events bracketing every GEMM / attention launch on its own
    # stream (per-kernel durations for the roofline); the events cost ~3-4 % so they stay out of `value`, and
    # the visual branch is folded back onto the main stream so that a kernel's duration is its own
    ms_prof, _ = timed_region(True)
    nk = 4
    ms_sum, fl_sum, cnt = (C.c_double * nk)(), (C.c_double * nk)(), (C.c_longlong * nk)()
    L.check(lib.rgqa_prof_collect(ms_sum, fl_sum, cnt), "rgqa_prof_collect")
    peak_tf, peak_gbs, peak_src = peaks()
    gemm_tf = (fl_sum[0] / (ms_sum[0] * 1e-3)) / 1e12 if ms_sum[0] > 0 else 0.0
    gq = GFLOP_PER_Q.get(Lq, GFLOP_PER_Q[20]) + MAHA_GFLOP_PER_Q
    roofline = {"bound": "tensor", "kernel": "all tcgen05 GEMM-class launches of a step: gemm_bf16_pair_kernel (FFN1, FFN2 / attention-output producers, cross Q|K|V), qkv_attn_pair_kernel (Q|K|V projection + self-attention in its epilogue; its attention time is inside the denominator), visn_embed_kernel, gemm_bf16_tcgen05_kernel (head)",
                "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf, "traffic": measured_gemm_traffic() if (B == 1024 and Lq == 20) else None,
                "traffic_unit": "bytes of DRAM traffic per GEMM-class launch (ncu, profiles/r02b_launch_summary.json)",
                "algorithmic_flops_per_launch": fl_sum[0] / max(cnt[0], 1),
                "peak_source": peak_src, "launches_per_step": int(cnt[0] // max(a.steps, 1)),
                "gemm_ms_per_step": ms_sum[0] / a.steps, "gemm_share_of_step": (ms_sum[0] / a.steps) / (ms_prof / a.steps),
                "attention_ms_per_step": ms_sum[2] / a.steps, "fp32_gemm_ms_per_step": ms_sum[1] / a.steps,
                "profiled_region_ms_per_step": ms_prof / a.steps,
                "whole_step_tflops": (value / world) * gq * 1e9 / 1e12,
                "whole_step_frac": (value / world) * gq * 1e9 / 1e12 / peak_tf}

    # ---------------- parity of the timed configuration ----------------
    parity = None
    if not a.no_parity and rank == 0:
        parity = parity_block(model, eng, dev, resident[0], host[0], mu, P, a.precision)
    # ---------------- N ranks == 1 rank ----------------
    sharded = None
    if world > 1:
        mine = [dev_step(i) for i in range(2)]            # this rank's two batches through the timed path
        torch.cuda.synchronize()
        gathered = [gather_results(torch.cat([k[m][0] for k in mine]), torch.cat([k[m][1] for k in mine])) for m in range(len(methods))]
        if rank == 0:
            same = True
            for r in range(world):                         # rank 0 recomputes every rank's shard on its own GPU
                for i in range(2):
                    inp = syn.make_inputs(B, Lq, seed=1234 + 17 * i + 1000 * r)
                    t = [inp["feat"].to(dev), inp["pos"].to(dev)] + [inp[k].to(torch.int32).to(dev) for k in ("ids", "mask", "seg")]
                    res = runner.score_device_batch(*t)
                    for m in range(len(methods)):
                        sl = slice((2 * r + i) * B, (2 * r + i + 1) * B)
                        same &= bool(torch.equal(res[m][0].cpu(), gathered[m][0][sl])) and bool(torch.equal(res[m][1].cpu(), gathered[m][1][sl]))
            sharded = same
    extra = None
    if rank == 0 and world == 1 and not a.no_extra and B == 1024:
        extra = uniter_extra(dev)
        if a.precision == "bf16" and "state_f16=0" not in os.environ.get("RGQA_OPTS", ""):
            # the same step with every GEMM operand in bf16 (the deferred-LayerNorm state as a bf16 hi/lo pair instead of one fp16
            # word per element): a second handle finalized with state_f16 = 0, same weights, same resident batches
            L.check(lib.rgqa_set_option(b"state_f16", 0), "rgqa_set_option")
            try:
                model_hl = GQAModel_maha(1842, args=Ns, max_seq_length=a.seq_len)
                model_hl.load_state_dict(syn.lxmert_state_dict(seed=9595, variant="init"), strict=True)
                model_hl = model_hl.to(dev)
                model_hl.engine().set_maha(mu, P)
                runner_hl = SweepRunner(model_hl, methods, device=dev)
                for i in range(4 + a.warmup):
                    runner_hl.score_device_batch(*resident[i & 1])
                torch.cuda.synchronize()
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                for i in range(a.steps):
                    runner_hl.score_device_batch(*resident[i & 1])
                h1.record()
                torch.cuda.synchronize()
                ms_hl = h0.elapsed_time(h1) / a.steps
                extra["all_bf16_hilo_state"] = {"value": B / (ms_hl * 1e-3), "unit": "questions/s", "ms_per_step": ms_hl,
                                                "what": "same workload with rgqa_set_option('state_f16', 0): hidden state as a bf16 hi/lo "
                                                        "pair, every GEMM bf16 x bf16 (the round-2a build)"}
                del runner_hl, model_hl
            finally:
                L.check(lib.rgqa_set_option(b"state_f16", 1), "rgqa_set_option")

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = {k: v for k, v in cpu_reference_rate(Lq, a.ref_batch, 16, 1, budget_s=25).items()
               if k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "questions/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": a.precision, "dtype_detail": DTYPE_DETAIL.get(a.precision, a.precision), "data": "synthetic",
                "config": {"workload": "LXMERT-base 9/5/5 GQA eval forward + energy(top-2) + Mahalanobis scoring "
                                       f"(BASELINE.json configs[1]), B={B} questions per GPU per step, L={Lq}, 36x2048 RoI, "
                                       "random-init seed 9595",
                           "questions_per_step": world * B, "parallelism": f"batch-sharded x{world}, one all-gather per sweep",
                           "l2": "inputs larger than L2: two alternating 302 MB fp32 feature batches per GPU"},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
                "parity": parity, "extra": extra}
        if world > 1:
            line["sharded_equals_single"] = sharded
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            print("bench.py: PARITY FAILED on the timed configuration: " + json.dumps(parity), file=sys.stderr, flush=True)
            sys.exit(3)
        if world > 1 and sharded is False:
            print("bench.py: the N-rank sweep does not reproduce the 1-rank results bit for bit", file=sys.stderr, flush=True)
            sys.exit(4)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# --workload train: BASELINE.json configs[4], the RP-finetune step, data parallel (NOT the headline metric)
# ------------------------------------------------------------------------------------------------
def run_train(a):
    """Per rank: B = 64 questions + 64 random-pairing negatives = 128 rows (src/tasks/gqa_conf.py:148-202: forward, BCE x A,
    backward, clip 5.0, BertAdam); gradients all-reduced (mean) in per-layer buckets overlapped with the backward.  Mixed
    precision: bf16 tensor-core GEMMs, fp32 master weights / activations / gradients / optimiser."""
    import random
    import torch.distributed as dist
    from rgqa_b200 import synthetic as syn
    from rgqa_b200.lxrt.trainer import LXMERTTrainer
    from rgqa_b200.tasks.rp import build_rp_batch
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, Lq, A = 64, a.seq_len, 1842
    tr = LXMERTTrainer(syn.lxmert_state_dict(variant="init"), device=dev, lr=1e-5, warmup=0.1, t_total=10000,
                       gemm="bf16" if a.precision == "bf16" else "fp32", use_graph=not a.no_graph)
    rng = random.Random(1 + rank)

    def batch(step):
        inp = syn.make_inputs(B, Lq, seed=10_000 * rank + step)
        qids = [f"q{rank}_{step}_{i}" for i in range(B)]
        img = {q: f"img{i}" for i, q in enumerate(qids)}
        g = torch.Generator().manual_seed(77 * rank + step)
        target = torch.zeros(B, A)
        target[torch.arange(B), torch.randint(0, A, (B,), generator=g)] = 1.0
        q2, f2, b2, s2, t2 = build_rp_batch(qids, inp["feat"], inp["pos"], list(range(B)), target,
                                            lambda q: img[q.split("+")[-1]] if "+" in q else img[q], rng)
        rows = torch.tensor(s2)
        return [t.to(dev) for t in (f2, b2, inp["ids"][rows], inp["seg"][rows], inp["mask"][rows], t2)]
    batches = [batch(i) for i in range(4)]     # resident: the step is timed, not the synthetic data generator

    def timed(k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = [tr.train_step(*batches[i & 3]) for i in range(k)]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / k, out
    for i in range(max(a.warmup, 3)):     # eager, capture, first replay
        tr.train_step(*batches[i & 3])
    ms_step, out = timed(a.steps)
    eager_overlap = eager_blocking = None
    if world > 1:            # A/B without the graph: per-layer buckets overlapped with the backward vs one blocking all-reduce
        tr.use_graph = False
        for mode in (True, False):
            tr.bucketed = mode
            for i in range(2):
                tr.train_step(*batches[i & 3])
            ms, _ = timed(a.steps)
            if mode:
                eager_overlap = ms
            else:
                eager_blocking = ms
        tr.bucketed, tr.use_graph = True, not a.no_graph
        chk = tr.flat_p.double().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        identical = all(bool(torch.equal(c, allc[0])) for c in allc)
    if rank == 0:
        line = {"metric": "RP-finetune questions/sec (BASELINE.json configs[4]; side workload, not the headline)", "value": world * B / (ms_step * 1e-3),
                "unit": "questions/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 2), "ms_per_step": ms_step,
                "steps_per_s": 1e3 / ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16 GEMMs / fp32 master" if a.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "LXMERT-base 9/5/5 RP-finetune step: 64 questions + 64 random-pairing negatives per rank, forward + BCE x A + "
                                       "backward + clip 5.0 + BertAdam, dropout off; gradient pass replayed as THREE CUDA graphs cut at layer boundaries, each finished segment's fp32 gradients all-reduced (mean) on NCCL's stream under the next segment's replay (eager mode: per-layer buckets overlapped with the backward)",
                           "rows_per_step": world * 2 * B, "grad_buckets": tr.comm_stats["buckets"], "grad_bytes": tr.comm_stats["bytes"]},
                "loss_last": out[-1][0], "grad_norm_last": out[-1][1], "cuda_graph": bool(tr.use_graph)}
        if world > 1:
            line["ms_per_step_eager_bucketed_overlap"] = eager_overlap
            line["ms_per_step_eager_blocking_allreduce"] = eager_blocking
            line["ranks_identical"] = identical
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)
