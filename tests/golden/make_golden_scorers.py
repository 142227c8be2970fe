"""Golden vectors for the rejection-scorer DRIVERS, produced by the UNMODIFIED reference at /root/reference.

    python tests/golden/make_golden_scorers.py        # build container only; writes scorers_reference.npz / .json

The reference's task drivers are executed as they are — ``GQA.ood_evaluate`` of ``src/tasks/gqa_conf.py`` (MSP, :296-330),
``gqa_energy.py`` (:194-213), ``gqa_maha.py`` (``sample_estimator`` :120-189, ``ood_evaluate`` :191-276 with ``--noise`` 0 and
> 0), ``gqa_odin.py`` (:85-137) and ``gqa_ensemble.py`` (:95-110, and its ``__main__`` combine step :135-215, whose statements are
compiled from the reference file at run time) — as unbound methods on a stand-in ``self`` that carries a
tiny differentiable model (``tiny_vqa.TinyVQA``) and an in-memory loader, so that every line between "batch comes out of
the loader" and "``quesid2ans`` goes into the evaluator" is the reference's own.  The real ``GQAOODEvaluator`` then produces
the result dict.  What is stubbed, and why it does not touch the lines under test:
  * ``boto3`` / ``botocore`` (absent; imported by lxrt/file_utils.py for S3 downloads), ``vilt.vilt`` / ``vilt.modules.vilt_utils``
    (need ``timm`` / a ``transformers`` that still exports ``AdamW``; the ViLT backbone is out of scope) — empty modules with
    the two imported names;
  * ``Tensor.cuda`` / ``Module.cuda`` are no-ops (no GPU in the build container; the drivers hard-code ``.cuda()``);
  * ``GQAOODEvaluator.evaluate`` is wrapped by a spy that records ``quesid2ans`` before calling the original.
The outputs pin ``oracle/scorers_oracle.py`` and ``oracle/odin_oracle.py`` (tests/test_oracle_cpu.py)."""
import json
import os
import sys
import tempfile
import types
import collections

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, HERE)

C, D, FEAT, V = 12, 16, 32, 4          # classes (without 'UQ'), pooled width, RoI feature width, RoIs per image
N_TRAIN, N_TEST, BS = 64, 48, 16
NOISE, TEMPERATURE = 0.05, 2.0


def shim():
    # `vilt` itself stays the reference's (namespace) package: gqa_data.py needs its real `vilt.transforms`
    for name in ("boto3", "botocore", "botocore.exceptions", "vilt.vilt", "vilt.modules", "vilt.modules.vilt_utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    sys.modules["vilt.vilt"].GQAViLT = type("GQAViLT", (), {})
    sys.modules["vilt.modules.vilt_utils"].set_schedule = lambda *a, **k: None
    for p in (os.path.join(REF, "src", "tasks"), os.path.join(REF, "src"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.argv = ["x", "--llayers", "9", "--xlayers", "5", "--rlayers", "5", "--fromScratch"]
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def make_data(seed, n):
    g = torch.Generator().manual_seed(seed)
    feats = torch.rand(n, V, FEAT, generator=g)
    boxes = torch.rand(n, V, 4, generator=g)
    words = "is the a red what color dog on table left of which man holding".split()
    sents = [" ".join(words[int(k)] for k in torch.randint(0, len(words), (int(m),), generator=g))
             for m in torch.randint(3, 12, (n,), generator=g)]
    target = torch.zeros(n, C + 1)
    labels = []
    for i in range(n):
        if torch.rand((), generator=g) < 0.3:                       # unanswerable question
            target[i, C] = 1.0
            labels.append({"UQ": 1.0})
        else:
            a = int(torch.randint(0, C, (), generator=g))
            target[i, a] = 1.0
            lab = {f"ans{a}": 1.0}
            if torch.rand((), generator=g) < 0.4:                   # a second, lower-scored answer (multi-label GQA/VQA style)
                b = int(torch.randint(0, C, (), generator=g))
                if b != a:
                    target[i, b] = 0.3
                    lab[f"ans{b}"] = 0.3
            labels.append(lab)
    qids = [f"q{seed}_{i}" for i in range(n)]
    return qids, feats, boxes, sents, target, labels


def loader_of(qids, feats, boxes, sents, target):
    return [(qids[i:i + BS], feats[i:i + BS].clone(), boxes[i:i + BS].clone(), sents[i:i + BS], target[i:i + BS].clone())
            for i in range(0, len(qids), BS)]


def main():
    shim()
    from tiny_vqa import TinyVQA
    from param import args
    import gqa_data                                                   # the module name the drivers import inside ood_evaluate
    import tasks.gqa_conf as conf
    import tasks.gqa_energy as energy
    import tasks.gqa_maha as maha
    import tasks.gqa_odin as odin
    import tasks.gqa_ensemble as ens

    captured = {}
    orig_eval = gqa_data.GQAOODEvaluator.evaluate

    def spy(self, quesid2ans):
        captured["q"] = dict(quesid2ans)
        return orig_eval(self, quesid2ans)
    gqa_data.GQAOODEvaluator.evaluate = spy

    torch.manual_seed(2024)
    model = TinyVQA(FEAT, D, C)
    for p in model.parameters():
        p.data.normal_(0, 0.5)
    model2 = TinyVQA(FEAT, D, C)                                       # second "checkpoint" for the ensemble
    for p in model2.parameters():
        p.data.normal_(0, 0.5)

    tr = make_data(11, N_TRAIN)
    te = make_data(22, N_TEST)
    DataTuple = collections.namedtuple("DataTuple", "dataset loader evaluator")
    Dset = collections.namedtuple("Dset", "num_answers label2ans id2datum")
    label2ans = [f"ans{j}" for j in range(C)] + ["UQ"]
    dset_tr = Dset(C + 1, label2ans, {q: {"label": l} for q, l in zip(tr[0], tr[5])})
    dset_te = Dset(C + 1, label2ans, {q: {"label": l} for q, l in zip(te[0], te[5])})
    train_tuple = DataTuple(dset_tr, loader_of(*tr[:5]), None)
    test_tuple = DataTuple(dset_te, loader_of(*te[:5]), None)

    args.backbone, args.tau, args.tqdm = "lxmert", 0.5, False
    out = {"feat_train": tr[1], "box_train": tr[2], "target_train": tr[4], "feat_test": te[1], "box_test": te[2],
           "target_test": te[4]}
    out.update({f"model.{k}": v for k, v in model.state_dict().items()})
    out.update({f"model2.{k}": v for k, v in model2.state_dict().items()})
    meta = {"sents_train": tr[3], "sents_test": te[3], "qids_test": te[0],
            # label dicts as ordered [answer, score] pairs: the evaluator takes the FIRST key as the classification target
            # (gqa_data.py:567), so the order must survive the JSON round trip
            "labels_test": [list(map(list, l.items())) for l in te[5]], "label2ans": label2ans,
            "noise": NOISE, "temperature": TEMPERATURE, "results": {}}

    def record(name, results):
        q = captured.pop("q")
        assert list(q.keys()) == te[0]
        out[f"{name}.score"] = np.asarray([float(q[k][1]) for k in te[0]], dtype=np.float64)
        out[f"{name}.label"] = np.asarray([label2ans.index(q[k][0]) for k in te[0]], dtype=np.int64)
        meta["results"][name] = {k: float(v) for k, v in results.items()}

    me = types.SimpleNamespace(model=model, train_tuple=train_tuple)
    # ---- MSP (gqa_conf.py:296-330) and energy (gqa_energy.py:194-213)
    model.maha = False
    record("msp", conf.GQA.ood_evaluate(me, test_tuple))
    record("energy", energy.GQA.ood_evaluate(me, test_tuple))
    # ---- ODIN (gqa_odin.py:85-137)
    args.noise, args.temperature = NOISE, TEMPERATURE
    record("odin", odin.GQA.ood_evaluate(me, test_tuple))
    args.noise, args.temperature = 0.0, 1.0
    record("odin_plain", odin.GQA.ood_evaluate(me, test_tuple))
    # ---- ensemble member vectors (gqa_ensemble.py:95-110)
    for tag, m in (("ens0", model), ("ens1", model2)):
        q = ens.GQA.ood_evaluate(types.SimpleNamespace(model=m), test_tuple)
        out[f"{tag}.sigmoid"] = np.stack([q[k] for k in te[0]])
    # ---- the ensemble's combine step (gqa_ensemble.py:135-215) sits in the script's `__main__` block: its statements are read
    # from the reference file at run time and executed, unmodified, in the module's own namespace, where only `GQA` (-> a
    # stand-in whose `load` switches between the two tiny checkpoints and whose `ood_evaluate` / `get_target` are the reference
    # methods) and `get_tuple` (-> the in-memory test tuple) are replaced
    import ast
    src = open(ens.__file__).read()
    main_if = [n for n in ast.parse(src).body if isinstance(n, ast.If) and "__main__" in ast.get_source_segment(src, n.test)][0]
    main_code = compile(ast.Module(body=main_if.body, type_ignores=[]), ens.__file__, "exec")

    class EnsGQA:
        ood_evaluate = ens.GQA.ood_evaluate
        get_target = ens.GQA.get_target

        def load(self, path):
            self.model = {"ckpt0": model, "ckpt1": model2}[path]
    for method in ("mean", "multiply"):
        with tempfile.TemporaryDirectory() as tmp:
            ns = dict(vars(ens), GQA=EnsGQA, get_tuple=lambda *a, **k: test_tuple)
            args.load, args.test, args.batch_size, args.ensemble_method = "ckpt0,ckpt1", "testdev", BS, method
            args.predict, args.predict_closeset, args.output = False, False, tmp
            exec(main_code, ns)
            record(f"ens_{method}", json.load(open(os.path.join(tmp, "testdev_result.json"))))
    # ---- the two training-time energies (gqa_energy.py:131 logsumexp with --mceLoss, :135 sum of softplus): locals of the
    # UNMODIFIED `GQA.train` loop, read with a trace function while one epoch runs on a throw-away copy of the tiny model
    import copy
    from lxrt.optimization import BertAdam

    def run_energy_train(mce):
        seen, rec = set(), []

        def tracer(frame, event, arg):
            if frame.f_code.co_name != "train" or not frame.f_code.co_filename.endswith("gqa_energy.py"):
                return None

            def local(fr, ev, a):
                e = fr.f_locals.get("E_f")
                if ev == "line" and e is not None and id(e) not in seen:
                    seen.add(id(e))
                    rec.append((fr.f_locals["logit"].detach().clone(), e.detach().clone()))
                return local
            return local
        m = copy.deepcopy(model)
        m.maha = False
        args.mce_loss, args.epochs, args.chart = mce, 1, False
        with tempfile.TemporaryDirectory() as tmp:
            me_tr = types.SimpleNamespace(model=m, optim=BertAdam(list(m.parameters()), lr=1e-3, warmup=0.1, t_total=4),
                                          bce_loss=torch.nn.BCEWithLogitsLoss(), mce_loss=torch.nn.CrossEntropyLoss(ignore_index=-1),
                                          valid_tuple=None, output=tmp, save=lambda name: None)
            tr_tuple = DataTuple(dset_tr, train_tuple.loader, types.SimpleNamespace(evaluate=lambda q: 0.0, dataset=dset_tr))
            sys.settrace(tracer)
            try:
                energy.GQA.train(me_tr, tr_tuple, None)
            finally:
                sys.settrace(None)
        assert len(rec) == len(train_tuple.loader)
        return torch.cat([r[0] for r in rec]).numpy(), torch.cat([r[1] for r in rec]).numpy()
    out["train_energy_sum.logit"], out["train_energy_sum.E_f"] = run_energy_train(False)
    out["train_lse.logit"], out["train_lse.E_f"] = run_energy_train(True)
    # ---- Mahalanobis: sample_estimator (gqa_maha.py:120-189) then ood_evaluate (:191-276)
    model.maha = True
    with tempfile.TemporaryDirectory() as tmp:
        args.output = tmp
        mean, precision = maha.GQA.sample_estimator(me)
    out["maha.mean"] = mean.numpy()
    out["maha.precision"] = np.asarray(precision)                      # float64, sklearn EmpiricalCovariance.precision_
    args.noise = 0.0
    record("maha", maha.GQA.ood_evaluate(me, test_tuple, (mean, precision)))
    args.noise = NOISE
    record("maha_noise", maha.GQA.ood_evaluate(me, test_tuple, (mean, precision)))

    import sklearn
    meta["versions"] = {"torch": torch.__version__, "sklearn": sklearn.__version__, "numpy": np.__version__}
    np.savez_compressed(os.path.join(HERE, "scorers_reference.npz"),
                        **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    with open(os.path.join(HERE, "scorers_reference.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for k, v in meta["results"].items():
        print(k, {a: round(b, 4) for a, b in v.items()})


if __name__ == "__main__":
    main()
