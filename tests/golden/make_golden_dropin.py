"""End-to-end golden of the reference's evaluation flow, produced by the UNMODIFIED reference stack at /root/reference:

    real ``tasks.gqa_model.GQAModel`` / ``GQAModel_maha`` (own constructor: ``LXRTEncoder`` -> ``BertTokenizer`` +
    ``VisualBertForLXRFeature.from_pretrained``, 2 / 1 / 1 layers, MAX_GQA_LENGTH = 30)  ->  ``GQA.ood_evaluate`` of
    ``gqa_conf.py`` (MSP), ``gqa_energy.py`` (top-2 energy), ``gqa_maha.py`` (Mahalanobis, --noise 0 and 1e-3), ``gqa_odin.py``
    (ODIN, --noise 1e-3)  ->  real
    ``GQAOODEvaluator.evaluate`` + ``dump_result``  ->  ``<method>_predict.json`` as the reference writes it.

    python tests/golden/make_golden_dropin.py         # build container only; writes dropin_reference.json.gz

Weights: ``rgqa_b200.synthetic.lxmert_state_dict(variant="rich", head_sigma=0.05)`` loaded with strict=True; inputs:
``synthetic.make_inputs`` features / boxes + sentences drawn from ``tokenizer_vocab.txt``; Mahalanobis statistics:
``synthetic.maha_stats``.  ``from_pretrained("bert-base-uncased")`` is pointed at a local directory (bert_config.json, empty
pytorch_model.bin, that vocab) instead of S3 — SURVEY.md Appendix C; other stubs as in ``make_golden_scorers.py``."""
import collections
import gzip
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import make_golden_scorers as G   # noqa: E402

N, BS, A, SEED = 40, 16, 1842, 2718
NOISE = 1e-3
HEAD_SIGMA = 0.05          # logits of a few units: sigmoids stay distinct (0.2 saturates fp32 sigmoid to 1.0: every MSP label ties)
LAYERS = dict(l_layers=2, x_layers=1, r_layers=1)
BERT_BASE = {"attention_probs_dropout_prob": 0.1, "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "hidden_size": 768,
             "initializer_range": 0.02, "intermediate_size": 3072, "max_position_embeddings": 512, "num_attention_heads": 12,
             "num_hidden_layers": 12, "type_vocab_size": 2, "vocab_size": 30522}


def sentences(n, seed):
    words = [w for w in open(os.path.join(HERE, "tokenizer_vocab.txt"), encoding="utf-8").read().split("\n")
             if w.isalpha() and len(w) > 2]
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        m = int(torch.randint(3, 15, (), generator=g))
        s = " ".join(words[int(k)] for k in torch.randint(0, len(words), (m,), generator=g))
        out.append(s.capitalize() + "?")
    out[3] = ""                                                      # empty question: [CLS] [SEP] only
    out[7] = " ".join(words[:45])                                    # longer than MAX_GQA_LENGTH - 2: truncated (entry.py:45-47)
    return out


def main():
    G.shim()
    sys.argv = ["x", "--llayers", "2", "--xlayers", "1", "--rlayers", "1", "--fromScratch"]
    from rgqa_b200 import synthetic as syn
    with tempfile.TemporaryDirectory() as arch:
        json.dump(BERT_BASE, open(os.path.join(arch, "bert_config.json"), "w"))
        torch.save({}, os.path.join(arch, "pytorch_model.bin"))
        with open(os.path.join(arch, "vocab.txt"), "w", encoding="utf-8") as f:
            f.write(open(os.path.join(HERE, "tokenizer_vocab.txt"), encoding="utf-8").read())
        import lxrt.modeling as M
        import lxrt.tokenization as T
        M.PRETRAINED_MODEL_ARCHIVE_MAP["bert-base-uncased"] = arch
        T.PRETRAINED_VOCAB_ARCHIVE_MAP["bert-base-uncased"] = os.path.join(arch, "vocab.txt")
        from param import args
        import gqa_data
        import tasks.gqa_conf as conf
        import tasks.gqa_energy as energy
        import tasks.gqa_maha as maha
        import tasks.gqa_odin as odin
        from tasks.gqa_model import GQAModel, GQAModel_maha
        torch.manual_seed(0)
        model = GQAModel(A)
        model_maha = GQAModel_maha(A)
    sd = syn.lxmert_state_dict(variant="rich", head_sigma=HEAD_SIGMA, **LAYERS)
    model.load_state_dict(sd, strict=True)
    model_maha.load_state_dict(sd, strict=True)

    inp = syn.make_inputs(N, 20, seed=SEED)
    feats, boxes = inp["feat"], inp["pos"]
    sents = sentences(N, SEED)
    qids = [f"{20000 + i}" for i in range(N)]
    label2ans = [f"answer{j}" for j in range(A)] + ["UQ"]
    # synthetic ground truth: 35 % unanswerable; of the rest 60 % carry the answer the model itself gives (so that the
    # accuracy-vs-FPR curve is not identically zero), the others a random answer
    model.eval()
    with torch.no_grad():
        pred = torch.cat([torch.sigmoid(model(feats[i:i + BS], boxes[i:i + BS], sents[i:i + BS])).max(1)[1] for i in range(0, N, BS)])
    g = torch.Generator().manual_seed(SEED + 1)
    labels = []
    for i in range(N):
        if torch.rand((), generator=g) < 0.35:
            labels.append([["UQ", 1.0]])
        elif torch.rand((), generator=g) < 0.6:
            labels.append([[label2ans[int(pred[i])], 1.0]])
        else:
            labels.append([[label2ans[int(torch.randint(0, A, (), generator=g))], 1.0]])
    Dset = collections.namedtuple("Dset", "num_answers label2ans id2datum")
    dset = Dset(A + 1, label2ans, {q: {"label": dict(l)} for q, l in zip(qids, labels)})
    target = torch.zeros(N, A + 1)
    loader = [(qids[i:i + BS], feats[i:i + BS].clone(), boxes[i:i + BS].clone(), sents[i:i + BS], target[i:i + BS])
              for i in range(0, N, BS)]
    DataTuple = collections.namedtuple("DataTuple", "dataset loader evaluator")
    tup = DataTuple(dset, loader, None)
    args.backbone, args.tau, args.tqdm, args.noise = "lxmert", 0.5, False, 0.0

    out = {"qids": qids, "sents": sents, "labels": labels, "inputs_seed": SEED, "layers": LAYERS, "max_seq_length": 30,
           "num_answers": A, "batch": BS, "head_sigma": HEAD_SIGMA, "predict_json": {}, "results": {}, "versions": {"torch": torch.__version__}}
    mu, P = syn.maha_stats()
    out["scores"], out["answers"] = {}, {}
    captured = {}
    orig_eval = gqa_data.GQAOODEvaluator.evaluate

    def spy(self, quesid2ans):                                        # full-precision scores (the wire file rounds to 4 decimals)
        captured["q"] = dict(quesid2ans)
        return orig_eval(self, quesid2ans)
    gqa_data.GQAOODEvaluator.evaluate = spy
    with tempfile.TemporaryDirectory() as tmp:
        def run(name, fn):
            path = os.path.join(tmp, f"{name}_predict.json")
            res = fn(path)
            q = captured.pop("q")
            out["scores"][name] = [float(q[k][1]) for k in qids]
            out["answers"][name] = [q[k][0] for k in qids]
            out["predict_json"][name] = open(path).read()
            out["results"][name] = {k: float(v) for k, v in res.items()}
            print(name, {k: round(float(v), 4) for k, v in res.items()})
        run("msp", lambda p: conf.GQA.ood_evaluate(types.SimpleNamespace(model=model), tup, dump=p))
        run("energy", lambda p: energy.GQA.ood_evaluate(types.SimpleNamespace(model=model), tup, dump=p))
        me = types.SimpleNamespace(model=model_maha, train_tuple=tup)
        run("maha", lambda p: maha.GQA.ood_evaluate(me, tup, (mu.clone(), P.clone()), dump=p))
        # the two input-gradient scorers on the real model (autograd through the reference LXMERT)
        args.noise, args.temperature = NOISE, 1.0
        out["noise"] = NOISE
        run("odin", lambda p: odin.GQA.ood_evaluate(types.SimpleNamespace(model=model), tup, dump=p))
        run("maha_noise", lambda p: maha.GQA.ood_evaluate(me, tup, (mu.clone(), P.clone()), dump=p))
    with gzip.open(os.path.join(HERE, "dropin_reference.json.gz"), "wt", compresslevel=9) as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
