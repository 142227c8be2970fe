"""BASELINE config 3 end to end on the UNMODIFIED reference: the real ``uniter.uniter.GQAUNITER`` (own constructor:
``UniterEncoder`` -> ``uniter.tokenization.BertTokenizer`` + ``UniterFeatureExtraction.from_pretrained("bert-base-cased")``, 12 layers)
driven by ``gqa_conf.GQA.ood_evaluate`` (MSP) and ``gqa_energy.GQA.ood_evaluate``, sentences in, ``*_predict.json`` out.

    python tests/golden/make_golden_dropin_uniter.py      # build container only; writes dropin_uniter_reference.json.gz

Weights ``synthetic.uniter_state_dict(variant="rich", head_sigma=0.05)`` (strict=True), features / 7-d boxes of
``synthetic.make_inputs(vocab=28996)``, sentences from ``tokenizer_vocab.txt``.  Same stubs and local-archive redirection as
``make_golden_dropin.py``."""
import collections
import gzip
import json
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import make_golden_scorers as G   # noqa: E402
import make_golden_dropin as DI   # noqa: E402

N, BS, A, SEED, HEAD_SIGMA = 24, 16, 1842, 1618, 0.05


def main():
    G.shim()
    sys.argv = ["x", "--fromScratch"]
    from rgqa_b200 import synthetic as syn
    with tempfile.TemporaryDirectory() as arch:
        cfg = dict(DI.BERT_BASE, vocab_size=28996)
        json.dump(cfg, open(os.path.join(arch, "bert_config.json"), "w"))
        torch.save({}, os.path.join(arch, "pytorch_model.bin"))
        with open(os.path.join(arch, "vocab.txt"), "w", encoding="utf-8") as f:
            f.write(open(os.path.join(HERE, "tokenizer_vocab.txt"), encoding="utf-8").read())
        import uniter.modeling as UM
        import uniter.tokenization as UT
        UM.PRETRAINED_MODEL_ARCHIVE_MAP["bert-base-cased"] = arch
        UT.PRETRAINED_VOCAB_ARCHIVE_MAP["bert-base-cased"] = os.path.join(arch, "vocab.txt")
        from param import args
        import gqa_data  # noqa: F401
        import tasks.gqa_conf as conf
        import tasks.gqa_energy as energy
        from uniter.uniter import GQAUNITER
        torch.manual_seed(0)
        model = GQAUNITER(A)
    model.load_state_dict(syn.uniter_state_dict(variant="rich", head_sigma=HEAD_SIGMA), strict=True)
    model.eval()

    inp = syn.make_inputs(N, 20, seed=SEED, vocab=28996)
    feats, boxes = inp["feat"], inp["pos7"]
    sents = DI.sentences(N, SEED)
    qids = [f"{30000 + i}" for i in range(N)]
    label2ans = [f"answer{j}" for j in range(A)] + ["UQ"]
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
    tup = collections.namedtuple("DataTuple", "dataset loader evaluator")(dset, loader, None)
    args.backbone, args.tau, args.tqdm = "uniter", 0.5, False
    out = {"qids": qids, "sents": sents, "labels": labels, "inputs_seed": SEED, "num_answers": A, "batch": BS,
           "head_sigma": HEAD_SIGMA, "max_seq_length": 20, "predict_json": {}, "results": {}, "versions": {"torch": torch.__version__}}
    with tempfile.TemporaryDirectory() as tmp:
        for name, drv in (("msp", conf), ("energy", energy)):
            path = os.path.join(tmp, f"{name}_predict.json")
            res = drv.GQA.ood_evaluate(types.SimpleNamespace(model=model), tup, dump=path)
            out["predict_json"][name] = open(path).read()
            out["results"][name] = {k: float(v) for k, v in res.items()}
            print(name, {k: round(float(v), 4) for k, v in res.items()})
    with gzip.open(os.path.join(HERE, "dropin_uniter_reference.json.gz"), "wt", compresslevel=9) as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
