"""Golden trajectory of the reference's RP-finetune LOOP, produced by the UNMODIFIED ``GQA.train`` of
``/root/reference/src/tasks/gqa_conf.py`` (:140-245) with ``--samplePair``.

    python tests/golden/make_golden_rp_train.py       # build container only; writes rp_train_reference.npz / .json

``GQA.train`` runs as an unbound method on a stand-in ``self`` (tiny model of ``tiny_vqa.py``, the reference's own
``lxrt.optimization.BertAdam``, ``nn.BCEWithLogitsLoss``), two epochs over three batches, ``random.seed(7)``.  Recorded per step:
the sentences the model was called with (= the random pairing, :155-170), the loss x A (:193-194), and the parameters after
``optim.step()`` (clip 5.0, :197-198).  ``make_golden_train.py`` pins the same recipe on the real LXMERT modules but restates the
loop body; this file pins the loop body itself.  Stubs: see ``make_golden_scorers.py`` (same shim)."""
import collections
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
import make_golden_scorers as G   # noqa: E402

EPOCHS, LR, SEED = 2, 2e-3, 7


def main():
    G.shim()
    from tiny_vqa import TinyVQA
    from param import args
    import tasks.gqa_conf as conf
    from lxrt.optimization import BertAdam

    torch.manual_seed(31)
    model = TinyVQA(G.FEAT, G.D, G.C)
    for p in model.parameters():
        p.data.normal_(0, 0.5)
    init = {k: v.clone() for k, v in model.state_dict().items()}
    qids, feats, boxes, sents, target, labels = G.make_data(33, 48)
    img = {q: f"img{i // 3}" for i, q in enumerate(qids)}            # three questions per image: the redraw loop (:159-160) runs
    Dset = collections.namedtuple("Dset", "num_answers label2ans id2datum")
    label2ans = [f"ans{j}" for j in range(G.C)] + ["UQ"]
    dset = Dset(G.C + 1, label2ans, {q: {"label": l, "img_id": img[q]} for q, l in zip(qids, labels)})
    loader = G.loader_of(qids, feats, boxes, sents, target)
    DataTuple = collections.namedtuple("DataTuple", "dataset loader evaluator")

    steps = []
    bce = torch.nn.BCEWithLogitsLoss()

    def bce_spy(logit, tgt):
        steps[-1]["loss_times_A"] = float(bce(logit, tgt).item() * logit.size(1))
        steps[-1]["target_rows_nonzero"] = int((tgt.abs().sum(1) > 0).sum())
        return bce(logit, tgt)

    fwd = model.forward

    def fwd_spy(f, b, s):
        steps.append({"sent": list(s), "rows": int(f.shape[0])})
        return fwd(f, b, s)
    model.forward = fwd_spy

    optim = BertAdam(list(model.parameters()), lr=LR, warmup=0.1, t_total=len(loader) * EPOCHS)
    step0 = optim.step
    params_after = []

    def step_spy(*a, **k):
        r = step0(*a, **k)
        params_after.append(torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone())
        return r
    optim.step = step_spy

    args.backbone, args.sample_pair, args.mce_loss, args.chart, args.save_all, args.tqdm = "lxmert", True, False, False, False, False
    args.epochs = EPOCHS
    with tempfile.TemporaryDirectory() as tmp:
        me = types.SimpleNamespace(model=model, optim=optim, bce_loss=bce_spy, mce_loss=None, valid_tuple=None, output=tmp,
                                   save=lambda name: None)
        random.seed(SEED)
        conf.GQA.train(me, DataTuple(dset, loader, None), None)
    assert len(steps) == len(params_after) == len(loader) * EPOCHS

    out = {"feat": feats.numpy(), "box": boxes.numpy(), "target": target.numpy(), "params_after": torch.stack(params_after).numpy()}
    out.update({f"init.{k}": v.numpy() for k, v in init.items()})
    meta = {"qids": qids, "sents": sents, "img_of": img, "steps": steps, "epochs": EPOCHS, "lr": LR, "seed": SEED,
            "batch": G.BS, "param_order": [k for k, _ in model.named_parameters()],
            "versions": {"torch": torch.__version__}}
    np.savez_compressed(os.path.join(HERE, "rp_train_reference.npz"), **out)
    with open(os.path.join(HERE, "rp_train_reference.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print([round(s["loss_times_A"], 4) for s in steps], [s["rows"] for s in steps])


if __name__ == "__main__":
    main()
