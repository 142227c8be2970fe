"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference at /root/reference.

Runs only in the build container (the GPU box has no /root/reference); the outputs are committed.
    python tests/golden/make_golden.py [--only model|uniter|metrics|tokenizer]

What it does
  model      imports the reference's ``lxrt.modeling.LXRTFeatureExtraction`` and rebuilds the 4-module
             ``logit_fc`` of ``tasks/gqa_model.py:22-27``, loads rgqa_b200.synthetic's seeded state dict
             (strict=True: proves the key schema), runs synthetic batches on the CPU in fp32 and stores
             pooled / logits (+ a few intermediate rows) as .npz.
  uniter     same for ``uniter.modeling.UniterFeatureExtraction``.
  metrics    runs the reference's own ``compute_accfpr/compute_accfpr.py`` evaluator and
             ``src/tasks/gqa_data.py:GQAOODEvaluator`` on ``compute_accfpr/example*.json`` and stores
             the printed dict; also writes a compact copy of the fixture rows (qid, label, prediction,
             confidence) so the known-answer test can run on the GPU box.
  tokenizer  runs the reference ``BertTokenizer`` + ``convert_sents_to_features`` (lxrt/entry.py:36-71)
             with a small synthetic vocab over fixture questions and stores the ids.
Import shims follow SURVEY.md Appendix C (boto3/botocore stubs, sys.path, sys.argv).
"""
import argparse
import gzip
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def _shim_reference():
    for name in ("boto3", "botocore", "botocore.exceptions"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["botocore.exceptions"].ClientError = type("ClientError", (Exception,), {})
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    for p in (os.path.join(REF, "src"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.argv = ["x", "--llayers", "9", "--xlayers", "5", "--rlayers", "5", "--fromScratch"]


def golden_model():
    from rgqa_b200 import synthetic as syn
    _shim_reference()
    import lxrt.modeling as M
    M.VISUAL_CONFIG.l_layers, M.VISUAL_CONFIG.x_layers, M.VISUAL_CONFIG.r_layers = 9, 5, 5
    torch.manual_seed(0)
    enc = M.LXRTFeatureExtraction(M.BertConfig(30522), mode='x')
    head = torch.nn.Sequential(torch.nn.Linear(768, 1536), M.GeLU(), M.BertLayerNorm(1536, eps=1e-12),
                               torch.nn.Linear(1536, 1842))
    enc.eval(); head.eval()
    for variant, L, B, head_sigma, seed_in in (("init", 20, 4, None, 1234), ("rich", 30, 4, None, 77),
                                               ("rich", 20, 6, 0.2, 99)):
        sd = syn.lxmert_state_dict(variant=variant, head_sigma=head_sigma)
        enc_sd = {k[len(syn.LXMERT_PREFIX):]: v for k, v in sd.items() if k.startswith(syn.LXMERT_PREFIX)}
        head_sd = {k[len("logit_fc."):]: v for k, v in sd.items() if k.startswith("logit_fc.")}
        enc.load_state_dict(enc_sd, strict=True)
        head.load_state_dict(head_sd, strict=True)
        inp = syn.make_inputs(B, L, seed=seed_in)
        with torch.no_grad():
            (lang, visn), pooled = enc.bert(inp["ids"], inp["seg"], inp["mask"], visual_feats=(inp["feat"], inp["pos"]))
            pooled2 = enc(inp["ids"], inp["seg"], inp["mask"], visual_feats=(inp["feat"], inp["pos"]))
            assert torch.equal(pooled, pooled2)
            logits = head(pooled)
            emb = enc.bert.embeddings(inp["ids"], inp["seg"])
            vfc = enc.bert.encoder.visn_fc((inp["feat"], inp["pos"]))
        tag = f"lxmert_{variant}_L{L}_B{B}" + ("_head02" if head_sigma else "")
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), pooled=pooled.numpy(), logits=logits.numpy(),
                            lang0=lang[0].numpy(), visn0=visn[0].numpy(), emb0=emb[0].numpy(), vfc0=vfc[0].numpy(),
                            meta=np.array(json.dumps({"variant": variant, "L": L, "B": B, "head_sigma": head_sigma,
                                                      "weights_seed": 9595, "inputs_seed": seed_in,
                                                      "torch": torch.__version__})))
        print(tag, "logits", tuple(logits.shape), float(logits.abs().max()))


def golden_uniter():
    from rgqa_b200 import synthetic as syn
    _shim_reference()
    import uniter.modeling as U
    torch.manual_seed(0)
    enc = U.UniterFeatureExtraction(U.BertConfig(28996))
    head = torch.nn.Sequential(torch.nn.Linear(768, 1536), U.GeLU(), U.BertLayerNorm(1536, eps=1e-12),
                               torch.nn.Linear(1536, 1842))
    enc.eval(); head.eval()
    for variant, B, seed_in in (("init", 4, 1234), ("rich", 4, 55)):
        sd = syn.uniter_state_dict(variant=variant)
        enc.load_state_dict({k[len(syn.UNITER_PREFIX):]: v for k, v in sd.items() if k.startswith(syn.UNITER_PREFIX)},
                            strict=True)
        head.load_state_dict({k[len("logit_fc."):]: v for k, v in sd.items() if k.startswith("logit_fc.")}, strict=True)
        inp = syn.make_inputs(B, 20, seed=seed_in, vocab=28996)
        V = inp["feat"].shape[1]
        with torch.no_grad():  # driven exactly as UniterEncoder.forward does (uniter/entry.py:83-99)
            pooled = enc(input_ids=inp["ids"], token_type_ids=inp["seg"], attention_mask=inp["mask"],
                         visual_feats=inp["feat"], visual_token_type_ids=torch.ones(B, V, dtype=torch.long),
                         visual_attention_mask=torch.ones(B, V, dtype=torch.long), img_pos_feat=inp["pos7"])
            logits = head(pooled)
        tag = f"uniter_{variant}_L20_B{B}"
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), pooled=pooled.numpy(), logits=logits.numpy(),
                            meta=np.array(json.dumps({"variant": variant, "L": 20, "B": B, "weights_seed": 9595,
                                                      "inputs_seed": seed_in, "torch": torch.__version__})))
        print(tag, "logits", tuple(logits.shape), float(logits.abs().max()))


def golden_metrics():
    _shim_reference()
    import importlib.util
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "compute_accfpr"))
    sys.path.insert(0, os.path.join(REF, "compute_accfpr"))
    try:
        spec = importlib.util.spec_from_file_location("ref_compute_accfpr", os.path.join(REF, "compute_accfpr", "compute_accfpr.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ds = json.load(open("example.json"))
        pred = json.load(open("example_predict.json"))
        qid2label = {q['question_id']: q['label'] for q in ds}
        res_cli = mod.ood_evaluate(qid2label, pred)
    finally:
        os.chdir(cwd)
        sys.path.pop(0)
    for m in [k for k in sys.modules if k == "ood_metrics" or k.startswith("accfpr")]:
        del sys.modules[m]
    from tasks.gqa_data import GQAOODEvaluator  # the src/ variant (adds fpr88.9)

    class _DS:
        id2datum = {q['question_id']: q for q in ds}
    res_src = GQAOODEvaluator(_DS(), 0.5).evaluate({p['questionId']: (p['prediction'], p['confidence']) for p in pred})
    out = {"compute_accfpr": {k: float(v) for k, v in res_cli.items()}, "src_gqa_data": {k: float(v) for k, v in res_src.items()},
           "numpy": np.__version__, "sklearn": __import__("sklearn").__version__, "rows": len(pred)}
    json.dump(out, open(os.path.join(HERE, "accfpr_fixture_expected.json"), "w"), indent=1, sort_keys=True)
    rows = [[p['questionId'], list(qid2label[p['questionId']].items())[0][0], list(qid2label[p['questionId']].items())[0][1],
             p['prediction'], p['confidence']] for p in pred]
    assert all(len(qid2label[p['questionId']]) == 1 for p in pred)
    with gzip.open(os.path.join(HERE, "accfpr_fixture_rows.json.gz"), "wt", compresslevel=9) as f:
        json.dump(rows, f, separators=(",", ":"))
    print(json.dumps(out["src_gqa_data"], indent=1))


def golden_tokenizer():
    _shim_reference()
    from lxrt.tokenization import BertTokenizer
    from lxrt.entry import convert_sents_to_features
    ds = json.load(open(os.path.join(REF, "compute_accfpr", "example.json")))
    sents = [q['sent'] for q in ds[::37]][:300]
    sents += ["  What's the  colour of the café's  sign?? ", "", "unaffable supercalifragilistic 12345 x-ray (test) #hash",
              "IS THIS ÜBER-cool naïve façade 北京 ok？", "a " * 40]
    # synthetic vocab: specials at the bert-base-uncased positions, then pieces harvested from the questions
    words = set()
    for s in sents:
        for w in s.lower().replace("?", " ? ").replace(",", " , ").replace("'", " ' ").split():
            words.add(w)
    pieces = set()
    for w in sorted(words):
        if len(w) <= 4 or __import__('zlib').crc32(w.encode()) % 3:
            pieces.add(w)
        else:  # force sub-word splitting for a third of the longer words
            pieces.add(w[:3]); pieces.add("##" + w[3:])
    for ch in "abcdefghijklmnopqrstuvwxyz0123456789?,.'-()#":
        pieces.add(ch); pieces.add("##" + ch)
    vocab = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"] + sorted(pieces)
    vpath = os.path.join(HERE, "tokenizer_vocab.txt")
    with open(vpath, "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    tok = BertTokenizer(vpath, do_lower_case=True)
    out = {}
    for L in (20, 30):
        feats = convert_sents_to_features(sents, L, tok)
        out[str(L)] = {"input_ids": [f.input_ids for f in feats], "input_mask": [f.input_mask for f in feats],
                       "segment_ids": [f.segment_ids for f in feats]}
    out["tokens"] = [tok.tokenize(s.strip()) for s in sents]
    out["sents"] = sents
    with gzip.open(os.path.join(HERE, "tokenizer_golden.json.gz"), "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("tokenizer golden:", len(sents), "sentences; vocab", len(vocab))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    steps = {"model": golden_model, "uniter": golden_uniter, "metrics": golden_metrics, "tokenizer": golden_tokenizer}
    for k, fn in steps.items():
        if a.only in (None, k):
            fn()
