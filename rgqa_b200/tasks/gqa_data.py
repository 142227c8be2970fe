"""Evaluator and prediction-JSON wire format of the realistic-GQA flows.

Host-side mirror of ``GQAOODEvaluator`` in the reference's ``src/tasks/gqa_data.py:531-605`` (and of the
stand-alone variant in ``compute_accfpr/compute_accfpr.py:13-76``): same constructor, ``evaluate`` and
``dump_result`` signatures, same result keys, same JSON layout, so ``<output>/<test>_predict.json`` and
``<test>_result.json`` written here are interchangeable with the reference's.
"""
from __future__ import annotations

import json

import numpy as np

from ..ood_metrics import accfpr_metrics, f1_binary, ood_performance


class LabelTable:
    """Minimal stand-in for the reference's ``GQADataset``: just ``id2datum[qid]['label']``.

    Build it from the RGQA dataset JSON rows ``{"question_id", "label": {ans: 1.0} | {"UQ": 1.0}, ...}``
    (``compute_accfpr/compute_accfpr.py:80-90``)."""

    def __init__(self, rows):
        self.id2datum = {r['question_id']: r for r in rows}

    @classmethod
    def from_json(cls, path):
        with open(path) as f:
            return cls(json.load(f))


TINY_IMG_NUM = 512     # gqa_data.py:21-22
FAST_IMG_NUM = 5000


class GQADataset:
    """``GQADataset`` of the reference (gqa_data.py:25-60), same files under ``root`` (the reference hard-codes ``data``):
    ``<root>/gqa/<split>.json`` rows ``{img_id, label, question_id, sent}`` for every comma-separated split,
    ``<root>/gqa/trainval_ans2label.json`` / ``trainval_label2ans.json``; attributes ``data``, ``id2datum``, ``ans2label``,
    ``label2ans``, ``num_answers``."""

    def __init__(self, splits: str, root: str = "data"):
        import os
        self.name = splits
        self.splits = splits.split(',')
        self.data = []
        for split in self.splits:
            with open(os.path.join(root, "gqa", "%s.json" % split)) as f:
                self.data.extend(json.load(f))
        print("Load %d data from split(s) %s." % (len(self.data), self.name))
        self.id2datum = {datum['question_id']: datum for datum in self.data}
        with open(os.path.join(root, "gqa", "trainval_ans2label.json")) as f:
            self.ans2label = json.load(f)
        with open(os.path.join(root, "gqa", "trainval_label2ans.json")) as f:
            self.label2ans = json.load(f)
        assert len(self.ans2label) == len(self.label2ans)
        for ans, label in self.ans2label.items():
            assert self.label2ans[label] == ans
        self.num_answers = len(self.ans2label)

    def __len__(self):
        return len(self.data)

    def add_uq_class(self):
        """get_tuple's "one more class" for gqa-uq (gqa_conf.py:36-40)."""
        self.ans2label['UQ'] = self.num_answers
        self.label2ans.append('UQ')
        self.num_answers = len(self.label2ans)
        print(f"One more answer added for {self.name}. Now there are {self.num_answers} answers.")


def feature_tsv_path(splits, root="data"):
    """GQABufferLoader.load_data (gqa_data.py:115-119): testdev splits read their own TSV, everything else Visual Genome's."""
    import os
    testdev = all('testdev' in s for s in splits)
    return os.path.join(root, "vg_gqa_imgfeat", "gqa_testdev_obj36.tsv" if testdev else "vg_gqa_obj36.tsv")


def open_feature_store(splits, root="data", backbone="lxmert", dtype="bf16", topk=None, store_dir=None):
    """The image features of ``splits`` as a memory-mapped store (rgqa_b200.ingest): ``store_dir`` if given, else a
    ``<tsv>.store_<dtype>_<backbone>[_topK]`` directory next to the reference's TSV, converted from it on first use (one
    streaming pass instead of the reference's decode-everything-into-dicts on every start, src/utils.py:16-54)."""
    import os
    from ..ingest import FeatureStore, convert_obj_tsv
    if store_dir is None:
        tsv = feature_tsv_path(splits, root)
        store_dir = "%s.store_%s_%s%s" % (tsv, dtype, backbone, "" if topk in (None, -1) else "_top%d" % topk)
        if not os.path.exists(os.path.join(store_dir, "meta.json")):
            if not os.path.exists(tsv):
                raise FileNotFoundError(f"image features not found: neither {store_dir} nor {tsv} exists")
            print("Converting %s -> %s" % (tsv, store_dir))
            return convert_obj_tsv(tsv, store_dir, topk=None if topk in (None, -1) else topk, dtype=dtype, backbone=backbone)
    store = FeatureStore(store_dir)
    if store.meta["backbone"] != backbone:
        raise ValueError(f"{store_dir} was converted for backbone {store.meta['backbone']!r}, not {backbone!r} (box format differs)")
    return store


class GQAOODEvaluator:
    def __init__(self, dataset, tau=0.5):
        # `dataset` is anything with .id2datum (GQADataset / LabelTable) or a plain {qid: label-dict}
        self.dataset = dataset
        self.tau = tau

    def _label(self, quesid):
        if hasattr(self.dataset, 'id2datum'):
            return self.dataset.id2datum[quesid]['label']
        return self.dataset[quesid]

    def evaluate(self, quesid2ans):
        """quesid2ans: {qid: (answer str, score)} (gqa_data.py:536-581) or the loaded prediction JSON,
        a list of {"questionId", "prediction", "confidence"} (compute_accfpr.py:18-63)."""
        if isinstance(quesid2ans, dict):
            triples = [(qid, ans, score) for qid, (ans, score) in quesid2ans.items()]
        else:
            triples = [(d['questionId'], d['prediction'], d['confidence']) for d in quesid2ans]
        id_acc = 0.
        id_num = 0.
        acc_acc = 0.
        acc_num = 0.
        all_acc = 0.
        ood_preds, ood_targets, clf_targets, clf_preds = [], [], [], []
        for quesid, ans, score in triples:
            label = self._label(quesid)
            if score > self.tau:
                acc_num += 1
                if ans in label:
                    acc_acc += label[ans]
            if 'UQ' not in label:
                id_num += 1
                if ans in label:
                    id_acc += label[ans]
            if ans in label:
                all_acc += label[ans]
            ood_preds.append(score)
            ood_targets.append(int('UQ' not in label))
            clf_preds.append(ans)
            clf_targets.append(list(label.keys())[0])
        ood_preds = np.asarray(ood_preds, dtype=float)
        clf_preds = np.asarray(clf_preds, dtype=str)
        clf_targets = np.asarray(clf_targets, dtype=str)
        results = {}
        results.update(accfpr_metrics(ood_preds, clf_preds, clf_targets, 'UQ', 0.95))
        if acc_num < len(triples):
            results.update(ood_performance(ood_targets, ood_preds))
            results['accuracy_accept'] = acc_acc / acc_num if acc_num > 0 else 0
            results['accuracy'] = all_acc / len(triples)
            results['f1'] = f1_binary(ood_targets, ood_preds > self.tau)
        results['accuracy_indomain'] = id_acc / id_num
        return results

    def dump_result(self, quesid2ans: dict, path):
        """``[{"questionId", "prediction", "confidence": round(conf, 4)}]``, indent 4, sorted keys,
        insertion (= loader) order — gqa_data.py:583-605."""
        with open(path, 'w') as f:
            json.dump(prediction_rows(quesid2ans), f, indent=4, sort_keys=True)


def prediction_rows(quesid2ans: dict):
    return [{'questionId': qid, 'prediction': ans, 'confidence': round(float(conf), 4)}
            for qid, (ans, conf) in quesid2ans.items()]


def dump_results_json(results: dict, path):
    """``<output>/<test>_result.json`` (gqa_conf.py:414-415): numpy scalars become plain floats."""
    with open(path, 'w') as f:
        json.dump({k: (float(v) if isinstance(v, (np.floating, np.integer)) else v) for k, v in results.items()}, f)
