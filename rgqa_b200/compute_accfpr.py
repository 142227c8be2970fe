"""Stand-alone metric CLI, mirror of the reference's ``compute_accfpr/compute_accfpr.py`` (README.md:39-47):

    python -m rgqa_b200.compute_accfpr example.json example_predict.json [--tau 0.5]

``example.json``: RGQA dataset rows ``{"question_id", "label": {answer: 1.0} | {"UQ": 1.0}, ...}``;
``example_predict.json``: ``[{"questionId", "prediction", "confidence"}, ...]``.  Prints the result dict."""
from __future__ import annotations

import argparse
import json

from .tasks.gqa_data import GQAOODEvaluator


def ood_evaluate(qid2label, quesid2ans, tau=0.5):
    return GQAOODEvaluator(qid2label, tau).evaluate(quesid2ans)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("dataset", nargs="?", default="example.json")
    ap.add_argument("predictions", nargs="?", default="example_predict.json")
    ap.add_argument("--tau", type=float, default=0.5)
    a = ap.parse_args(argv)
    with open(a.dataset) as f:
        qid2label = {q['question_id']: q['label'] for q in json.load(f)}
    with open(a.predictions) as f:
        pred = json.load(f)
    result = ood_evaluate(qid2label, pred, a.tau)
    print({k: float(v) for k, v in result.items()})
    return result


if __name__ == "__main__":
    main()
