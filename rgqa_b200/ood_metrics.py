"""AUAF / FF95 / FACC and the OOD-detection metrics — numpy only.

Host-side mirror of the reference's ``src/ood_metrics.py`` (``accfpr_metrics`` :118-151,
``ood_performance`` :103-116 and the helpers above it), same names and result keys.  The reference
leans on scikit-learn (``auc``, ``roc_curve``, ``precision_recall_curve``); those three are restated
here from their published algorithm (scikit-learn 1.9.0 ``sklearn/metrics/_ranking.py`` — the version
in this image; the reference pins 1.1.1, environment.yml:111) so the metric path has no third-party
arithmetic.  Bit-exactness is pinned by tests/test_metrics_cpu.py on the reference's own fixture.
"""
from __future__ import annotations

import numpy as np

from .accfpr.accfpr import accfpr_curve


def auc(x, y):
    """Trapezoidal area under a monotonic curve (sklearn.metrics.auc)."""
    x, y = np.asarray(x, dtype=np.float64).ravel(), np.asarray(y, dtype=np.float64).ravel()
    if x.shape[0] < 2:
        raise ValueError("At least 2 points are needed to compute area under curve, but x.shape = %s" % (x.shape,))
    direction = 1
    dx = np.diff(x)
    if np.any(dx < 0):
        if np.all(dx <= 0):
            direction = -1
        else:
            raise ValueError("x is neither increasing nor decreasing : {}.".format(x))
    return float(direction * np.trapezoid(y, x))


def _binary_counts(labels, preds, pos_label=None):
    """Cumulative (fps, tps, thresholds) for a binary problem, thresholds decreasing
    (sklearn confusion_matrix_at_thresholds)."""
    y_true = np.asarray(labels).ravel()
    y_score = np.asarray(preds, dtype=np.float64).ravel()
    if pos_label is None:
        classes = np.unique(y_true)
        if not (classes.size <= 2 and set(classes.tolist()) <= {0, 1} or set(classes.tolist()) <= {-1, 1}):
            raise ValueError("y_true takes values outside {0, 1} / {-1, 1} and pos_label is not specified")
        pos_label = 1
    y_true = (y_true == pos_label)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[order], y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    threshold_idxs = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true.astype(np.float64), dtype=np.float64)[threshold_idxs]
    fps = 1 + threshold_idxs.astype(np.float64) - tps
    return fps, tps, y_score[threshold_idxs]


def roc_curve(labels, preds, pos_label=None, drop_intermediate=True):
    fps, tps, thresholds = _binary_counts(labels, preds, pos_label)
    if drop_intermediate and fps.shape[0] > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thresholds = fps[keep], tps[keep], thresholds[keep]
    tps = np.r_[0.0, tps]
    fps = np.r_[0.0, fps]
    thresholds = np.r_[np.inf, thresholds]
    fpr = fps / fps[-1] if fps[-1] > 0 else np.full(fps.shape, np.nan)
    tpr = tps / tps[-1] if tps[-1] > 0 else np.full(tps.shape, np.nan)
    return fpr, tpr, thresholds


def precision_recall_curve(labels, preds, pos_label=None):
    fps, tps, thresholds = _binary_counts(labels, preds, pos_label)
    ps = tps + fps
    precision = np.where(ps != 0, np.divide(tps, ps, out=np.zeros_like(tps), where=ps != 0), 0.0)
    recall = np.full(tps.shape, 1.0) if tps[-1] == 0 else tps / tps[-1]
    return np.r_[precision[::-1], 1.0], np.r_[recall[::-1], 0.0], thresholds[::-1]


def auroc(preds, labels):
    fpr, tpr, _ = roc_curve(labels, preds)
    return auc(fpr, tpr)


def aupr(preds, labels):
    precision, recall, _ = precision_recall_curve(labels, preds)
    return auc(recall, precision)


def fpr_at_any_tpr(preds, labels, min_tpr=0.95):
    """FPR at the first operating point with TPR >= min_tpr, linearly interpolated (ood_metrics.py:57-75)."""
    fpr, tpr, _ = roc_curve(labels, preds)
    if all(tpr < min_tpr):
        return 0
    elif all(tpr >= min_tpr):
        idxs = [i for i, x in enumerate(tpr) if x >= min_tpr]
        return min(map(lambda idx: fpr[idx], idxs))
    else:
        return np.interp(min_tpr, tpr, fpr)


def fpr_at_95_tpr(preds, labels):
    return fpr_at_any_tpr(preds, labels, 0.95)


def detection_error(preds, labels):
    """Minimum misclassification probability over operating points with TPR >= 95 % (ood_metrics.py:77-100)."""
    fpr, tpr, _ = roc_curve(labels, preds)
    neg_ratio = sum(np.array(labels) == 1) / len(labels)
    pos_ratio = 1 - neg_ratio
    idxs = [i for i, x in enumerate(tpr) if x >= 0.95]
    return min(neg_ratio * (1 - tpr[i]) + pos_ratio * fpr[i] for i in idxs)


def ood_performance(gts, scores):
    """ood_metrics.py:103-116."""
    gts = np.asarray(gts)
    scores = np.asarray(scores, dtype=np.float64)
    measures = {}
    measures['fpr95'] = fpr_at_any_tpr(scores, gts)
    measures['fpr81.4'] = fpr_at_any_tpr(scores, gts, min_tpr=0.814)
    measures['fpr88.9'] = fpr_at_any_tpr(scores, gts, min_tpr=0.889)
    measures['auroc'] = auroc(scores, gts)
    measures['aupr_in'] = aupr(scores, gts)
    measures['aupr_out'] = aupr(-scores, 1 - gts)
    try:
        measures['detection_error'] = detection_error(scores, gts)
    except Exception:
        measures['detection_error'] = np.nan
    return measures


def f1_binary(y_true, y_pred):
    """sklearn.metrics.f1_score for binary {0,1} inputs: 2 tp / (true positives+false negatives + predicted positives)."""
    y_true = np.asarray(y_true).astype(bool)
    y_pred = np.asarray(y_pred).astype(bool)
    tp = float(np.sum(y_true & y_pred))
    denom = float(np.sum(y_true)) + float(np.sum(y_pred))
    return 0.0 if denom == 0 else float(2.0 * tp / denom)


def accfpr_metrics(scores, preds, labels, neg_label, target_acc_ratio=0.95, target_acc=None):
    """AUAF, FF95 (``fpr@0.95acc``) and FACC (``full_acc``).  reference: src/ood_metrics.py:118-151."""
    fpr, acc, _ = accfpr_curve(labels, preds, scores, neg_label)
    if target_acc is None:
        target_acc = acc[-1] * target_acc_ratio
    else:
        target_acc_ratio = round(target_acc / acc[-1], 2)
    key = f'fpr@{target_acc_ratio:.2f}acc'
    if all(acc >= target_acc):
        idxs = [i for i, x in enumerate(acc) if x >= target_acc]
        min_idx = np.argmin(list(map(lambda idx: fpr[idx], idxs)))
        return {'auaf': auc(fpr, acc), key: fpr[min_idx], 'full_acc': acc[-1]}
    return {'auaf': auc(fpr, acc), key: np.interp(target_acc, acc, fpr), 'full_acc': acc[-1]}
