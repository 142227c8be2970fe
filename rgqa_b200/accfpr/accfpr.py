"""Accuracy-vs-FPR curve for selective (realistic) VQA — numpy only.

Host-side mirror of the reference's ``accfpr/accfpr.py`` (``_multi_clf_curve`` :16-84,
``accfpr_curve`` :87-167) with the same function names, argument meaning and return order, so
``from accfpr.accfpr import accfpr_curve`` call sites keep working.  The curve is integer counting in
float64, so the results are bit-identical to the reference's (pinned by the known-answer fixture
``compute_accfpr/example*.json``, see tests/test_metrics_cpu.py).

A question is a *positive* when its true label is not ``neg_label`` ('UQ'); it counts towards accuracy
only when it is accepted (score >= threshold) AND its prediction equals its label; accepted negatives
are false positives.  Thresholds sweep the distinct scores from high to low.
"""
from __future__ import annotations

import warnings

import numpy as np


def _as_1d(a, name):
    a = np.asarray(a)
    if a.ndim == 2 and a.shape[1] == 1:
        a = a.ravel()
    if a.ndim != 1:
        raise ValueError(f"{name} should be a 1d array, got an array of shape {a.shape} instead.")
    return a


def _multi_clf_curve(y_true, y_pred, y_score, neg_label, sample_weight=None):
    """Cumulative (false positives, correct accepted positives) per distinct score threshold.

    Returns (fps, accs, thresholds), thresholds decreasing.  reference: accfpr/accfpr.py:16-84."""
    y_true, y_pred, y_score = _as_1d(y_true, "y_true"), _as_1d(y_pred, "y_pred"), _as_1d(y_score, "y_score")
    if not (len(y_true) == len(y_pred) == len(y_score)):
        raise ValueError("Found input variables with inconsistent numbers of samples: %r"
                         % [len(y_true), len(y_pred), len(y_score)])
    if not np.all(np.isfinite(y_score)):
        raise ValueError("Input y_score contains NaN or infinity.")
    if sample_weight is not None:
        sample_weight = _as_1d(sample_weight, "sample_weight").astype(np.float64)
        if len(sample_weight) != len(y_true):
            raise ValueError("sample_weight has an inconsistent number of samples")
        keep = sample_weight != 0
        y_true, y_pred, y_score, sample_weight = y_true[keep], y_pred[keep], y_score[keep], sample_weight[keep]
    assert neg_label is not None, "You need to specify a negative label!"

    # stable sort, descending score (:66)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true, y_pred = y_score[order], y_true[order], y_pred[order]
    weight = sample_weight[order] if sample_weight is not None else 1.0

    # last index of every run of tied scores, plus the end of the curve (:78-79)
    distinct_value_indices = np.where(np.diff(y_score))[0]
    threshold_idxs = np.r_[distinct_value_indices, y_true.size - 1]

    correct = np.logical_and(y_true != neg_label, y_true == y_pred) * weight
    negative = (y_true == neg_label) * weight
    accs = np.cumsum(correct, dtype=np.float64)[threshold_idxs]
    fps = np.cumsum(negative, dtype=np.float64)[threshold_idxs]
    return fps, accs, y_score[threshold_idxs]


def accfpr_curve(y_true, y_pred, y_score, neg_label, *, sample_weight=None, drop_intermediate=True):
    """(fpr, acc, thresholds) of the acc-fpr curve.  reference: accfpr/accfpr.py:87-167."""
    y_true_arr = _as_1d(y_true, "y_true")
    fps, accs, thresholds = _multi_clf_curve(y_true_arr, y_pred, y_score, neg_label, sample_weight=sample_weight)

    # drop points collinear with their neighbours (second difference zero on both axes) (:134-140)
    if drop_intermediate and len(fps) > 2:
        optimal_idxs = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(accs, 2)), True])[0]
        fps, accs, thresholds = fps[optimal_idxs], accs[optimal_idxs], thresholds[optimal_idxs]

    # start the curve at (0, 0) (:144-146)
    accs = np.r_[0, accs]
    fps = np.r_[0, fps]
    thresholds = np.r_[thresholds[0] + 1, thresholds]

    if fps[-1] <= 0:
        warnings.warn("No negative samples in y_true, false positive value should be meaningless", RuntimeWarning)
        fpr = np.repeat(np.nan, fps.shape)
    else:
        fpr = fps / fps[-1]

    num_pos = (y_true_arr != neg_label).sum()
    if num_pos <= 0:
        warnings.warn("No positive samples in y_true, accuracy should be meaningless", RuntimeWarning)
        acc = np.repeat(np.nan, accs.shape)
    else:
        acc = accs / num_pos
    return fpr, acc, thresholds
