"""Golden vectors of the reference's ConfusionMeter / LossMeter (test infrastructure).

Run in the build container only (needs /root/reference):

    python oracle/make_golden_metrics.py        # writes tests/golden/metrics.pt

Feeds seeded random (probabilities, targets) batches -- more batches than the window -- through the
UNMODIFIED reference meters (metrics.py:25-137) and records the confusion matrix, per-class precision
and recall after every batch, for a windowed (64) and an unbounded meter, plus the scenario of the
reference's own tests/test_metrics.py:6-37 (identity predictions with one error).
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import OUT, _import_reference  # noqa: E402


def main() -> None:
    _import_reference()
    from marl_classification.metrics import ConfusionMeter, LossMeter

    g = torch.Generator().manual_seed(2024)
    out = {"cases": []}
    for nb_class, window, n_batches, bs in ((10, 4, 9, 16), (45, 64, 70, 8), (7, None, 12, 5)):
        meter = ConfusionMeter(nb_class, window)
        batches, mats, precs, recs = [], [], [], []
        for _ in range(n_batches):
            proba = torch.rand(bs, nb_class, generator=g)
            y = torch.randint(nb_class, (bs,), generator=g)
            meter.add(proba, y)
            batches.append((proba, y))
            mats.append(meter.conf_mat().to(torch.int16))  # counts <= window * batch size
            precs.append(meter.precision().clone())
            recs.append(meter.recall().clone())
        out["cases"].append(dict(nb_class=nb_class, window=window, batches=batches, conf_mat=mats, precision=precs,
                                 recall=recs))
    # tests/test_metrics.py:6-29
    nb_class = 13
    y_pred = torch.eye(nb_class)
    y_pred[0, 0], y_pred[0, 1] = 0.0, 1.0
    meter = ConfusionMeter(nb_class, None)
    meter.add(y_pred, torch.arange(nb_class))
    out["identity_one_error"] = dict(nb_class=nb_class, y_pred=y_pred, conf_mat=meter.conf_mat().clone(),
                                     precision=meter.precision().clone(), recall=meter.recall().clone())
    lm = LossMeter(3)
    vals = [0.5, 0.25, 0.75, 0.5, 1.5]
    means = []
    for v in vals:
        lm.add(v)
        means.append(lm.loss())
    out["loss_meter"] = dict(window=3, values=vals, means=means)
    torch.save(out, os.path.join(OUT, "metrics.pt"))
    print("wrote", os.path.join(OUT, "metrics.pt"))


if __name__ == "__main__":
    main()
