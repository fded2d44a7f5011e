"""Probability calibrator of the smoother stage (reference src/Smooth/Calibration.py:19-69):
one isotonic regression per class on a subsample, then renormalisation.  Host-side (it is off
by default -- config.yaml:24 -- and outside the accelerated path for now); the plotting and
calibration-error helpers of the reference file are out of scope."""
from __future__ import annotations

import numpy as np


class Calibrator:

    def __init__(self, n_classes, method="Isotonic"):
        self.method = method
        self.n_classes = n_classes
        self.models = [None] * n_classes

    def normalize(self, proba):
        if self.n_classes == 2:
            proba[:, 0] = 1. - proba[:, 1]
        else:
            proba /= np.sum(proba, axis=1)[:, np.newaxis]
        proba[np.isnan(proba)] = 1. / self.n_classes
        proba[(1.0 < proba) & (proba <= 1.0 + 1e-5)] = 1.0
        return proba

    def fit(self, proba, y):
        from sklearn.isotonic import IsotonicRegression
        if self.method == "Platt":
            print("Not implemented yet. Using Isotonic regression..")
            self.method = "Isotonic"
        y = np.asarray(y).reshape(-1)
        classes = np.unique(y)
        onehot = (y[:, None] == classes[None, :]).astype(float)      # OneHotEncoder().fit_transform
        for i in range(self.n_classes):
            self.models[i] = IsotonicRegression(out_of_bounds="clip").fit(proba[:, i], onehot[:, i])

    def transform(self, proba):
        if np.any([model is None for model in self.models]):
            print("Warning: No trained calibrator found. Returning original probabilities.")
            return proba
        shape = proba.shape
        flat = proba.reshape(-1, self.n_classes)
        iso = np.zeros((flat.shape[0], self.n_classes))
        for i in range(self.n_classes):
            iso[:, i] = self.models[i].transform(flat[:, i])
        return self.normalize(iso).reshape(*shape)
