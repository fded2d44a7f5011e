"""Probability calibrator of the smoother stage (reference src/Smooth/Calibration.py:19-69):
one isotonic regression per class on a subsample, then renormalisation.  `fit` is host-side
scikit-learn, as in the reference; `transform` (the inference side, src/Smooth/smooth.py:48-52)
runs on the device (K7, csrc/calibrate.cu) and is bit-identical to the reference's
scikit-learn path.  The plotting and calibration-error helpers of the reference file are out
of scope."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


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

    # -- device side --------------------------------------------------------------
    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_handles", None)
        return d

    def thresholds(self):
        """[(X_thresholds_, y_thresholds_)] of the fitted per-class models."""
        return [(np.asarray(m.X_thresholds_), np.asarray(m.y_thresholds_)) for m in self.models]

    def handle(self):
        """gnx_cal_t* on the current CUDA device."""
        import torch
        _lib.require_gpu()
        hs = self.__dict__.setdefault("_handles", {})
        key = torch.cuda.current_device()
        h = hs.get(key)
        if h is None:
            thr = self.thresholds()
            dts = {t[0].dtype for t in thr}
            assert len(dts) == 1 and dts <= {np.dtype(np.float32), np.dtype(np.float64)}, "calibrator thresholds must share one float dtype"
            is_f32 = int(thr[0][0].dtype == np.float32)
            n = np.array([len(t[0]) for t in thr], dtype=np.int32)
            x = np.ascontiguousarray(np.concatenate([t[0] for t in thr]), dtype=np.float64)
            y = np.ascontiguousarray(np.concatenate([t[1].astype(t[0].dtype) for t in thr]), dtype=np.float64)
            out = C.c_void_p()
            _lib.check(_lib.lib().gnx_cal_model_create(C.byref(out), self.n_classes, is_f32, n.ctypes.data, x.ctypes.data, y.ctypes.data),
                       "gnx_cal_model_create")
            from .gbt import _Handle
            h = _Handle(out, _lib.lib().gnx_cal_model_destroy)
            hs[key] = h
        return h.ptr

    def transform_device(self, proba, want_proba=True, want_label=False):
        """proba: cuda tensor [..., A] float32 / float64 -> (calibrated float64 cuda tensor or None,
        int32 argmax labels [...] or None)."""
        import torch
        p = proba.contiguous()
        assert p.is_cuda and p.dtype in (torch.float32, torch.float64) and p.shape[-1] == self.n_classes
        rows = p.numel() // self.n_classes
        out = torch.empty(p.shape, dtype=torch.float64, device=p.device) if want_proba else None
        lab = torch.empty(p.shape[:-1], dtype=torch.int32, device=p.device) if want_label else None
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().gnx_calibrate(self.handle(), p.data_ptr(), int(p.dtype == torch.float32), rows,
                                            out.data_ptr() if want_proba else None, lab.data_ptr() if want_label else None, st),
                   "gnx_calibrate")
        return out, lab

    def transform(self, proba):
        if np.any([model is None for model in self.models]):
            print("Warning: No trained calibrator found. Returning original probabilities.")
            return proba
        import torch
        if hasattr(proba, "is_cuda"):
            return self.transform_device(proba.cuda())[0]
        a = np.asarray(proba)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        out, _ = self.transform_device(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy().reshape(a.shape)
