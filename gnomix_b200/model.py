"""Gnomix model API (reference src/model.py:12-214): same constructor, attributes and
methods, so gnomix.py-style drivers and whole-object pickles keep working; the stages
it composes run on the GPU."""
from __future__ import annotations

import pickle
import sys
from copy import deepcopy
from time import time

import numpy as np

from .base import LogisticRegressionBase, CovRSKBase
from .smooth import XGB_Smoother


class Gnomix:

    def __init__(self, C, M, A, S,
                 base=None, smooth=None, mode="default",
                 snp_pos=None, snp_ref=None, snp_alt=None, population_order=None, missing_encoding=2,
                 n_jobs=None, path=None,
                 calibrate=False, context_ratio=0.5, mode_filter=False,
                 seed=94305, verbose=False):
        self.C = C
        self.M = M
        self.A = A
        self.S = S
        self.W = self.C // self.M

        self.path = path
        self.n_jobs = n_jobs
        self.seed = seed
        self.verbose = verbose

        self.snp_pos = snp_pos
        self.snp_ref = snp_ref
        self.snp_alt = snp_alt
        self.population_order = population_order

        self.context = int(self.M * context_ratio)
        self.calibrate = calibrate

        # plugin choice, src/model.py:50-73
        if base is None:
            base = CovRSKBase if mode == "best" else LogisticRegressionBase
            if verbose:
                print("Base models:", base)
        if smooth is None:
            if mode == "fast":
                from .smooth import CRF_Smoother
                smooth = CRF_Smoother
            elif mode == "large":
                raise NotImplementedError("mode='large' (CNN smoother) is outside the accelerated path")
            else:
                smooth = XGB_Smoother
            if verbose:
                print("Smoother:", smooth)

        self.base = base(chm_len=self.C, window_size=self.M, num_ancestry=self.A,
                         missing_encoding=missing_encoding, context=self.context,
                         n_jobs=self.n_jobs, seed=self.seed, verbose=self.verbose)

        self.smooth = smooth(n_windows=self.W, num_ancestry=self.A, smooth_window_size=self.S,
                             n_jobs=self.n_jobs, calibrate=self.calibrate, mode_filter=mode_filter,
                             seed=self.seed, verbose=self.verbose)

        self.time = {}
        self.accuracies = {}
        self.gen_map_df = {}

    def write_gen_map_df(self, gen_map_df):
        self.gen_map_df = deepcopy(gen_map_df)

    def conf_matrix(self, y, y_pred):
        from sklearn.metrics import confusion_matrix
        cm = confusion_matrix(y.reshape(-1), y_pred.reshape(-1))
        indices = sorted(np.unique(np.concatenate((y.reshape(-1), y_pred.reshape(-1)))))
        return cm, indices

    def save(self):
        if self.path is not None:
            pickle.dump(self, open(self.path + "model.pkl", "wb"))

    def train(self, data, retrain_base=True, evaluate=True, verbose=True):
        """src/model.py:104-167, same sequence."""
        train_time_begin = time()
        (X_t1, y_t1), (X_t2, y_t2), (X_v, y_v) = data

        if verbose:
            print("Training base models...")
        self.base.train(X_t1, y_t1)

        if verbose:
            print("Training smoother...")
        B_t2 = self.base.predict_proba(X_t2)
        self.smooth.train(B_t2, y_t2)

        if self.calibrate:  # src/model.py:118-123: calibrated w.r.t. the train1 class distribution
            if verbose:
                print("Fitting calibrator...")
            B_t1 = self.base.predict_proba(X_t1)
            self.smooth.train_calibrator(B_t1, y_t1)

        if evaluate:
            if verbose:
                print("Evaluating model...")
            Acc, CM = {}, {}
            B_t1 = self.base.predict_proba(X_t1)
            y_t1_pred = self.smooth.predict(B_t1)
            y_t2_pred = self.smooth.predict(B_t2)
            Acc["base_train_acc"], Acc["base_train_acc_bal"] = self.base.evaluate(X=None, y=y_t1, B=B_t1)
            Acc["smooth_train_acc"], Acc["smooth_train_acc_bal"] = self.smooth.evaluate(B=None, y=y_t2, y_pred=y_t2_pred)
            CM["train"] = self.conf_matrix(y=y_t1, y_pred=y_t1_pred)
            if X_v is not None:
                B_v = self.base.predict_proba(X_v)
                y_v_pred = self.smooth.predict(B_v)
                Acc["base_val_acc"], Acc["base_val_acc_bal"] = self.base.evaluate(X=None, y=y_v, B=B_v)
                Acc["smooth_val_acc"], Acc["smooth_val_acc_bal"] = self.smooth.evaluate(B=None, y=y_v, y_pred=y_v_pred)
                CM["val"] = self.conf_matrix(y=y_v, y_pred=y_v_pred)
            self.accuracies = Acc
            self.Confusion_Matrices = CM

        if retrain_base:
            if X_v is not None:
                X_t, y_t = np.concatenate([X_t1, X_t2, X_v]), np.concatenate([y_t1, y_t2, y_v])
            else:
                X_t, y_t = np.concatenate([X_t1, X_t2]), np.concatenate([y_t1, y_t2])
            if verbose:
                print("Re-training base models...")
            self.base.train(X_t, y_t)

        self.save()
        self.time["training"] = round(time() - train_time_begin, 2)

    def _base_on_device(self, X):
        """Base stage with the result left in HBM in the dtype the smoother consumes (float32 for the tree
        smoother, float64 for the CRF): the same values the reference's two numpy hand-offs produce, without
        the device -> host -> device round trip of B.  Host matrices are uploaded through the packed transfer."""
        from .base import to_device_haplotypes
        Xd, _ = to_device_haplotypes(X)
        if getattr(self.smooth, "b_dtype", "float32") == "float64" and hasattr(self.base, "predict_proba_f64"):
            return self.base.predict_proba_f64(Xd)
        return self.base.predict_proba(Xd)

    def predict(self, X):
        """src/model.py:169-173."""
        if hasattr(X, "is_cuda") and X.is_cuda:
            return self.smooth.predict(self.base.predict_proba(X))
        import torch
        y = self.smooth.predict(self._base_on_device(X))
        torch.cuda.current_stream().synchronize()
        return y.cpu().numpy().astype(np.int64)

    def predict_proba(self, X):
        """src/model.py:175-179."""
        if hasattr(X, "is_cuda") and X.is_cuda:
            return self.smooth.predict_proba(self.base.predict_proba(X))
        import torch
        p = self.smooth.predict_proba(self._base_on_device(X))
        torch.cuda.current_stream().synchronize()
        return p.cpu().numpy()

    def write_config(self, fname):
        with open(fname, "w") as f:
            for attr in dir(self):
                val = getattr(self, attr)
                if type(val) in [int, float, str, bool, np.float64, np.float32, np.int64]:
                    f.write("{}\t{}\n".format(attr, val))

    def phase(self, X, B=None, verbose=False, want_tracker=False):
        """Gnofix over all individuals (src/model.py:188-214): one launch instead of a
        Python loop over individuals."""
        assert self.smooth is not None, "Smoother is not trained, returning original haplotypes"
        assert self.smooth.gnofix, "Type of Smoother ({}) does not currently support re-phasing".format(self.smooth)
        from .gnofix import phase_all
        return phase_all(self, X, B=B, verbose=verbose, want_tracker=want_tracker)

    # -- host-buffer fast path (include/gnx.h gnx_infer_host) ------------------
    def predict_host(self, X, want_proba=False, chunk_haps=0):
        """Streams a host int8 matrix through Base -> Smoother with overlapped copies.
        Returns labels [N, W] int32 (and proba float32 if asked).  Logistic base + tree smoother
        without a calibrator only (what gnx_infer_host carries); anything else raises."""
        import ctypes as C
        from . import _lib
        from .gbt import GBTForest
        _lib.require_gpu()
        if not isinstance(self.base, LogisticRegressionBase):
            raise TypeError("predict_host needs a LogisticRegressionBase, not %s; use predict()" % type(self.base).__name__)
        if not isinstance(getattr(self.smooth, "model", None), GBTForest):
            raise TypeError("predict_host needs an XGB_Smoother holding a GBTForest; use predict()")
        if getattr(self.smooth, "calibrate", False) and getattr(self.smooth, "calibrator", None) is not None:
            raise NotImplementedError("predict_host does not apply the calibrator; use predict() / predict_proba()")
        if hasattr(X, "data_ptr"):
            import torch
            if X.is_cuda or X.dtype != torch.int8 or X.dim() != 2 or X.stride(1) != 1:
                raise TypeError("predict_host takes a HOST int8 matrix [N, >=C] with unit column stride")
            N, ld, xp = X.shape[0], X.stride(0), X.data_ptr()
        else:
            X = np.ascontiguousarray(X, dtype=np.int8)
            if X.ndim != 2:
                raise TypeError("predict_host takes a 2-D int8 matrix")
            N, ld, xp = X.shape[0], X.strides[0], X.ctypes.data
        if X.shape[1] < self.C:
            raise ValueError("X has %d columns, the model was built for C=%d SNPs" % (X.shape[1], self.C))
        labels = np.empty((N, self.W), dtype=np.int32)
        proba = np.empty((N, self.W, self.A), dtype=np.float32) if want_proba else None
        _lib.check(_lib.lib().gnx_infer_host(self.base.handle(), self.smooth.model.handle(self.smooth.S), xp, N, ld,
                                             proba.ctypes.data if want_proba else None, labels.ctypes.data, int(chunk_haps)),
                   "gnx_infer_host")
        return (labels, proba) if want_proba else labels
