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

    def phase(self, X, B=None, verbose=False, want_tracker=False, crf_extension=False):
        """Gnofix over all individuals (src/model.py:188-214): one launch instead of a
        Python loop over individuals.  `crf_extension=True` additionally admits a CRF smoother: the reference
        refuses that combination (its assert below), the extension is defined in include/gnx.h (gnx_gnofix_crf)
        and has no reference oracle."""
        assert self.smooth is not None, "Smoother is not trained, returning original haplotypes"
        from .smooth import CRFModel
        is_crf_ext = crf_extension and isinstance(getattr(self.smooth, "model", None), CRFModel)
        assert self.smooth.gnofix or is_crf_ext, "Type of Smoother ({}) does not currently support re-phasing".format(self.smooth)
        from .gnofix import phase_all
        return phase_all(self, X, B=B, verbose=verbose, want_tracker=want_tracker)

    # -- host-buffer fast path (include/gnx.h gnx_infer_host_ex) ------------------
    def predict_host(self, X, want_proba=False, chunk_haps=0, phase=False, want_phased=False, crf_extension=False):
        """One call from a HOST haplotype matrix to labels: chunks stream through Base -> [Gnofix] -> Smoother ->
        [Calibrator] on two device slots with the copies overlapped (gnomix.py:48-72 as one pipeline).
        X: numpy / torch-CPU int8 [N, >=C] (pageable or pinned), or a `gnomix_b200.io.PackedHaplotypes`
        (2-bit planes, e.g. from `vcf_to_packed`: a quarter of the bytes cross PCIe, no packing on the way).
        Every plugin combination of the accelerated path is carried: logistic / CovRSK base, XGB / CRF smoother,
        calibrator when `smooth.calibrate` is set, Gnofix with `phase=True` (tree smoother only).
        Returns labels [N, W] int32; with want_proba also proba [N, W, A] (float32 for the tree smoother without
        calibrator, float64 otherwise -- the reference's dtypes); with phase and want_phased also X_phased int8 [N, C]."""
        import ctypes as Ct
        from . import _lib
        from .gbt import GBTForest
        from .smooth import CRFModel
        from .io import PackedHaplotypes
        _lib.require_gpu()
        pipe = _lib.Pipeline()
        if isinstance(self.base, LogisticRegressionBase):
            pipe.lr = self.base.handle()
        elif isinstance(self.base, CovRSKBase):
            pipe.svc = self.base.handle()
        else:
            raise TypeError("predict_host: base %s is not on the accelerated path" % type(self.base).__name__)
        sm = getattr(self.smooth, "model", None)
        if sm is not None and not isinstance(sm, (GBTForest, CRFModel)):
            from .pickle_compat import adopt_smoother_model
            sm = self.smooth.model = adopt_smoother_model(sm, self.smooth.A, self.smooth.S)
        if isinstance(sm, GBTForest):
            pipe.gbt = sm.handle(self.smooth.S)
        elif isinstance(sm, CRFModel):
            pipe.crf = sm.handle()
        else:
            raise TypeError("predict_host: the smoother holds no trained forest / CRF")
        cal = getattr(self.smooth, "calibrator", None)
        use_cal = bool(getattr(self.smooth, "calibrate", False)) and cal is not None
        if use_cal:
            pipe.cal = cal.handle()
        if getattr(self.smooth, "mode_filter", 0) not in (0, 1, False, True, None):
            raise NotImplementedError("predict_host does not apply smooth.mode_filter; use predict()")
        if phase:
            # the reference's refusal (src/model.py:194) stands unless the CRF + Gnofix extension is asked for by name
            ext = bool(crf_extension) and bool(pipe.crf)
            assert (getattr(self.smooth, "gnofix", False) and pipe.gbt) or ext, \
                "Type of Smoother ({}) does not currently support re-phasing".format(self.smooth)
            pipe.phase = 1
            if ext:
                pipe.crf_phase_S = int(self.smooth.S)
        if isinstance(X, PackedHaplotypes):
            N, Cx, ld, xp = X.N, X.C, X.pitch_words, X.words.ctypes.data
            pipe.x_packed = 1
            keep = X
        elif hasattr(X, "data_ptr"):
            import torch
            if X.is_cuda or X.dtype != torch.int8 or X.dim() != 2 or X.stride(1) != 1:
                raise TypeError("predict_host takes a HOST int8 matrix [N, >=C] with unit column stride")
            N, Cx, ld, xp = X.shape[0], X.shape[1], X.stride(0), X.data_ptr()
            keep = X
        else:
            keep = np.ascontiguousarray(X, dtype=np.int8)
            if keep.ndim != 2:
                raise TypeError("predict_host takes a 2-D int8 matrix")
            N, Cx, ld, xp = keep.shape[0], keep.shape[1], (keep.strides[0] if keep.shape[0] else keep.shape[1]), keep.ctypes.data
        if Cx < self.C:
            raise ValueError("X has %d columns, the model was built for C=%d SNPs" % (Cx, self.C))
        if phase and N % 2:
            N -= 1   # the reference's N // 2 reshape drops an odd trailing haplotype
        labels = np.empty((N, self.W), dtype=np.int32)
        p_dtype = np.float64 if (pipe.crf or use_cal) else np.float32
        proba = np.empty((N, self.W, self.A), dtype=p_dtype) if want_proba else None
        xph = np.empty((N, self.C), dtype=np.int8) if (phase and want_phased) else None
        _lib.check(_lib.lib().gnx_infer_host_ex(Ct.byref(pipe), xp, N, ld, proba.ctypes.data if want_proba else None, labels.ctypes.data,
                                                xph.ctypes.data if xph is not None else None, int(chunk_haps)), "gnx_infer_host_ex")
        del keep
        out = (labels,)
        if want_proba:
            out += (proba,)
        if xph is not None:
            out += (xph,)
        return out if len(out) > 1 else labels
