"""Smoother-stage plugins with the reference's surface (src/Smooth/smooth.py:7-92,
src/Smooth/models.py:8-32).  `predict_proba` / `predict` run on the GPU through
include/gnx.h; `slide_window`'s [N*W, S*A] matrix is never materialised."""
from __future__ import annotations

import ctypes as C
from time import time

import numpy as np

from . import _lib
from .gbt import GBTForest, _Handle


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Smoother:
    """src/Smooth/smooth.py:7-27 (same constructor and attributes)."""

    def __init__(self, n_windows, num_ancestry, smooth_window_size=75, model=None,
                 calibrate=None, n_jobs=None, seed=None, mode_filter=0, verbose=False):
        self.W = n_windows
        self.A = num_ancestry
        self.S = smooth_window_size if smooth_window_size % 2 else smooth_window_size - 1
        self.model = model
        self.calibrate = calibrate
        self.calibrator = None
        self.mode_filter = mode_filter
        self.n_jobs = n_jobs
        self.seed = seed
        self.verbose = verbose
        self.gnofix = False
        self.time = {}

    def process_base_proba(self, B, y=None):
        return B, y

    def train(self, B, y):
        assert len(np.unique(y)) == self.A, "Smoother training data does not include all populations"
        t = time()
        self._fit(np.asarray(B), np.asarray(y))
        self.time["train"] = time() - t

    # device-side evaluation: B cuda tensor -> (proba cuda [N,W,A] or None, label cuda int32 [N,W] or None)
    def _device_smooth(self, Bd, want_proba=True, want_label=True):
        raise NotImplementedError

    def _to_device_B(self, B, dtype):
        import torch
        if _is_torch(B):
            return B.to(device="cuda", dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(B), dtype=np.float32 if dtype == torch.float32 else np.float64)).cuda()

    def predict_proba(self, B):
        """B [N, W, A] -> proba [N, W, A] (src/Smooth/smooth.py:40-56)."""
        _lib.require_gpu()
        import torch
        t = time()
        proba, _ = self._device_smooth(B, want_proba=True, want_label=False)
        if self.calibrate:
            if self.calibrator is None:
                print("No calibrator found, returning original probabilities.")
            else:
                proba, _ = self.calibrator.transform_device(proba)   # K7, float64 as the reference returns
        if not (_is_torch(B) and B.is_cuda):
            torch.cuda.current_stream().synchronize()
            proba = proba.cpu().numpy()
        self.time["inference"] = time() - t
        return proba

    def predict(self, B):
        """argmax of predict_proba (src/Smooth/smooth.py:58-65); the argmax is fused
        into the kernel (first maximum wins, as np.argmax)."""
        _lib.require_gpu()
        import torch
        if self.calibrate and self.calibrator is not None:
            proba, _ = self._device_smooth(B, want_proba=True, want_label=False)
            _, label = self.calibrator.transform_device(proba, want_proba=False, want_label=True)
        else:
            _, label = self._device_smooth(B, want_proba=False, want_label=True)
        if self.mode_filter:
            label = mode_filter_device(label, self.mode_filter, self.A)
        if _is_torch(B) and B.is_cuda:
            return label
        torch.cuda.current_stream().synchronize()
        return label.cpu().numpy().astype(np.int64)

    def evaluate(self, B=None, y=None, y_pred=None):
        from sklearn.metrics import accuracy_score, balanced_accuracy_score
        round_accr = lambda accr: round(np.mean(accr) * 100, 2)
        if B is not None:
            y_pred = self.predict(B)
        elif y_pred is None:
            print("Error: Need either Base probabilities or y predictions.")
        accr = round_accr(accuracy_score(y.reshape(-1), y_pred.reshape(-1)))
        accr_bal = round_accr(balanced_accuracy_score(y.reshape(-1), y_pred.reshape(-1)))
        return accr, accr_bal


def mode_filter_device(label, size, A):
    """mode_filter of src/Smooth/utils.py:31-46 over the window axis of a cuda label tensor [N, W]: positions
    ends <= i < W - ends (ends = size // 2) take the most frequent label of pred[i-ends : i+ends+1] when it is
    unique among the most frequent ones, else keep the centre value; the borders are untouched (size 1 / True
    is therefore a no-op, as in the reference).  Host-side glue in torch ops, not a hot kernel."""
    import torch
    ends = int(size) // 2
    N, W = label.shape
    if ends == 0 or W <= 2 * ends:
        return label
    lab = label.long()
    oh = torch.nn.functional.one_hot(lab, A).to(torch.int32)
    cs = torch.cat([torch.zeros((N, 1, A), dtype=torch.int32, device=label.device), oh.cumsum(1, dtype=torch.int32)], dim=1)
    win = cs[:, 2 * ends + 1:] - cs[:, :W - 2 * ends]          # counts of window i-ends..i+ends for i = ends..W-ends-1
    mx = win.max(dim=-1, keepdim=True).values
    at_max = win == mx
    unique = at_max.sum(-1) == 1
    mode = at_max.to(torch.int8).argmax(-1)
    out = label.clone()
    mid = label[:, ends:W - ends]
    out[:, ends:W - ends] = torch.where(unique, mode.to(label.dtype), mid)
    return out


def _train_calibrator(self, B, y, frac=0.05):
    """src/Smooth/smooth.py:81-92."""
    from .calibration import Calibrator
    calibrate = self.calibrate
    self.calibrate = False
    idxs = np.random.choice(len(B), int(frac * len(B)), replace=False)
    proba = np.asarray(self.predict_proba(np.asarray(B)[idxs])).reshape(-1, self.A)
    self.calibrate = calibrate
    self.calibrator = Calibrator(self.A)
    self.calibrator.fit(proba, np.asarray(y)[idxs].reshape(-1))


Smoother.train_calibrator = _train_calibrator


def host_slide_window(B, S):
    """Training-time twin of slide_window (src/Smooth/utils.py:4-29): only Smoother.train
    needs the materialised matrix (on a subsample), inference never does."""
    N, W, A = B.shape
    pad = (S + 1) // 2
    Bp = np.concatenate([np.flip(B[:, 0:pad, :], axis=1), B, np.flip(B[:, -pad:, :], axis=1)], axis=1).astype(np.float32)
    view = np.lib.stride_tricks.sliding_window_view(Bp.reshape(N, -1), S * A, axis=1)[:, ::A][:, :W]
    return np.ascontiguousarray(view).reshape(N * W, S * A)


class XGB_Smoother(Smoother):
    """src/Smooth/models.py:8-24.  `model` is a GBTForest (xgboost multi:softprob semantics); a model pickled by
    the reference (an xgboost XGBClassifier) is converted on first use (pickle_compat).  Inference is the
    accelerated path.  Training is NOT xgboost: `_fit` uses scikit-learn's HistGradientBoostingClassifier (255-bin
    histogram splits) with the reference's hyper-parameters (100 rounds x A trees, depth 4, lr 0.1, lambda 1) on at
    most `max_rows` (default 400 000) randomly chosen rows of the slid matrix, exported to the same forest form --
    a stand-in so that train -> predict runs end to end, not a reproduction of xgboost's exact-split training."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.gnofix = True
        assert self.W >= 2 * self.S, "Smoother size to large for given window size. "
        self.model = None

    def process_base_proba(self, B, y=None):
        return host_slide_window(np.asarray(B), self.S), (None if y is None else np.asarray(y).reshape(-1))

    def _fit(self, B, y, max_rows=400_000):
        Xs, ys = self.process_base_proba(B, y)
        if len(Xs) > max_rows:
            rng = np.random.default_rng(self.seed)
            idx = rng.choice(len(Xs), max_rows, replace=False)
            Xs, ys = Xs[idx], ys[idx]
        from sklearn.ensemble import HistGradientBoostingClassifier
        hgb = HistGradientBoostingClassifier(max_iter=100, max_depth=4, learning_rate=0.1, l2_regularization=1.0,
                                             max_leaf_nodes=None, early_stopping=False, random_state=self.seed)
        hgb.fit(Xs, ys)
        self.model = GBTForest.from_hgb(hgb, self.S * self.A)

    def _device_smooth(self, B, want_proba=True, want_label=True):
        import torch
        if self.model is not None and not isinstance(self.model, GBTForest):
            from .pickle_compat import adopt_smoother_model
            self.model = adopt_smoother_model(self.model, self.A, self.S)   # a foreign (reference-pickled) model
        assert isinstance(self.model, GBTForest), "XGB_Smoother has no trained forest"
        Bd = self._to_device_B(B, torch.float32)
        N, W, A = Bd.shape
        assert W == self.W and A == self.A
        proba = torch.empty((N, W, A), dtype=torch.float32, device="cuda") if want_proba else None
        label = torch.empty((N, W), dtype=torch.int32, device="cuda") if want_label else None
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().gnx_gbt_smooth(self.model.handle(self.S), Bd.data_ptr(), N, W,
                                              proba.data_ptr() if want_proba else None,
                                              label.data_ptr() if want_label else None, st), "gnx_gbt_smooth")
        return proba, label


class CRFModel:
    """Linear-chain CRF weights in the form sklearn_crfsuite exposes them
    (state_features_, transition_features_): state_w[a, y], trans_w[i, j]."""

    def __init__(self, state_w, trans_w):
        self.state_w = np.ascontiguousarray(state_w, dtype=np.float64)
        self.trans_w = np.ascontiguousarray(trans_w, dtype=np.float64)
        self.classes_ = [str(i) for i in range(self.state_w.shape[1])]
        self._handles = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handles"] = {}
        return d

    def handle(self):
        import torch
        key = torch.cuda.current_device()
        h = self._handles.get(key)
        if h is None:
            out = C.c_void_p()
            A, L = self.state_w.shape
            _lib.check(_lib.lib().gnx_crf_model_create(C.byref(out), A, L, self.state_w.ctypes.data_as(C.c_void_p),
                                                       self.trans_w.ctypes.data_as(C.c_void_p)), "gnx_crf_model_create")
            h = _Handle(out, _lib.lib().gnx_crf_model_destroy)
            self._handles[key] = h
        return h.ptr


class CRF_Smoother(Smoother):
    """src/Smooth/models.py:27-32 (+ src/Smooth/crf.py).  proba = CRF marginals (float64)."""

    b_dtype = "float64"   # the reference hands the CRF the base's float64 probabilities (Gnomix keeps them so on the device)

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.model = None

    def _fit(self, B, y):
        from .crf_train import fit_crf
        sw, tw = fit_crf(B, y, self.A)
        self.model = CRFModel(sw, tw)

    def _device_smooth(self, B, want_proba=True, want_label=True):
        import torch
        if self.model is not None and not isinstance(self.model, CRFModel):
            from .pickle_compat import adopt_smoother_model
            self.model = adopt_smoother_model(self.model, self.A, self.S)
        assert isinstance(self.model, CRFModel), "CRF_Smoother has no trained model"
        Bd = self._to_device_B(B, torch.float64)
        N, W, A = Bd.shape
        proba = torch.empty((N, W, A), dtype=torch.float64, device="cuda") if want_proba else None
        label = torch.empty((N, W), dtype=torch.int32, device="cuda") if want_label else None
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().gnx_crf_smooth(self.model.handle(), Bd.data_ptr(), N, W,
                                              proba.data_ptr() if want_proba else None,
                                              label.data_ptr() if want_label else None, st), "gnx_crf_smooth")
        return proba, label
