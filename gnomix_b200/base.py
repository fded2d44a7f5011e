"""Base-stage plugins with the reference's surface (src/Base/base.py:10-228,
src/Base/models.py:12-21,195-215).  Training stays where the reference has it
(scikit-learn on the host: it is not the accelerated path); `predict_proba` runs all
W windows in one launch on the GPU through include/gnx.h."""
from __future__ import annotations

import ctypes as C
from time import time

import numpy as np

from . import _lib
from .gbt import _Handle


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def to_device_haplotypes(X):
    """numpy/torch [N, C] integer matrix -> (int8 cuda tensor [N, ld] view with 128-byte
    aligned pitch, ld).  The padded pitch is what lets TMA address the rows."""
    import torch
    if _is_torch(X) and X.is_cuda:
        assert X.dtype in (torch.int8, torch.uint8) and X.dim() == 2
        if X.stride(1) == 1 and X.stride(0) % 16 == 0 and X.data_ptr() % 16 == 0:
            return X, X.stride(0)
        src = X
    else:
        Xn = np.asarray(X)
        if Xn.dtype != np.int8:
            Xn = Xn.astype(np.int8)
        src = torch.from_numpy(np.ascontiguousarray(Xn))
    N, Cc = src.shape
    ld = (Cc + 127) // 128 * 128
    buf = torch.empty((N, ld), dtype=torch.int8, device="cuda")
    if not src.is_cuda and N > 0 and src.stride(1) == 1:
        # host matrix: packed transfer (2 bits per SNP over PCIe, gnx_upload_haplotypes)
        torch.cuda.current_stream().synchronize()
        _lib.check(_lib.lib().gnx_upload_haplotypes(src.data_ptr(), N, src.stride(0), Cc, buf.data_ptr(), ld), "gnx_upload_haplotypes")
    else:
        buf[:, :Cc].copy_(src.view(torch.int8) if src.dtype == torch.uint8 else src, non_blocking=False)
    return buf[:, :Cc], ld


class Base:
    """Same constructor, attributes and methods as the reference's Base
    (src/Base/base.py:10-39).  Subclasses provide `models` (one fitted estimator per
    window) and a `_device_predict(X_dev, ld, N)` that evaluates all windows."""

    def __init__(self, chm_len, window_size, num_ancestry, missing_encoding=2,
                 context=0.5, train_admix=True, n_jobs=None, seed=94305, verbose=False):
        self.C = chm_len
        self.M = window_size
        self.W = self.C // self.M
        self.A = num_ancestry
        self.missing_encoding = missing_encoding
        self.context = context
        self.train_admix = train_admix
        self.n_jobs = n_jobs
        self.seed = seed
        self.verbose = verbose
        self.base_multithread = False
        self.log_inference = False
        self.vectorize = True
        self.time = {}
        self._handles = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handles"] = {}
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._handles = {}

    def init_base_models(self, model_factory):
        self.models = [model_factory() for _ in range(self.W)]

    # -- geometry (src/Base/base.py:41-44, 157-164) ---------------------------
    def pad(self, X):
        pad_left = np.flip(X[:, 0:self.context], axis=1)
        pad_right = np.flip(X[:, -self.context:], axis=1)
        return np.concatenate([pad_left, X, pad_right], axis=1)

    def window_slices(self):
        """padded-coordinate (lo, hi) of every window, the last one absorbing the remainder"""
        rem = self.C - self.M * self.W
        M_ = self.M + 2 * self.context
        out = [(w * self.M, w * self.M + M_) for w in range(self.W - 1)]
        out.append((self.C + 2 * self.context - (M_ + rem), self.C + 2 * self.context))
        return out

    # -- training: host side, as in the reference (src/Base/base.py:99-127) ----
    def train(self, X, y, verbose=True):
        t = time()
        Xp = self.pad(np.asarray(X)) if self.context != 0 else np.asarray(X)
        for w, (lo, hi) in enumerate(self.window_slices()):
            self.models[w] = self.train_base_model(self.models[w], Xp[:, lo:hi], y[:, w])
        self._handles = {}
        self.time["train"] = time() - t

    def train_base_model(self, b, X, y):
        return b.fit(X, y)

    # -- inference --------------------------------------------------------------
    def predict_proba(self, X):
        """X [N, C] int8 (numpy, or a torch cuda tensor for HBM-resident pipelines)
        -> B [N, W, A]: numpy float64 for numpy input (the reference's dtype), a torch
        cuda float32 tensor for cuda input (float32 is what every smoother consumes)."""
        _lib.require_gpu()
        import torch
        t = time()
        Xd, ld = to_device_haplotypes(X)
        assert Xd.shape[1] == self.C, "expected %d SNPs, got %d" % (self.C, Xd.shape[1])
        if _is_torch(X) and X.is_cuda:
            out = self._device_predict(Xd, ld)
        else:
            # host callers get what the reference returns: float64 (the float64 epilogue's own values,
            # not float32 widened); rounding them to float32 gives exactly the device-resident result
            Bd = self._device_predict(Xd, ld, dtype=torch.float64)
            torch.cuda.current_stream().synchronize()
            out = Bd.cpu().numpy()
        self.time["inference"] = time() - t
        return out

    def predict(self, X):
        B = self.predict_proba(X)
        if _is_torch(B):
            return B.argmax(dim=-1)
        return np.argmax(B, axis=-1)

    def evaluate(self, X=None, y=None, B=None):
        from sklearn.metrics import accuracy_score, balanced_accuracy_score
        round_accr = lambda accr: round(np.mean(accr) * 100, 2)
        if X is not None:
            y_pred = self.predict(X)
        elif B is not None:
            y_pred = np.argmax(B, axis=-1)
        else:
            print("Error: Need either SNP input or estimated probabilities to evaluate.")
        accr = round_accr(accuracy_score(y.reshape(-1), y_pred.reshape(-1)))
        accr_bal = round_accr(balanced_accuracy_score(y.reshape(-1), y_pred.reshape(-1)))
        return accr, accr_bal


def _make_liblinear_lr():
    """The reference's estimator (src/Base/models.py:19-21).  scikit-learn >= 1.8 refuses
    liblinear on > 2 classes; OneVsRest(LogisticRegression(liblinear)) is the same
    sigmoid-per-class, row-normalised model (SURVEY.md headline fact 3)."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.multiclass import OneVsRestClassifier
    return OneVsRestClassifier(LogisticRegression(penalty="l2", C=3., solver="liblinear", max_iter=1000))


def lr_weights_of(model, A):
    """(coef [A_rows, M_w], intercept [A_rows]) float64 of a fitted per-window model."""
    if hasattr(model, "estimators_"):
        coef = np.concatenate([e.coef_ for e in model.estimators_], axis=0)
        icpt = np.concatenate([np.atleast_1d(e.intercept_) for e in model.estimators_])
    else:
        coef, icpt = np.asarray(model.coef_), np.atleast_1d(model.intercept_)
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    icpt = np.ascontiguousarray(icpt, dtype=np.float64)
    rows = 1 if A == 2 else A
    assert coef.shape[0] == rows, "window model saw %d classes, need all %d (reference base.py:174 ignores classes_)" % (coef.shape[0], A)
    return coef, icpt


class LinearWindowModel:
    """Minimal fitted per-window estimator (coef_/intercept_/classes_) for models that
    were not trained through scikit-learn in this process (synthetic or imported)."""

    def __init__(self, coef, intercept, A):
        self.coef_ = np.asarray(coef, dtype=np.float64)
        self.intercept_ = np.asarray(intercept, dtype=np.float64)
        self.classes_ = np.arange(A)


class LogisticRegressionBase(Base):
    """src/Base/models.py:12-21."""

    limbs = 0  # fixed-point limbs per weight (0 = library default)
    kernel = 0  # 0 = tcgen05 tensor-core kernel, 1 = dp4a cross-check kernel

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.base_multithread = True
        self.init_base_models(_make_liblinear_lr)

    def set_window_weights(self, coefs, intercepts):
        """Install per-window weights directly (coefs[w] is [A_rows, M_w])."""
        self.models = [LinearWindowModel(c, b, self.A) for c, b in zip(coefs, intercepts)]
        self._handles = {}

    def packed_weights(self):
        cs, bs = [], []
        for (lo, hi), mdl in zip(self.window_slices(), self.models):
            c, b = lr_weights_of(mdl, self.A)
            assert c.shape[1] == hi - lo, "window weight length %d != %d" % (c.shape[1], hi - lo)
            cs.append(c.ravel())
            bs.append(b)
        return np.concatenate(cs), np.ascontiguousarray(np.stack(bs))

    def handle(self):
        import torch
        key = (torch.cuda.current_device(), self.limbs)
        h = self._handles.get(key)
        if h is None:
            coef, icpt = self.packed_weights()
            out = C.c_void_p()
            _lib.check(_lib.lib().gnx_lr_model_create(
                C.byref(out), int(self.A), int(self.C), int(self.M), int(self.context),
                coef.ctypes.data_as(C.c_void_p), icpt.ctypes.data_as(C.c_void_p), int(self.limbs)), "gnx_lr_model_create")
            h = _Handle(out, _lib.lib().gnx_lr_model_destroy)
            self._handles[key] = h
        _lib.check(_lib.lib().gnx_lr_set_kernel(h.ptr, int(self.kernel)), "gnx_lr_set_kernel")
        return h.ptr

    def fixed_point_scale(self):
        return int(_lib.lib().gnx_lr_model_scale(self.handle()))

    def _device_predict(self, Xd, ld, dtype=None):
        import torch
        dtype = dtype or torch.float32
        N = Xd.shape[0]
        Bd = torch.empty((N, self.W, self.A), dtype=dtype, device=Xd.device)
        st = torch.cuda.current_stream().cuda_stream
        fn = _lib.lib().gnx_lr_predict if dtype == torch.float32 else _lib.lib().gnx_lr_predict_f64
        # one call handles <= 65535*128 haplotypes; chunk beyond that
        step = 4_000_000
        for n0 in range(0, N, step):
            n = min(step, N - n0)
            _lib.check(fn(self.handle(), Xd.data_ptr() + n0 * ld, n, ld, Bd[n0:].data_ptr(), st), "gnx_lr_predict")
        return Bd

    def predict_proba_f64(self, X):
        """float64 B on the device (feeds the CRF smoother, which the reference gives float64)."""
        import torch
        _lib.require_gpu()
        Xd, ld = to_device_haplotypes(X)
        return self._device_predict(Xd, ld, dtype=torch.float64)


class CovRSKBase(Base):
    """src/Base/models.py:195-215: one SVC(kernel=CovRSK string kernel, probability=True) per
    window.  Inference runs K2 (bit-plane string kernel against the window's support
    vectors) + K3 (libsvm probability epilogue) on the GPU.  Training stays on the host
    (libsvm through scikit-learn, as in the reference) but takes its Gram matrices from K2
    instead of the reference's Python DP (src/Base/string_kernel.py:91-111)."""

    MS_ALPHA, MS_BETA, MS_SEED = 0.6, 1.0, 37  # CovRSK_DP_triangular_numbers defaults

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.log_inference = True
        self.train_admix = False
        self.models = [None] * self.W
        self.sv_rows = [None] * self.W      # int8 [nSV_w, M_w]: training rows at support_, grouped by class

    @staticmethod
    def cov_sample(M, alpha=0.6, beta=1.0, seed=37):
        """CovSample (src/Base/string_kernel.py:80-89) without touching the global RNG."""
        rs = np.random.RandomState(seed)
        u = rs.rand(max(M - 1, 0))
        Ms = [1]
        for i, m in enumerate(range(2, M + 1)):
            if (1 - (alpha ** (m - Ms[-1] + 1))) * (m ** (-beta)) >= u[i]:
                Ms.append(m)
        return Ms

    def _ms(self):
        lens = [hi - lo for lo, hi in self.window_slices()]
        return np.asarray(self.cov_sample(max(lens), self.MS_ALPHA, self.MS_BETA, self.MS_SEED), dtype=np.int32)

    def set_window_svcs(self, sv_rows, n_support, dual_coef, intercept, probA, probB):
        """Install fitted per-window SVCs directly: sv_rows[w] int8 [nSV_w, M_w] (grouped by
        class), n_support[w] [A], dual_coef[w] [A-1, nSV_w] (SVC._dual_coef_), intercept /
        probA / probB [w] [A(A-1)/2] (SVC._intercept_, probA_, probB_)."""
        self.sv_rows = [np.ascontiguousarray(s, dtype=np.int8) for s in sv_rows]
        self._fitted = dict(n_support=[np.asarray(v, dtype=np.int32) for v in n_support],
                            dual_coef=[np.ascontiguousarray(v, dtype=np.float64) for v in dual_coef],
                            intercept=[np.asarray(v, dtype=np.float64) for v in intercept],
                            probA=[np.asarray(v, dtype=np.float64) for v in probA],
                            probB=[np.asarray(v, dtype=np.float64) for v in probB])
        self._handles = {}

    def _make_handle(self, sv_rows, n_support, dual_coef, intercept, probA, probB):
        P = self.A * (self.A - 1) // 2
        sv = np.concatenate([s.ravel() for s in sv_rows]) if len(sv_rows) else np.zeros(0, np.int8)
        ns = np.ascontiguousarray(np.stack(n_support), dtype=np.int32)
        dc = np.concatenate([d.ravel() for d in dual_coef])
        ic, pa, pb = (np.ascontiguousarray(np.stack(v).reshape(self.W, P), dtype=np.float64) for v in (intercept, probA, probB))
        Ms = self._ms()
        out = C.c_void_p()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().gnx_svc_model_create(C.byref(out), int(self.A), int(self.C), int(self.M), int(self.context), p(sv), p(ns),
                                                   p(dc), p(ic), p(pa), p(pb), p(Ms), len(Ms)), "gnx_svc_model_create")
        return _Handle(out, _lib.lib().gnx_svc_model_destroy)

    def handle(self):
        import torch
        key = torch.cuda.current_device()
        h = self._handles.get(key)
        if h is None:
            if getattr(self, "_fitted", None) is None:   # unpickled from the reference: derive the arrays from its SVCs
                from .pickle_compat import adopt_covrsk
                adopt_covrsk(self)
            f = self._fitted
            h = self._make_handle(self.sv_rows, f["n_support"], f["dual_coef"], f["intercept"], f["probA"], f["probB"])
            self._handles[key] = h
        return h.ptr

    def kernel_window(self, w, X, handle=None):
        """Raw string-kernel values of window w against its support vectors: int32 [N, nSV_w]."""
        import torch
        _lib.require_gpu()
        Xd, ld = to_device_haplotypes(X)
        nsv = len(self.sv_rows[w]) if handle is None else handle[1]
        K = torch.empty((Xd.shape[0], nsv), dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().gnx_svc_kernel_window(self.handle() if handle is None else handle[0].ptr, int(w), Xd.data_ptr(),
                                                    Xd.shape[0], ld, K.data_ptr(), st), "gnx_svc_kernel_window")
        return K

    def _device_predict(self, Xd, ld, dtype=None):
        import torch
        N = Xd.shape[0]
        Bd = torch.empty((N, self.W, self.A), dtype=torch.float64, device=Xd.device)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().gnx_svc_predict(self.handle(), Xd.data_ptr(), N, ld, Bd.data_ptr(), st), "gnx_svc_predict")
        return Bd

    def predict_proba(self, X):
        """[N, W, A] float64 (numpy in -> numpy out, cuda in -> cuda out), as the reference's SVC."""
        _lib.require_gpu()
        import torch
        t = time()
        Xd, ld = to_device_haplotypes(X)
        assert Xd.shape[1] == self.C
        Bd = self._device_predict(Xd, ld)
        if _is_torch(X) and X.is_cuda:
            out = Bd
        else:
            torch.cuda.current_stream().synchronize()
            out = Bd.cpu().numpy()
        self.time["inference"] = time() - t
        return out

    def train(self, X, y, verbose=True):
        """Host training (src/Base/base.py:99-127 with SVC windows): Gram matrices from K2,
        libsvm + Platt scaling through scikit-learn's SVC(kernel="precomputed")."""
        from sklearn import svm
        t = time()
        X = np.asarray(X, dtype=np.int8)
        n = len(X)
        Xp = self.pad(X) if self.context != 0 else X
        # a throw-away model whose "support vectors" are all training rows gives the Gram matrices
        sl = self.window_slices()
        rows = [np.ascontiguousarray(Xp[:, lo:hi]) for lo, hi in sl]
        P = self.A * (self.A - 1) // 2
        ns0 = np.zeros(self.A, np.int32); ns0[0] = n
        tmp = self._make_handle(rows, [ns0] * self.W, [np.zeros((self.A - 1, n))] * self.W, [np.zeros(P)] * self.W,
                                [np.zeros(P)] * self.W, [np.zeros(P)] * self.W)
        fitted = dict(n_support=[], dual_coef=[], intercept=[], probA=[], probB=[])
        svs = []
        for w in range(self.W):
            G = self.kernel_window(w, X, handle=(tmp, n)).cpu().numpy().astype(np.float64)
            mdl = svm.SVC(kernel="precomputed", probability=True, random_state=self.seed).fit(G, y[:, w])
            assert len(mdl.classes_) == self.A, "window %d saw %d classes, need all %d" % (w, len(mdl.classes_), self.A)
            self.models[w] = mdl
            svs.append(rows[w][mdl.support_])
            fitted["n_support"].append(mdl.n_support_)
            fitted["dual_coef"].append(mdl._dual_coef_)
            fitted["intercept"].append(mdl._intercept_)
            fitted["probA"].append(mdl._probA if hasattr(mdl, "_probA") else mdl.probA_)
            fitted["probB"].append(mdl._probB if hasattr(mdl, "_probB") else mdl.probB_)
        self.set_window_svcs(svs, fitted["n_support"], fitted["dual_coef"], fitted["intercept"], fitted["probA"], fitted["probB"])
        self.time["train"] = time() - t
