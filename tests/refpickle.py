"""Builds pickles with the object graph the REFERENCE writes (gnomix.py `model.save()`: a `src.model.Gnomix` holding
scikit-learn estimators, an `xgboost.sklearn.XGBClassifier` whose Booster pickles its serialised buffer, or an
`sklearn_crfsuite.CRF` whose model file pickles its bytes) -- without those libraries: stand-in classes are
registered under the libraries' module paths only while `pickle.dumps` runs, so the byte stream names
`sklearn.linear_model._logistic.LogisticRegression`, `xgboost.core.Booster`, ... exactly as a real one does.
Test infrastructure for gnomix_b200/pickle_compat.py."""
from __future__ import annotations

import contextlib
import pickle
import sys
import types

import numpy as np


def _cls(module, name, **body):
    c = type(name, (object,), dict(body))
    c.__module__ = module
    c.__qualname__ = name
    return c


def _make_classes():
    def booster_getstate(self):
        return {"handle": bytearray(self._raw), "feature_names": None, "feature_types": None}

    def fileres_getstate(self):
        return {"name": None, "keep_tempfiles": False, "suffix": ".crfsuite", "prefix": "model", "__FILE_CONTENT__": self._content}

    return {
        "Gnomix": _cls("src.model", "Gnomix"),
        "Base": _cls("src.Base.base", "Base"),
        "LogisticRegressionBase": _cls("src.Base.models", "LogisticRegressionBase"),
        "CovRSKBase": _cls("src.Base.models", "CovRSKBase"),
        "XGB_Smoother": _cls("src.Smooth.models", "XGB_Smoother"),
        "CRF_Smoother": _cls("src.Smooth.models", "CRF_Smoother"),
        "CRFwrap": _cls("src.Smooth.crf", "CRF"),
        "Calibrator": _cls("src.Smooth.Calibration", "Calibrator"),
        "LogisticRegression": _cls("sklearn.linear_model._logistic", "LogisticRegression"),
        "SVC": _cls("sklearn.svm._classes", "SVC"),
        "IsotonicRegression": _cls("sklearn.isotonic", "IsotonicRegression"),
        "XGBClassifier": _cls("xgboost.sklearn", "XGBClassifier"),
        "Booster": _cls("xgboost.core", "Booster", __getstate__=booster_getstate),
        "CRF": _cls("sklearn_crfsuite.estimator", "CRF"),
        "FileResource": _cls("sklearn_crfsuite._fileresource", "FileResource", __getstate__=fileres_getstate),
    }


def _kernel_fn(X, Y):   # stands in for src.Base.string_kernel.CovRSK_DP_triangular_numbers_multithread
    raise RuntimeError("placeholder")


@contextlib.contextmanager
def _registered(classes, extra=()):
    saved = {}
    names = {}
    for c in list(classes.values()) + list(extra):
        names.setdefault(c.__module__, []).append(c)
    try:
        import importlib.util
        for modname in list(names):   # parents of libraries that are not installed here
            parts = modname.split(".")
            for i in range(1, len(parts)):
                parent = ".".join(parts[:i])
                if parent in sys.modules or parent in saved:
                    continue
                try:
                    found = importlib.util.find_spec(parent) is not None
                except (ImportError, ValueError):
                    found = False
                if not found:
                    saved[parent] = None
                    pm = types.ModuleType(parent)
                    pm.__path__ = []
                    sys.modules[parent] = pm
        for modname, objs in names.items():
            saved[modname] = sys.modules.get(modname)
            m = types.ModuleType(modname)
            for o in objs:
                setattr(m, o.__qualname__, o)
            sys.modules[modname] = m
        yield
    finally:
        for modname, old in saved.items():
            if old is None:
                sys.modules.pop(modname, None)
            else:
                sys.modules[modname] = old


def _obj(cls, **attrs):
    o = cls.__new__(cls)
    o.__dict__.update(attrs)
    return o


def reference_pickle(C, M, A, S, context, *, lr=None, svc=None, booster_bytes=None, crfsuite_bytes=None,
                     calibrator=None, mode_filter=False, snp_pos=None, protocol=4):
    """bytes of a reference-style model pickle.
    lr:  (coefs[w] [A_rows, M_w], intercepts[w] [A_rows])            -> LogisticRegressionBase
    svc: (Xfit[w] int8 [n, M_w], support[w], n_support[w], dual_coef[w], intercept[w], probA[w], probB[w]) -> CovRSKBase
    booster_bytes: what xgboost's Booster.__getstate__ puts in state['handle']  -> XGB_Smoother
    crfsuite_bytes: the CRFsuite model file                                     -> CRF_Smoother
    calibrator: [(X_thresholds_, y_thresholds_)] per class."""
    K = _make_classes()
    W = C // M
    common = dict(C=C, M=M, W=W, A=A, missing_encoding=2, context=context, n_jobs=None, seed=94305, verbose=False,
                  log_inference=False, vectorize=True, time={})
    if lr is not None:
        coefs, icpts = lr
        models = [_obj(K["LogisticRegression"], penalty="l2", C=3.0, solver="liblinear", max_iter=1000, coef_=np.asarray(c),
                       intercept_=np.asarray(b), classes_=np.arange(A), n_iter_=np.array([7], dtype=np.int32), n_features_in_=np.asarray(c).shape[1],
                       _sklearn_version="1.0.1") for c, b in zip(coefs, icpts)]
        base = _obj(K["LogisticRegressionBase"], train_admix=True, base_multithread=True, models=models, **common)
    else:
        Xfit, support, n_support, dual, icpt, pA, pB = svc
        fn = _kernel_fn
        fn.__module__, fn.__qualname__ = "src.Base.string_kernel", "CovRSK_DP_triangular_numbers_multithread"
        models = [_obj(K["SVC"], kernel=fn, probability=True, support_=np.asarray(support[w], dtype=np.int32), _n_support=np.asarray(n_support[w], dtype=np.int32),
                       _dual_coef_=np.asarray(dual[w]), dual_coef_=np.asarray(dual[w]), _intercept_=np.asarray(icpt[w]), intercept_=-np.asarray(icpt[w]),
                       probA_=np.asarray(pA[w]), probB_=np.asarray(pB[w]), classes_=np.arange(A), support_vectors_=np.empty((0, 0)),
                       _BaseLibSVM__Xfit=np.asarray(Xfit[w]), _sklearn_version="1.0.1") for w in range(W)]
        base = _obj(K["CovRSKBase"], train_admix=False, base_multithread=False, kernel=fn, models=models, **common)
        base.log_inference = True
    sm_common = dict(W=W, A=A, S=S if S % 2 else S - 1, calibrate=bool(calibrator), mode_filter=mode_filter, n_jobs=None, seed=94305,
                     verbose=False, time={})
    cal = None
    if calibrator is not None:
        cal = _obj(K["Calibrator"], method="Isotonic", n_classes=A,
                   models=[_obj(K["IsotonicRegression"], out_of_bounds="clip", increasing=True, X_thresholds_=np.asarray(x), y_thresholds_=np.asarray(y),
                                X_min_=float(x[0]), X_max_=float(x[-1])) for x, y in calibrator])
    if booster_bytes is not None:
        bst = _obj(K["Booster"], _raw=bytes(booster_bytes))
        xgb = _obj(K["XGBClassifier"], n_estimators=100, max_depth=4, learning_rate=0.1, objective="multi:softprob", missing=float("nan"),
                   n_classes_=A, classes_=np.arange(A), _Booster=bst, kwargs={"num_class": A, "use_label_encoder": False})
        smooth = _obj(K["XGB_Smoother"], gnofix=True, model=xgb, calibrator=cal, **sm_common)
    else:
        fr = _obj(K["FileResource"], _content=bytes(crfsuite_bytes))
        inner = _obj(K["CRF"], algorithm="lbfgs", max_iterations=10000, all_possible_transitions=True, all_possible_states=True, verbose=False,
                     modelfile=fr, _tagger=None, _info_cached=None, training_log_=None)
        smooth = _obj(K["CRF_Smoother"], gnofix=False, model=_obj(K["CRFwrap"], CRF=inner, classes_=[str(a) for a in range(A)]), calibrator=cal, **sm_common)
    model = _obj(K["Gnomix"], C=C, M=M, A=A, S=S, W=W, path=None, n_jobs=None, seed=94305, verbose=False,
                 snp_pos=np.arange(C) * 10 + 5 if snp_pos is None else snp_pos, snp_ref=np.array(["A"] * C), snp_alt=np.array(["G"] * C),
                 population_order=np.array(["P%d" % a for a in range(A)]), context=context, calibrate=bool(calibrator), base=base, smooth=smooth,
                 time={"training": 1.0}, accuracies={}, gen_map_df={})
    extra = [_kernel_fn] if svc is not None else []
    with _registered(K, extra):
        return pickle.dumps(model, protocol=protocol)
