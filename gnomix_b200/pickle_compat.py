"""Loading a model pickled by the reference (`gnomix.py:26-35`: `pickle.load` of a whole `src.model.Gnomix`
object holding scikit-learn 1.0.1 estimators, an `xgboost.sklearn.XGBClassifier` and, for mode "fast", an
`sklearn_crfsuite.CRF`) in a process that has none of those libraries.

`load_model(path)` unpickles with a `find_class` that
  * maps the reference's own classes (`src.model.Gnomix`, `src.Base.models.*`, `src.Smooth.models.*`,
    `src.Smooth.Calibration.Calibrator`) onto this package's plugins,
  * replaces every class of scikit-learn / scipy / xgboost / sklearn-crfsuite by a state-holding placeholder
    (only fitted attributes are read: `coef_`, `intercept_`, SVC dual coefficients, isotonic thresholds, the
    booster's serialised buffer, the CRFsuite model file) -- so the library versions that wrote the pickle do
    not matter and nothing of them is executed,
and then converts the placeholders into the device-side model forms (`LinearWindowModel`, `GBTForest`,
`CRFModel`).  Pickles written by this package load through the same entry point."""
from __future__ import annotations

import gzip
import io
import pickle
import struct
import warnings

import numpy as np

_STUB_TOPS = ("sklearn", "scipy", "xgboost", "sklearn_crfsuite", "pycrfsuite", "calibration", "lightgbm", "joblib")
_stub_cache = {}


class ForeignObject:
    """Placeholder for an instance of a class that is not imported here: keeps constructor arguments and state."""

    _foreign_module = "?"
    _foreign_name = "?"

    def __new__(cls, *args, **kwargs):
        self = object.__new__(cls)
        if args or kwargs:
            self.__dict__["_ctor_args"] = (args, kwargs)
        return self

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[0], (dict, type(None))):
            for part in state:
                if part:
                    self.__dict__.update(part)
        elif isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    def __reduce__(self):   # placeholders that survive conversion (e.g. isotonic models) re-pickle as placeholders
        return _rebuild_foreign, (self._foreign_module, self._foreign_name, dict(self.__dict__))

    def __repr__(self):
        return "<foreign %s.%s>" % (self._foreign_module, self._foreign_name)


def _rebuild_foreign(module, name, state):
    o = foreign_class(module, name)()
    o.__dict__.update(state)
    return o


def foreign_class(module, name):
    key = (module, name)
    c = _stub_cache.get(key)
    if c is None:
        c = type(name.split(".")[-1], (ForeignObject,), {"_foreign_module": module, "_foreign_name": name, "__module__": __name__})
        _stub_cache[key] = c
    return c


def is_foreign(obj, name=None):
    return isinstance(obj, ForeignObject) and (name is None or obj._foreign_name.split(".")[-1] == name)


def _src_class(module, name):
    from . import base, smooth, model, calibration
    table = {
        "Gnomix": model.Gnomix,
        "Base": base.Base, "LogisticRegressionBase": base.LogisticRegressionBase, "CovRSKBase": base.CovRSKBase,
        "Smoother": smooth.Smoother, "XGB_Smoother": smooth.XGB_Smoother, "CRF_Smoother": smooth.CRF_Smoother,
        "Calibrator": calibration.Calibrator,
    }
    if name in table:
        return table[name]
    # everything else of the reference tree (src.Smooth.crf.CRF, string-kernel functions, experimental bases ...)
    return foreign_class(module, name)


try:
    from pandas.compat.pickle_compat import Unpickler as _BaseUnpickler   # renames of old pandas internals
except Exception:  # pragma: no cover
    _BaseUnpickler = pickle.Unpickler


class ReferenceUnpickler(_BaseUnpickler):
    stub_pandas = False

    def find_class(self, module, name):
        top = module.split(".")[0]
        if top == "src":
            return _src_class(module, name)
        if top in _STUB_TOPS or (self.stub_pandas and top == "pandas"):
            return foreign_class(module, name)
        return super().find_class(module, name)


def _read(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        return f.read()


def loads(data):
    """bytes of a reference (or gnomix_b200) model pickle -> gnomix_b200.Gnomix, converted for the device."""
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            obj = ReferenceUnpickler(io.BytesIO(data)).load()
    except Exception as first:
        # an old pandas DataFrame (model.gen_map_df) that this pandas cannot rebuild: load without it
        class _NoPandas(ReferenceUnpickler):
            stub_pandas = True
        try:
            obj = _NoPandas(io.BytesIO(data)).load()
        except Exception:
            raise first
        if isinstance(getattr(obj, "gen_map_df", None), ForeignObject):
            warnings.warn("gen_map_df of the pickled model could not be rebuilt with this pandas (%s); .msp files will "
                          "lack genetic positions until write_gen_map_df() is called" % first)
            obj.gen_map_df = {}
    return adopt(obj)


def load_model(path, verbose=False):
    """Drop-in for the reference's `load_model` (gnomix.py:26-35): `.pkl` or `.pkl.gz`."""
    if verbose:
        print("Loading model...")
    return loads(_read(path))


# ------------------------------------------------------------------------------------------ conversion
def adopt(model):
    """Turn whatever the unpickler produced into a usable gnomix_b200.Gnomix (in place)."""
    from .model import Gnomix
    from .base import Base, LogisticRegressionBase, CovRSKBase
    from .smooth import Smoother
    if not isinstance(model, Gnomix):
        raise TypeError("the pickle holds a %r, not a Gnomix model" % type(model).__name__)
    d = model.__dict__
    d.setdefault("W", d["C"] // d["M"])
    d.setdefault("time", {})
    d.setdefault("accuracies", {})
    d.setdefault("gen_map_df", {})
    d.setdefault("calibrate", False)
    b = d.get("base")
    if isinstance(b, ForeignObject):
        raise NotImplementedError("base model %s.%s is outside the accelerated path (logistic and CovRSK bases are supported)"
                                  % (b._foreign_module, b._foreign_name))
    if not isinstance(b, Base):
        raise TypeError("model.base is a %r" % type(b).__name__)
    b.__dict__.setdefault("_handles", {})
    if isinstance(b, LogisticRegressionBase):
        adopt_logistic(b)
    elif isinstance(b, CovRSKBase):
        adopt_covrsk(b)
    s = d.get("smooth")
    if isinstance(s, ForeignObject):
        raise NotImplementedError("smoother %s.%s is outside the accelerated path (XGB and CRF smoothers are supported)"
                                  % (s._foreign_module, s._foreign_name))
    if isinstance(s, Smoother):
        s.__dict__.setdefault("calibrator", None)
        s.__dict__.setdefault("mode_filter", 0)
        s.__dict__.setdefault("time", {})
        s.model = adopt_smoother_model(s.model, s.A, s.S)
    return model


def adopt_logistic(base):
    """Per-window scikit-learn LogisticRegression placeholders -> LinearWindowModel (coef_, intercept_)."""
    from .base import LinearWindowModel
    out = []
    for w, m in enumerate(base.models):
        if isinstance(m, ForeignObject) and not hasattr(m, "estimators_"):   # (a OneVsRest wrapper is read as it is)
            if not hasattr(m, "coef_"):
                raise ValueError("window %d of the logistic base is not fitted" % w)
            lw = LinearWindowModel(np.asarray(m.coef_, dtype=np.float64), np.atleast_1d(np.asarray(m.intercept_, dtype=np.float64)), base.A)
            if hasattr(m, "classes_"):
                lw.classes_ = np.asarray(m.classes_)
            out.append(lw)
        else:
            out.append(m)
    base.models = out
    base._handles = {}


def adopt_covrsk(base):
    """Per-window scikit-learn SVC(kernel=callable, probability=True) placeholders -> the arrays K2 + K3 consume.
    scikit-learn keeps the training matrix of a callable-kernel SVC in `_BaseLibSVM__Xfit`; its rows at `support_`
    are the support vectors, already grouped by class (sklearn/svm/_base.py `BaseLibSVM.fit`)."""
    if getattr(base, "_fitted", None) is not None and all(s is not None for s in getattr(base, "sv_rows", [None])):
        return
    svs, fitted = [], dict(n_support=[], dual_coef=[], intercept=[], probA=[], probB=[])
    for w, m in enumerate(base.models):
        g = lambda *names: next((getattr(m, n) for n in names if hasattr(m, n)), None)
        xfit, support = g("_BaseLibSVM__Xfit"), g("support_")
        if xfit is None or support is None:
            raise ValueError("window %d of the CovRSK base holds no fitted SVC (no training matrix / support indices)" % w)
        svs.append(np.ascontiguousarray(np.asarray(xfit)[np.asarray(support)], dtype=np.int8))
        fitted["n_support"].append(np.asarray(g("_n_support", "n_support_"), dtype=np.int32))
        fitted["dual_coef"].append(np.asarray(g("_dual_coef_", "dual_coef_"), dtype=np.float64))
        fitted["intercept"].append(np.asarray(g("_intercept_", "intercept_"), dtype=np.float64))
        fitted["probA"].append(np.asarray(g("_probA", "probA_"), dtype=np.float64))
        fitted["probB"].append(np.asarray(g("_probB", "probB_"), dtype=np.float64))
    base.set_window_svcs(svs, fitted["n_support"], fitted["dual_coef"], fitted["intercept"], fitted["probA"], fitted["probB"])
    base.__dict__.pop("kernel", None)   # the reference's Python string-kernel function (a placeholder here)


def adopt_smoother_model(m, A, S):
    """smoother.model as pickled -> GBTForest / CRFModel (objects of this package pass through)."""
    from .gbt import GBTForest
    from .smooth import CRFModel
    if m is None or isinstance(m, (GBTForest, CRFModel)):
        return m
    if is_foreign(m, "XGBClassifier") or is_foreign(m, "XGBModel") or is_foreign(m, "Booster"):
        return forest_from_foreign_xgb(m, A, S)
    if is_foreign(m, "CRF"):
        return crf_from_foreign(m, A)
    try:   # a real xgboost model, when xgboost happens to be importable
        booster = m.get_booster() if hasattr(m, "get_booster") else m
        if hasattr(booster, "save_raw"):
            from .xgb_io import forest_from_booster_bytes
            return forest_from_booster_bytes(bytes(booster.save_raw()), num_class=A, n_features=S * A)
    except Exception:
        pass
    raise NotImplementedError("smoother model of type %r cannot be converted for the device" % type(m).__name__)


def forest_from_foreign_xgb(m, A, S):
    from .xgb_io import forest_from_booster_bytes
    booster = m if is_foreign(m, "Booster") else getattr(m, "_Booster", None)
    if booster is None:
        raise ValueError("the pickled XGBClassifier holds no fitted booster")
    raw = booster if isinstance(booster, (bytes, bytearray)) else getattr(booster, "handle", None)
    if not isinstance(raw, (bytes, bytearray)):
        raise ValueError("the pickled xgboost Booster holds no serialised model buffer")
    return forest_from_booster_bytes(bytes(raw), num_class=A, n_features=S * A)


# ------------------------------------------------------------------------------------------ CRFsuite model file
def parse_crfsuite_model(buf):
    """CRFsuite 0.12 model file ("lCRF" container, crfsuite/lib/crf/src/crf1d_model.c) -> (labels [L] str,
    attributes [A] str, state_w [A, L], trans_w [L, L]).  Header (48 bytes): magic "lCRF", u32 size, type "FOMC",
    u32 version, num_features, num_labels, num_attrs, off_features, off_labels, off_attrs, off_labelrefs,
    off_attrrefs.  Feature chunk at off_features: "FEAT", u32 size, u32 num, then per feature u32 type (0 = state:
    src attribute -> dst label, 1 = transition: src label -> dst label), u32 src, u32 dst, f64 weight.  Label /
    attribute strings live in CQDB chunks ("CQDB", u32 size, flag, byteorder, bwd_size, bwd_offset, 256 x {u32
    offset, u32 num} table references; record = u32 id, u32 key size, key bytes with NUL; bwd[id] = record offset)."""
    b = bytes(buf)
    if b[:4] != b"lCRF":
        raise ValueError("not a CRFsuite model (magic %r)" % b[:4])
    size, = struct.unpack_from("<I", b, 4)
    if b[8:12] != b"FOMC":
        raise ValueError("CRFsuite model type %r is not a first-order Markov CRF" % b[8:12])
    _ver, nfeat, nlab, nattr, off_feat, off_lab, off_attr, _olr, _oar = struct.unpack_from("<9I", b, 12)

    def cqdb_strings(off, n):
        if b[off:off + 4] != b"CQDB":
            raise ValueError("CRFsuite model: no CQDB chunk at %d" % off)
        _size, _flag, _bo, bwd_size, bwd_off = struct.unpack_from("<5I", b, off + 4)
        out = [None] * n
        for i in range(min(n, bwd_size)):
            (ro,) = struct.unpack_from("<I", b, off + bwd_off + 4 * i)
            if ro == 0:
                continue
            rid, ks = struct.unpack_from("<II", b, off + ro)
            out[i] = b[off + ro + 8: off + ro + 8 + ks].split(b"\0")[0].decode()
        return out

    labels, attrs = cqdb_strings(off_lab, nlab), cqdb_strings(off_attr, nattr)
    if b[off_feat:off_feat + 4] != b"FEAT":
        raise ValueError("CRFsuite model: no FEAT chunk at %d" % off_feat)
    _fsize, fnum = struct.unpack_from("<II", b, off_feat + 4)
    rec = np.frombuffer(b, dtype=np.dtype([("type", "<u4"), ("src", "<u4"), ("dst", "<u4"), ("w", "<f8")]), count=fnum, offset=off_feat + 12)
    state_w, trans_w = np.zeros((nattr, nlab)), np.zeros((nlab, nlab))
    st, tr = rec[rec["type"] == 0], rec[rec["type"] == 1]
    state_w[st["src"], st["dst"]] = st["w"]
    trans_w[tr["src"], tr["dst"]] = tr["w"]
    return labels, attrs, state_w, trans_w


def write_crfsuite_model(labels, attrs, state_w, trans_w):
    """The inverse of parse_crfsuite_model (fixture writer for the round-trip tests; hash tables left empty --
    only the backward arrays, which the parser reads, are filled)."""
    def cqdb(strings):
        recs, offs = bytearray(), []
        base = 24 + 256 * 8
        for i, s in enumerate(strings):
            offs.append(base + len(recs))
            k = s.encode() + b"\0"
            recs += struct.pack("<II", i, len(k)) + k
        bwd_off = base + len(recs)
        body = bytes(recs) + b"".join(struct.pack("<I", o) for o in offs)
        size = base + len(body)
        return b"CQDB" + struct.pack("<5I", size, 0, 0x62445371, len(strings), bwd_off) + b"\0" * (256 * 8) + body

    feats = []
    A, L = np.asarray(state_w).shape
    for a in range(A):
        for y in range(L):
            feats.append(struct.pack("<IIId", 0, a, y, float(state_w[a][y])))
    for i in range(L):
        for j in range(L):
            feats.append(struct.pack("<IIId", 1, i, j, float(trans_w[i][j])))
    feat = b"".join(feats)
    feat_chunk = b"FEAT" + struct.pack("<II", 12 + len(feat), len(feats)) + feat
    lab, att = cqdb(list(labels)), cqdb(list(attrs))
    off_feat = 48
    off_lab = off_feat + len(feat_chunk)
    off_att = off_lab + len(lab)
    end = off_att + len(att)
    hdr = b"lCRF" + struct.pack("<I", end) + b"FOMC" + struct.pack("<9I", 100, len(feats), L, A, off_feat, off_lab, off_att, 0, 0)
    return hdr + feat_chunk + lab + att


def crf_from_foreign(m, A):
    """`src.Smooth.crf.CRF` placeholder (its `.CRF` is an sklearn_crfsuite.CRF whose `modelfile` pickles the
    CRFsuite model bytes as `__FILE_CONTENT__`, sklearn_crfsuite/_fileresource.py) -> CRFModel with
    state_w[a, y] / trans_w[i, j] ordered by the integer value of the attribute / label strings
    (src/Smooth/crf.py:17-33 names attribute a `str(a)` and label y `str(y)`)."""
    from .smooth import CRFModel
    inner = getattr(m, "CRF", m)
    mf = getattr(inner, "modelfile", None)
    content = None
    if mf is not None:
        content = getattr(mf, "__dict__", {}).get("__FILE_CONTENT__")
        if content is None and isinstance(getattr(mf, "_state", None), dict):
            content = mf._state.get("__FILE_CONTENT__")
    if content is None:
        raise ValueError("the pickled CRF holds no CRFsuite model file")
    labels, attrs, sw, tw = parse_crfsuite_model(content)
    try:
        lo = np.argsort([int(s) for s in labels])
        ao = np.argsort([int(s) for s in attrs])
    except (TypeError, ValueError):
        raise ValueError("CRFsuite labels / attributes are not the integers the reference writes: %r / %r" % (labels, attrs))
    if len(lo) != A or len(ao) != A:
        raise ValueError("CRFsuite model has %d labels / %d attributes, the smoother was built for %d" % (len(lo), len(ao), A))
    return CRFModel(sw[np.ix_(ao, lo)], tw[np.ix_(lo, lo)])
