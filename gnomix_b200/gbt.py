"""Flat gradient-boosted forest with xgboost multi:softprob semantics (the model
behind XGB_Smoother, reference src/Smooth/models.py:14-20) and its device handle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class GBTForest:
    """Node arrays concatenated over trees (xgboost dump order).  Tree t owns nodes
    tree_offsets[t]:tree_offsets[t+1]; left/right are relative to the tree's first
    node; feat < 0 marks a leaf; split test `x[feat] < thr` -> left, NaN ->
    default_left; tree t adds its leaf to class t % A."""

    def __init__(self, A, n_features, feat, thr, left, right, default_left, leaf, tree_offsets, base_margin):
        self.A = int(A)
        self.n_features = int(n_features)
        self.feat = np.ascontiguousarray(feat, dtype=np.int32)
        self.thr = np.ascontiguousarray(thr, dtype=np.float32)
        self.left = np.ascontiguousarray(left, dtype=np.int32)
        self.right = np.ascontiguousarray(right, dtype=np.int32)
        self.default_left = np.ascontiguousarray(default_left, dtype=np.uint8)
        self.leaf = np.ascontiguousarray(leaf, dtype=np.float32)
        self.tree_offsets = np.ascontiguousarray(tree_offsets, dtype=np.int32)
        self.base_margin = np.ascontiguousarray(base_margin, dtype=np.float32)
        self.classes_ = np.arange(self.A)
        self.kernel = 0  # 0 = rank-form kernel when eligible, 1 = generic float traversal
        self._handles = {}

    @property
    def n_trees(self):
        return len(self.tree_offsets) - 1

    # -- pickling: device handles are a derived cache, never pickled ---------
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handles"] = {}
        return d

    def to_npz_dict(self):
        return dict(A=self.A, n_features=self.n_features, feat=self.feat, thr=self.thr, left=self.left,
                    right=self.right, default_left=self.default_left, leaf=self.leaf,
                    tree_offsets=self.tree_offsets, base_margin=self.base_margin)

    @classmethod
    def from_npz_dict(cls, d):
        return cls(int(d["A"]), int(d["n_features"]), d["feat"], d["thr"], d["left"], d["right"],
                   d["default_left"], d["leaf"], d["tree_offsets"], d["base_margin"])

    # -- construction from a fitted sklearn HistGradientBoostingClassifier ---
    @classmethod
    def from_hgb(cls, hgb, n_features):
        """Export an sklearn HGB model to xgboost form:
        tree index = iteration*A + class; HGB's `x <= t64` becomes `x < cond32` with
        cond32 = nextafter32(floor32(t64), +inf) -- identical on every float32 x."""
        preds = hgb._predictors
        A = len(preds[0])
        binary = (A == 1)   # HGB fits ONE tree per iteration for two classes (sigmoid of the raw score)
        if binary:
            A = 2           # multi:softprob form: class-0 trees are a single zero leaf, class-1 trees are HGB's,
                            # softmax([0, raw]) = [1 - sigmoid(raw), sigmoid(raw)]
        feat, thr, left, right, dl, leaf, offs = [], [], [], [], [], [], [0]
        for it in range(len(preds)):
            for k in range(A):
                if binary and k == 0:
                    feat.append(-1); thr.append(0.0); left.append(0); right.append(0); dl.append(0); leaf.append(np.float32(0))
                    offs.append(len(feat))
                    continue
                nodes = preds[it][0 if binary else k].nodes
                for nd in nodes:
                    if nd["is_leaf"]:
                        feat.append(-1); thr.append(0.0); left.append(0); right.append(0); dl.append(0)
                        leaf.append(np.float32(nd["value"]))
                    else:
                        t64 = float(nd["num_threshold"])
                        f32 = np.float32(t64)
                        if float(f32) > t64:
                            f32 = np.nextafter(f32, np.float32(-np.inf))
                        cond = np.nextafter(f32, np.float32(np.inf))
                        feat.append(int(nd["feature_idx"])); thr.append(cond)
                        left.append(int(nd["left"])); right.append(int(nd["right"]))
                        dl.append(1 if nd["missing_go_to_left"] else 0); leaf.append(0.0)
                offs.append(len(feat))
        base = np.asarray(hgb._baseline_prediction, dtype=np.float64).reshape(-1)[:A].astype(np.float32)
        if binary:
            base = np.array([0.0, base[0]], dtype=np.float32)
        return cls(A, n_features, feat, thr, left, right, dl, leaf, offs, base)

    @classmethod
    def random(cls, rng, A, S, n_rounds=100, depth=4):
        """Random complete forest of the reference's shape (100 rounds x A trees, depth 4)."""
        F = S * A
        feat, thr, left, right, dl, leaf, offs = [], [], [], [], [], [], [0]
        n_split = 2 ** depth - 1
        for _ in range(n_rounds * A):
            for i in range(n_split):
                feat.append(int(rng.integers(0, F))); thr.append(float(rng.random()))
                left.append(2 * i + 1); right.append(2 * i + 2); dl.append(int(rng.integers(0, 2))); leaf.append(0.0)
            for i in range(2 ** depth):
                feat.append(-1); thr.append(0.0); left.append(0); right.append(0); dl.append(0)
                leaf.append(float(rng.normal() * 0.1))
            offs.append(len(feat))
        return cls(A, F, feat, thr, left, right, dl, leaf, offs, np.full(A, 0.5, dtype=np.float32))

    # -- device side ----------------------------------------------------------
    def handle(self, S):
        """gnx_gbt_t* for smoother width S on the current CUDA device."""
        import torch
        _lib.require_gpu()
        key = (torch.cuda.current_device(), int(S))
        h = self._handles.get(key)
        if h is None:
            assert S * self.A == self.n_features, "forest was trained for %d features, S*A=%d" % (self.n_features, S * self.A)
            out = C.c_void_p()
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            _lib.check(_lib.lib().gnx_gbt_model_create(
                C.byref(out), self.A, int(S), self.n_trees, p(self.feat), p(self.thr), p(self.left), p(self.right),
                p(self.default_left), p(self.leaf), p(self.tree_offsets), p(self.base_margin)), "gnx_gbt_model_create")
            h = _Handle(out, _lib.lib().gnx_gbt_model_destroy)
            self._handles[key] = h
        _lib.check(_lib.lib().gnx_gbt_set_kernel(h.ptr, int(getattr(self, "kernel", 0))), "gnx_gbt_set_kernel")
        return h.ptr

    def predict_proba(self, rows):
        """smoother.model.predict_proba(rows[k, S*A]) -- what gnofix calls
        (reference src/Gnofix/gnofix.py:157).  numpy in -> numpy float32 out."""
        import torch
        rows_t = torch.as_tensor(np.ascontiguousarray(rows, dtype=np.float32)).cuda()
        k, F = rows_t.shape
        assert F == self.n_features
        out = torch.empty((k, self.A), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().gnx_gbt_rows(self.handle(F // self.A), rows_t.data_ptr(), k, out.data_ptr(), st), "gnx_gbt_rows")
        return out.cpu().numpy()


class _Handle:
    def __init__(self, ptr, destroy):
        self.ptr = ptr
        self._destroy = destroy

    def __del__(self):
        try:
            if self.ptr:
                self._destroy(self.ptr)
        except Exception:
            pass
