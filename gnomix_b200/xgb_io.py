"""xgboost booster buffers -> GBTForest, without xgboost.

A reference `.pkl` (gnomix.py:26-35) embeds an `xgboost.sklearn.XGBClassifier` whose `Booster` pickles
itself as the byte buffer `XGBoosterSerializeToBuffer` returns (python-package/xgboost/core.py,
`Booster.__getstate__`).  Depending on the xgboost version that wrote it, that buffer is

  * xgboost 1.0 - 1.5:  b"CONFIG-offset:" + int64 offset + <legacy binary model, `offset` bytes> + JSON config
                        (src/learner.cc `LearnerIO::Save`, the reference pins 1.1.1);
  * xgboost >= 1.6:     UBJSON `{"Model": {...}, "Config": {...}}`;
  * or, for `save_model` / `save_raw` buffers, the bare legacy binary model (optionally behind b"binf"),
    the JSON model `{"learner": ...}` or its UBJSON form.

Legacy binary model (little endian; src/learner.cc `LearnerIO::SaveModel(dmlc::Stream*)`,
src/gbm/gbtree_model.h `GBTreeModel::Save`, include/xgboost/tree_model.h `RegTree::Save`):

    [b"binf"]                                   optional 4-byte header
    LearnerModelParamLegacy  136 bytes          float base_score; u32 num_feature; i32 num_class;
                                                i32 contain_extra_attrs; i32 contain_eval_metrics;
                                                u32 major, minor; i32 reserved[27]
    string objective, string booster            dmlc strings: u64 length + bytes
    GBTreeModelParam         160 bytes          i32 num_trees; i32 x3 (deprecated / pad); i64 deprecated;
                                                i32 deprecated; i32 size_leaf_vector; i32 reserved[32]
    num_trees x RegTree:
        TreeParam            148 bytes          i32 num_roots(deprecated); i32 num_nodes; i32 num_deleted;
                                                i32 max_depth(deprecated); i32 num_feature;
                                                i32 size_leaf_vector; i32 reserved[31]
        num_nodes x Node      20 bytes          i32 parent (bit 31: is left child); i32 cleft; i32 cright;
                                                u32 sindex (bit 31: default left); float leaf value | split cond
        num_nodes x NodeStat  16 bytes          float loss_chg, sum_hess, base_weight; i32 leaf_child_cnt
        [u64 n + n floats]                      leaf vector, only when size_leaf_vector != 0 (pre-1.0)
    num_trees x i32 tree_info                   class (output group) of every tree
    ... attributes / metric names (not needed here)

The writers in this module (`write_legacy_binary`, `wrap_serialized`, `ubjson_dumps`) exist so that the parsers
are round-trip tested offline (tests/test_pickle_compat_cpu.py); no xgboost is available in this environment, so
agreement with a buffer written by xgboost itself is not verified here."""
from __future__ import annotations

import json
import struct

import numpy as np

from .gbt import GBTForest

SER_HEADER = b"CONFIG-offset:"
_NODE = np.dtype([("parent", "<i4"), ("cleft", "<i4"), ("cright", "<i4"), ("sindex", "<u4"), ("value", "<f4")])
_STAT = np.dtype([("loss_chg", "<f4"), ("sum_hess", "<f4"), ("base_weight", "<f4"), ("leaf_child_cnt", "<i4")])
LEARNER_PARAM_BYTES, GBTREE_PARAM_BYTES, TREE_PARAM_BYTES = 136, 160, 148


# ------------------------------------------------------------------------------------------ UBJSON
def ubjson_loads(buf):
    """Universal Binary JSON (draft 12, big endian) -> Python objects; handles the optimised containers
    (`$` type / `#` count) xgboost >= 2.0 writes for its typed arrays."""
    mv = memoryview(bytes(buf))
    pos = 0
    _FIX = {"i": (">b", 1), "U": (">B", 1), "I": (">h", 2), "l": (">i", 4), "L": (">q", 8), "d": (">f", 4), "D": (">d", 8)}
    _NP = {"i": ">i1", "U": ">u1", "I": ">i2", "l": ">i4", "L": ">i8", "d": ">f4", "D": ">f8"}

    def marker():
        nonlocal pos
        m = chr(mv[pos])
        pos += 1
        return m

    def fixed(m):
        nonlocal pos
        fmt, n = _FIX[m]
        v = struct.unpack_from(fmt, mv, pos)[0]
        pos += n
        return v

    def length():
        m = marker()
        if m not in "iUIlL":
            raise ValueError("UBJSON: bad length marker %r at %d" % (m, pos - 1))
        return fixed(m)

    def string():
        nonlocal pos
        n = length()
        s = bytes(mv[pos:pos + n]).decode("utf-8")
        pos += n
        return s

    def value(m=None):
        nonlocal pos
        m = m or marker()
        while m == "N":
            m = marker()
        if m == "Z":
            return None
        if m == "T":
            return True
        if m == "F":
            return False
        if m in _FIX:
            return fixed(m)
        if m == "C":
            c = chr(mv[pos])
            pos += 1
            return c
        if m in "SH":
            return string()
        if m == "[":
            typ = cnt = None
            if chr(mv[pos]) == "$":
                pos += 1
                typ = marker()
            if chr(mv[pos]) == "#":
                pos += 1
                cnt = length()
            if typ is not None:
                if cnt is None:
                    raise ValueError("UBJSON: typed array without a count")
                if typ in _NP:
                    dt = np.dtype(_NP[typ])
                    arr = np.frombuffer(mv, dtype=dt, count=cnt, offset=pos)
                    pos += cnt * dt.itemsize
                    return arr.astype(dt.newbyteorder("=")).tolist()
                return [value(typ) for _ in range(cnt)]
            if cnt is not None:
                return [value() for _ in range(cnt)]
            out = []
            while chr(mv[pos]) != "]":
                out.append(value())
            pos += 1
            return out
        if m == "{":
            typ = cnt = None
            if chr(mv[pos]) == "$":
                pos += 1
                typ = marker()
            if chr(mv[pos]) == "#":
                pos += 1
                cnt = length()
            out = {}
            if cnt is not None:
                for _ in range(cnt):
                    k = string()
                    out[k] = value(typ)
                return out
            while chr(mv[pos]) != "}":
                k = string()
                out[k] = value(typ)
            pos += 1
            return out
        raise ValueError("UBJSON: unknown marker %r at %d" % (m, pos - 1))

    return value()


def ubjson_dumps(obj, typed_arrays=True) -> bytes:
    """Python objects -> UBJSON (for the round-trip tests): ints as the smallest type, floats as float32 when
    exact else float64, homogeneous numeric lists as optimised typed arrays when `typed_arrays`."""
    out = bytearray()

    def w_int(v, with_marker=True):
        for m, fmt, lo, hi in (("i", ">b", -128, 127), ("U", ">B", 0, 255), ("I", ">h", -2 ** 15, 2 ** 15 - 1),
                               ("l", ">i", -2 ** 31, 2 ** 31 - 1), ("L", ">q", -2 ** 63, 2 ** 63 - 1)):
            if lo <= v <= hi:
                if with_marker:
                    out.extend(m.encode())
                out.extend(struct.pack(fmt, v))
                return m
        raise ValueError("integer out of range")

    def w_str(s):
        b = s.encode("utf-8")
        w_int(len(b))
        out.extend(b)

    def w(v):
        if v is None:
            out.extend(b"Z")
        elif v is True:
            out.extend(b"T")
        elif v is False:
            out.extend(b"F")
        elif isinstance(v, (int, np.integer)):
            w_int(int(v))
        elif isinstance(v, (float, np.floating)):
            f = float(v)
            if float(np.float32(f)) == f or f != f:
                out.extend(b"d" + struct.pack(">f", f))
            else:
                out.extend(b"D" + struct.pack(">d", f))
        elif isinstance(v, str):
            out.extend(b"S")
            w_str(v)
        elif isinstance(v, (list, tuple, np.ndarray)):
            seq = list(v)
            if typed_arrays and seq and all(isinstance(x, (float, np.floating)) for x in seq) and \
                    all(float(np.float32(x)) == float(x) for x in seq):
                out.extend(b"[$d#")
                w_int(len(seq))
                out.extend(np.asarray(seq, dtype=">f4").tobytes())
            elif typed_arrays and seq and all(isinstance(x, (int, np.integer)) and not isinstance(x, bool) for x in seq) and \
                    all(-2 ** 31 <= int(x) < 2 ** 31 for x in seq):
                out.extend(b"[$l#")
                w_int(len(seq))
                out.extend(np.asarray(seq, dtype=">i4").tobytes())
            else:
                out.extend(b"[")
                for x in seq:
                    w(x)
                out.extend(b"]")
        elif isinstance(v, dict):
            out.extend(b"{")
            for k, x in v.items():
                w_str(str(k))
                w(x)
            out.extend(b"}")
        else:
            raise TypeError("cannot encode %r" % type(v))

    w(obj)
    return bytes(out)


# ------------------------------------------------------------------------------------------ legacy binary
def _dmlc_string(buf, pos):
    (n,) = struct.unpack_from("<Q", buf, pos)
    pos += 8
    if n > len(buf) - pos:
        raise ValueError("xgboost binary model: string length %d runs past the buffer" % n)
    return bytes(buf[pos:pos + n]).decode("utf-8", "replace"), pos + n


def parse_legacy_binary(buf, num_class=None, n_features=None) -> GBTForest:
    """Legacy binary model (layout in the module docstring) -> GBTForest."""
    try:
        return _parse_legacy_binary(buf, num_class, n_features)
    except struct.error as e:
        raise ValueError("xgboost binary model: truncated buffer (%s)" % e)


def _parse_legacy_binary(buf, num_class, n_features):
    buf = memoryview(bytes(buf))
    pos = 4 if bytes(buf[:4]) == b"binf" else 0
    if len(buf) - pos < LEARNER_PARAM_BYTES:
        raise ValueError("xgboost binary model: buffer too short")
    base_score, num_feature, n_class, _extra, _metrics, major, minor = struct.unpack_from("<fIiiiII", buf, pos)
    pos += LEARNER_PARAM_BYTES
    objective, pos = _dmlc_string(buf, pos)
    booster, pos = _dmlc_string(buf, pos)
    if booster not in ("gbtree",):
        raise ValueError("xgboost binary model: booster %r is not supported (the reference trains gbtree)" % booster)
    A = int(num_class or n_class)
    if A < 2:
        raise ValueError("xgboost binary model: num_class=%d; the smoother is a multi:softprob forest (objective %r)" % (n_class, objective))
    num_trees = struct.unpack_from("<i", buf, pos)[0]
    gb_size_leaf_vector = struct.unpack_from("<i", buf, pos + 28)[0]
    pos += GBTREE_PARAM_BYTES
    if num_trees < 0 or num_trees > 10_000_000:
        raise ValueError("xgboost binary model: implausible tree count %d" % num_trees)
    feat, thr, left, right, dl, leaf, offs = [], [], [], [], [], [], [0]
    for t in range(num_trees):
        _roots, num_nodes, _deleted, _depth, _nf, size_leaf_vector = struct.unpack_from("<6i", buf, pos)
        pos += TREE_PARAM_BYTES
        if num_nodes <= 0 or pos + num_nodes * (_NODE.itemsize + _STAT.itemsize) > len(buf):
            raise ValueError("xgboost binary model: tree %d has %d nodes past the end of the buffer" % (t, num_nodes))
        nodes = np.frombuffer(buf, dtype=_NODE, count=num_nodes, offset=pos)
        pos += num_nodes * (_NODE.itemsize + _STAT.itemsize)
        if size_leaf_vector != 0 or gb_size_leaf_vector != 0:
            (n,) = struct.unpack_from("<Q", buf, pos)
            pos += 8 + 4 * n
        is_leaf = nodes["cleft"] == -1
        feat.append(np.where(is_leaf, -1, (nodes["sindex"] & 0x7FFFFFFF).astype(np.int64)).astype(np.int32))
        thr.append(np.where(is_leaf, np.float32(0), nodes["value"]).astype(np.float32))
        left.append(np.where(is_leaf, 0, nodes["cleft"]).astype(np.int32))
        right.append(np.where(is_leaf, 0, nodes["cright"]).astype(np.int32))
        dl.append(np.where(is_leaf, 0, nodes["sindex"] >> 31).astype(np.uint8))
        leaf.append(np.where(is_leaf, nodes["value"], np.float32(0)).astype(np.float32))
        offs.append(offs[-1] + num_nodes)
    info = np.frombuffer(buf, dtype="<i4", count=num_trees, offset=pos) if num_trees else np.zeros(0, np.int32)
    _check_tree_info(info, A)
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    F = int(n_features or num_feature)
    return GBTForest(A, F, cat(feat, np.int32), cat(thr, np.float32), cat(left, np.int32), cat(right, np.int32), cat(dl, np.uint8),
                     cat(leaf, np.float32), offs, np.full(A, base_score, dtype=np.float32))


def _check_tree_info(info, A):
    info = np.asarray(info, dtype=np.int64)
    if len(info) % A != 0 or not np.array_equal(info, np.arange(len(info)) % A):
        raise ValueError("tree t must belong to class t % A (xgboost multi:softprob layout); got tree_info " + str(info[:2 * A].tolist()))


def write_legacy_binary(forest: GBTForest, binf=True, version=(1, 1), attributes=()) -> bytes:
    """GBTForest -> legacy binary model, the byte layout `parse_legacy_binary` documents (fixture writer)."""
    out = bytearray()
    if binf:
        out += b"binf"
    A, F, T = forest.A, forest.n_features, forest.n_trees
    out += struct.pack("<fIiiiII", float(forest.base_margin[0]), F, A, 1 if attributes else 0, 0, version[0], version[1])
    out += b"\0" * (4 * 27)
    for s in ("multi:softprob", "gbtree"):
        out += struct.pack("<Q", len(s)) + s.encode()
    out += struct.pack("<iiiiqii", T, 1, F, 0, 0, A, 0) + b"\0" * (4 * 32)
    for t in range(T):
        o, e = int(forest.tree_offsets[t]), int(forest.tree_offsets[t + 1])
        n = e - o
        out += struct.pack("<6i", 1, n, 0, 0, F, 0) + b"\0" * (4 * 31)
        nodes = np.zeros(n, dtype=_NODE)
        is_leaf = forest.feat[o:e] < 0
        nodes["parent"] = -1
        for i in range(n):
            if not is_leaf[i]:
                nodes["parent"][forest.left[o + i]] = np.int32(np.uint32(i) | np.uint32(0x80000000))
                nodes["parent"][forest.right[o + i]] = i
        nodes["cleft"] = np.where(is_leaf, -1, forest.left[o:e])
        nodes["cright"] = np.where(is_leaf, -1, forest.right[o:e])
        nodes["sindex"] = np.where(is_leaf, 0, forest.feat[o:e].astype(np.int64) | (forest.default_left[o:e].astype(np.int64) << 31)).astype(np.uint32)
        nodes["value"] = np.where(is_leaf, forest.leaf[o:e], forest.thr[o:e])
        stats = np.zeros(n, dtype=_STAT)
        out += nodes.tobytes() + stats.tobytes()
    out += (np.arange(T, dtype="<i4") % A).astype("<i4").tobytes()
    if attributes:
        out += struct.pack("<Q", len(attributes))
        for k, v in attributes:
            out += struct.pack("<Q", len(k)) + k.encode() + struct.pack("<Q", len(v)) + v.encode()
    return bytes(out)


def wrap_serialized(binary_model: bytes, config=None) -> bytes:
    """`LearnerIO::Save` envelope of xgboost 1.0 - 1.5 around a legacy binary model."""
    cfg = json.dumps(config if config is not None else {"learner": {"objective": {"name": "multi:softprob"}}, "version": [1, 1, 1]})
    return SER_HEADER + struct.pack("<q", len(binary_model)) + binary_model + cfg.encode()


# ------------------------------------------------------------------------------------------ JSON / UBJSON model
def _num(v, default=None):
    """xgboost stores scalars as strings ("5E-1"), newer versions as one-element vectors ("[5E-1]" / [0.5])."""
    if v is None:
        return default
    if isinstance(v, (list, tuple)):
        return _num(v[0], default) if len(v) else default
    if isinstance(v, str):
        s = v.strip().strip("[]").split(",")[0]
        return float(s) if s else default
    return float(v)


def forest_from_model_dict(obj, num_class=None, n_features=None) -> GBTForest:
    """JSON-schema model (`{"learner": ...}`, or the `{"Model": ..., "Config": ...}` snapshot a pickled booster of
    xgboost >= 1.6 holds) -> GBTForest.  doc/model.schema: learner.gradient_booster.model.trees[t] has parallel arrays
    left_children / right_children (-1 = leaf), split_indices, split_conditions (threshold, or the leaf value at a
    leaf), default_left; tree_info[t] = class of tree t; learner_model_param has base_score / num_class / num_feature."""
    if "Model" in obj and "learner" not in obj:
        obj = obj["Model"]
    learner = obj["learner"]
    lmp = learner["learner_model_param"]
    A = int(num_class or max(int(_num(lmp.get("num_class"), 0)), 1))
    F = int(n_features or _num(lmp.get("num_feature"), 0))
    base_score = np.float32(_num(lmp.get("base_score"), 0.5))
    gb = learner["gradient_booster"]
    if gb.get("name", "gbtree") not in ("gbtree",):
        raise ValueError("booster %r is not supported (the reference trains gbtree)" % gb.get("name"))
    model = gb["model"] if "model" in gb else gb["gbtree"]["model"]
    trees, info = model["trees"], [int(v) for v in model["tree_info"]]
    if A < 2:
        raise ValueError("binary:logistic boosters are not a multi:softprob forest (the reference trains with num_class=A)")
    _check_tree_info(info, A)
    feat, thr, left, right, dl, leaf, offs = [], [], [], [], [], [], [0]
    for tr in trees:
        if any(int(x) != 0 for x in tr.get("split_type", ())) or len(tr.get("categories", ())):
            raise ValueError("categorical splits are not supported")
        lc = np.asarray(tr["left_children"], dtype=np.int32)
        rc = np.asarray(tr["right_children"], dtype=np.int32)
        cond = np.asarray(tr["split_conditions"], dtype=np.float32)
        idx = np.asarray(tr["split_indices"], dtype=np.int32)
        dfl = np.asarray([1 if v else 0 for v in tr["default_left"]], dtype=np.uint8)
        is_leaf = lc == -1
        feat.append(np.where(is_leaf, -1, idx).astype(np.int32))
        thr.append(np.where(is_leaf, np.float32(0), cond).astype(np.float32))
        left.append(np.where(is_leaf, 0, lc).astype(np.int32))
        right.append(np.where(is_leaf, 0, rc).astype(np.int32))
        dl.append(np.where(is_leaf, 0, dfl).astype(np.uint8))
        leaf.append(np.where(is_leaf, cond, np.float32(0)).astype(np.float32))
        offs.append(offs[-1] + len(lc))
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return GBTForest(A, F, cat(feat, np.int32), cat(thr, np.float32), cat(left, np.int32), cat(right, np.int32), cat(dl, np.uint8),
                     cat(leaf, np.float32), offs, np.full(A, base_score, dtype=np.float32))


# ------------------------------------------------------------------------------------------ dispatch
def forest_from_booster_bytes(buf, num_class=None, n_features=None) -> GBTForest:
    """Any of the booster buffers listed in the module docstring -> GBTForest."""
    b = bytes(buf)
    if b.startswith(SER_HEADER):
        (off,) = struct.unpack_from("<q", b, len(SER_HEADER))
        start = len(SER_HEADER) + 8
        if off < 0 or start + off > len(b):
            raise ValueError("xgboost serialised booster: bad CONFIG offset %d" % off)
        return parse_legacy_binary(b[start:start + off], num_class, n_features)
    if b.startswith(b"bs64"):
        raise ValueError("base64 xgboost models (pre-0.7) are not supported")
    if b[:1] == b"{":
        head = b[1:2]
        # '{' followed by '"' or whitespace is JSON text; UBJSON continues with a length / type marker
        if head in (b'"', b" ", b"\n", b"\r", b"\t", b"}"):
            return forest_from_model_dict(json.loads(b.decode("utf-8")), num_class, n_features)
        return forest_from_model_dict(ubjson_loads(b), num_class, n_features)
    return parse_legacy_binary(b, num_class, n_features)
