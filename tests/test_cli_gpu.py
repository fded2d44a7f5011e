"""End to end through the driver (BASELINE config 1 in miniature): VCF in -> trained model ->
.msp/.fb (+ phased VCF) out, GPU labels equal to the oracle's."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _write_vcf(path, pos, ref, alt, X, samples):
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n##contig=<ID=22>\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n")
        for j in range(len(pos)):
            gts = "\t".join("%d|%d" % (X[2 * i, j], X[2 * i + 1, j]) for i in range(len(samples)))
            f.write("22\t%d\trs%d\t%s\t%s\t.\tPASS\t.\tGT\t%s\n" % (pos[j], j, ref[j], alt[j], gts))


def test_train_pickle_infer_roundtrip(tmp_path):
    import warnings
    import pandas as pd
    from gnomix_b200 import Gnomix, cli, synth
    from oracle import np_oracle as npo, c_oracle as co
    rng = np.random.default_rng(12)
    C, M, A, S = 6007, 300, 3, 9
    pos = np.sort(rng.choice(np.arange(16_000_000, 20_000_000), C, replace=False))
    ref = rng.choice(np.array(["A", "C", "G", "T"]), C)
    alt = np.where(ref == "A", "G", "A")
    freqs = synth.population_frequencies(rng, C, A, fst=0.25)
    fx, fpop = synth.founders(rng, freqs, per_pop=30)
    W = C // M
    model = Gnomix(C, M, A, S, snp_pos=pos, snp_ref=ref, snp_alt=alt, population_order=["P0", "P1", "P2"], path=str(tmp_path) + "/")
    model.write_gen_map_df(pd.DataFrame({"chm": ["22"] * 50, "pos": np.linspace(15_900_000, 20_100_000, 50).astype(int),
                                         "pos_cm": np.linspace(0.0, 6.0, 50)}))
    y = np.repeat(fpop[:, None], W, axis=1)
    half = len(fx) // 2
    idx = rng.permutation(len(fx))
    t1, t2 = idx[:half], idx[half:]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.train(((fx[t1], y[t1]), (fx[t2], y[t2]), (None, None)), retrain_base=False, evaluate=False, verbose=False)
    assert os.path.exists(tmp_path / "model.pkl")
    # query: 5 individuals, mosaics of founders
    Xq, _ = synth.admix_host(rng, fx, fpop, 10, morgans=0.3, missing=0.0)
    samples = ["Q%d" % i for i in range(5)]
    _write_vcf(str(tmp_path / "q.vcf"), pos, ref, alt, Xq, samples)
    rc = cli.main(["cli", str(tmp_path / "q.vcf"), str(tmp_path / "out"), "22", "False", str(tmp_path / "model.pkl")])
    assert rc == 0
    msp = open(tmp_path / "out" / "query_results.msp").read().splitlines()
    assert msp[0] == "#Subpopulation order/codes: P0=0\tP1=1\tP2=2" and len(msp) == 2 + W
    labels = np.array([ln.split("\t")[6:] for ln in msp[2:]], dtype=int).T        # [N, W]
    # oracle on the same pickled model
    m2 = pickle.load(open(tmp_path / "model.pkl", "rb"))
    coefs, icpts = [], []
    from gnomix_b200.base import lr_weights_of
    for mdl in m2.base.models:
        c, b = lr_weights_of(mdl, A)
        coefs.append(c); icpts.append(b)
    ctx = m2.context
    s = npo.lr_choose_scale(coefs, C, M, ctx, 7)
    B_o = co.lr_fixed(Xq, npo.lr_quantize_fold(coefs, C, M, ctx, s), np.stack(icpts), C, M, ctx, A, s)
    p_o, l_o = co.gbt_smooth(m2.smooth.model, B_o, m2.smooth.S)
    assert np.array_equal(labels, l_o)
    fb = open(tmp_path / "out" / "query_results.fb").read().splitlines()
    assert len(fb) == 2 + W
    first = np.array(fb[2].split("\t")[4:], dtype=np.float32).reshape(10, A)
    assert np.array_equal(first, p_o[:, 0, :])
    # the smoother learnt something: labels mostly equal the true single-ancestry founders' populations
    acc = (m2.predict(fx[t2]) == y[t2]).mean()
    assert acc > 0.8, acc
    # phase=True path writes the phased VCF and stays consistent
    rc = cli.main(["cli", str(tmp_path / "q.vcf"), str(tmp_path / "out2"), "22", "True", str(tmp_path / "model.pkl")])
    assert rc == 0 and os.path.exists(tmp_path / "out2" / "query_file_phased.vcf")
    from gnomix_b200 import io as gio
    vp = gio.read_vcf(str(tmp_path / "out2" / "query_file_phased.vcf"), chm="22")
    Xp = gio.vcf_to_npy(vp, verbose=False)
    assert Xp.shape == Xq.shape
    # phasing only exchanges alleles between the two haplotypes of an individual
    assert np.array_equal(np.sort(np.stack([Xp[0::2], Xp[1::2]]), axis=0), np.sort(np.stack([Xq[0::2], Xq[1::2]]), axis=0))


def test_gnomix_predict_numpy_equals_two_step_handoff():
    """Gnomix.predict / predict_proba on a numpy matrix (src/model.py:169-179) keep B in HBM between the stages;
    the results equal the reference's two numpy hand-offs (base.predict_proba -> smooth.predict_proba) bit for bit,
    for the tree smoother (float32 B), the CRF smoother (float64 B) and with a calibrator."""
    import numpy as np
    from gnomix_b200 import Gnomix, GBTForest
    from gnomix_b200.smooth import CRF_Smoother, CRFModel
    from tests import util
    rng = np.random.default_rng(17)
    C, M, A, S, N = 30_011, 500, 7, 15, 333
    model = Gnomix(C, M, A, S)
    coefs, icpts, _ = util.random_lr(rng, C, M, A)
    model.base.set_window_weights(coefs, icpts)
    model.smooth.model = GBTForest.random(rng, A, model.smooth.S, n_rounds=20, depth=4)
    X = util.random_haplotypes(rng, N, C)
    B = model.base.predict_proba(X)
    assert B.dtype == np.float64
    p2, y2 = model.smooth.predict_proba(B), model.smooth.predict(B)
    p1, y1 = model.predict_proba(X), model.predict(X)
    assert p1.dtype == p2.dtype == np.float32 and np.array_equal(p1, p2)
    assert y1.dtype == np.int64 and np.array_equal(y1, y2)
    # calibrated
    np.random.seed(1)
    model.smooth.calibrate = True
    model.smooth.train_calibrator(B[:200], np.argmax(B[:200], axis=-1), frac=0.5)
    assert np.array_equal(model.predict_proba(X), model.smooth.predict_proba(B)) and np.array_equal(model.predict(X), model.smooth.predict(B))
    # CRF smoother: float64 hand-off
    crf = CRF_Smoother(n_windows=C // M, num_ancestry=A, smooth_window_size=S)
    crf.model = CRFModel(rng.normal(0, 1.5, (A, A)), rng.normal(0, 1.0, (A, A)))
    model.smooth = crf
    p2, y2 = crf.predict_proba(B), crf.predict(B)
    p1, y1 = model.predict_proba(X), model.predict(X)
    assert p1.dtype == np.float64 and np.array_equal(p1, p2) and np.array_equal(y1, y2)
