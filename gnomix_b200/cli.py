"""Inference driver with the argv contract and outputs of the reference's gnomix.py in
pre-trained mode (gnomix.py:37-100, 318-370):

    python -m gnomix_b200.cli <query_file> <output_basename> <chr_nr> <phase> <path_to_model>

writes <output_basename>/query_results.msp and .fb (and query_file_phased.vcf when phase is
True).  Base, Smoother and Gnofix run on the GPU; reading and writing are gnomix_b200.io /
gnomix_b200.postprocess."""
from __future__ import annotations

import gzip
import os
import pickle
import sys

import numpy as np

from . import io as gio
from . import postprocess as pp


def load_model(path_to_model, verbose=True):
    """gnomix.py:26-35 (plain or gzip pickle of the whole Gnomix object) -- reference-written pickles included
    (scikit-learn / xgboost / sklearn-crfsuite objects are converted, not imported: gnomix_b200/pickle_compat.py)."""
    from .pickle_compat import load_model as _load
    return _load(path_to_model, verbose=verbose)


def read_headers(vcf_file):
    """src/utils.py:332-348."""
    header = ""
    opener = gzip.open if vcf_file.endswith(".gz") else open
    with opener(vcf_file, "rb") as f:
        for line in f:
            if not line.startswith(b"#"):
                break
            if line.startswith(b"##"):
                header += line.decode("utf-8")
    return header


def update_vcf(vcf_data, mask=None, Updates=None):
    """src/utils.py:227-241."""
    out = dict(vcf_data)
    if mask is not None:
        for key in vcf_data:
            if key != "samples":
                out[key] = vcf_data[key][mask]
    if Updates is not None:
        for key in Updates:
            if key != "samples":
                out[key] = Updates[key]
    return out


def _vcf_preamble(f, headers, cols):
    f.write(headers.encode())
    f.write(b"##fileformat=VCFv4.1\n##source=gnomix.py\n")
    f.write(b'##FORMAT=<ID=GT,Number=1,Type=String,Description="Phased Genotype">\n')
    f.write(("#" + "\t".join(cols) + "\n").encode())


def _qual_str(q):
    return "" if (isinstance(q, float) or isinstance(q, np.floating)) and np.isnan(q) else str(q)


def npy_to_vcf(data, npy, results_file, headers=""):
    """The light VCF writer of src/utils.py:247-329: metadata from `data`, genotypes `maternal|paternal`
    from the int matrix [2n, C].  The records are written by the library's host threads
    (gnx_write_vcf_body); npy_to_vcf_py is the per-record Python loop it replaces."""
    import ctypes as C
    from . import _lib
    if results_file.split(".")[-1] not in [".vcf", ".bcf"]:
        results_file += ".vcf"
    hap = np.ascontiguousarray(np.asarray(npy).astype(int).astype(np.int8))
    chmlen = data["calldata/GT"].shape[0]
    h, c = hap.shape
    n = h // 2
    assert chmlen == c, "reference (" + str(chmlen) + ") and numpy matrix (" + str(c) + ") not compatible"
    samples = list(data["samples"]) if "samples" in data and len(data["samples"]) == n else ["sample%d" % i for i in range(n)]
    cols = ["CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + [str(s) for s in samples]
    with open(results_file, "wb") as f:
        _vcf_preamble(f, headers, cols)

    def blob(values):
        b = ("\n".join(values) + "\n").encode() if c else b""
        return b, len(b)

    pos = np.ascontiguousarray(np.asarray(data["variants/POS"]), dtype=np.int64)
    chrom, chrom_l = blob([str(v) for v in data["variants/CHROM"]])
    ident, ident_l = blob([str(v) for v in data["variants/ID"]])
    ref, ref_l = blob([str(v) for v in data["variants/REF"]])
    alt, alt_l = blob([str(v[0]) for v in data["variants/ALT"]])
    qual, qual_l = blob([_qual_str(q) for q in data["variants/QUAL"]])
    _lib.check(_lib.lib().gnx_write_vcf_body(results_file.encode(), 1, c, 2 * n, hap.ctypes.data, c, pos.ctypes.data, chrom, chrom_l,
                                             ident, ident_l, ref, ref_l, alt, alt_l, qual, qual_l, 0), "gnx_write_vcf_body")
    return results_file


def npy_to_vcf_py(data, npy, results_file, headers=""):
    """src/utils.py:247-329 with a Python loop over the records (cross-check of the native writer)."""
    if results_file.split(".")[-1] not in [".vcf", ".bcf"]:
        results_file += ".vcf"
    npy = np.asarray(npy).astype(int)
    chmlen = data["calldata/GT"].shape[0]
    h, c = npy.shape
    n = h // 2
    assert chmlen == c, "reference (" + str(chmlen) + ") and numpy matrix (" + str(c) + ") not compatible"
    samples = list(data["samples"]) if "samples" in data and len(data["samples"]) == n else ["sample%d" % i for i in range(n)]
    cols = ["CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + [str(s) for s in samples]
    digits = (npy.T + ord("0")).astype(np.uint8)                      # [C, 2n]
    row = np.empty((c, 4 * n), dtype=np.uint8)
    row[:, 0::4] = ord("\t")
    row[:, 1::4] = digits[:, 0::2]
    row[:, 2::4] = ord("|")
    row[:, 3::4] = digits[:, 1::2]
    with open(results_file, "wb") as f:
        _vcf_preamble(f, headers, cols)
        for i in range(c):
            f.write("\t".join([str(data["variants/CHROM"][i]), str(data["variants/POS"][i]), str(data["variants/ID"][i]),
                               str(data["variants/REF"][i]), str(data["variants/ALT"][i][0]), _qual_str(data["variants/QUAL"][i]),
                               "PASS", ".", "GT"]).encode())
            f.write(row[i].tobytes())
            f.write(b"\n")
    return results_file


def run_inference(base_args, model, snp_level=False, bed_file_output=False, verbose=False):
    """gnomix.py:37-100 without the plotting branch."""
    query_file, chm, output_path = base_args["query_file"], base_args["chm"], base_args["output_basename"]
    os.makedirs(output_path, exist_ok=True)
    gen_map_df = model.gen_map_df
    if verbose:
        print("Loading and processing query file...")
    vcf = gio.read_vcf(query_file, chm=chm, fields="*")
    # gnomix.py:48-72 as ONE host-buffer pipeline: the genotype calls are gathered straight into 2-bit planes in
    # pinned memory (a quarter of the int8 matrix's bytes cross PCIe, nothing is packed on the way) and
    # Gnomix.predict_host streams them through Base -> [Gnofix] -> Smoother -> [Calibrator] with overlapped copies.
    # Same labels / probabilities as the staged calls (tests/test_pipeline_gpu.py).
    X_query, vcf_idx, fmt_idx = gio.vcf_to_packed(vcf, model.snp_pos, model.snp_ref, return_idx=True, verbose=verbose)
    if verbose:
        print("Inferring ancestry on query data...")
    if not base_args["phase"]:
        y_pred, y_proba = model.predict_host(X_query, want_proba=True)
    else:
        y_pred, y_proba, X_phased = model.predict_host(X_query, want_proba=True, phase=True, want_phased=True)
        U = {"variants/REF": np.asarray(model.snp_ref)[fmt_idx],
             "variants/ALT": np.asarray(model.snp_alt)[fmt_idx].reshape(len(fmt_idx), 1)}
        vcf_phase = update_vcf(vcf, mask=vcf_idx, Updates=U)
        npy_to_vcf(vcf_phase, X_phased[:, fmt_idx], output_path + "/" + "query_file_phased", headers=read_headers(query_file))
    if verbose:
        print("Saving results...")
    meta = pp.get_meta_data(chm, model.snp_pos, vcf["variants/POS"], model.W, model.M, gen_map_df)
    out_prefix = output_path + "/" + "query_results"
    pp.write_msp(out_prefix, meta, y_pred, model.population_order, vcf["samples"])
    pp.write_fb(out_prefix, meta, y_proba, model.population_order, vcf["samples"])
    if snp_level:                                   # gnomix.py:83-84 (BETA)
        pp.msp_to_lai(msp_file=out_prefix + ".msp", positions=vcf["variants/POS"], lai_file=out_prefix + ".lai")
    if bed_file_output:                             # gnomix.py:86-90
        bed_root = output_path + "/" + "query_results_bed"
        os.makedirs(bed_root, exist_ok=True)
        pp.msp_to_bed(msp_file=out_prefix + ".msp", root=bed_root, pop_order=model.population_order)
    return out_prefix


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) != 6:
        print("usage: python -m gnomix_b200.cli <query_file> <output_basename> <chr_nr> <phase> <path_to_model>")
        return 2
    base_args = {"query_file": argv[1], "output_basename": argv[2], "chm": argv[3], "phase": argv[4] in ("True", "true", "1")}
    model = load_model(argv[5])
    model.base.vectorize = True
    run_inference(base_args, model, verbose=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
