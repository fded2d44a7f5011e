"""CPU tests of the rows next to the hot path: window table, .msp/.fb writers,
VCF -> int8 block -- against files and arrays produced by the reference's own functions."""
import gzip
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")


def _gen_map(d):
    import pandas as pd
    return pd.DataFrame({"chm": ["22"] * len(d["gm_pos"]), "pos": d["gm_pos"], "pos_cm": d["gm_cm"]})


def test_meta_table_and_writers_byte_identical(tmp_path):
    from gnomix_b200 import postprocess as pp
    d = np.load(os.path.join(G, "meta.npz"))
    meta = pp.get_meta_data("22", d["pos"], d["qpos"], int(d["W"]), int(d["M"]), _gen_map(d))
    assert list(meta.columns) == d["columns"].tolist()
    assert np.array_equal(np.asarray(meta.values).astype(str), d["table"])
    pops, samples = d["pops"].tolist(), d["samples"].tolist()
    pp.write_msp(str(tmp_path / "q"), meta, d["labels"], pops, samples)
    pp.write_fb(str(tmp_path / "q"), meta, d["proba"], pops, samples)
    assert open(tmp_path / "q.msp", "rb").read() == d["msp"].tobytes()
    assert open(tmp_path / "q.fb", "rb").read() == d["fb"].tobytes()


def test_meta_demo_known_answer():
    """demo chr22 .msp geometry -- 370 windows of 857 SNPs."""
    from gnomix_b200 import postprocess as pp
    import pandas as pd
    rng = np.random.default_rng(0)
    C, M = 317_408, 857
    pos = np.sort(rng.choice(np.arange(16_050_000, 51_240_000), C, replace=False))
    gm = pd.DataFrame({"chm": ["22"] * 1000, "pos": np.linspace(16e6, 51.3e6, 1000).astype(int), "pos_cm": np.linspace(0, 74.1, 1000)})
    meta = pp.get_meta_data("22", pos, pos, C // M, M, gm)
    assert meta.shape == (370, 6)
    n = np.asarray(meta["n snps"]).astype(float).astype(int)
    assert n[:-1].tolist() == [857] * 369 and n[-1] == C - 857 * 369 and n.sum() == C
    assert int(meta["spos"].iloc[0]) == pos[0] and int(meta["epos"].iloc[-1]) == pos[-1]


def test_vcf_to_npy_against_reference():
    from gnomix_b200 import io as gio
    d = np.load(os.path.join(G, "vcf_to_npy.npz"))
    vcf = {"calldata/GT": d["gt"].copy(), "variants/POS": d["vpos"], "variants/REF": d["vref"]}
    X = gio.vcf_to_npy(vcf, d["model_pos"], d["model_ref"], verbose=False)
    assert X.dtype == np.int8 and np.array_equal(X, d["X"])
    X0 = gio.vcf_to_npy({"calldata/GT": d["gt"].copy(), "variants/POS": d["vpos"], "variants/REF": d["vref"]}, verbose=False)
    assert np.array_equal(X0, d["X_noformat"])


VCF_TEXT = """##fileformat=VCFv4.2
##contig=<ID=22>
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3
22\t100\trs1\tA\tG\t.\tPASS\t.\tGT\t0|1\t1|1\t0|0
22\t250\trs2\tC\tT\t.\tPASS\t.\tGT\t1|0\t.|.\t0|1
21\t300\trs3\tG\tA\t.\tPASS\t.\tGT\t1|1\t1|1\t1|1
22\t900\trs4\tT\tC\t.\tPASS\t.\tGT:DS\t0|0:0.1\t1|0:1.0\t1/1:2.0
"""


@pytest.mark.parametrize("gz", [False, True])
def test_read_vcf_small(tmp_path, gz):
    from gnomix_b200 import io as gio
    path = str(tmp_path / ("q.vcf.gz" if gz else "q.vcf"))
    if gz:
        with gzip.open(path, "wt") as f:
            f.write(VCF_TEXT)
    else:
        open(path, "w").write(VCF_TEXT)
    v = gio.read_vcf(path, chm="22")
    assert v["samples"].tolist() == ["S1", "S2", "S3"]
    assert v["variants/POS"].tolist() == [100, 250, 900]
    assert v["variants/REF"].tolist() == ["A", "C", "T"] and v["variants/ALT"][:, 0].tolist() == ["G", "T", "C"]
    assert v["variants/CHROM"].tolist() == ["22"] * 3
    gt = v["calldata/GT"]
    assert gt.shape == (3, 3, 2) and gt.dtype == np.int8
    assert gt[0].tolist() == [[0, 1], [1, 1], [0, 0]]
    assert gt[1].tolist() == [[1, 0], [-1, -1], [0, 1]]
    assert gt[2].tolist() == [[0, 0], [1, 0], [1, 1]]
    # region that does not exist -> the reference falls back to the whole file
    assert len(gio.read_vcf(path, chm="7")["variants/POS"]) == 4
    X = gio.vcf_to_npy(v, np.array([100, 250, 500, 900]), np.array(["A", "T", "G", "T"]), verbose=False)
    assert X.tolist() == [[0, 0, 2, 0], [1, 1, 2, 0], [1, 2, 2, 1], [1, 2, 2, 0], [0, 1, 2, 1], [0, 0, 2, 1]]


def test_read_vcf_demo_if_present():
    from gnomix_b200 import io as gio
    path = "/root/reference/demo/data/small_query_chr22.vcf.gz"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    v = gio.read_vcf(path, chm="22")
    assert v["calldata/GT"].shape == (317_408, 9, 2)
    assert set(np.unique(v["calldata/GT"]).tolist()) <= {0, 1}
    assert v["variants/POS"][0] == 16_050_075 or v["variants/POS"][0] > 16_000_000


def test_genetic_map_reader(tmp_path):
    from gnomix_b200 import io as gio
    p = tmp_path / "m.gmap"
    p.write_text("chr22\t100\t0.0\nchr22\t200\t0.5\nchr1\t10\t0.1\n")
    df = gio.read_genetic_map(str(p), chm="22")
    assert df["pos"].tolist() == [100, 200] and df["pos_cm"].tolist() == [0.0, 0.5]


def test_write_fb_native_body_matches_numpy_formatting(tmp_path):
    """gnx_write_fb_body (csrc/host_io.cpp) against the reference's per-number numpy str formatting
    (src/postprocess.py:112-126) on awkward values, both dtypes, several thread counts."""
    import pandas as pd
    from gnomix_b200 import postprocess as pp, _lib
    rng = np.random.default_rng(5)
    N, W, A = 14, 9, 3
    meta = pd.DataFrame(np.array([["7"] * W, np.arange(W) * 1000 + 5, np.arange(W) * 1000 + 900, np.round(np.arange(W) * 0.2, 5),
                                  np.round(np.arange(W) * 0.2 + 0.19, 5), np.full(W, 857)]).T,
                        columns=["chm", "spos", "epos", "sgpos", "egpos", "n snps"])
    pops, samples = ["AFR", "EUR", "EAS"], ["S%d" % i for i in range(N // 2)]
    for dt in (np.float32, np.float64):
        proba = rng.dirichlet(np.full(A, 0.2), (N, W)).astype(dt)
        proba[0, 0] = [0.0, 1.0, 1e-4]
        proba[1, 0] = [1e-5, 9.9999e-5, 0.5]
        proba[2, 1] = [1e-30, 1.0 / 3.0, 2.0 / 3.0]
        proba[3, 2] = [np.nan, 1e6, 123456.7]
        for th in ("1", "3", ""):
            if th:
                os.environ["GNX_HOST_THREADS"] = th
            else:
                os.environ.pop("GNX_HOST_THREADS", None)
            pp.write_fb(str(tmp_path / "n"), meta, proba, pops, samples)
            lines = open(tmp_path / "n.fb").read().split("\n")
            assert len(lines) == W + 3 and lines[-1] == ""
            fb_prob = np.swapaxes(proba, 1, 2).reshape(-1, W).T          # the reference's [W, N*A] view
            for l in range(W):
                assert lines[2 + l].split("\t")[4:] == list(fb_prob[l].astype(str)), (dt, th, l)
    # the float formatter alone, wide range
    v = np.concatenate([10.0 ** rng.uniform(-40, 30, 5000) * rng.choice([-1, 1], 5000), rng.random(5000), [0.0, -0.0, 1e16, 1e-4, 999999.94, 1e6]])
    for dt in (np.float32, np.float64):
        a = np.ascontiguousarray(v.astype(dt))
        import ctypes as C
        out = C.create_string_buffer(a.size * 34 + 8)
        n = _lib.lib().gnx_format_floats(a.ctypes.data, int(dt == np.float64), a.size, out, len(out))
        assert out.raw[:n].decode().split("\n")[:-1] == list(a.astype(str))


def _same_vcf(a, b):
    assert set(a) == set(b)
    for k in b:
        if a[k].dtype == np.float32:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
        else:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k


def test_native_vcf_reader_matches_python_parser(tmp_path):
    """gnx_vcf_open (csrc/host_vcf.cpp) against the pure-Python restatement of allel.read_vcf's fields
    (src/utils.py:55-81) on awkward records: missing and haploid calls, unphased separators, extra FORMAT
    keys, multi-allelic ALT, multi-digit alleles, CRLF, a short line, two chromosomes, no final newline, gzip."""
    from gnomix_b200 import io as gio
    hdr = "##fileformat=VCFv4.2\n##contig=<ID=7>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\tS3\n"
    rows = [
        "7\t100\trs1\tA\tG\t.\tPASS\t.\tGT\t0|1\t1|1\t0|0",
        "7\t150\t.\tC\tT,G\t30.5\tq10\tDP=4\tGT:DP\t1/0:3\t.|.:0\t2|1:9",
        "7\t160\trs3\tG\tA,C,T,GG\t7\t.\t.\tGT\t0\t.\t1",
        "8\t170\trs4\tT\tC\t.\t.\t.\tGT\t1|1\t1|1\t1|1",
        "7\t180\tshort\tA\tC",
        "7\t190\trs5\tT\tC\t99\t.\t.\tGT:GQ\t10|0:1\t0|.\t./1:5",
        "7\t200\trs6\tAT\tA\t.\t.\t.\tGT\t0|0\t0|1",
    ]
    plain = tmp_path / "a.vcf"
    plain.write_text(hdr + "\n".join(rows))                       # no final newline
    crlf = tmp_path / "b.vcf"
    crlf.write_bytes((hdr + "\n".join(rows) + "\n").replace("\n", "\r\n").encode())
    gz = tmp_path / "c.vcf.gz"
    with gzip.open(gz, "wb") as f:
        f.write((hdr + "\n".join(rows) + "\n").encode())
    for path in (plain, gz):
        for chm in ("7", "8", None, "9"):
            _same_vcf(gio.read_vcf(str(path), chm), gio.read_vcf_py(str(path), chm))
    d = gio.read_vcf(str(plain), "7")
    assert d["calldata/GT"].shape == (5, 3, 2) and list(d["variants/POS"]) == [100, 150, 160, 190, 200]
    assert d["calldata/GT"][1].tolist() == [[1, 0], [-1, -1], [2, 1]] and d["calldata/GT"][2].tolist() == [[0, -1], [-1, -1], [1, -1]]
    assert d["calldata/GT"][3].tolist() == [[10, 0], [0, -1], [-1, 1]] and d["calldata/GT"][4].tolist() == [[0, 0], [0, 1], [-1, -1]]
    assert list(d["variants/ALT"][2]) == ["A", "C", "T"] and list(d["samples"]) == ["S1", "S2", "S3"]
    c = gio.read_vcf(str(crlf), "7")
    assert np.array_equal(c["calldata/GT"], d["calldata/GT"]) and list(c["samples"]) == ["S1", "S2", "S3"]
    # a larger random cohort, threads or not
    rng = np.random.default_rng(2)
    lut = np.array(["0|0", "0|1", "1|0", "1|1", ".|.", "0/1"])
    big = tmp_path / "big.vcf.gz"
    with gzip.open(big, "wt") as f:
        f.write(hdr.replace("S1\tS2\tS3", "\t".join("X%d" % i for i in range(40))))
        for r in range(6000):
            f.write("7\t%d\tr%d\tA\tG\t.\t.\t.\tGT\t%s\n" % (10 + 3 * r, r, "\t".join(lut[rng.integers(0, 6, 40)])))
    for th in ("1", "5"):
        os.environ["GNX_HOST_THREADS"] = th
        _same_vcf(gio.read_vcf(str(big), "7"), gio.read_vcf_py(str(big), "7"))
    os.environ.pop("GNX_HOST_THREADS", None)
    with pytest.raises(Exception):
        gio.read_vcf(str(tmp_path / "missing.vcf"))


def test_native_vcf_to_npy_matches_numpy_statement():
    """gnx_vcf_to_haplotypes against vcf_to_npy_py (src/utils.py:104-159 statement by statement): SNP intersection,
    reference-allele flips, missing / multi-allelic calls -> miss_fill, with and without a model format."""
    from gnomix_b200 import io as gio
    rng = np.random.default_rng(12)
    for R, S in [(1, 1), (300, 3), (2000, 70), (777, 33)]:
        gt = rng.integers(-1, 4, (R, S, 2), dtype=np.int8)
        pos = np.sort(rng.choice(np.arange(10, 50 * R + 100), R, replace=False)).astype(np.int32)
        ref = rng.choice(np.array(["A", "C", "G", "T"], dtype=object), R)
        d = {"calldata/GT": gt, "variants/POS": pos, "variants/REF": ref}
        keep = np.sort(rng.choice(R, max(1, R * 3 // 4), replace=False))
        extra = np.setdiff1d(rng.choice(np.arange(10, 50 * R + 100), max(1, R // 5)), pos)
        mp = np.sort(np.concatenate([pos[keep], extra]))
        mr = rng.choice(np.array(["A", "C", "G", "T"], dtype=object), len(mp))
        for fill in (2, 9):
            a = gio.vcf_to_npy(d, mp, mr, miss_fill=fill, verbose=False)
            b = gio.vcf_to_npy_py({k: (v.copy() if hasattr(v, "copy") else v) for k, v in d.items()}, mp, mr, miss_fill=fill, verbose=False)
            assert a.dtype == np.int8 and np.array_equal(a, b), (R, S, fill)
        a, vi, fi = gio.vcf_to_npy(d, mp, None, return_idx=True, verbose=False)
        b, vj, fj = gio.vcf_to_npy_py(d, mp, None, return_idx=True, verbose=False)
        assert np.array_equal(a, b) and np.array_equal(vi, vj) and np.array_equal(fi, fj)
        assert np.array_equal(gio.vcf_to_npy(d, verbose=False), gio.vcf_to_npy_py(d, verbose=False))


def test_lai_and_bed_outputs_byte_identical(tmp_path):
    """msp_to_lai / msp_to_bed against files written by the reference's own functions (src/postprocess.py:128-210,
    tests/golden/lai_bed.npz) from the .msp of meta.npz -- including the reference's "\\n.bed" name of the last column."""
    from gnomix_b200 import postprocess as pp
    d = np.load(os.path.join(G, "meta.npz"))
    g = np.load(os.path.join(G, "lai_bed.npz"))
    msp = tmp_path / "q.msp"
    msp.write_bytes(d["msp"].tobytes())
    df = pp.msp_to_lai(str(msp), d["qpos"], lai_file=str(tmp_path / "q.lai"))
    assert df.shape == (len(d["qpos"]), d["labels"].shape[0])
    assert open(tmp_path / "q.lai", "rb").read() == g["lai"].tobytes()
    for tag, pop_order in (("num", None), ("pop", d["pops"].tolist())):
        root = tmp_path / ("bed_" + tag)
        root.mkdir()
        pp.msp_to_bed(str(msp), str(root), pop_order=pop_order)
        names = sorted(os.listdir(root))
        assert names == g["bed_%s_names" % tag].tolist()
        for i, nm in enumerate(names):
            assert open(root / nm, "rb").read() == g["bed_%s_%d" % (tag, i)].tobytes(), (tag, nm)


def _write_bgzf(path, data, blk=4000):
    import struct, zlib
    with open(path, "wb") as f:
        for i in list(range(0, len(data), blk)) + [None]:          # the last member is the empty BGZF EOF marker
            chunk = b"" if i is None else data[i:i + blk]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp
                    + struct.pack("<II", zlib.crc32(chunk), len(chunk)))


def test_native_vcf_reader_bgzf_members_in_parallel(tmp_path):
    """bgzip / tabix files (BGZF) are inflated member by member in parallel; a gzip file that is not BGZF, or a BGZF
    file with a damaged member, falls back to / fails like the single zlib stream."""
    from gnomix_b200 import io as gio
    rng = np.random.default_rng(4)
    hdr = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join("X%d" % i for i in range(25)) + "\n"
    lut = np.array(["0|0", "0|1", "1|0", "1|1", ".|."])
    text = hdr + "".join("3\t%d\tr%d\tC\tT\t.\t.\t.\tGT\t%s\n" % (5 + 2 * r, r, "\t".join(lut[rng.integers(0, 5, 25)])) for r in range(3000))
    plain = tmp_path / "p.vcf"
    plain.write_text(text)
    bg = tmp_path / "b.vcf.gz"
    _write_bgzf(str(bg), text.encode())
    assert gzip.open(bg, "rb").read() == text.encode()               # a valid multi-member gzip file
    ref = gio.read_vcf_py(str(plain), "3")
    for th in ("1", "4"):
        os.environ["GNX_HOST_THREADS"] = th
        _same_vcf(gio.read_vcf(str(bg), "3"), ref)
    os.environ.pop("GNX_HOST_THREADS", None)
    # multi-member gzip without the BC field (e.g. `cat a.gz b.gz`): single-stream path, same result
    cat = tmp_path / "c.vcf.gz"
    half = len(text) // 2
    cat.write_bytes(gzip.compress(text[:half].encode()) + gzip.compress(text[half:].encode()))
    _same_vcf(gio.read_vcf(str(cat), "3"), ref)
    # a damaged member is an error, not silent garbage
    raw = bytearray(bg.read_bytes())
    raw[len(raw) // 2] ^= 0xFF
    bad = tmp_path / "bad.vcf.gz"
    bad.write_bytes(bytes(raw))
    try:
        d = gio.read_vcf(str(bad), "3")
        damaged_ok = d is not None and np.array_equal(d["calldata/GT"], ref["calldata/GT"])
    except Exception:
        damaged_ok = False
    assert not damaged_ok


def test_native_phased_vcf_writer_matches_python_loop(tmp_path):
    """gnx_write_vcf_body behind npy_to_vcf (src/utils.py:247-329) against the per-record Python loop: same bytes,
    including NaN QUAL, multi-allelic ALT (first allele kept), missing calls written as '2', header passthrough."""
    from gnomix_b200 import cli
    rng = np.random.default_rng(9)
    for R, n in [(1, 1), (257, 3), (1500, 40)]:
        data = {
            "calldata/GT": np.zeros((R, n, 2), dtype=np.int8),
            "samples": np.array(["S%d" % i for i in range(n)], dtype=object),
            "variants/CHROM": np.array(["22"] * R, dtype=object),
            "variants/POS": np.sort(rng.choice(np.arange(1, 10 ** 9), R, replace=False)).astype(np.int32),
            "variants/ID": np.array(["rs%d" % i if i % 7 else "." for i in range(R)], dtype=object),
            "variants/REF": rng.choice(np.array(["A", "C", "GT", "T"], dtype=object), R),
            "variants/ALT": np.stack([rng.choice(np.array(["A", "C", "G", "TTA"], dtype=object), R), np.array([""] * R, dtype=object),
                                      np.array([""] * R, dtype=object)], axis=1),
            "variants/QUAL": np.where(rng.random(R) < 0.5, np.nan, rng.random(R) * 100).astype(np.float32),
        }
        X = rng.integers(0, 3, size=(2 * n, R))
        a = cli.npy_to_vcf(data, X, str(tmp_path / ("a%d" % R)), headers="##contig=<ID=22>\n")
        b = cli.npy_to_vcf_py(data, X, str(tmp_path / ("b%d" % R)), headers="##contig=<ID=22>\n")
        assert open(a, "rb").read() == open(b, "rb").read(), (R, n)


def test_vcf_to_packed_equals_vcf_to_npy():
    """gnx_vcf_to_haplotypes_packed writes the 2-bit planes of exactly the matrix vcf_to_npy returns (reference golden
    included), for ascending and for shuffled index lists, every miss_fill in 0..3, C not a multiple of 64."""
    from gnomix_b200 import _lib, io as gio
    d = np.load(os.path.join(G, "vcf_to_npy.npz"))
    vcf = {"calldata/GT": d["gt"], "variants/POS": d["vpos"], "variants/REF": d["vref"]}
    P = gio.vcf_to_packed(vcf, d["model_pos"], d["model_ref"], verbose=False, pinned=False)
    assert P.shape == d["X"].shape and np.array_equal(P.to_numpy(), d["X"])
    assert np.array_equal(gio.PackedHaplotypes.from_numpy(d["X"], pinned=False).words, P.words)
    rng = np.random.default_rng(3)
    R, S, Cm = 700, 21, 1000 + 37
    gt = rng.integers(-1, 4, (R, S, 2)).astype(np.int8)
    for fill in (0, 1, 2, 3):
        for shuffle in (False, True):
            fmt = np.sort(rng.choice(Cm, 500, replace=False)).astype(np.int64)
            vi = rng.choice(R, 500, replace=False).astype(np.int64)
            if shuffle:
                perm = rng.permutation(500)
                fmt, vi = fmt[perm], vi[perm]
            swap = (rng.random(500) < 0.3).astype(np.uint8)
            X = np.empty((2 * S, Cm), np.int8)
            _lib.check(_lib.lib().gnx_vcf_to_haplotypes(gt.ctypes.data, R, S, vi.ctypes.data, fmt.ctypes.data, swap.ctypes.data, 500, Cm, fill,
                                                        X.ctypes.data, Cm, 3))
            out = gio.PackedHaplotypes(2 * S, Cm, pinned=False)
            out.words[:] = np.uint64(0xDEADBEEFDEADBEEF)
            _lib.check(_lib.lib().gnx_vcf_to_haplotypes_packed(gt.ctypes.data, R, S, vi.ctypes.data, fmt.ctypes.data, swap.ctypes.data, 500, Cm,
                                                               fill, out.words.ctypes.data, out.pitch_words, 3))
            assert np.array_equal(out.to_numpy(), X), (fill, shuffle)
            assert np.array_equal(out.words, gio.PackedHaplotypes.from_numpy(X, pinned=False).words), (fill, shuffle)
    assert _lib.lib().gnx_vcf_to_haplotypes_packed(gt.ctypes.data, R, S, vi.ctypes.data, fmt.ctypes.data, None, 500, Cm, 7,
                                                   out.words.ctypes.data, out.pitch_words, 1) != 0   # miss_fill must fit 2 bits
