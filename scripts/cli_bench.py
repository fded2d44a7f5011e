"""End-to-end driver timing: VCF in -> .msp / .fb out (gnomix.py run_inference, phase=False) for a synthetic cohort,
with the time of every stage, next to the pure-Python / numpy host stages (read_vcf_py, vcf_to_npy_py, per-number
.fb formatting).  python scripts/cli_bench.py [n_samples] [workdir]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from gnomix_b200 import Gnomix, GBTForest, synth, io as gio, postprocess as pp

n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
work = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp()
rng = np.random.default_rng(94305)
C, M, A, S = 317_408, 857, 7, 75            # the reference's demo chr22 geometry (370 windows)
W = C // M
model = Gnomix(C, M, A, S)
freqs = synth.population_frequencies(rng, C, A)
fx, fpop = synth.founders(rng, freqs, per_pop=8)
coefs, icpts = synth.discriminant_lr_weights(freqs, C, M, model.context)
model.base.set_window_weights(coefs, icpts)
model.smooth.model = GBTForest.random(rng, A, model.smooth.S, n_rounds=100, depth=4)
pos = np.sort(rng.choice(np.arange(16_050_000, 51_240_000), C, replace=False)).astype(np.int64)
model.snp_pos, model.snp_ref, model.snp_alt = pos, np.array(["A"] * C, dtype=object), np.array(["G"] * C, dtype=object)
model.population_order = ["P%d" % i for i in range(A)]
model.gen_map_df = pd.DataFrame({"chm": ["22"] * 1000, "pos": np.linspace(16e6, 51.3e6, 1000).astype(int), "pos_cm": np.linspace(0, 74.1, 1000)})
X, _ = synth.admix_host(rng, fx, fpop, 2 * n_samples, morgans=0.74)
X[X > 1] = 0
# VCF text built as a byte matrix (fast): one record per SNP
t = time.perf_counter()
vcf_path = os.path.join(work, "cohort.vcf")
digits = (X.T + ord("0")).astype(np.uint8)
row = np.empty((C, 4 * n_samples), dtype=np.uint8)
row[:, 0::4] = ord("\t"); row[:, 1::4] = digits[:, 0::2]; row[:, 2::4] = ord("|"); row[:, 3::4] = digits[:, 1::2]
with open(vcf_path, "wb") as f:
    f.write(b"##fileformat=VCFv4.2\n##contig=<ID=22>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join("S%d" % i for i in range(n_samples)).encode() + b"\n")
    for j in range(C):
        f.write(b"22\t%d\trs%d\tA\tG\t.\tPASS\t.\tGT" % (pos[j], j)); f.write(row[j].tobytes()); f.write(b"\n")
print("wrote %s: %.0f MB in %.1f s" % (vcf_path, os.path.getsize(vcf_path) / 1e6, time.perf_counter() - t), flush=True)

def stage(name, fn, out):
    t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out[name] = time.perf_counter() - t0; return r

res = {}
model.predict_proba(X[:64])                                   # warm the library (module load, workspace)
vcf = stage("read_vcf", lambda: gio.read_vcf(vcf_path, "22"), res)
Xq, vi, fi = stage("vcf_to_npy", lambda: gio.vcf_to_npy(vcf, model.snp_pos, model.snp_ref, return_idx=True, verbose=False), res)
assert np.array_equal(Xq, X)
proba = stage("inference (upload + K1 + K4 + download proba)", lambda: model.predict_proba(Xq), res)
y = np.argmax(proba, axis=-1)
meta = stage("get_meta_data", lambda: pp.get_meta_data("22", model.snp_pos, vcf["variants/POS"], model.W, model.M, model.gen_map_df), res)
stage("write_msp", lambda: pp.write_msp(os.path.join(work, "q"), meta, y, model.population_order, vcf["samples"]), res)
stage("write_fb", lambda: pp.write_fb(os.path.join(work, "q"), meta, proba, model.population_order, vcf["samples"]), res)
print("native host stages, %d samples x %d SNPs:" % (n_samples, C))
for k, v in res.items(): print("  %-50s %7.2f s" % (k, v))
print("  %-50s %7.2f s" % ("total", sum(res.values())), flush=True)
old = {}
v2 = stage("read_vcf_py", lambda: gio.read_vcf_py(vcf_path, "22"), old)
stage("vcf_to_npy_py", lambda: gio.vcf_to_npy_py(v2, model.snp_pos, model.snp_ref, return_idx=True, verbose=False), old)
def fb_py():
    fb_prob = np.swapaxes(proba, 1, 2).reshape(-1, W).T
    with open(os.path.join(work, "q_py.fb"), "w") as f:
        for l in range(W):
            f.write("\t".join(fb_prob[l].astype(str))); f.write("\n")
stage("write_fb body, per-number Python formatting", fb_py, old)
print("the same host stages in Python / numpy:")
for k, v in old.items(): print("  %-50s %7.2f s" % (k, v))
