"""Per-kernel SASS evidence of the Blackwell paths (run here, no GPU): python scripts/sass_markers.py > profiles/sass_markers.txt
Counts the mnemonics that prove tcgen05 / TMA / TMEM use (B200_PROFILING.md) in every kernel of gnomix_b200/libgnx.so."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gnomix_b200", "libgnx.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MARK = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDS", "STS", "LDG", "STG",
        "POPC", "FLO", "SHF", "LOP3", "DFMA", "DADD", "DMUL", "IDP", "BAR", "ATOM", "RED", "LDC", "ULDC"]
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        sizes[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1).split(".")[0]
        sizes[cur] += 1
        for k in MARK:
            if op == k or (k in ("UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS") and op.startswith(k)):
                counts[cur][k] += 1
names = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("SASS markers per kernel of gnomix_b200/libgnx.so (cuobjdump -sass, sm_100a); sizes in instructions")
print("arch:", sorted(set(re.findall(r"arch = (sm_\w+)", out))))
for (fn, c), nm in zip(counts.items(), names):
    depth, cut = 0, len(nm)          # drop the trailing parameter list, keep template arguments
    for i in range(len(nm) - 1, -1, -1):
        if nm[i] == ")":
            depth += 1
        elif nm[i] == "(":
            depth -= 1
            if depth == 0:
                cut = i
                break
    nm = nm[:cut].replace("gnx::", "").replace("(int)", "").replace("(bool)", "")
    print("%-62s %6d  %s" % (nm[:62], sizes[fn], "  ".join("%s=%d" % (k, c[k]) for k in MARK if c[k])))
