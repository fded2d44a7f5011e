"""Extract the per-kernel facts bench.py quotes (DRAM bytes per launch, issue / LSU utilisation, instruction and
shared-memory wavefront counts) from tracked ncu captures into profiles/ncu_facts.json.  Runs here, no GPU needed:

    python scripts/ncu_facts.py profiles/ncu_facts.json  KEY=capture.ncu-rep:kernel-name-regex[:workload note] ...

Every entry records which capture it came from; bench.py reads the JSON and never carries such numbers as literals."""
import csv, io, json, re, subprocess, sys

M = {
    "duration_ms": ("gpu__time_duration.sum", 1e-6),   # raw page reports ns
    "dram_read_bytes": ("dram__bytes_read.sum", None),
    "dram_write_bytes": ("dram__bytes_write.sum", None),
    "dram_pct_of_peak": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "inst_executed": ("smsp__inst_executed.sum", 1),
    "shared_wavefronts": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1),
    "shared_wavefronts_pct_of_peak": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    "shared_bank_conflict_wavefronts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1),
    "lsu_pipe_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    "alu_pipe_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    "fma_pipe_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    "fp64_pipe_pct": ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "tensor_pipe_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "registers_per_thread": ("launch__registers_per_thread", 1),
    "grid": ("launch__grid_size", 1),
    "block": ("launch__block_size", 1),
    "sm_cycles_elapsed_max": ("sm__cycles_elapsed.max", 1),
    "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
        "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def main():
    out_path = sys.argv[1]
    try:
        facts = json.load(open(out_path))
    except Exception:
        facts = {}
    cache = {}
    for spec in sys.argv[2:]:
        key, rest = spec.split("=", 1)
        parts = rest.split(":", 2)
        rep, pat = parts[0], parts[1]
        note = parts[2] if len(parts) > 2 else ""
        if rep not in cache:
            cache[rep] = rows_of(rep)
        hdr, units, rows = cache[rep]
        kn = hdr.index("Kernel Name")
        sel = [r for r in rows if re.search(pat, r[kn])]
        if not sel:
            print("no kernel matching %r in %s" % (pat, rep))
            continue
        r = sel[-1]   # the last matching launch (earlier ones may be warm-up sizes)
        d = {"kernel": r[kn].split("(")[0], "capture": rep, "launches_in_capture": len(sel)}
        if note:
            d["workload"] = note
        for name, (metric, scale) in M.items():
            if metric not in hdr:
                continue
            i = hdr.index(metric)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if name == "duration_ms":
                v = v * UNIT.get(u, 1.0) * 1e-6
            elif scale is None:
                v = v * UNIT.get(u, 1.0)
            d[name] = v
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        facts[key] = d
        print(key, json.dumps(d))
    json.dump(facts, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
