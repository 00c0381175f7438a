"""profiles/*.csv (scripts/ncu_summary.py output of `ncu --set full` captures) -> profiles/ncu_summary.json:
per kernel the per-launch DRAM bytes, issue-slot utilisation, lanes per instruction and duration that bench.py
quotes next to its own live timings (roofline.traffic).  Later captures override earlier ones."""
import csv, glob, json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_ncu_full_*.csv"))):
    rows = {r[0]: r for r in csv.reader(open(path)) if r}
    if "Kernel Name" not in rows:
        continue
    names = rows["Kernel Name"][2:]
    def col(metric, i, scale=None):
        r = rows.get(metric)
        if not r or i + 2 >= len(r) or r[i + 2] in ("", "n/a"):
            return None
        v = float(r[i + 2].replace(",", ""))
        return v * (scale or {}).get(r[1], 1.0) if scale else v
    seen = set()
    for i, full in enumerate(names):
        m = re.search(r"(\w+_kernel)(<[^>(]*>)?\(", full) or re.search(r"(\w+_kernel)(<[^>(]*>)?", full)
        if not m:
            continue
        short = m.group(1)
        tmpl = m.group(2) or ""
        key = short + (tmpl if short == "cm_lists_kernel" else "")
        if (path, key) in seen:
            continue          # first launch of a kernel in a capture
        seen.add((path, key))
        rd, wr = col("dram__bytes_read.sum", i, SCALE), col("dram__bytes_write.sum", i, SCALE)
        out[key] = {"dram_bytes": None if rd is None else int(rd + (wr or 0)),
                    "issue_slots_busy_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active", i),
                    "threads_per_inst": col("smsp__thread_inst_executed_per_inst_executed.ratio", i),
                    "kernel_ms": col("gpu__time_duration.sum", i, TIME), "source": os.path.basename(path)}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1, sort_keys=True)
print(json.dumps({k: (v["dram_bytes"], v["source"]) for k, v in out.items()}, indent=1))
