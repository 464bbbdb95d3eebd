"""Extract the judged counters from gpurun_out/prof_*.ncu-rep into profiles/<prefix>_ncu_summary.json (+ raw csv pages)."""
import csv
import json
import subprocess
import sys

prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
names = sys.argv[2:] or ["swiglu", "attn_vit", "attn_llm", "layernorm", "rmsnorm", "gather_ln", "frames", "im2col"]
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct2",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "sm__icc_request_hit_rate.pct": "icache_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
}
out = {}
for n in names:
    rep = f"gpurun_out/prof_{n}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        print("no data for", n)
        continue
    open(f"profiles/{prefix}_ncu_full_{n}.csv", "w").write(raw)
    hdr, units, vals = rows[0], rows[1], rows[2]
    rec = {"kernel": vals[hdr.index("Kernel Name")][:110]}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                rec[WANT[h]] = float(v.replace(",", ""))
            except ValueError:
                rec[WANT[h]] = v
            rec[WANT[h] + "_unit"] = u
    if "dram_read" in rec and "duration" in rec:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        rd = rec["dram_read"] * scale.get(rec.get("dram_read_unit", "byte"), 1.0)
        wr = rec["dram_write"] * scale.get(rec.get("dram_write_unit", "byte"), 1.0)
        dur = rec["duration"] * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(rec.get("duration_unit", "ns"), 1e-9)
        rec["dram_bytes_total"] = rd + wr
        rec["dram_GBps"] = (rd + wr) / dur / 1e9
    out[n] = rec
    print(n, {k: v for k, v in rec.items() if not k.endswith("_unit")})
json.dump(out, open(f"profiles/{prefix}_ncu_summary.json", "w"), indent=1)
