"""Pull the per-kernel numbers bench.py quotes next to its roofline out of ncu reports:
    python scripts/ncu_metrics.py workload=path.ncu-rep [...]  -> profiles/ncu_metrics.json
(dram bytes per launch, FP32-pipe, shared-memory/l1tex data-pipe and issue utilisation, duration under ncu)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "dram_bytes": ("dram__bytes_read.sum", "dram__bytes_write.sum"),
    "fp32_pipe_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",),
    "l1tex_pct": ("l1tex__throughput.avg.pct_of_peak_sustained_active",),
    "smem_wavefront_pct": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",),
    "issue_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active",),
    "duration_us_under_ncu": ("gpu__time_duration.sum",),
    "registers": ("launch__registers_per_thread",),
    "smem_wavefronts": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",),
    "smem_bank_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",),
    "warp_instructions": ("smsp__inst_executed.sum",),
}
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def main():
    path = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        wl, rep = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            kind = "fwd" if "fwd" in name else ("bwd" if "bwd" in name else None)
            if kind is None:
                continue
            rec = {"kernel": name.split("(")[0].strip(), "source": os.path.basename(rep)}
            for k, cols in KEYS.items():
                tot = 0.0
                for c in cols:
                    if c in idx:
                        tot += float(r[idx[c]].replace(",", "")) * UNIT_SCALE.get(units[idx[c]], 1.0)
                rec[k] = round(tot, 3)
            out.setdefault(wl, {})[kind] = rec
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
