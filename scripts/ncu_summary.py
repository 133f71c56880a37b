"""Condenses an `ncu --page raw --csv` export into the per-kernel entries of profiles/r02_ncu_summary.json.

    python scripts/ncu_summary.py profiles/r02e_tc_fwd_bwd2.raw.csv r02e --coords 303104 --T 20 [--out profiles/r02_ncu_summary.json]
"""
import argparse
import csv
import json
import os

KEYS = {
    "duration_ms": ("gpu__time_duration.sum", 1.0),
    "dram_read_bytes": ("dram__bytes_read.sum", 1.0),
    "dram_write_bytes": ("dram__bytes_write.sum", 1.0),
    "tensor_pipe_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "xu_pipe_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    "fma_pipe_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1.0),
    "lsu_pipe_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "registers": ("launch__registers_per_thread", 1.0),
    "warp_inst": ("smsp__inst_executed.sum", 1.0),
    "smem_bank_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("tag")
    ap.add_argument("--coords", type=int, required=True)
    ap.add_argument("--T", type=int, required=True)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "profiles", "r02_ncu_summary.json"))
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = json.load(open(a.out)) if os.path.exists(a.out) else {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = name.split("<")[0].replace("void ", "").strip()
        e = {"kernel": name[:100], "coords": a.coords, "T": a.T, "source": os.path.relpath(a.csv)}
        for k, (m, _) in KEYS.items():
            if m in idx and r[idx[m]] != "":
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                if k.endswith("_bytes") or k == "duration_ms":
                    v *= UNIT_SCALE.get(u, 1.0)
                e[k] = v
        stalls = {}
        for h in hdr:
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[idx[h]])
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        e["stall_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]}
        cu = a.coords * a.T
        e["dram_bytes_per_coord_update"] = (e.get("dram_read_bytes", 0) + e.get("dram_write_bytes", 0)) / cu
        e["warp_inst_per_coord_update"] = e.get("warp_inst", 0) / cu
        e["coord_updates_per_s_under_ncu"] = cu / (e["duration_ms"] * 1e-3)
        out["%s_%s" % (a.tag, short.replace("::", "_"))] = e
    out["_note"] = ("ncu --set full --clock-control none; per-launch numbers under the profiler (cold caches, serialised): "
                    "use for shares / pipe utilisation / stall mix, not as bench values")
    json.dump(out, open(a.out, "w"), indent=1)
    print("wrote", a.out, [k for k in out if k.startswith(a.tag)])


if __name__ == "__main__":
    main()
