#!/usr/bin/env python
"""profiles/scaling_rNN.txt from bench.py lines: scripts/scaling_table.py <label>=<bench.json> ... (first entry = the N=1 line of its config)"""
import json, sys
rows = []
for a in sys.argv[1:]:
    label, path = a.rsplit("=", 1)
    d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    rows.append((label, d))
base = {}
for label, d in rows:
    cfg = d["config"]["workload"][:2]
    if d["n_gpus"] == 1: base[cfg] = d
print("# device-timed value and end-to-end value per N (bench.py, one rank per GPU, max over ranks); efficiency = value / (N x the N=1 value of the same config)")
print(f"{'run':28s} {'N':>2s} {'scaling':>7s} {'value reg/s':>12s} {'eff':>5s} {'e2e reg/s':>11s} {'eff':>5s} {'step ms':>8s} {'e2e ms':>7s}  per-rank e2e H2D GB/s | plain H2D copy, all ranks at once GB/s")
for label, d in rows:
    cfg = d["config"]["workload"][:2]; n = d["n_gpus"]; b = base.get(cfg)
    ev = d["value"] / (n * b["value"]) if b else float("nan")
    ee = d["e2e"]["value"] / (n * b["e2e"]["value"]) if b else float("nan")
    cols = d["per_rank"]["columns"]; pr = d["per_rank"]["rows"]
    i_h = cols.index("e2e_h2d_GBps"); i_c = cols.index("plain_h2d_copy_GBps_all_ranks_at_once") if "plain_h2d_copy_GBps_all_ranks_at_once" in cols else None
    h = "/".join(f"{r[i_h]:.0f}" for r in pr); c = "/".join(f"{r[i_c]:.0f}" for r in pr) if i_c is not None else "-"
    print(f"{label:28s} {n:2d} {d['scaling']:>7s} {d['value']:12.0f} {ev:5.2f} {d['e2e']['value']:11.0f} {ee:5.2f} {d['ms_per_step']:8.2f} {d['e2e']['ms_per_step']:7.2f}  {h} | {c}")
