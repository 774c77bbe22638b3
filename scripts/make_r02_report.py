#!/usr/bin/env python
"""Regenerate the driver-style table of profiles/r02_measurements.md (section 1) from the bench lines under gpurun_out/.
usage: make_r02_report.py n1.json n2.json n4.json n8.json   (prints the table)"""
import json, sys


def line(f):
    return json.loads(open(f).read().strip().splitlines()[-1])


rows = [line(f) for f in sys.argv[1:]]
n1 = rows[0]
print("| GPUs | sweeps/s (all ranks) | e2e sweeps/s (20-sweep stateless calls) | nsig wall s | speed-up | BIC-optimal nsig (planted 20) | sharded ms/sweep | speed-up | sharded_parity_ok |")
print("|---|---|---|---|---|---|---|---|---|")
for d in rows:
    print("| %d | %.0f | %.0f | %.2f | x%.2f | %d | %.3f | x%.2f | %s |" % (
        d["n_gpus"], d["value"], d["e2e"]["value"], d["nsig_wall_s"], n1["nsig_wall_s"] / d["nsig_wall_s"], d["bic_optimal_nsig"],
        d["sharded_ms_per_sweep"], n1["sharded_ms_per_sweep"] / d["sharded_ms_per_sweep"], d["sharded_parity_ok"]))
