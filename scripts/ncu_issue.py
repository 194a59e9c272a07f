#!/usr/bin/env python
"""ncu launch list of bench.py's k_rollout_private launches -> profiles/rollout_issue.json
and profiles/rollout_traffic.json.

    ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum \
        --clock-control none -k regex:k_rollout_private --csv --log-file gpurun_out/issue.csv \
        python bench.py --main-only --steps 20 --warmup 3
    python scripts/ncu_issue.py gpurun_out/issue.csv r02
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, tag = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if r[0] == "ID")
col = {name: hdr.index(name) for name in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
by_id = {}
for r in rows:
    if r[0] == "ID" or "k_rollout_private" not in r[col["Kernel Name"]]:
        continue
    by_id.setdefault(int(r[col["ID"]]), {})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
launches = [by_id[k] for k in sorted(by_id)]
inst = [l["smsp__inst_executed.sum"] for l in launches]
issue = {"kernel": "k_rollout_private<boat,philox>",
         "source": "profiles/%s_launches_rollout_issue.csv: ncu --metrics smsp__inst_executed.sum ... python bench.py --main-only "
                   "--steps 20 --warmup 3 (launch i of the list = bench launch i: 3 warm-up, 20 timed, 2 + 20 host-buffer)" % tag,
         "warp_instructions_by_launch": inst,
         "ncu_ms_by_launch": [l["gpu__time_duration.sum"] / 1e6 for l in launches]}
for pipe in ("alu", "fma"):                                  # warp-instructions through each half-rate pipe
    name = "sm__inst_executed_pipe_%s.sum" % pipe
    if all(name in l for l in launches):
        issue["%s_pipe_warp_instructions_by_launch" % pipe] = [l[name] for l in launches]
json.dump(issue, open(os.path.join(ROOT, "profiles", "rollout_issue.json"), "w"), indent=1)
timed = launches[3:23] if len(launches) >= 23 else launches
rd = sum(l["dram__bytes_read.sum"] for l in timed) / len(timed)
wr = sum(l["dram__bytes_write.sum"] for l in timed) / len(timed)
traffic = {"kernel": "k_rollout_private<boat,philox>", "source": issue["source"], "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "env_steps_per_launch": 655360000, "algorithmic_bytes_per_launch": 86507520000}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "rollout_traffic.json"), "w"), indent=1)
print("launches", len(launches), "mean warp-inst", sum(inst) / len(inst), "dram bytes/launch", rd + wr)
