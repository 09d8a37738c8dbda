"""profiles/traffic.json from an ncu csv (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of ONE launch).

usage: ncu_traffic.py <csv> <leaf> <explores> <games> <group_lanes> "<how>" [config] [commit]
bench.py reads the record whose (config, leaf, explores, games, group_lanes) and kernel match its workload for `roofline.traffic`,
and ignores it when its own launch time differs from the capture's gpu_time_ns by more than 3 %."""
import csv
import json
import os
import sys

path, leaf, explores, games, gl, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
config = int(sys.argv[7]) if len(sys.argv) > 7 else (1 if leaf == "nn" else 0)
commit = sys.argv[8] if len(sys.argv) > 8 else ""
vals, kernel = {}, None
for r in csv.reader(open(path)):
    if len(r) > 10 and r[0].isdigit():
        vals[r[-3]] = int(float(r[-1]))
        kernel = r[4]
rec = {"config": config, "commit": commit, "leaf": leaf, "explores": explores, "games": games, "group_lanes": gl, "kernel": kernel,
       "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
       "dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], "gpu_time_ns": vals["gpu__time_duration.sum"],
       "how": how}
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
recs = []
if os.path.exists(out):
    recs = [x for x in json.load(open(out)) if (x.get("config", 1 if x["leaf"] == "nn" else 0), x["leaf"], x["explores"], x["games"], x["group_lanes"]) != (config, leaf, explores, games, gl)]
recs.append(rec)
json.dump(recs, open(out, "w"), indent=1)
print(json.dumps(rec))
