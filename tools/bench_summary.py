import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["value"], d["config"]["seconds_per_layer"], d["config"]["quarter_ms"]); print({k:v for k,v in d["phases_ms"].items() if k.startswith("boot_") or k.startswith("att")}); print(sorted(d["kernels_ms"].items(), key=lambda kv:-kv[1][0])[:8])
