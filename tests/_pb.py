import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(round(d["value"]), round(d["ms_per_step"],1), round(d["e2e"]["value"]), {k:round(v,1) for k,v in d["phases_ms_per_step"].items()}, d["gpu_launches"])
    elif "Error" in l or "error" in l: print(l.rstrip()[:300])
