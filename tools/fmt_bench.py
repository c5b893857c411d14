import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d["value"]), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["detail_ms_per_step"].items()}, round(d["e2e"]["value"]) if d.get("e2e") else None)
