"""print the side measurements of a bench line: python tools/show_side_lines.py gpurun_out/x.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
if "lba_c3" in d:
    c = d["lba_c3"]
    print("lba_c3", c.get("value"), c.get("e2e"), c.get("unavailable"))
if "c4_dense" in d:
    c = d["c4_dense"]
    print("c4_dense", c.get("value"), c.get("ms_per_step"), c.get("unavailable"))
