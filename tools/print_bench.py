"""Key fields of bench.py JSON lines read from stdin (one per line)."""
import json
import sys

for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    e2e = d.get("e2e") or {}
    r = d.get("roofline") or {}
    cb = d.get("cpu_baseline") or {}
    print(f"{d.get('impl', 'b200'):9s} {d['metric']:26s} value={d['value']:.1f} ms/step={d.get('ms_per_step', 0):.3f} "
          f"e2e={e2e.get('value', 0):.1f} roofline={r.get('achieved', 0):.1f}/{r.get('peak', 0):.0f} {r.get('unit', '')} "
          f"cpu={cb.get('value', 0):.1f} [{cb.get('sample', '')[:40]}]")
