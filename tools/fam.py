import sys, json
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    j = json.loads(line)
    fam = {k: round(v, 2) for k, v in j["roofline"]["family_ms_per_step"].items()}
    print(sys.argv[1] if len(sys.argv) > 1 else "", round(j["ms_per_step"], 2), "ms", fam)
