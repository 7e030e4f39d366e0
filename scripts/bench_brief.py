#!/usr/bin/env python
"""one-line digest of a bench.py JSON line on stdin (tuning runs)"""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
st = d["roofline"]["stage_ms_per_step"]
print("%s value %.0f e2e %.0f | %s | %s" % (sys.argv[1] if len(sys.argv) > 1 else "", d["value"], d["e2e"]["value"], {k: v for k, v in d.get("kernel_S_last_step", {}).items() if k.startswith("rows")},
                                        " ".join("%s %.2f" % (k, v) for k, v in st.items() if v > 0.05)))
