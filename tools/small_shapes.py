"""Scratch: device-resident step time of the small named shapes under a few static-schedule settings."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sched_sweep import run
for (B, N) in [(1, 30000), (2, 30000), (8, 30000), (16, 30000), (32, 30000), (64, 5000), (16, 5000)]:
    for sc in ["0.6,16,24,4", "0.6,16,24,8", "0.6,16,24,12", "0.6,16,24,16", "0.6,16,24,24", "0.6,16,24,32"]:
        try:
            print(json.dumps(run(B, N, sc, 0)), flush=True)
        except Exception as e:
            print("failed", B, N, sc, e, flush=True)
