import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sched_sweep import run
for (B, N) in [(256, 30000), (1024, 30000)]:
    for v, scheds in [(0, ["0.6,16", "0.7,16", "0.5,16", "0.6,20"]), (1, ["1.0,0", "0.6,16", "0.6,32", "0.6,8"]), (6, ["1.0,0", "0.6,16", "0.6,32"]),
                      (2, ["1.0,0", "0.6,8", "0.6,16", "0.6,4"]), (3, ["1.0,0", "0.6,16"]), (7, ["1.0,0", "0.6,8", "0.6,16"]), (5, ["1.0,0", "0.6,16", "0.6,32"])]:
        for sc in scheds:
            try:
                print(json.dumps(run(B, N, sc, v)), flush=True)
            except Exception as e:
                print("failed", B, N, v, sc, e, flush=True)
