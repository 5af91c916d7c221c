import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import torch
import config_runs as cr
dev = torch.device("cuda:0")
for g in ("auto", "int8", "dmma"):
    print(json.dumps(cr.run_linear(dev, reps=2, check=8, gemm=g)), flush=True)
    torch.cuda.empty_cache()
