import os, sys, json, time
sys.path.insert(0, ".")
os.environ["TB_E2E_STAGES"] = "1"
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for rep in range(2):
    r = bench.e2e_mapmaker("c4", 128, 2160000, 50, dev, 0, 1)
    print({k: v for k, v in r.items()})
