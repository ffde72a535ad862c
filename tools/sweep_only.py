import json, sys, torch
sys.path.insert(0, "/root/repo")
import bench
print(json.dumps(bench.sweep_record(torch.device("cuda", 0))))
