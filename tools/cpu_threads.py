import os, sys, time, json, torch
sys.path.insert(0, os.getcwd())
import bench
res = {}
for nt in (16, 32, 64, 128):
    if nt > (os.cpu_count() or 1):
        continue
    torch.set_num_threads(nt)
    v, per = bench.cpu_step_rate(4, 256, 2, 1)
    res[nt] = {"median_step_s": sorted(per)[len(per) // 2], "value": v}
    print(nt, res[nt], flush=True)
print(json.dumps({"cpus": os.cpu_count(), "default_torch_threads": None, "by_threads": res}))
