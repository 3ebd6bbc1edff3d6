import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dune_grid_b200 as dg
PARAMS = [1.0, 1.0, 1.0, 0.0, 0.25, 0.25, 0.25, 0.125, 0.2]
os.environ["DGB_FV_STAGING"] = sys.argv[1]
g = dg.YaspGrid(dim=3, N=(512, 512, 512), upper=(1.0, 1.0, 1.0), overlap=1)
fv = dg.FiniteVolume(g, 0, 0, PARAMS, "separable")
s = [fv.initial(), None]; s[1] = torch.empty_like(s[0])
def step():
    fv.step(s[0], s[1]); s.reverse()
for _ in range(10): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): step()
e1.record(); torch.cuda.synchronize()
print(sys.argv[1], os.environ.get("DGB_FV_TMA_SHAPE", "0"), os.environ.get("DGB_FV_ZCHUNK", "auto"), round(e0.elapsed_time(e1)/200, 4), fv.kernel_info()["staging"])
