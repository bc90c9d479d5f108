"""development probe: a few eager project iterations at a given size (for ncu launch lists)"""
import sys
sys.path.insert(0, '.')
import torch
from gaussian_fluids_code_b200 import timestep3d, gsr3d
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
gsr3d.device = torch.device('cuda', 0)
ts = timestep3d.LeapfrogTimestep(n=n, iters=iters, test_res=int(sys.argv[3]) if len(sys.argv) > 3 else 32, check_iter=1000, use_graph=False)
ts.step()
torch.cuda.synchronize()
print('ok')
