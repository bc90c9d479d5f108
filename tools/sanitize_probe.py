"""small end-to-end time step for compute-sanitizer (development tool): eager iterations, two lattice test passes, output passes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import timestep3d, gsr3d
gsr3d.device = torch.device('cuda', 0)
ts = timestep3d.LeapfrogTimestep(n=int(sys.argv[1]) if len(sys.argv) > 1 else 10, iters=4, test_res=64, check_iter=2, use_graph=len(sys.argv) > 2 and sys.argv[2] == 'graph')
for _ in range(2):
	ts.step()
torch.cuda.synchronize()
print('ok')
