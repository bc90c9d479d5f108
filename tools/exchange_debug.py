import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from gaussian_fluids_code_b200 import gsr3d, timestep3d
gsr3d.device = torch.device('cuda', local)
iters, res, graph = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] == 'graph'
ts = timestep3d.LeapfrogTimestep(n=10, iters=iters, test_res=res, check_iter=100 if iters >= 100 else iters, rank=rank, world=world, use_graph=graph)
for k in range(3):
	ts.reset()
	ts.step()
	torch.cuda.synchronize()
	if rank == 0:
		print('step', k, 'ok', flush=True)
dist.barrier()
os._exit(0)
