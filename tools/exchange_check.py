"""2-rank check of the peer-memory exchange against NCCL (run under torch.distributed.run): same parameters after a short step, and
bit-identical replicas.  python -m torch.distributed.run --nproc-per-node 2 tools/exchange_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from gaussian_fluids_code_b200 import gsr3d, timestep3d
gsr3d.device = torch.device('cuda', local)
res = {}
for mode in ('nccl', 'p2p'):
	os.environ['GSR_EXCHANGE'] = mode
	ts = timestep3d.LeapfrogTimestep(n=10, iters=40, test_res=32, check_iter=20, rank=rank, world=world)
	for _ in range(2):
		ts.step()
	torch.cuda.synchronize()
	params = torch.cat([p.detach().flatten() for p in ts.cur._params()])
	others = [torch.empty_like(params) for _ in range(world)]
	dist.all_gather(others, params)
	same = all(torch.equal(others[0], o) for o in others)
	res[mode] = params
	ex = [e['fp'].exchange for e in ts._proj.values()]
	if rank == 0:
		print(mode, 'exchange used:', ex, 'replicas bit-identical:', same, 'finite:', bool(torch.isfinite(params).all()), flush=True)
	assert same
if rank == 0:
	d = (res['p2p'] - res['nccl']).abs().max().item()
	print('max |p2p - nccl| over all parameters:', d, 'scale', res['nccl'].abs().max().item(), flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
