"""
Worker of tests/test_gpu_multirank.py — one process per GPU under torch.distributed.run (NCCL).  Cases:
  weak     sample-sharded optimisation (the one exchange per iteration): replicas bit-identical after two frames in both exchange
           modes (peer-memory kernel csrc/xrank.cu, and NCCL all-reduce), and the two modes agree to rounding
  strong   fixed 128^3-style job split by lattice planes, training replicated: every rank's trained field equals the 1-rank run's
           BITWISE, the lattice shards add up to the full lattice's losses
  timeout  rank 1 never publishes: rank 0's exchange kernel gives up after GSR_XRANK_SPIN_LIMIT polls, hands the step zeros
           and raises at the next check
Prints one JSON line per case on rank 0 and exits 0 when every assertion held.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

case = sys.argv[1]
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
if case == 'timeout':
	os.environ['GSR_XRANK_SPIN_LIMIT'] = str(1 << 16)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from gaussian_fluids_code_b200 import _lib, gsr3d, timestep3d  # noqa: E402
gsr3d.device = torch.device('cuda', local)


def flat(ts):
	return torch.cat([p.detach().flatten() for f in (ts.cur, ts.new) for p in (f.positions, f.scalings, f.rotations, f.values)])


def replicas_equal(t):
	others = [torch.empty_like(t) for _ in range(world)]
	dist.all_gather(others, t)
	return all(torch.equal(others[0], o) for o in others)


out = {'case': case}
if case == 'weak':
	res = {}
	for mode in ('nccl', 'p2p'):
		os.environ['GSR_EXCHANGE'] = mode
		ts = timestep3d.LeapfrogTimestep(n=10, iters=40, test_res=32, check_iter=20, rank=rank, world=world)
		for _ in range(2):
			ts.step()
		torch.cuda.synchronize()
		res[mode] = flat(ts)
		used = sorted({fp.exchange for f in (ts.cur, ts.new) for fp in f.__dict__.get('_pipelines', {}).values()})
		assert used == [mode], used
		assert bool(torch.isfinite(res[mode]).all())
		assert replicas_equal(res[mode]), mode
		out[mode + '_replicas_bit_identical'] = True
	d = (res['p2p'] - res['nccl']).abs().max().item() / res['nccl'].abs().max().item()
	out['p2p_vs_nccl_rel'] = d
	assert d < 1e-4, d	# different summation order of the same f32 partial sums, amplified over 80 Adam steps
elif case == 'strong':
	ts = timestep3d.LeapfrogTimestep(n=10, iters=40, test_res=32, check_iter=20, rank=rank, world=world, scaling='strong')
	one = timestep3d.LeapfrogTimestep(n=10, iters=40, test_res=32, check_iter=20, rank=0, world=1)
	for _ in range(2):
		ts.step(); one.step()
	torch.cuda.synchronize()
	assert torch.equal(flat(ts), flat(one))	# the training is replicated: not a bit may differ from the single-GPU job
	assert replicas_equal(flat(ts))
	rel = ((ts.last_test - one.last_test).abs() / one.last_test.abs()).max().item()
	out['lattice_losses_rel'] = rel
	assert rel < 1e-5, rel	# the shards' loss sums, all-reduced, against the full lattice's (summation order only)
elif case == 'timeout':
	ex = timestep3d.PeerExchange(1024, gsr3d.device)
	it = torch.zeros(1, device=gsr3d.device)
	ex.bufs[0].fill_(1.)
	ex.out.fill_(7.)
	torch.cuda.synchronize(); dist.barrier()
	if rank == 0:
		ex.sum(0, it)	# the peers never call: the wait must end by itself
		torch.cuda.synchronize()
		assert int(ex.err.item()) == 1 and float(ex.out.abs().max()) == 0.	# zeros, not a sum over stale buffers
		try:
			ex.check()
			raise AssertionError('no error raised')
		except _lib.GsrError:
			out['raised'] = True
	dist.barrier()
else:
	raise SystemExit('unknown case ' + case)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
	print('RESULT ' + json.dumps(out), flush=True)
# symmetric-memory mappings hold peer allocations: drop everything that owns them before the process group goes away
del out
import gc  # noqa: E402
for name in ('ts', 'one', 'ex'):
	globals().pop(name, None)
gc.collect()
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
