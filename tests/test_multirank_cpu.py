"""
World-size-2 tests of the multi-GPU design on CPU (gloo): the sharding bookkeeping of timestep3d and the identity the
single all-reduce relies on — with the loss normalisers taken over the GLOBAL sample count, the sum over ranks of the
per-shard gradients equals the gradient of the unsharded batch.  The CPU oracle stands in for the engine here (test
infrastructure); the CUDA path is checked against the same oracle in the -m gpu tests.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import NAMES, load_golden, oracle_from_golden, rel_err


def _free_port():
	with socket.socket() as s:
		s.bind(('127.0.0.1', 0))
		return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
	os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
	dist.init_process_group('gloo', rank=rank, world_size=world)
	try:
		from gaussian_fluids_code_b200 import timestep3d
		# 1. lattice shards: disjoint, union = the (res, res, world * res) lattice, equal work per rank
		res = 6
		mine, total = timestep3d.shard_lattice(res, rank, world, 'cpu')
		assert mine.shape == (res ** 3, 3) and total == world * res ** 3
		gathered = [torch.empty_like(mine) for _ in range(world)]
		dist.all_gather(gathered, mine)
		allpts = torch.cat(gathered)
		full = torch.stack(torch.meshgrid(torch.linspace(0, 1, res), torch.linspace(0, 1, res), torch.linspace(0, 1, res * world), indexing='ij'), -1).reshape(-1, 3)
		key = lambda t: sorted(map(tuple, (t * 1e6).round().to(torch.int64).tolist()))
		assert key(allpts) == key(full)
		# 1b. strong scaling: the reference's own res^3 lattice, split by z planes — disjoint, union = the lattice
		res2 = 6
		mine2, total2 = timestep3d.shard_lattice(res2, rank, world, 'cpu', scaling='strong')
		assert total2 == res2 ** 3 and mine2.shape[0] == res2 ** 3 // world
		gathered = [torch.empty_like(mine2) for _ in range(world)]
		dist.all_gather(gathered, mine2)
		full2 = torch.stack(torch.meshgrid(*[torch.linspace(0, 1, res2)] * 3, indexing='ij'), -1).reshape(-1, 3)
		assert key(torch.cat(gathered)) == key(full2)
		# 1c. the test losses of the whole lattice from the shards: sum of the shards' loss sums over the total point count
		loss = (mine2.double() ** 2).sum(-1)
		sums = loss.sum()[None]
		dist.all_reduce(sums)
		assert abs(float(sums) / total2 - float((full2.double() ** 2).sum(-1).mean())) < 1e-9
		# 2. flat all-reduce buffer layout
		lay = timestep3d.flat_layout(10, 3, 5)
		assert lay['acc'] == (0, 360) and lay['lp'] == (360, 384) and lay['lpb'] == (384, 424) and lay['total'] == 424
		# 3. sum of shard gradients (global normaliser) == gradient of the whole batch
		g = load_golden('ref3d_kernels_f64.npz')
		o = oracle_from_golden(g, 3, 'f64')
		X = g['in_x']
		Q = X.shape[0]
		sl = slice(rank * Q // world, (rank + 1) * Q // world)
		val, grad = o.forward(X[sl])
		scale = (sl.stop - sl.start) / Q	# oracle normalises by the shard size; the engine's Q_norm is the global Q
		_, vor, div = o.backward3d(X[sl], val, grad, ref_vor=g['in_ref_vor'][sl], weight_vor=scale, ref_hel=g['in_ref_hel'][sl], weight_hel=scale,
								   weight_div=scale, direct=o.zero_grads(), vor=o.zero_grads(), div=o.zero_grads())
		flat = torch.from_numpy(np.concatenate([a.ravel() for a in vor + div]))
		dist.all_reduce(flat)
		if rank == 0:
			fv, fg = o.forward(X)
			_, rvor, rdiv = o.backward3d(X, fv, fg, ref_vor=g['in_ref_vor'], weight_vor=1., ref_hel=g['in_ref_hel'], weight_hel=1., weight_div=1.,
										 direct=o.zero_grads(), vor=o.zero_grads(), div=o.zero_grads())
			ref = np.concatenate([a.ravel() for a in rvor + rdiv])
			assert rel_err(flat.numpy(), ref) < 1e-12
		# 4. replicas that apply the same reduced buffer stay bit-identical: every rank holds the same bytes after the all-reduce
		digest = torch.tensor([float(flat.double().sum()), float(flat.double().abs().max())], dtype=torch.float64)
		both = [torch.empty_like(digest) for _ in range(world)]
		dist.all_gather(both, digest)
		assert all(torch.equal(both[0], b) for b in both)
		open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
	finally:
		dist.destroy_process_group()


def test_world2_gloo(tmp_path):
	world = 2
	mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
	assert all(os.path.exists(tmp_path / f'ok{r}') for r in range(world))
