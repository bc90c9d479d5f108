"""
TEST INFRASTRUCTURE.  Writes tests/golden/ref3d_checkpoint.pt and ref2d_checkpoint.pt with the REFERENCE'S OWN classes
(GaussianSplatting3DFast.save of /root/reference/3D/GSR.py:81-82, :179-188 and GaussianSplattingFast.save of 2D/GSR.py:81-82,
:231-240, run through tests/golden/ti_shim.py) and tests/golden/ref_checkpoint_expect.npz with what the reference's own kernels
return for the saved fields at a few points (value and gradient).  The GPU tests load the files with the CUDA classes
(`load_file=`, `.load()`), compare the fields, and compare the dict the CUDA classes save with the reference's, key by key.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_checkpoint.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

if __name__ == '__main__':
	out = {}
	rng = np.random.default_rng(97)
	# ---- 3D
	from make_golden_project3d import load as load3
	mod = load3()
	n = 3
	P = (np.stack(np.meshgrid(*[np.linspace(.12, .88, n)] * 3, indexing='ij'), -1).reshape(-1, 3) + rng.uniform(-.06, .06, (n ** 3, 3))).astype(np.float32)
	gv = mod.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., P, dim=3)
	with torch.no_grad():
		gv.scalings += torch.tensor(rng.uniform(-.15, .15, gv.scalings.shape).astype(np.float32))
		gv.rotations.copy_(torch.tensor(rng.normal(size=gv.rotations.shape).astype(np.float32)))
		gv.values.copy_(torch.tensor(rng.normal(scale=.3, size=gv.values.shape).astype(np.float32)))
	gv.reinitialize_grid()
	gv.save(os.path.join(HERE, 'ref3d_checkpoint.pt'))
	x = torch.tensor(rng.uniform(0., 1., (12, 3)).astype(np.float32))
	grad, val = gv.gradient(x, need_val=True)
	out.update(x3=x.numpy(), val3=val.detach().numpy(), grad3=grad.detach().numpy(), grid_scale3=np.float64(gv.grid_scale), grid_size3=np.array(gv.grid_size))
	back = mod.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., np.zeros((1, 3)), dim=3, load_file=os.path.join(HERE, 'ref3d_checkpoint.pt'))
	assert back.N == gv.N and back.grid_size == gv.grid_size
	# ---- 2D
	from make_golden_project2d import DOM, load as load2
	mod2 = load2()
	n = 5
	P2 = (np.stack(np.meshgrid(*[np.linspace(-4., 4., n)] * 2, indexing='ij'), -1).reshape(-1, 2) + rng.uniform(-.4, .4, (n * n, 2))).astype(np.float32)
	g2 = mod2.GaussianSplattingFast(*DOM, P2, dim=2)
	with torch.no_grad():
		g2.scalings += torch.tensor(rng.uniform(-.15, .15, g2.scalings.shape).astype(np.float32))
		g2.rotations.copy_(torch.tensor(rng.uniform(-np.pi, np.pi, g2.rotations.shape).astype(np.float32)))
		g2.values.copy_(torch.tensor(rng.normal(scale=.3, size=g2.values.shape).astype(np.float32)))
	g2.reinitialize_grid()
	g2.save(os.path.join(HERE, 'ref2d_checkpoint.pt'))
	x2 = torch.tensor(rng.uniform(-5., 5., (12, 2)).astype(np.float32))
	grad2, val2 = g2.gradient(x2, need_val=True)
	out.update(x2=x2.numpy(), val2=val2.detach().numpy(), grad2=grad2.detach().numpy(), grid_scale2=np.float64(g2.grid_scale), grid_size2=np.array(g2.grid_size))
	np.savez_compressed(os.path.join(HERE, 'ref_checkpoint_expect.npz'), **out)
	for f in ('ref3d_checkpoint.pt', 'ref2d_checkpoint.pt'):
		d = torch.load(os.path.join(HERE, f))
		print(f, {k: (tuple(v.shape), str(v.dtype), v.requires_grad) if isinstance(v, torch.Tensor) else v for k, v in d.items()})
