"""
Synthetic Gaussian fields of the benchmark (BASELINE.md §3 / SURVEY 8d): generated on the CPU with
torch.Generator().manual_seed(seed) so that every device and the CPU oracle see identical bits.
"""
import numpy as np
import torch

SIZES = {'S1': 10, 'S2': 40, 'S3': 100, 'S4': 160, 'S5': 256}	# lattice points per axis (N = n^3)


def synthetic_field(n, seed=42, tau=5e-3):
	"""
	Domain [0,1]^3; positions: n^3 lattice (3D/GSR.py:719-725) + U(-h/4, h/4) jitter; scalings s0 + N(0, .1^2) clipped to
	+-.2 (axis ratio < 1.5), s0 = 1/2 ln(-2 ln tau) - ln(min_grid_scale) (3D/GSR.py:166); rotations ~ N(0,1)^4
	(unnormalised); values ~ N(0, .1^2)^3.  Returns numpy arrays, min_grid_scale and the generator (for the samples).
	"""
	gen = torch.Generator().manual_seed(seed)
	N = n ** 3
	ax = torch.linspace(0., 1., n)
	P = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
	h = 1. / (n - 1)
	P = (P + (torch.rand(P.shape, generator=gen) - .5) * .5 * h).clamp(0., 1.)
	mgs = 2. * N ** (-1. / 3.)
	s0 = .5 * np.log(-2. * np.log(tau)) - np.log(mgs)
	S = s0 + (torch.randn((N, 3), generator=gen) * .1).clamp(-.2, .2)
	R = torch.randn((N, 4), generator=gen)
	V = torch.randn((N, 3), generator=gen) * .1
	return P.numpy(), S.numpy(), R.numpy(), V.numpy(), mgs, gen


def make_fast3d(P, S, R, V, tau, mgs):
	"""a GaussianSplatting3DFast on [0,1]^3 holding the given parameters"""
	from . import gsr3d
	o = gsr3d.GaussianSplatting3DFast(0., 1., 0., 1., 0., 1., np.asarray(P, np.float32), min_grid_scale=mgs, clamp_threshold=tau, dim=3)
	dev = gsr3d.device
	with torch.no_grad():
		o.scalings.copy_(torch.tensor(np.asarray(S, np.float32), device=dev))
		o.rotations.copy_(torch.tensor(np.asarray(R, np.float32), device=dev))
		o.values.copy_(torch.tensor(np.asarray(V, np.float32), device=dev))
	o.zero_grad()
	return o
