"""Shared helpers for the tests (test infrastructure; may use oracle/)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
NAMES = ('positions', 'scalings', 'rotations', 'values')


def load_golden(name):
	return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def rel_err(a, b):
	"""max|a-b| / max|b|  (the per-tensor relative error of SURVEY 8c)"""
	a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
	den = np.abs(b).max()
	return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a).max())


def oracle_from_golden(g, D, precision, **kw):
	from oracle.oracle import OracleGSR, extended_bounds
	mgs = float(g['in_min_grid_scale'])
	ext = extended_bounds(D, (0., 1.) * D, mgs)
	return OracleGSR(D, ext, g['in_positions'], g['in_scalings'], g['in_rotations'], g['in_values'], float(g['in_tau']), mgs, precision=precision, **kw)


def ring_particles_np(ring, real=np.float64):
	"""the n vortex particles of a ring as the reference builds them (3D/init_cond.py:147-156; torch ops in the default dtype —
	float32 in the reference, float64 in the f64 golden run): x0 (n,3), w (n,3) strength-scaled tangents, U = radius / (2 n),
	a = thickness.  Arithmetic in `real`."""
	import torch
	dt = torch.float32 if real == np.float32 else torch.float64
	normal, center = torch.tensor(ring['normal'], dtype=dt), torch.tensor(ring['center'], dtype=dt)
	axis_x = torch.tensor([1., 0., 0.], dtype=dt)
	if torch.linalg.cross(axis_x, normal).norm() < 1e-5:
		axis_x = torch.tensor([0., 1., 0.], dtype=dt)
	axis_y = torch.linalg.cross(normal, axis_x)
	axis_y = axis_y / axis_y.norm()
	axis_x = torch.linalg.cross(axis_y, normal)
	theta = torch.linspace(0., 2. * torch.pi, ring['n'] + 1, dtype=dt)[:-1]
	x0 = (axis_x[None] * torch.cos(theta)[:, None] + axis_y[None] * torch.sin(theta)[:, None]) * ring['radius'] + center
	w = (axis_x[None] * -torch.sin(theta)[:, None] + axis_y[None] * torch.cos(theta)[:, None]) * ring['strength']
	return x0.numpy().astype(real), w.numpy().astype(real), ring['radius'] / (2 * ring['n']), ring['thickness']
