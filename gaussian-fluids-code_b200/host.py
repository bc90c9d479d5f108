"""
Host-side scalar logic of the spatial hash — pure Python, no device work, unit-tested on CPU.
Follows the reference's constructors and reinitialize_grid (3D/GSR.py:156-177, :247-252; 2D/GSR.py:173-192, :224-229).
"""
import math

import numpy as np

from ._lib import GridDesc


def default_min_grid_scale(D, bounds, N):
	"""2 (V/N)^(1/3) in 3D (3D/GSR.py:160); 3 (A/N)^(1/2) in 2D (2D/GSR.py:177)."""
	ext = [bounds[2 * k + 1] - bounds[2 * k] for k in range(D)]
	return (ext[0] * ext[1] * ext[2] / N) ** (1. / 3.) * 2. if D == 3 else (ext[0] * ext[1] / N) ** .5 * 3.


def extend(D, bounds, min_grid_scale):
	"""the domain grown by one min_grid_scale on every side (3D/GSR.py:162-164)"""
	out = []
	for k in range(D):
		out += [bounds[2 * k] - min_grid_scale, bounds[2 * k + 1] + min_grid_scale]
	return out


def initial_scaling(tau, min_grid_scale):
	"""log inverse radius such that the truncation radius equals the cell size (3D/GSR.py:166)"""
	return .5 * np.log(-2. * np.log(tau)) - np.log(min_grid_scale)


def grid_size(D, ext_bounds, min_grid_scale):
	"""create_grid_data: 3D/GSR.py:173; 2D/GSR.py:188 (the 2D y axis uses true division)"""
	if D == 3:
		return [int((ext_bounds[2 * k + 1] - ext_bounds[2 * k]) // min_grid_scale) + 1 for k in range(3)]
	return [int((ext_bounds[1] - ext_bounds[0]) // min_grid_scale) + 1, int((ext_bounds[3] - ext_bounds[2]) / min_grid_scale) + 1]


def grid_scale(tau, min_scaling, min_grid_scale, ext_bounds):
	"""reinitialize_grid's host-double formula (3D/GSR.py:248-251)"""
	if tau:
		return max(np.sqrt(-2. * np.log(tau)) * np.exp(-min_scaling), min_grid_scale)
	D = len(ext_bounds) // 2
	return max(ext_bounds[2 * k + 1] - ext_bounds[2 * k] for k in range(D))


def make_desc(D, ext_bounds, dims, gscale, tau):
	"""the kernel constants, rounded to f32 the way the reference's JIT bakes / passes them"""
	d = GridDesc()
	d.D = D
	for k in range(3):
		d.dims[k] = int(dims[k]) if k < D else 1
		d.lo[k] = float(ext_bounds[2 * k]) if k < D else 0.
		d.hi[k] = float(ext_bounds[2 * k + 1]) if k < D else 0.
	d.grid_scale = float(gscale)
	d.tau = float(tau)
	return d


def n_cells(D, dims):
	return int(math.prod(dims[:D]))
