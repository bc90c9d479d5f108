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
