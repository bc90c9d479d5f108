"""
CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every function that include/gsr_b200.h
declares (no compute call is made: there is no GPU here), the host-side mirrors of the reference's scalar formulas agree
with the oracle's restatement, and the product package refuses to run without CUDA instead of falling back.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
	src = open(os.path.join(ROOT, 'include', 'gsr_b200.h')).read()
	src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
	return sorted(set(re.findall(r'\b(gsr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
	from gaussian_fluids_code_b200 import _lib
	lib = _lib.lib()
	names = declared_functions()
	assert len(names) >= 25
	missing = [n for n in names if not hasattr(lib, n)]
	assert not missing, missing
	assert b'sm_100a' in lib.gsr_version()


def test_ctypes_structs_match_header_sizes():
	"""the ctypes mirrors must have the C layout: compile a tiny probe with the system compiler and compare sizeof"""
	import subprocess
	import tempfile
	from gaussian_fluids_code_b200 import _lib
	with tempfile.TemporaryDirectory() as d:
		c = os.path.join(d, 'p.c')
		open(c, 'w').write('#include <stdio.h>\n#include "gsr_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(gsr_grid_desc), sizeof(gsr_loss_cfg), sizeof(gsr_step_cfg), sizeof(gsr_loss_src));return 0;}\n')
		exe = os.path.join(d, 'p')
		cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
		subprocess.run([cc, '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
		sizes = [int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
	assert sizes == [C.sizeof(_lib.GridDesc), C.sizeof(_lib.LossCfg), C.sizeof(_lib.StepCfg), C.sizeof(_lib.LossSrc)]


def test_argument_validation_needs_no_gpu():
	"""bad arguments are rejected before any CUDA call (GSR_EINVAL = -1)"""
	from gaussian_fluids_code_b200 import _lib, host
	lib = _lib.lib()
	desc = host.make_desc(3, [-.2, 1.2] * 3, [7, 7, 7], .2, 5e-3)
	assert lib.gsr_forward(C.byref(desc), None, None, None, None, C.c_int64(10), None, None, None, None, None, C.c_int(0), None) == -1
	assert lib.gsr_build_grid(None, None, C.c_int64(1), None, None, None, None, None, None, None, None, None, None, C.c_size_t(0), None) == -1
	assert lib.gsr_set_tuning(C.c_int(99), C.c_int(0)) == -1
	bad = host.make_desc(3, [-.2, 1.2] * 3, [7, 7, 7], .2, 5e-3)
	bad.D = 4
	assert lib.gsr_tile_slots(C.byref(bad), C.c_int64(10)) == -1


def test_host_formulas_match_oracle():
	"""grid_size / extended domain / grid_scale / initial scaling (3D/GSR.py:160-177, 247-252) — host mirror vs oracle restatement"""
	from gaussian_fluids_code_b200 import host
	from oracle import oracle as orc
	for D, N in ((3, 1000), (3, 64000), (2, 576), (2, 5041)):
		bounds = (0., 1.) * D if D == 3 else (0., 10.) * D
		mgs = host.default_min_grid_scale(D, bounds, N)
		ext = host.extend(D, bounds, mgs)
		assert np.allclose(ext, orc.extended_bounds(D, bounds, mgs), rtol=0, atol=0)
		assert host.grid_size(D, ext, mgs) == orc.grid_dims(D, ext, mgs)
		for tau in (5e-3, 1e-3):
			s0 = host.initial_scaling(tau, mgs)
			assert np.float32(host.grid_scale(tau, s0, mgs, ext)) == np.float32(orc.grid_scale_of(tau, np.array([s0, s0 + .1]), mgs, ext))
			assert np.float32(host.grid_scale(tau, s0 - .3, mgs, ext)) == np.float32(orc.grid_scale_of(tau, np.array([s0, s0 - .3]), mgs, ext))


def test_no_cpu_fallback():
	import torch
	if torch.cuda.is_available():
		pytest.skip('CUDA present')
	from gaussian_fluids_code_b200 import _lib, engine
	with pytest.raises(_lib.GsrError):
		engine.HashEngine(3, 'cpu')
	with pytest.raises(_lib.GsrError):
		_lib.ptr(torch.zeros(3))


def test_obj_parser_of_the_mesh_sampler():
	"""host side of mesh_sampler.MeshSampler (3D/mesh_sampler.py:23-36): v / vn / f records, `a/b/c` and `a//c` face items, 1-based"""
	from gaussian_fluids_code_b200 import _lib
	from gaussian_fluids_code_b200.mesh_sampler import parse_obj
	text = '# comment\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvn 1 0 0\nvt 0.5 0.5\nf 1//1 2//1 3//1\nf 1/7/2 3/8/2 4/9/2\n'
	v, n, f, fn = parse_obj(text)
	assert v == [[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]] and n == [[0., 0., 1.], [1., 0., 0.]]
	assert f == [[0, 1, 2], [0, 2, 3]] and fn == [[0, 0, 0], [1, 1, 1]]
	with pytest.raises(_lib.GsrError):
		parse_obj('v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\n')	# quads: not supported, as in the reference
	with pytest.raises(_lib.GsrError):
		parse_obj('v 0 0 0\nf 1 1 1\n')	# no normals
