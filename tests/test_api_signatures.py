"""
The drop-in boundary, name by name (way A): every public function and class method of the reference's 3D/GSR.py,
3D/advance.py, 2D/GSR.py and 2D/advance.py (signatures recorded from the reference's sources by tests/golden/make_golden_api.py)
exists in the mirror module with the reference's positional parameters, in the reference's order, with the reference's defaults —
so a driver written against the reference calls the CUDA classes unchanged.  Mirrors may ADD keyword parameters after the
reference's (fused=, normals=, seed=, ...).  The Taichi kernels themselves (`*_ti`, @ti.kernel) are the replaced implementation, not
API: they are exempt.  CPU test: nothing here touches the GPU or the CUDA library.
"""
import ast
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
API = json.load(open(os.path.join(HERE, 'golden', 'ref_api_signatures.json')))
# helpers of the reference that are plotting / matplotlib only, outside the hot path's boundary (out of scope)
EXEMPT = {'3D/GSR.py': set(), '2D/GSR.py': {'show_field', 'draw_ellipses', 'generate_blue_noise'},
		  '3D/advance.py': set(), '2D/advance.py': set()}


def norm(src):
	"""default values compared as Python values where they are literals, else as source text"""
	if src is None:
		return None
	try:
		return ast.literal_eval(src)
	except (ValueError, SyntaxError):
		return src.replace(' ', '')


def check(ref, fn, where):
	sig = inspect.signature(fn)
	params = [p for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
	names = [p.name for p in params]
	want = ref['args']
	assert names[:len(want)] == want, f'{where}: parameters {names} do not start with the reference\'s {want}'
	for p, d in zip(params, ref['defaults']):
		if d is None:
			assert p.default is inspect.Parameter.empty, f'{where}: {p.name} has a default the reference does not have'
			continue
		assert p.default is not inspect.Parameter.empty, f'{where}: {p.name} lost its default {d}'
		got = p.default
		if isinstance(norm(d), str):	# an expression (np.log(...), device=...): only its presence is checked
			continue
		assert got == norm(d) or (isinstance(got, float) and got == pytest.approx(norm(d), rel=1e-12)), f'{where}: default of {p.name} is {got!r}, reference {d}'
	for p in params[len(want):]:
		assert p.default is not inspect.Parameter.empty, f'{where}: extra parameter {p.name} must be optional'


@pytest.mark.parametrize('rel', sorted(API))
def test_mirror_module_exposes_the_reference_signatures(rel):
	mod = importlib.import_module('gaussian_fluids_code_b200.' + API[rel]['mirror'])
	missing = []
	for name, ref in API[rel]['functions'].items():
		if name in EXEMPT[rel] or ref['taichi_kernel']:
			continue
		if not hasattr(mod, name):
			missing.append(name)
			continue
		check(ref, getattr(mod, name), f'{rel}:{ref["line"]} {name}')
	for cname, c in API[rel]['classes'].items():
		if not hasattr(mod, cname):
			missing.append(cname)
			continue
		cls = getattr(mod, cname)
		for mname, ref in c['methods'].items():
			if ref['taichi_kernel'] or mname.endswith('_ti'):
				continue
			if not hasattr(cls, mname):
				missing.append(f'{cname}.{mname}')
				continue
			check(ref, getattr(cls, mname), f'{rel}:{ref["line"]} {cname}.{mname}')
	assert not missing, f'{rel}: missing in {mod.__name__}: {missing}'
