"""
TEST INFRASTRUCTURE.  Records the call signatures of the reference's public surface on the hot path — every top-level function and
every class method of 3D/GSR.py, 3D/advance.py, 2D/GSR.py, 2D/advance.py — by parsing the files (ast: nothing is imported or run),
into tests/golden/ref_api_signatures.json.  tests/test_api_signatures.py holds the drop-in modules to it.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_api.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {'3D/GSR.py': 'gsr3d', '3D/advance.py': 'advance3d', '2D/GSR.py': 'gsr2d', '2D/advance.py': 'advance2d'}


def signature(fn):
	a = fn.args
	names = [x.arg for x in a.posonlyargs + a.args]
	defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
	return {'args': names, 'defaults': defaults, 'vararg': a.vararg.arg if a.vararg else None, 'kwarg': a.kwarg.arg if a.kwarg else None,
			'kwonly': [x.arg for x in a.kwonlyargs], 'line': fn.lineno,
			'taichi_kernel': any('ti.kernel' in ast.unparse(d) or 'ti.func' in ast.unparse(d) for d in fn.decorator_list)}


if __name__ == '__main__':
	out = {}
	for rel, mine in FILES.items():
		tree = ast.parse(open(os.path.join('/root/reference', rel)).read())
		mod = {'functions': {}, 'classes': {}}
		for node in tree.body:
			if isinstance(node, ast.FunctionDef):
				mod['functions'][node.name] = signature(node)
			elif isinstance(node, ast.ClassDef):
				mod['classes'][node.name] = {'bases': [ast.unparse(b) for b in node.bases],
											 'methods': {m.name: signature(m) for m in node.body if isinstance(m, ast.FunctionDef)}}
		out[rel] = {'mirror': mine, **mod}
		print(rel, len(mod['functions']), 'functions,', {k: len(v['methods']) for k, v in mod['classes'].items()})
	json.dump(out, open(os.path.join(HERE, 'ref_api_signatures.json'), 'w'), indent=1, sort_keys=True)
