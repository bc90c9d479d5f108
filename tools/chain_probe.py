"""development probe: marginal device time of each part of one S1 project iteration INSIDE the captured graph, by ablation —
the iteration is captured with one part replaced by a no-op and the replay time compared with the full iteration."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import timestep3d, gsr3d
gsr3d.device = torch.device('cuda', 0)
UNIT, REPS = 10, 30
for kv in filter(None, os.environ.get('GSR_TUNE', '').split(',')):	# e.g. GSR_TUNE=7=0,8=8192
	import ctypes
	from gaussian_fluids_code_b200 import _lib
	k, v = kv.split('=')
	assert _lib.lib().gsr_set_tuning(ctypes.c_int(int(k)), ctypes.c_int(int(v))) == 0
ts = timestep3d.LeapfrogTimestep(n=int(sys.argv[1]) if len(sys.argv) > 1 else 10, iters=20, test_res=32, check_iter=10)
ts.step(); ts.step()
torch.cuda.synchronize()
cur, new = ts.cur, ts.new
fp = ts._projector(new, cur)['fp']
e, ce = new._engine, cur._engine
real = {'fwd': e.forward, 'gather': e.backward_gather, 'rk4': ce.advected_vorticity, 'step': fp.stepper.step, 'rebuild': fp._rebuild, 'bin': e.bin_samples}
masks = {}
def gather_spy(*a, **k):
	r = real['gather'](*a, **k)
	masks[k.get('tag', 'main')] = r[1]
	return r
e.backward_gather = gather_spy
sample_fns = [lambda: ts._samples(fp), (lambda: ts._boundary(fp)) if os.environ.get('SEPARATE_BOUNDARY_GEN') else (lambda: ts._boundary_binned(fp))]
fp.set_samplers(*sample_fns); fp.prime(); fp.iterate(None)
torch.cuda.synchronize()
bins_cache = {}
def bin_spy(x, need_cells, tag='x'):
	b = real['bin'](x, need_cells, tag=tag)
	bins_cache[tag] = b
	return b
e.bin_samples = bin_spy
fp.iterate(None); torch.cuda.synchronize()
if not os.environ.get('SEPARATE_BOUNDARY_GEN'):
	bins_cache['pb'] = ts._boundary_binned(fp)[1]
tick = torch.zeros(8, device='cuda')

def variant(off, boundary=True):
	e.forward = (lambda *a, **k: None) if 'fwd' in off else real['fwd']
	e.backward_gather = (lambda *a, **k: (None, masks[k.get('tag', 'main')])) if 'gather' in off else real['gather']
	ce.advected_vorticity = (lambda *a, **k: None) if 'rk4' in off else real['rk4']
	fp.stepper.step = (lambda *a, **k: tick.add_(1.)) if 'step' in off else real['step']
	fp._rebuild = (lambda: None) if 'rebuild' in off else real['rebuild']
	if 'prep' in off:
		e.bin_samples = lambda x, need_cells, tag='x': bins_cache[tag]
		fns = [lambda: ts._x, (lambda: (ts._xb, ts._nb)) if os.environ.get('SEPARATE_BOUNDARY_GEN') else (lambda: ((ts._xb, ts._nb), bins_cache['pb']))]
	else:
		e.bin_samples = real['bin']
		fns = sample_fns
	ts.reset(); fp.restart()
	fp.set_samplers(fns[0], fns[1] if boundary else None)
	fp.prime()
	for _ in range(2):
		fp.iterate(None)
	torch.cuda.synchronize()
	g = torch.cuda.CUDAGraph()
	with torch.cuda.graph(g):
		for k in range(UNIT):
			fp.iterate(None, join_all=(k == UNIT - 1))
	for _ in range(3):
		g.replay()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(REPS):
		g.replay()
	e1.record(); torch.cuda.synchronize()
	return e0.elapsed_time(e1) * 1e3 / (REPS * UNIT)

out = {}
out['full'] = variant(())
out['no_boundary'] = variant((), boundary=False)
for part in ('rk4', 'step', 'rebuild', 'prep', 'fwd', 'gather'):
	out['no_' + part] = variant((part,))
out['only_step_rebuild'] = variant(('fwd', 'gather', 'rk4', 'prep'))
out['only_step'] = variant(('fwd', 'gather', 'rk4', 'prep', 'rebuild'))
out['only_rebuild'] = variant(('fwd', 'gather', 'rk4', 'prep', 'step'))
out['only_prep'] = variant(('fwd', 'gather', 'rk4', 'step', 'rebuild'))
out['only_fwd_gather'] = variant(('rk4', 'prep', 'step', 'rebuild'))
out['only_fwd'] = variant(('rk4', 'prep', 'step', 'rebuild', 'gather'))
out['nothing'] = variant(('fwd', 'gather', 'rk4', 'prep', 'step', 'rebuild'))
out['full_again'] = variant(())
variant(())	# restores the real functions
def loop_us(fn, reps=300):
	for _ in range(10):
		fn()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(reps):
		fn()
	e1.record(); torch.cuda.synchronize()
	return e0.elapsed_time(e1) * 1e3 / reps
print(json.dumps({k: round(v, 2) for k, v in out.items()}))
