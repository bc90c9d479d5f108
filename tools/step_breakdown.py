"""where a fixed-work time step spends its time (development tool): python tools/step_breakdown.py [n] [iters]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import timestep3d, gsr3d, advance3d, engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 600
gsr3d.device = torch.device('cuda', 0)
ts = timestep3d.LeapfrogTimestep(n=n, iters=iters)
acc = {}
def timed(name, fn):
	def w(*a, **k):
		if torch.cuda.is_current_stream_capturing():
			return fn(*a, **k)
		torch.cuda.synchronize(); t0 = time.perf_counter()
		r = fn(*a, **k)
		torch.cuda.synchronize(); acc.setdefault(name, []).append(time.perf_counter() - t0)
		return r
	return w
advance3d.FusedProjector.evaluate = timed('evaluate(lattice)', advance3d.FusedProjector.evaluate)
gsr3d.GaussianSplatting3DFast.gradient = timed('gradient(lattice)', gsr3d.GaussianSplatting3DFast.gradient)
advance3d.advect_covector_field = timed('advect', advance3d.advect_covector_field)
engine.HashEngine.bin_samples = timed('bin_samples', engine.HashEngine.bin_samples)
advance3d.clone_velocity_field = timed('clone', advance3d.clone_velocity_field)
advance3d._project_pipelined = timed('project (all of it)', advance3d._project_pipelined)
for name in ('run_iterations', 'restart', 'begin', 'evaluate_global'):
	setattr(timestep3d.ShardedProjector, name, timed(name, getattr(timestep3d.ShardedProjector, name)))
advance3d.FusedProjector.finish = timed('finish', advance3d.FusedProjector.finish)
advance3d.curl = timed('curl', advance3d.curl)
for rep in range(3):
	acc.clear()
	ts.reset()
	torch.cuda.synchronize(); t0 = time.perf_counter()
	ts.step()
	torch.cuda.synchronize(); tot = time.perf_counter() - t0
	out = {'rep': rep, 'total_ms': tot * 1e3}
	for k, v in acc.items():
		out[k] = {'calls': len(v), 'sum_ms': sum(v) * 1e3, 'max_ms': max(v) * 1e3}
	print(json.dumps(out))
