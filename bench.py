#!/usr/bin/env python
"""
bench.py — measures the Gaussian-Fluids hot path on B200 (contract: the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S1|S2|S3] [--iters 600] [--scaling weak|strong|both]
                  [--dim 3|2 --scene taylor_green|leapfrog|karman] [--workload timestep|density] [--impl ours|reference]

A "step" is one fixed-work 3D leapfrog time step run through the public API, advance3d.advance_frame: clone + advect +
`iters` project iterations (+ boundary passes) + a test pass every 100 iterations + the output-field pass (|vorticity| and divergence from one gradient).
`value` = Gaussian-sample pair evaluations per second, whole job (all ranks); `timesteps_per_s` rides along.
`e2e` is the same call with the Gaussian parameters coming from pinned HOST buffers each step and the updated parameters plus the
two output fields copied back to the host inside the timed region.
`roofline`: the dominant kernel timed live with CUDA events inside the timed region, plus `step_frac` (the whole step's
algorithmic flops over step time x FP32 peak), `kernels` (every kernel class of the step timed in isolation on the step's own
inputs: algorithmic flops or bytes, achieved rate, fraction of its roofline) and `hbm_kernels` (hash build + pack, sample binning
and the fused optimiser step at N = 10^6 against the measured HBM bandwidth).
`--impl reference` times the CPU restatement of the reference's kernels (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'gaussian_sample_pair_evals_per_s'
UNIT = 'pair-evals/s'
SIZES = {'S1': 10, 'S2': 40, 'S3': 100, 'S4': 160, 'S5': 256}


def parse():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=3)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', type=str, default='ours')
	ap.add_argument('--size', type=str, default='S1', help='S1 = 10^3 Gaussians (the reference\'s 3D leapfrog size), S2 = 40^3, S3 = 100^3')
	ap.add_argument('--iters', type=int, default=600, help='project iterations per time step (reference minimum: 600)')
	ap.add_argument('--test-res', type=int, default=128)
	ap.add_argument('--scaling', type=str, default='both', choices=('weak', 'strong', 'both'),
					help='multi-GPU: weak = samples and a world-times-finer lattice sharded (the headline line); strong = the fixed frame, its lattice shared; '
						 'both = the weak line with the strong numbers nested (default)')
	ap.add_argument('--dim', type=int, default=3)
	ap.add_argument('--scene', type=str, default='taylor_green', help='2D scene (BASELINE configs 1-3): taylor_green | leapfrog | karman')
	ap.add_argument('--workload', type=str, default='timestep', choices=('timestep', 'density'))
	ap.add_argument('--density-res', type=int, default=512)
	ap.add_argument('--no-cpu-baseline', action='store_true')
	ap.add_argument('--no-unhoisted', action='store_true', help='skip the comparison run with the test reference recomputed at every test pass')
	ap.add_argument('--no-kernel-table', action='store_true')
	ap.add_argument('--no-graph', action='store_true', help='run the project iterations eagerly instead of replaying a CUDA graph')
	return ap.parse_args()


def workload_name(args, n):
	return (f'3D leapfrog fixed-work timestep through advance3d.advance_frame, N={n ** 3} Gaussians ({args.size}), Q=N samples/rank/iter, {args.iters} project iters, '
			f'boundary 8192/rank, test lattice {args.test_res}^3')


def load_synth():
	spec = __import__('importlib.util').util.spec_from_file_location('synth_b200', os.path.join(ROOT, 'gaussian-fluids-code_b200', 'synth.py'))
	synth = __import__('importlib.util').util.module_from_spec(spec)
	spec.loader.exec_module(synth)
	return synth


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C restatement of the reference's kernels), all host threads, bounded samples
# ---------------------------------------------------------------------------------------------------------

def cpu_sample(n, nthreads, lattice_points=16384, reps=1, seed=42):
	"""
	B1 — one bounded sample of the timestep's kernel work on the CPU: {RK4 pull-back (5 evals) + forward + backward with the
	project weights on Q = N samples; boundary forward + backward on 8192 samples; RK4 pull-back + forward on
	`lattice_points` lattice points}.  Returns (candidate visits, seconds).
	"""
	import numpy as np
	import torch
	from oracle.oracle import OracleGSR, extended_bounds
	P, S, R, V, mgs, gen = load_synth().synthetic_field(n, seed)
	N = n ** 3
	ext = extended_bounds(3, (0., 1.) * 3, mgs)
	orc = OracleGSR(3, ext, P, S, R, V, 5e-3, mgs, precision='f32', nthreads=nthreads)
	x = torch.rand((N, 3), generator=gen).numpy()
	xb = torch.rand((8192, 3), generator=gen).numpy()
	face = np.arange(8192) % 6
	xb[np.arange(8192), face // 2] = (face % 2).astype(np.float32)
	nb = np.zeros((8192, 3), np.float32)
	nb[np.arange(8192), face // 2] = 1. - 2. * (face % 2)
	xl = torch.rand((lattice_points, 3), generator=gen).numpy()
	C_x, C_b, C_l = orc.count_candidates(x), orc.count_candidates(xb), orc.count_candidates(xl)
	cands = reps * (7 * C_x + 2 * C_b + 6 * C_l)
	t0 = time.perf_counter()
	for _ in range(reps):
		_, dpsi, pv, pdv = orc.rk4(x, -.02, pos_only=False)
		pb_vor = np.stack((pdv[:, 2, 1] - pdv[:, 1, 2], pdv[:, 0, 2] - pdv[:, 2, 0], pdv[:, 1, 0] - pdv[:, 0, 1]), -1)
		ref_hel = (pv * pb_vor).sum(-1)
		ref_vor = np.linalg.solve(dpsi, pb_vor[..., None])[..., 0]
		val, grad = orc.forward(x)
		orc.backward3d(x, val, grad, ref_vor=ref_vor, weight_vor=1., ref_hel=ref_hel, weight_hel=1., weight_div=1.,
					   direct=orc.zero_grads(), vor=orc.zero_grads(), div=orc.zero_grads())
		valb, gradb = orc.forward(xb)
		orc.backward3d(xb, valb, gradb, normals=nb, weight_boundary=10.)
		orc.rk4(xl, -.02, pos_only=False)
		orc.forward(xl)
	return cands, time.perf_counter() - t0


def cpu_dense_torch(n, nthreads, Q=2048, seed=42):
	"""
	B2 — the reference's own DENSE torch path (GaussianSplatting3D.__call__ / gradient, 3D/GSR.py:93-130: every Gaussian at every
	point, no truncation, no hash), restated in oracle/oracle.py, on the CPU with torch's threads; N, Q <= 4096 (O(N Q) memory).
	Returns (pairs evaluated, seconds) for value + gradient.
	"""
	import torch
	from oracle.oracle import dense_torch_value_gradient
	P, S, R, V, mgs, gen = load_synth().synthetic_field(min(n, 16), seed)
	torch.set_num_threads(nthreads)
	x = torch.rand((Q, 3), generator=gen)
	args = [torch.tensor(a) for a in (P, S, R, V)]
	dense_torch_value_gradient(*args, x[:64])
	t0 = time.perf_counter()
	dense_torch_value_gradient(*args, x)
	return Q * args[0].shape[0], time.perf_counter() - t0


def cpu_whole_step(n, nthreads, iters, test_res, seed=42, sample_iters=3, lattice_points=131072):
	"""
	B3 — the whole fixed-work time step on the CPU, assembled from measured pieces: `sample_iters` full project iterations of
	oracle.OracleProjector3D (pull-back reference, forward, backward, PCGrad, regularisers, boundary pass, Adam x4 + schedulers,
	hash rebuild: 3D/advance.py:183-287) and one test pass on `lattice_points` lattice points, scaled to `iters` iterations and
	iters/100 + 2 lattice passes of test_res^3 points.  Returns seconds per time step and what was measured.
	"""
	import numpy as np
	import torch
	from oracle.oracle import OracleGSR, OracleProjector3D, extended_bounds
	P, S, R, V, mgs, gen = load_synth().synthetic_field(n, seed)
	N = n ** 3
	ext = extended_bounds(3, (0., 1.) * 3, mgs)
	prev = OracleGSR(3, ext, P, S, R, V, 5e-3, mgs, precision='f32', nthreads=nthreads)
	proj = OracleProjector3D((0., 1.) * 3, (P, S, R, V), prev, .02, 10., 5e-3, mgs, precision='f32', nthreads=nthreads)
	xb = torch.rand((8192, 3), generator=gen).numpy()
	face = np.arange(8192) % 6
	xb[np.arange(8192), face // 2] = (face % 2).astype(np.float32)
	nb = np.zeros((8192, 3), np.float32)
	nb[np.arange(8192), face // 2] = 1. - 2. * (face % 2)
	proj.iterate(torch.rand((N, 3), generator=gen).numpy(), (xb, nb))	# warm-up
	t0 = time.perf_counter()
	for _ in range(sample_iters):
		proj.iterate(torch.rand((N, 3), generator=gen).numpy(), (xb, nb))
	t_iter = (time.perf_counter() - t0) / sample_iters
	xl = torch.rand((lattice_points, 3), generator=gen).numpy()
	t0 = time.perf_counter()
	prev.rk4(xl, -.02, pos_only=False)
	prev.forward(xl)
	t_test = (time.perf_counter() - t0) * (test_res ** 3 / lattice_points)
	t0 = time.perf_counter()
	prev.forward(xl)
	t_field = (time.perf_counter() - t0) * (test_res ** 3 / lattice_points)
	sec = iters * t_iter + (iters // 100) * t_test + 2 * t_field
	return sec, {'iteration_s': t_iter, 'test_pass_s': t_test, 'field_pass_s': t_field}


def cpu_baseline(args, n, N):
	cores = os.cpu_count() or 1
	lattice_points = 131072 if n <= 40 else 16384
	cpu_sample(n, cores, lattice_points=2048)
	c, tsec = cpu_sample(n, cores, lattice_points=lattice_points, reps=2 if n <= 40 else 1)
	out = {'value': c / tsec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
		   'sample': f'B1: {2 if n <= 40 else 1} x (RK4 pull-back + fwd + bwd on Q=N={N}, boundary fwd+bwd on 8192, RK4 pull-back + fwd on {lattice_points} lattice points), oracle/ f32 OpenMP'}
	try:
		pairs, t2 = cpu_dense_torch(n, cores)
		out['B2_dense_torch'] = {'value': pairs / t2, 'unit': 'dense pairs/s (value + gradient, no truncation)', 'cores': cores,
								 'sample': 'the reference\'s dense torch class (3D/GSR.py:93-130, restated in oracle/oracle.py) on the CPU: 2048 points x min(N, 4096) Gaussians'}
	except Exception as ex:	# the dense path is a side note: never lose the line over it
		out['B2_dense_torch'] = {'unavailable': f'{type(ex).__name__}: {ex}'}
	if n <= 40:
		sec, parts = cpu_whole_step(n, cores, args.iters, args.test_res)
		out['B3_whole_step'] = {'timesteps_per_s': 1. / sec, 's_per_step': sec, 'cores': cores, 'measured': parts,
								'sample': f'3 full project iterations (oracle.OracleProjector3D, f32 kernels) + one test pass on 131072 lattice points, scaled to {args.iters} '
										  f'iterations and {args.iters // 100} + 2 lattice passes of {args.test_res}^3 points'}
	return out


def run_reference(args):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	n = SIZES[args.size]
	cores = os.cpu_count() or 1
	cpu_sample(n, cores, lattice_points=2048)	# warm-up (builds the oracle, pages memory in)
	lattice_points = 131072 if n <= 40 else 16384
	for _ in range(max(args.warmup - 1, 0)):
		cpu_sample(n, cores, lattice_points=lattice_points)
	t_tot, c_tot = 0., 0
	for _ in range(args.steps):
		c, t = cpu_sample(n, cores, lattice_points=lattice_points)
		c_tot += c
		t_tot += t
	v = c_tot / t_tot
	sample = f'per step: RK4 pull-back + fwd + bwd on Q=N={n ** 3} samples, boundary fwd+bwd on 8192, RK4 pull-back + fwd on {lattice_points} lattice points'
	line = {
		'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
		'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': workload_name(args, n), 'note': 'reference Taichi kernels cannot run (taichi not installed): CPU restatement oracle/ (C, OpenMP), bounded sample'},
		'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
		'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
	}
	if n <= 40:
		sec, parts = cpu_whole_step(n, cores, args.iters, args.test_res)
		line['timesteps_per_s'] = 1. / sec
		line['cpu_baseline']['B3_whole_step'] = {'timesteps_per_s': 1. / sec, 's_per_step': sec, 'measured': parts}
	print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
	"""samples SM clock / throttle reasons through NVML while the timed region runs"""

	def __init__(self, index):
		super().__init__(daemon=True)
		self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

	def run(self):
		try:
			import pynvml
			pynvml.nvmlInit()
			h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
			self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
			names = {pynvml.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown', pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
					 pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown', pynvml.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap'}
			while not self.stop_flag:
				self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
				r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
				for bit, nm in names.items():
					if r & bit:
						self.reasons.add(nm)
				time.sleep(.05)
		except Exception as ex:	# NVML not usable: report that instead of clocks
			self.reasons.add(f'nvml_unavailable:{type(ex).__name__}')

	def summary(self):
		import statistics
		return {'sm_mhz': statistics.median(self.samples) if self.samples else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def timeit(fn, reps=10, warm=2):
	"""median CUDA-event time of fn() in ms on the current stream"""
	import statistics
	import torch
	for _ in range(warm):
		fn()
	torch.cuda.synchronize()
	evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
	for a, b in evs:
		a.record()
		fn()
		b.record()
	torch.cuda.synchronize()
	return statistics.median(a.elapsed_time(b) for a, b in evs)


def timeit_graph(fn, reps=20, warm=2):
	"""GPU time of one fn() in ms with the host out of the way: `reps` back-to-back calls captured in one CUDA graph, one replay
	timed with events (tens-of-microsecond kernels would otherwise be timed as their eager launch overhead)"""
	import torch
	s = torch.cuda.Stream()
	s.wait_stream(torch.cuda.current_stream())
	with torch.cuda.stream(s):
		for _ in range(warm):
			fn()
	torch.cuda.current_stream().wait_stream(s)
	torch.cuda.synchronize()
	g = torch.cuda.CUDAGraph()
	from gaussian_fluids_code_b200.graphloop import capture_guard
	with capture_guard(), torch.cuda.graph(g):
		for _ in range(reps):
			fn()
	g.replay()
	torch.cuda.synchronize()
	best = None
	for _ in range(3):
		a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		a.record()
		g.replay()
		b.record()
		torch.cuda.synchronize()
		t = a.elapsed_time(b) / reps
		best = t if best is None else min(best, t)
	return best


def kernel_table(ts, args, peaks, step_ms):
	"""
	Every kernel class of the step, timed in isolation (20 back-to-back calls inside one CUDA graph) on the step's own inputs, with its algorithmic
	work (24 flop per candidate visit + 28 per accepted pair forward, + 115 per accepted pair and accumulator set
	backward; bytes for the hash / optimiser kernels) and the fraction of its roofline.  `per_step` = launches per time step;
	share = per_step x isolated time / step time (the pipelined step overlaps the small kernels, so shares need not sum to 1).
	"""
	import torch
	from gaussian_fluids_code_b200 import advance3d
	cur, new = ts.cur, ts.new
	e, ce = new._engine, cur._engine
	for f in (cur, new):
		f._engine.ensure_packed(f._params())
	N, Qb, iters = ts.N, ts.Qb, ts.iters
	fma, hbm = peaks['fp32_tflops_live'] * 1e12, (peaks.get('hbm_gbs') or 6650.) * 1e9
	dev = cur.positions.device
	gen = torch.Generator(device=dev).manual_seed(1)
	x = torch.rand((N, 3), device=dev, generator=gen)
	xb, nb = advance3d.BoxSurfaceSampler(0., 1., 0., 1., 0., 1.)(Qb)
	lat = ts.lattice

	def census(engine, pts):
		c = torch.zeros(2, dtype=torch.int64, device=dev)
		engine.count_pairs(pts, c, 1, True)
		return [int(v) for v in c.tolist()]

	rows = []

	def row(name, per_step, ms, flop=None, nbytes=None, note=None):
		r = {'kernel': name, 'per_step': per_step, 'ms': ms, 'share_of_step': per_step * ms / step_ms}
		if flop is not None:
			r.update(bound='fp32', algorithmic_flop=flop, achieved_tflops=flop / (ms * 1e-3) / 1e12, frac=flop / (ms * 1e-3) / fma)
		if nbytes is not None:
			r.update(bound='hbm', algorithmic_bytes=nbytes, achieved_gbs=nbytes / (ms * 1e-3) / 1e9, frac=nbytes / (ms * 1e-3) / hbm)
		if note:
			r['note'] = note
		rows.append(r)

	# lattice passes
	Cl, Pl = census(ce, lat)
	Ql = lat.shape[0]
	rv, rh = torch.empty((Ql, 3), device=dev), torch.empty((Ql,), device=dev)
	bins_l = ce.bin_samples(lat, False)
	row('RK4 pull-back on the test lattice (gsr_advected_vorticity, 5 evaluations)', 1 if advance3d.HOIST_TEST_REFERENCE else iters // 100, timeit_graph(lambda: ce.advected_vorticity(lat, -ts.dt, rv, rh, perm=bins_l)),
		flop=5 * (24 * Cl + 28 * Pl))
	val_l, grad_l = torch.empty((Ql, 3), device=dev), torch.empty((Ql, 3, 3), device=dev)
	row('forward u + grad u on the test lattice (gsr_forward)', iters // 100 + 1, timeit_graph(lambda: ce.forward(lat, val_l, grad_l, False, perm=bins_l)), flop=24 * Cl + 28 * Pl)
	del rv, rh, val_l, grad_l
	# training batch Q = N
	Cx, Px = census(e, x)
	bins = e.bin_samples(x, True)
	rv, rh = torch.empty((N, 3), device=dev), torch.empty((N,), device=dev)
	val, grad = torch.empty((N, 3), device=dev), torch.empty((N, 3, 3), device=dev)
	row('RK4 pull-back reference, Q = N (gsr_advected_vorticity)', iters, timeit_graph(lambda: ce.advected_vorticity(x, -ts.dt, rv, rh, perm=bins)), flop=5 * (24 * Cx + 28 * Px))
	row('forward, Q = N (gsr_forward)', iters, timeit_graph(lambda: e.forward(x, val, grad, False, perm=bins)), flop=24 * Cx + 28 * Px)
	row('backward adjoint + gather, Q = N, vor/hel + div sets (gsr_backward_gather)', iters,
		timeit_graph(lambda: e.backward_gather(x, bins.perm, bins.scs, val, grad, (0., 0., 0., 1., 1., 1.), {'ref_vor': rv, 'ref_hel': rh}, None, tag='kt')),
		flop=24 * Cx + 230 * Px)
	# boundary batch
	Cb, Pb = census(e, xb)
	bins_b = e.bin_samples(xb, True, tag='ktb')
	valb = torch.empty((Qb, 3), device=dev)
	row('boundary forward, 8192 samples (gsr_forward, value only)', iters, timeit_graph(lambda: e.forward(xb, valb, None, False, perm=bins_b)), flop=24 * Cb + 7 * Pb)
	row('boundary adjoint + gather, 8192 samples (gsr_backward_gather, direct set)', iters,
		timeit_graph(lambda: e.backward_gather(xb, bins_b.perm, bins_b.scs, valb, None, (0., 10., 0., 0., 0., 0.), {'normals': nb}, None, tag='ktb')), flop=24 * Cb + 115 * Pb)
	# sample generation + ordering, hash, optimiser step
	xs = torch.empty((N, 3), device=dev)
	it = torch.zeros(1, device=dev)
	row('training samples: draw + order (gsr_sample_box + gsr_bin_samples)', iters, timeit_graph(lambda: (e.sample_box((0., 1.) * 3, xs, 42, 0, it), e.bin_samples(xs, True, tag='kts'))),
		nbytes=12 * N + 12 * N + 4 * N + 4 * (e.lib.gsr_padded_cells(__import__('ctypes').byref(e.desc)) + 1), note='latency bound at this size')
	fp = next(iter(new.__dict__.get('_pipelines', {}).values()), None) or next(iter(cur.__dict__.get('_pipelines', {}).values()), None)
	if fp is not None:
		g = fp.gv
		acc = torch.zeros((3, N, 12), device=dev)
		lp = torch.zeros((fp.nblk, 8), device=dev)
		srcs = [(lp, fp.nblk, [1. / N, 0., 1. / N, 0., 0., 0., 0., 0.])]
		params = [p.detach().clone() for p in g._params()]

		def step():
			for p, q in zip(g._params(), params):
				p.detach().copy_(q)
			fp.stepper.step([p.detach() for p in g._params()], acc, 7, loss_srcs=srcs, rebuild=True)
		t_copy = timeit_graph(lambda: [p.detach().copy_(q) for p, q in zip(g._params(), params)])
		fp.stepper.init(g.scalings, keep_clock=False)
		row('fused optimiser step + hash rebuild + pack (gsr_step_rebuild)', iters, max(timeit_graph(step) - t_copy, 1e-4),
			nbytes=(312 + 48 * 3) * N + 16 * N + 8 * e.ncell + 104 * N, note='latency bound at this size' if N <= 4096 else None)
		fp.stepper.init(g.scalings, keep_clock=False)
		g.zero_grad()
	rows.sort(key=lambda r: -r['share_of_step'])
	return rows


def hbm_kernels(peaks, n=100):
	"""the HBM-bound stages at N = n^3 = 10^6 (north star: achieved GB/s for the sort and optimiser kernels): hash build + pack,
	sample binning, fused step; algorithmic bytes as SURVEY 8d states them"""
	import ctypes as C
	import torch
	from gaussian_fluids_code_b200 import advance3d, engine
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	hbm = (peaks.get('hbm_gbs') or 6650.) * 1e9
	P, S, R, V, mgs, _ = synthetic_field(n)
	o = make_fast3d(P, S, R, V, 5e-3, mgs)
	e = o._engine
	N = o.N
	dev = o.positions.device
	out = []

	def row(name, ms, nbytes, formula):
		out.append({'kernel': name, 'N': N, 'ms': ms, 'algorithmic_bytes': nbytes, 'formula': formula, 'achieved_gbs': nbytes / (ms * 1e-3) / 1e9,
					'frac_of_measured_hbm': nbytes / (ms * 1e-3) / hbm})
	params = [p.detach() for p in o._params()]
	row('gsr_build_grid + pack (radix sort of the cell keys, cell table, packed records)', timeit(lambda: e.build(o.positions.detach(), params=params)),
		12 * N + 4 * N + 8 * e.ncell + 52 * N + 52 * N, 'read positions 12N + write sorted_id 4N + 8 cells; pack: read 52N + write 52N')
	x = torch.rand((N, 3), device=dev)
	pcell = e.lib.gsr_padded_cells(C.byref(e.desc))
	row('gsr_bin_samples, Q = N (radix sort of the sample keys + cell table)', timeit(lambda: e.bin_samples(x, True, tag='hb')), 12 * N + 4 * N + 4 * (pcell + 1),
		'read samples 12Q + write perm 4Q + 4 (padded cells + 1)')
	st = engine.FusedStepper(e, [3e-4, 1e-5, 3e-4, 1e-5], 50, 10., 10., tau=o.clamp_threshold, min_grid_scale=o.min_grid_scale, ext_bounds=o._ext())
	st.init(o.scalings)
	acc = torch.zeros((3, N, 12), device=dev)
	acc.normal_(0., 1e-6)
	lp = torch.zeros((e.lib.gsr_loss_blocks(N), 8), device=dev)
	srcs = [(lp, lp.shape[0], [1. / N, 0., 1. / N, 0., 0., 0., 0., 0.])]
	row('gsr_step (chain rule of 2 sets, PCGrad dots, regularisers, Adam x4, scheduler, min s): 4 launches', timeit(lambda: st.step(params, acc, 6, loss_srcs=srcs)),
		(312 + 48 * 2) * N, '(312 + 48 sets) N: params 52N + moments 104N read and written, accumulators 48N per set read')
	del st, acc, o
	torch.cuda.empty_cache()
	return out


def host_overheads(ts):
	"""what the Python / ctypes layer above the C ABI costs on the host, next to the device time of the same call: an empty C-ABI call,
	and one public-API evaluation `field.gradient(x, need_val=True)` at Q = N (bin + forward: 2-3 C-ABI calls, a few torch allocations)"""
	import time
	import torch
	from gaussian_fluids_code_b200 import _lib
	lib = _lib.lib()
	n = 200000
	t0 = time.perf_counter()
	for _ in range(n):
		lib.gsr_launch_count()
	call_ns = (time.perf_counter() - t0) / n * 1e9
	gv = ts.cur
	x = torch.rand((gv.N, 3), device=gv.positions.device)
	for _ in range(5):
		gv.gradient(x, need_val=True)
	torch.cuda.synchronize()
	reps = 200
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	t0 = time.perf_counter()
	e0.record()
	for _ in range(reps):
		gv.gradient(x, need_val=True)
	e1.record()
	host_us = (time.perf_counter() - t0) / reps * 1e6	# time to ENQUEUE (no synchronisation inside the loop)
	torch.cuda.synchronize()
	dev_us = e0.elapsed_time(e1) / reps * 1e3
	return {'ctypes_empty_call_ns': call_ns, 'api_gradient_host_us_per_call': host_us, 'api_gradient_device_us_per_call': dev_us, 'Q': int(gv.N),
			'note': 'eager public-API call: the host enqueue time bounds the eager path at small N; the captured project() pipeline pays it once, at capture'}


def timed_steps(ts, args, barrier, dev, world, dist, host_params=None, host_out=None, host_fields=None):
	"""(ms for args.steps steps, max over ranks); with host buffers the copies are inside the timed region (the e2e number)"""
	import torch
	barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(args.steps):
		t_host = time.perf_counter()
		if host_params is None:
			ts.reset()
			ts.step()
		else:
			ts.reset([p.to(dev, non_blocking=True) for p in host_params])
			vor, div = ts.step()
			for dst, src in zip(host_out, (ts.cur.positions, ts.cur.scalings, ts.cur.rotations, ts.cur.values)):
				dst.copy_(src.detach(), non_blocking=True)
			host_fields[0].copy_(vor, non_blocking=True)
			host_fields[1].copy_(div, non_blocking=True)
			torch.cuda.synchronize()	# the caller owns the results only once they are on the host
			if os.environ.get('GSR_BENCH_TRACE'):
				print(f'[e2e step] {1e3 * (time.perf_counter() - t_host):.1f} ms host wall', file=sys.stderr, flush=True)
	e1.record()
	barrier()
	t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	return float(t.item())


def run_ours(args):
	import ctypes as C
	import torch
	import torch.distributed as dist
	rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
	if not torch.cuda.is_available():
		raise SystemExit('bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)')
	torch.cuda.set_device(local)
	if world > 1:
		dist.init_process_group('nccl', device_id=torch.device('cuda', local))
	from gaussian_fluids_code_b200 import _lib, advance3d, gsr3d, timestep3d
	gsr3d.device = torch.device('cuda', local)
	lib = _lib.lib()
	n = SIZES[args.size]
	dev = gsr3d.device

	def barrier():
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	def launches(tss):
		return lib.gsr_launch_count() + sum(t.graph_launches for t in tss)

	main_mode = 'strong' if (args.scaling == 'strong' and world > 1) else 'weak'
	ts = timestep3d.LeapfrogTimestep(n=n, iters=args.iters, test_res=args.test_res, rank=rank, world=world, use_graph=not args.no_graph, scaling=main_mode)
	N = ts.N
	# host-resident copies of the parameters (pinned) for the e2e path, and pinned landing buffers for its outputs
	host_params = [torch.as_tensor(a).clone().pin_memory() for a in ts.params0]
	host_out = [torch.empty_like(p).pin_memory() for p in host_params]
	host_fields = [torch.empty(ts.lattice.shape[0], dtype=torch.float32).pin_memory() for _ in range(2)]

	skipped_job = [0]

	def census_of(t):
		"""job totals (C, P): weak — every rank's visits; strong — one copy of the replicated training visits + all lattice shares"""
		cen = timestep3d.Census(dev)
		t.reset()
		t.step(cen)
		C_all, P_all = cen.value()
		C_lat, P_lat = cen.lattice_value()
		C_skip, _ = cen.skipped_value()	# lattice visits of the reference's algorithm this engine does not make (never part of C_all)
		if world == 1:
			skipped_job[0] = C_skip
			return C_all, P_all
		tot = torch.tensor([C_all, P_all, C_lat, P_lat, C_skip], dtype=torch.int64, device=dev)
		dist.all_reduce(tot)
		skipped_job[0] = int(tot[4])
		if t.scaling == 'weak':
			return int(tot[0]), int(tot[1])
		return C_all - C_lat + int(tot[2]), P_all - P_lat + int(tot[3])

	# ---- warm-up; the work census (candidate visits, accepted pairs) is taken on the last warm-up step --------------
	W = max(args.warmup, 3)
	for _ in range(W - 1):
		ts.reset()
		ts.step()
	C_job, P_job = census_of(ts)
	C_skipped = skipped_job[0]

	# ---- timed region: device-resident inputs -------------------------------------------------------------------
	sampler = ClockSampler(local)
	sampler.start()
	probe = []
	ts.probe = probe
	l0 = launches([ts])
	ms = timed_steps(ts, args, barrier, dev, world, dist)
	n_launches = launches([ts]) - l0
	ts.probe = None
	sampler.stop_flag = True
	sampler.join()
	value = C_job * args.steps / (ms * 1e-3)

	# ---- end-to-end: host buffers in, host buffers out, copies inside the timed region -----------------------------
	import argparse
	# untimed warm-up of the end-to-end path, like the W steps above: first use of the pinned buffers and copy paths, and a one-off host
	# hiccup (50-70 ms, only in the first processes on a fresh box, always in the third end-to-end step: profiles/README.md)
	timed_steps(ts, argparse.Namespace(steps=3), barrier, dev, world, dist, host_params, host_out, host_fields)
	ms_e2e = timed_steps(ts, args, barrier, dev, world, dist, host_params, host_out, host_fields)
	h2d = sum(p.numel() * 4 for p in host_params)
	d2h = sum(p.numel() * 4 for p in host_out) + sum(f.numel() * 4 for f in host_fields)

	# ---- the same steps with the test lattice's pull-back reference recomputed at every test pass, as the reference does (and as
	# this engine did until the hoist): same results, the round-1 work mix — kept beside the headline for comparison -------------
	unhoisted = None
	if advance3d.HOIST_TEST_REFERENCE and not args.no_unhoisted:
		advance3d.HOIST_TEST_REFERENCE = False
		try:
			ts.reset()
			ts.step()
			C_u, P_u = census_of(ts)
			ms_u = timed_steps(ts, args, barrier, dev, world, dist)
		finally:
			advance3d.HOIST_TEST_REFERENCE = True
		unhoisted = {'what': 'GSR_HOIST_TEST_REFERENCE=0: the RK4 pull-back reference on the fixed test lattice evaluated at each of the test passes of a frame '
							 '(3D/advance.py:290-291 recomputes it) instead of once per frame; bitwise the same losses and fields',
					 'ms_per_step': ms_u / args.steps, 'timesteps_per_s': args.steps / (ms_u * 1e-3), 'value': C_u * args.steps / (ms_u * 1e-3), 'unit': UNIT,
					 'pair_evals_per_step': C_u, 'accepted_pairs_per_step': P_u}

	# ---- the other scaling mode (N > 1): the fixed frame with its lattice shared between the ranks --------------------
	other = None
	if world > 1 and args.scaling == 'both':
		ts2 = timestep3d.LeapfrogTimestep(n=n, iters=args.iters, test_res=args.test_res, rank=rank, world=world, use_graph=not args.no_graph, scaling='strong')
		for _ in range(2):
			ts2.reset()
			ts2.step()
		C2, P2 = census_of(ts2)
		ms2 = timed_steps(ts2, args, barrier, dev, world, dist)
		other = {'scaling': 'strong', 'what': f'the fixed frame: N={N} training replicated on every rank (no exchange), the {args.test_res}^3 test / output lattice split over the ranks by z planes',
				 'ms_per_step': ms2 / args.steps, 'timesteps_per_s': args.steps / (ms2 * 1e-3), 'value': C2 * args.steps / (ms2 * 1e-3), 'unit': UNIT,
				 'pair_evals_per_step': C2}
		del ts2

	def finish():
		"""release the captured graphs and the peer-memory exchange before the process group goes away; if the teardown does not
		return (NCCL's destroy has been seen to block with communicator kernels captured in live graphs), leave without it —
		every rank has synchronised and rank 0 has printed by then"""
		sys.stdout.flush()
		sys.stderr.flush()
		if world > 1:
			import gc
			for f in (ts.cur, ts.new):
				f.__dict__.pop('_pipelines', None)
			gc.collect()
			torch.cuda.synchronize()
			dist.barrier()
			torch.cuda.synchronize()
			done = threading.Event()

			def destroy():
				try:
					dist.destroy_process_group()
				finally:
					done.set()
			threading.Thread(target=destroy, daemon=True).start()
			if not done.wait(20.):
				print('[bench] destroy_process_group did not return in 20 s: leaving without it', file=sys.stderr, flush=True)
				os._exit(0)

	if rank != 0:
		finish()
		return

	# ---- roofline of the dominant kernel: the RK4 pull-back on the test lattice ---------------------------------------------
	fma, mufu = C.c_double(0.), C.c_double(0.)
	lib.gsr_peak_fma(C.c_int(20000), C.byref(fma), _lib.stream())
	lib.gsr_peak_mufu(C.c_int(20000), C.byref(mufu), _lib.stream())
	peaks = {}
	try:
		peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
	except Exception:
		pass
	mpeaks = {'hbm_gbs': peaks.get('hbm_gbs'), 'hbm_source': 'MEASURED_PEAKS.json' if peaks.get('hbm_gbs') else 'fallback 6650 GB/s (B200_PROFILING.md)',
			  'fp32_tflops_live': fma.value, 'mufu_tops_live': mufu.value}
	cl = torch.zeros(2, dtype=torch.int64, device=dev)
	ts.cur._engine.ensure_packed(ts.cur._params())
	ts.cur._engine.count_pairs(ts.lattice, cl, 1, True)
	C_lat, P_lat = [int(v) for v in cl.tolist()]
	kernel_ms = [a.elapsed_time(b) for a, b in probe]
	k_ms = sum(kernel_ms) / max(len(kernel_ms), 1)
	flop_per_launch = 5 * (24 * C_lat + 28 * P_lat)
	achieved = flop_per_launch / (k_ms * 1e-3) / 1e12 if kernel_ms else None
	traffic = None
	try:	# DRAM bytes per launch of this kernel from a committed `ncu --set full` capture of this workload, if there is one
		tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
		traffic = tr.get(f'{args.size}/{args.test_res}/{world}', {}).get('dram_bytes_per_launch')
	except Exception:
		pass
	step_flop = 24 * C_job + 28 * P_job
	roofline = {'bound': 'fp32', 'kernel': ('rk4 pull-back of the previous field on the test lattice (gsr_advected_vorticity: rk4_tiled3s_kernel<2> at S1, 5 field evaluations per point): '
											'the largest single launch and the FP32-bound kernel of the step; share_of_step says how much of the step it is (one launch per frame since the '
											'test reference is hoisted) — most of the step is the latency-bound iteration cycle, see kernels[] and step_frac'),
				'achieved': achieved, 'peak': fma.value, 'unit': 'TFLOP/s', 'frac': (achieved / fma.value) if achieved else None,
				'traffic': traffic, 'traffic_source': 'profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of a committed ncu --set full capture of this kernel on this workload); null when this workload was not captured',
				'peak_source': 'FP32 FFMA peak measured live by gsr_peak_fma on this GPU (MEASURED_PEAKS.json holds only HBM and bf16 peaks); nominal 74.4 at 1965 MHz',
				'mufu_peak_Tops': mufu.value, 'mufu_achieved_Tops': (5 * P_lat / (k_ms * 1e-3) / 1e12) if kernel_ms else None,
				'pair_evals_per_s': (5 * C_lat / (k_ms * 1e-3)) if kernel_ms else None,
				'algorithmic_flop_per_launch': flop_per_launch, 'launches_timed': len(kernel_ms), 'avg_launch_ms': k_ms,
				'share_of_step': (sum(kernel_ms) / ms) if kernel_ms else None,
				'step_frac': step_flop / (ms / args.steps * 1e-3) / (fma.value * 1e12) / world,
				'step_frac_what': 'whole step, all kernels and gaps: (24 x candidate visits + 28 x accepted pairs of the step, job total) / (step time x FP32 peak x GPUs)'}
	if not args.no_kernel_table:
		roofline['kernels'] = kernel_table(ts, args, mpeaks, ms / args.steps)
		roofline['hbm_kernels'] = hbm_kernels(mpeaks)
		roofline['api_overhead'] = host_overheads(ts)

	out = {
		'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': W,
		'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': main_mode, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': workload_name(args, n),
				   'api': 'advance3d.advance_frame(clone_velocity_field -> advect_covector_field -> project) with the stock BoxSampler / LatticeGenerator / BoxSurfaceSampler generators',
				   'sharding': ((f'weak: samples sharded over {world} ranks — per rank N training + 8192 boundary samples per iteration (global Q = {world}*N, global normalisers) and '
								 f'{args.test_res}^3 of a {args.test_res}x{args.test_res}x{world * args.test_res} test lattice; parameters replicated; per iteration one exchange of the 3*N*12 '
								 'gradient accumulators + loss sums: ' + ('a single kernel summing all ranks\' buffers over NVLink peer memory (csrc/xrank.cu)'
																			 if os.environ.get('GSR_EXCHANGE', 'p2p') == 'p2p' else 'NCCL all-reduce'))
								if main_mode == 'weak' else f'strong: the fixed frame, its {args.test_res}^3 lattice split over {world} ranks, training replicated') if world > 1 else 'single GPU',
				   'l2_policy': 'every step sweeps > 126 MB (test lattice passes write 2.1M x 12 floats; the iterations rewrite all buffers), inputs regenerated each iteration',
				   'work_census': 'candidate visits EXECUTED, counted on the last warm-up step (gsr_count_pairs), RK4 counted as 4 or 5 evaluations of its start points',
				   'test_reference': ('the pull-back reference on the fixed test lattice is evaluated by the first test pass of every frame and reused by the other '
									  'passes of that frame (nothing is carried from one frame to the next); the visits this saves are NOT counted in value '
									  '(pair_evals_not_executed_per_step); "unhoisted" holds the same steps with the recomputation')
									 if advance3d.HOIST_TEST_REFERENCE else 'recomputed at every test pass (GSR_HOIST_TEST_REFERENCE=0)'},
		'timesteps_per_s': args.steps / (ms * 1e-3), 'project_iters_per_s': args.steps * args.iters / (ms * 1e-3),
		'pair_evals_per_step': C_job, 'accepted_pairs_per_step': P_job, 'pair_evals_not_executed_per_step': C_skipped,
		'e2e': {'value': C_job * args.steps / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
				'ms_per_step': ms_e2e / args.steps, 'timesteps_per_s': args.steps / (ms_e2e * 1e-3)},
		'gpu_launches': int(n_launches),
		'clocks': sampler.summary(),
		'roofline': roofline,
		'measured_peaks': mpeaks,
	}
	if other is not None:
		out['strong_scaling'] = other
	if unhoisted is not None:
		unhoisted['step_frac'] = (24 * unhoisted['pair_evals_per_step'] + 28 * unhoisted['accepted_pairs_per_step']) / (unhoisted['ms_per_step'] * 1e-3) / (fma.value * 1e12) / world
		out['unhoisted'] = unhoisted
	if world == 1 and not args.no_cpu_baseline:
		out['cpu_baseline'] = cpu_baseline(args, n, N)
	print(json.dumps(out))
	finish()


if __name__ == '__main__':
	args = parse()
	if args.impl == 'reference':
		run_reference(args)
	elif args.dim == 2:
		from bench_more import run_2d
		run_2d(args)
	elif args.workload == 'density':
		from bench_more import run_density
		run_density(args)
	else:
		run_ours(args)
