"""
bench_more.py — the secondary workloads of bench.py (same JSON contract, same CPU arm conventions):

  python bench.py --workload density [--density-res 512] [--size S2] [--gpus N]
      BASELINE config 5: passive density advection (3D/advance_density.py) of two smoke fields on a res^3 lattice through a
      40^3-Gaussian field; a step = one frame = RK4 back-trace (4 evaluations) + clamp + trilinear resampling of both fields,
      one kernel.  N > 1: the lattice is shared by slabs of x planes, halo planes exchanged point-to-point after every frame.

  python bench.py --dim 2 --scene taylor_green|leapfrog|karman
      BASELINE configs 1-3: one fixed-work 2D time step through advance2d's API (clone -> advect -> project with `iters`
      iterations -> test passes on the scene's visualisation grid) on a synthetic field of the scene's own size
      (24^2, 71^2, 400x60 Gaussians).
"""
import json
import os
import time

from bench import METRIC, ROOT, SIZES, UNIT, ClockSampler, load_synth, timeit


def _dist():
	import torch
	import torch.distributed as dist
	rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
	if not torch.cuda.is_available():
		raise SystemExit('bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)')
	torch.cuda.set_device(local)
	if world > 1:
		dist.init_process_group('nccl', device_id=torch.device('cuda', local))

	def barrier():
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()
	return rank, world, local, dist, barrier


def _peaks(lib):
	import ctypes as C
	from gaussian_fluids_code_b200 import _lib
	fma = C.c_double(0.)
	lib.gsr_peak_fma(C.c_int(20000), C.byref(fma), _lib.stream())
	peaks = {}
	try:
		peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
	except Exception:
		pass
	return fma.value, peaks.get('hbm_gbs')


# ---------------------------------------------------------------------------------------------------------
# density advection (config 5)
# ---------------------------------------------------------------------------------------------------------

def cpu_density_sample(n, nthreads, points=65536, seed=42):
	"""the oracle's RK4 back-trace (positions only) + trilinear resampling on a bounded sample of voxels; (candidate visits, seconds)"""
	import numpy as np
	import torch
	from oracle.oracle import OracleGSR, extended_bounds, interp_val
	P, S, R, V, mgs, gen = load_synth().synthetic_field(n, seed)
	orc = OracleGSR(3, extended_bounds(3, (0., 1.) * 3, mgs), P, S, R, V, 5e-3, mgs, precision='f32', nthreads=nthreads)
	x = torch.rand((points, 3), generator=gen).numpy()
	field = np.random.default_rng(seed).random((64, 64, 64), dtype=np.float32)
	C = orc.count_candidates(x)
	t0 = time.perf_counter()
	pos = np.clip(orc.rk4(x, -.02), 0., 1.)
	interp_val(field, pos, (0., 1.) * 3, real=np.float32, nthreads=nthreads)
	interp_val(field, pos, (0., 1.) * 3, real=np.float32, nthreads=nthreads)
	return 4 * C, time.perf_counter() - t0


def run_density(args):
	import torch
	rank, world, local, dist, barrier = _dist()
	from gaussian_fluids_code_b200 import _lib, advance_density, gsr3d, init_cond3d
	from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
	gsr3d.device = dev = torch.device('cuda', local)
	lib = _lib.lib()
	n = SIZES[args.size] if args.size != 'S1' else 40	# the scenes that carry smoke have 40^3 Gaussians (3D/init_cond.py:22-24)
	res = args.density_res
	P, S, R, V, mgs, _ = synthetic_field(n)
	gv = make_fast3d(P, S, R, V, 5e-3, mgs)
	adv = advance_density.DensityAdvector(0., 1., 0., 1., 0., 1., res=(res, res, res))
	info = init_cond3d.other_info['ring_collide']
	d = [adv.set_ring(info['ring1']), adv.set_ring(info['ring2'])]
	o = [torch.zeros_like(d[0]), torch.zeros_like(d[1])]
	slab = advance_density.SlabSharding(res, rank, world, halo=4)
	rng = slab.x_range() if world > 1 else None

	def frame():
		nonlocal d, o
		adv.advect(gv, .02, d[0], d[1], x_range=rng, out=o)
		d, o = o, d
		slab.exchange(*d)

	# census: candidate visits of this rank's voxels (4 evaluations of the start points), slab by slab
	cnt = torch.zeros(2, dtype=torch.int64, device=dev)
	gv._engine.ensure_packed(gv._params())
	x0, x1 = slab.x_range()
	xs, ys, zs = adv.axes
	for i0 in range(x0, x1, 16):
		pts = torch.stack(torch.meshgrid(xs[i0:min(i0 + 16, x1)], ys, zs, indexing='ij'), dim=-1).reshape(-1, 3).contiguous()
		gv._engine.count_pairs(pts, cnt, 4, True)
	del pts
	if world > 1:
		dist.all_reduce(cnt)
	C_job, P_job = [int(v) for v in cnt.tolist()]
	# pair tests the kernel really runs (a candidate that survives the warp's bounding-box culling is tested on the warp's 128 voxels)
	ex = torch.zeros(1, dtype=torch.int64, device=dev)
	gv._engine.advect_density(adv.axes, adv.domain, -.02, d[0], o[0], d[1], o[1], x_range=rng, executed=ex)
	if world > 1:
		dist.all_reduce(ex)
	E_job = int(ex.item())
	W = max(args.warmup, 3)
	for _ in range(W):
		frame()
	sampler = ClockSampler(local)
	sampler.start()
	l0 = lib.gsr_launch_count()
	barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(args.steps):
		frame()
	e1.record()
	barrier()
	nl = lib.gsr_launch_count() - l0
	sampler.stop_flag = True
	sampler.join()
	t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	ms = float(t.item())
	# e2e: both fields from pinned host memory each frame, both results back
	hd = [torch.empty(d[0].shape, dtype=torch.float32).pin_memory() for _ in range(2)]
	for h, s in zip(hd, d):
		h.copy_(s)
	ho = [torch.empty_like(h).pin_memory() for h in hd]
	barrier()
	e0.record()
	for _ in range(args.steps):
		for dst, h in zip(d, hd):
			dst.copy_(h, non_blocking=True)
		frame()
		for h, s in zip(ho, d):
			h[x0:x1].copy_(s[x0:x1], non_blocking=True)
		torch.cuda.synchronize()
	e1.record()
	barrier()
	t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	ms_e2e = float(t.item())
	if rank != 0:
		if world > 1:
			barrier()
			os._exit(0)
		return
	fma, hbm = _peaks(lib)
	kms = timeit(lambda: adv.advect(gv, .02, d[0], d[1], x_range=rng, out=o), reps=5)
	flop = (24 * E_job + 7 * P_job) / world	# executed pair tests (24 flop) + accepted pairs (u only: 7 flop), this rank's slab
	flop_stencil = (24 * C_job + 7 * P_job) / world	# SURVEY 8d's unit: every occupant of every voxel's 27-cell stencil
	out = {
		'metric': METRIC, 'value': C_job * args.steps / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': W, 'ms_per_step': ms / args.steps,
		'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': f'passive density advection (3D/advance_density.py), {res}^3 lattice, two smoke fields of ring_collide, N={n ** 3} Gaussians; a step = one frame',
				   'sharding': f'the lattice split into {world} slabs of x planes, 4 halo planes exchanged point-to-point with each neighbour after every frame' if world > 1 else 'single GPU',
				   'l2_policy': f'the two density fields are {2 * res ** 3 * 4 / 1e6:.0f} MB in and the same out per frame (> 126 MB L2 at 512^3)'},
		'frames_per_s': args.steps / (ms * 1e-3), 'pair_evals_per_step': C_job, 'accepted_pairs_per_step': P_job,
		'e2e': {'value': C_job * args.steps / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': 2 * res ** 3 * 4, 'd2h_bytes_per_step': 2 * (x1 - x0) * res * res * 4,
				'ms_per_step': ms_e2e / args.steps},
		'gpu_launches': int(nl), 'clocks': sampler.summary(),
		'roofline': {'bound': 'fp32', 'kernel': 'advect_density_kernel<2> (RK4 back-trace, 4 field evaluations per voxel, clamp, 2 x 8 trilinear taps)', 'achieved': flop / (kms * 1e-3) / 1e12,
					 'peak': fma, 'unit': 'TFLOP/s', 'frac': flop / (kms * 1e-3) / 1e12 / fma, 'traffic': None, 'avg_launch_ms': kms,
					 'algorithmic_flop_per_launch': flop, 'executed_pair_tests_per_step': E_job, 'stencil_flop_per_launch': flop_stencil,
					 'stencil_frac': flop_stencil / (kms * 1e-3) / 1e12 / fma,
					 'note': 'frac counts the pair tests that survive the warp-level culling (gsr_advect_density_census, same kernel with a counter) at 24 flop + the accepted '
					 'pairs at 7 flop; stencil_frac counts every occupant of every voxel\'s 27-cell stencil (SURVEY 8d\'s unit, what the reference evaluates) and may exceed 1: '
					 'culling skips work the stencil definition charges for', 'hbm_bytes_per_launch': 4 * res ** 3 * 4 / world, 'hbm_gbs': 4 * res ** 3 * 4 / world / (kms * 1e-3) / 1e9},
		'measured_peaks': {'hbm_gbs': hbm, 'fp32_tflops_live': fma},
	}
	if world == 1 and not args.no_cpu_baseline:
		cores = os.cpu_count() or 1
		cpu_density_sample(n, cores, points=4096)
		c, ts_ = cpu_density_sample(n, cores)
		out['cpu_baseline'] = {'value': c / ts_, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': 'RK4 back-trace + 2 trilinear resamplings of 65536 voxels, oracle/ f32 OpenMP'}
	print(json.dumps(out))
	if world > 1:
		barrier()
		os._exit(0)


# ---------------------------------------------------------------------------------------------------------
# 2D time step (configs 1-3)
# ---------------------------------------------------------------------------------------------------------

def synthetic_field2d(scene, seed=42):
	"""a field of the scene's own size: the scene's lattice of Gaussians in GSR space, jittered, log-radii near the initial value,
	random angles, values = the scene's analytic velocity at the centres scaled down (so that the flow moves the field a little)"""
	import numpy as np
	import torch
	from gaussian_fluids_code_b200 import gsr2d
	gen = torch.Generator().manual_seed(seed)
	x_min, x_max, y_min, y_max = scene.scaled(scene.initialize_domain)
	nx, ny = scene.particle_count
	pts = gsr2d.get_grid_points(x_min, x_max, y_min, y_max, nx, ny).cpu()
	hx, hy = (x_max - x_min) / (nx - 1), (y_max - y_min) / (ny - 1)
	pts = pts + (torch.rand(pts.shape, generator=gen) - .5) * .5 * torch.tensor([hx, hy])
	pts[:, 0].clamp_(x_min, x_max)
	pts[:, 1].clamp_(y_min, y_max)
	gv = gsr2d.GaussianSplattingFast(x_min, x_max, y_min, y_max, pts.numpy().astype(np.float32), dim=2)
	N = gv.N
	with torch.no_grad():
		gv.scalings += (torch.randn((N, 2), generator=gen) * .1).clamp(-.2, .2).to(gv.scalings.device)
		gv.rotations.copy_((torch.rand((N,), generator=gen) * 3.14159).to(gv.rotations.device))
		gv.values.copy_((torch.randn((N, 2), generator=gen) * .1).to(gv.values.device))
	gv.zero_grad()
	return gv


def run_2d(args):
	import numpy as np
	import torch
	rank, world, local, dist, barrier = _dist()
	if world > 1:
		raise SystemExit('the 2D workloads are single-GPU (the largest, karman, has 24000 Gaussians)')
	from gaussian_fluids_code_b200 import _lib, advance2d, gsr2d, init_cond2d
	gsr2d.device = dev = torch.device('cuda', local)
	lib = _lib.lib()
	scene = init_cond2d.Scene2D(args.scene)
	dt = {'taylor_green': .001, 'leapfrog': .025, 'karman': .05}.get(args.scene, .01)	# BASELINE.json configs 1-3
	cur, new = synthetic_field2d(scene), synthetic_field2d(scene)
	params0 = [p.detach().clone() for p in cur._params()]
	N = cur.N
	gen_fn = lambda n_, gs, restrict=None: scene.data_generator(gs)
	gen_fn.graph_safe = True	# a pure function of torch's CUDA random stream: project() may replay its iterations from a CUDA graph
	test_fn = lambda gs: scene.test_generator()
	b1, b2 = scene.boundary_samplers
	iters = args.iters

	def reset():
		with torch.no_grad():
			for f in (cur, new):
				for p, q in zip(f._params(), params0):
					p.copy_(q)
				f.zero_grad()

	state = {'cur': cur, 'new': new}

	def step():
		"""one frame of 2D/advance.py:354-365 with the iteration count pinned"""
		c, nw = state['cur'], state['new']
		advance2d.clone_velocity_field(nw, c, gen_fn, test_fn, max_epoch=iters, verbose=0)
		advance2d.advect_covector_field(nw, c, dt, extra_advector=None)
		ref = advance2d.AdvectedCovectorField(c, c, dt, domain=scene.scaled(scene.advance_domain))
		advance2d.project(nw, ref, gen_fn, test_fn, boundary_generator_1=b1, boundary_generator_2=b2, boundary_lambda=1., max_epoch=iters, patience=10 ** 9, verbose=0)
		state['cur'], state['new'] = nw, c

	# census on representative batches: per iteration 5 evaluations of the previous field (RK4 pull-back) + forward + backward on
	# Q = N samples, forward + backward on each boundary batch; per 100 iterations a test pass (5 + 1) on the visualisation grid
	cnt = torch.zeros(2, dtype=torch.int64, device=dev)
	e = cur._engine
	e.ensure_packed(cur._params())
	one = torch.zeros(2, dtype=torch.int64, device=dev)
	e.count_pairs(scene.data_generator(cur), one, 7, True)
	for b in (b1, b2):
		if b is not None:
			e.count_pairs(b(512)[0].contiguous(), one, 2, True)
	cnt += one * iters
	e.count_pairs(scene.test_generator().contiguous(), cnt, 6 * (iters // 100), True)
	e.count_pairs(cur.positions.detach(), cnt, 4, True)
	C_step, P_step = [int(v) for v in cnt.tolist()]
	W = max(args.warmup, 3)
	for _ in range(W):
		reset()
		step()
	sampler = ClockSampler(local)
	sampler.start()
	from gaussian_fluids_code_b200 import graphloop
	l0 = lib.gsr_launch_count() + graphloop.GRAPH_LAUNCHES
	barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
	e0.record()
	marks[0].record()
	for k in range(args.steps):
		reset()
		step()
		marks[k + 1].record()
	e1.record()
	barrier()
	ms = e0.elapsed_time(e1)
	ms_each = [marks[k].elapsed_time(marks[k + 1]) for k in range(args.steps)]
	nl = lib.gsr_launch_count() + graphloop.GRAPH_LAUNCHES - l0	# eager launches + those replayed from the captured iteration graphs
	sampler.stop_flag = True
	sampler.join()
	host_params = [p.cpu().pin_memory() for p in params0]
	host_out = [torch.empty_like(p).pin_memory() for p in host_params]
	barrier()
	e0.record()
	for _ in range(args.steps):
		with torch.no_grad():
			for f in (cur, new):
				for p, q in zip(f._params(), host_params):
					p.copy_(q, non_blocking=True)
				f.zero_grad()
		step()
		for dst, src in zip(host_out, state['cur']._params()):
			dst.copy_(src, non_blocking=True)
		torch.cuda.synchronize()
	e1.record()
	barrier()
	ms_e2e = e0.elapsed_time(e1)
	fma, hbm = _peaks(lib)
	# the dominant kernel class: the RK4 pull-back + forward on the visualisation grid
	grid_pts = scene.test_generator().contiguous()
	c = state['cur']
	c._engine.ensure_packed(c._params())
	cg = torch.zeros(2, dtype=torch.int64, device=dev)
	c._engine.count_pairs(grid_pts, cg, 1, True)
	Cg, Pg = [int(v) for v in cg.tolist()]
	rv = torch.empty((grid_pts.shape[0],), device=dev)
	kms = timeit(lambda: c._engine.advected_vorticity(grid_pts, -dt, rv, None, domain=scene.scaled(scene.advance_domain)))
	flop = 5 * (16 * Cg + 13 * Pg)	# 2D: d (2) + w = A d (6) + q (3) + compare ~ 16 flop per candidate; 13 per accepted pair (u, grad u)
	out = {
		'metric': METRIC, 'value': C_step * args.steps / (ms * 1e-3), 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': W, 'ms_per_step': ms / args.steps,
		'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': f'2D {args.scene} fixed-work timestep through advance2d (clone -> advect -> project), N={N} Gaussians ({scene.particle_count[0]}x{scene.particle_count[1]}), '
							   f'Q=N samples/iter, {iters} project iters, dt={dt}, test grid {scene.visualize_res[0]}x{scene.visualize_res[1]}',
				   'work_census': 'candidate visits on representative batches x the step\'s evaluation counts'},
		'ms_each_step': ms_each, 'timesteps_per_s': args.steps / (ms * 1e-3), 'project_iters_per_s': args.steps * iters / (ms * 1e-3), 'pair_evals_per_step': C_step, 'accepted_pairs_per_step': P_step,
		'e2e': {'value': C_step * args.steps / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': sum(p.numel() * 4 for p in host_params),
				'd2h_bytes_per_step': sum(p.numel() * 4 for p in host_out), 'ms_per_step': ms_e2e / args.steps},
		'gpu_launches': int(nl), 'clocks': sampler.summary(),
		'roofline': {'bound': 'fp32', 'kernel': 'rk4_2d_kernel<2> on the visualisation grid (RK4 pull-back, 5 evaluations)', 'achieved': flop / (kms * 1e-3) / 1e12, 'peak': fma, 'unit': 'TFLOP/s',
					 'frac': flop / (kms * 1e-3) / 1e12 / fma, 'traffic': None, 'avg_launch_ms': kms, 'algorithmic_flop_per_launch': flop,
					 'step_frac': (16 * C_step + 13 * P_step) / (ms / args.steps * 1e-3) / (fma * 1e12),
					 'note': 'the 2D step is latency bound at these sizes (a few thousand samples per kernel): ~25 short kernels per iteration, replayed 10 iterations per CUDA graph (graphloop.py)'},
		'measured_peaks': {'hbm_gbs': hbm, 'fp32_tflops_live': fma},
	}
	if not args.no_cpu_baseline:
		from oracle.oracle import OracleGSR, extended_bounds
		cores = os.cpu_count() or 1
		P_, S_, R_, V_ = [p.cpu().numpy() for p in params0]
		bounds = scene.scaled(scene.initialize_domain)
		orc = OracleGSR(2, extended_bounds(2, bounds, cur.min_grid_scale), P_, S_, R_, V_, cur.clamp_threshold, cur.min_grid_scale, precision='f32', nthreads=cores)
		x = scene.data_generator(cur).cpu().numpy()
		xg = grid_pts.cpu().numpy()[:65536]
		cc = torch.zeros(2, dtype=torch.int64, device=dev)	# candidate visits of the CPU sample, counted by the census kernel on the same hash
		cur._engine.ensure_packed(cur._params())
		cur._engine.count_pairs(torch.tensor(x, device=dev), cc, 7, False)
		cur._engine.count_pairs(torch.tensor(xg, device=dev), cc, 6, False)
		cands = int(cc[0].item())
		t0 = time.perf_counter()
		for _ in range(3):
			orc.rk4(x, -dt, pos_only=False)
			val, grad = orc.forward(x)
			orc.backward2d_grad(x, grad, ref_vor=np.zeros(x.shape[0], np.float32), weight_vor=1., weight_div=1.)
			orc.rk4(xg, -dt, pos_only=False)
			orc.forward(xg)
		out['cpu_baseline'] = {'value': 3 * cands / (time.perf_counter() - t0), 'unit': UNIT, 'cores': cores, 'kind': 'port',
							   'sample': f'3 x (RK4 pull-back + fwd + bwd on Q=N={N}, RK4 pull-back + fwd on {xg.shape[0]} grid points), oracle/ f32 OpenMP'}
	print(json.dumps(out))
