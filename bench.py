#!/usr/bin/env python
"""
bench.py — measures the Gaussian-Fluids hot path on B200 (contract: see the task statement and DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S1|S2|S3] [--iters 600] [--impl ours|reference]

A "step" is one fixed-work 3D leapfrog time step (gaussian-fluids-code_b200/timestep3d.py): advect + `iters`
project iterations (+ boundary passes) + a test pass every 100 iterations + the two output-field passes.
`value` = Gaussian-sample pair evaluations per second, whole job (all ranks); `timesteps_per_s` rides along.
`e2e` is the same metric with the Gaussian parameters coming from pinned HOST buffers each step and the updated
parameters plus the two output fields copied back to the host inside the timed region.
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

# DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture: (size, test_res, world) -> bytes
NCU_TRAFFIC = {('S1', 128, 1): 33694720 + 1133056}

METRIC = 'gaussian_sample_pair_evals_per_s'
UNIT = 'pair-evals/s'


def parse():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type=int, default=1)
	ap.add_argument('--steps', type=int, default=3)
	ap.add_argument('--warmup', type=int, default=3)
	ap.add_argument('--impl', type=str, default='ours')
	ap.add_argument('--size', type=str, default='S1', help='S1 = 10^3 Gaussians (the reference\'s 3D leapfrog size), S2 = 40^3, S3 = 100^3')
	ap.add_argument('--iters', type=int, default=600, help='project iterations per time step (reference minimum: 600)')
	ap.add_argument('--test-res', type=int, default=128)
	ap.add_argument('--no-cpu-baseline', action='store_true')
	ap.add_argument('--no-graph', action='store_true', help='run the project iterations eagerly instead of replaying a CUDA graph')
	return ap.parse_args()


def workload_name(args, n):
	return f'3D leapfrog fixed-work timestep, N={n ** 3} Gaussians ({args.size}), Q=N samples/rank/iter, {args.iters} project iters, boundary 8192/rank, test lattice {args.test_res}^3'


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C restatement of the reference's kernels), all host threads, bounded sample
# ---------------------------------------------------------------------------------------------------------

def cpu_sample(n, nthreads, lattice_points=16384, reps=1, seed=42):
	"""
	One bounded sample of the timestep's kernel work on the CPU: {RK4 pull-back (5 evals) + forward + backward with the
	project weights on Q = N samples; boundary forward + backward on 8192 samples; RK4 pull-back + forward on
	`lattice_points` lattice points}.  Returns (candidate visits, seconds).
	"""
	import numpy as np
	import torch
	from oracle.oracle import OracleGSR, extended_bounds
	spec = __import__('importlib.util').util.spec_from_file_location('synth_b200', os.path.join(ROOT, 'gaussian-fluids-code_b200', 'synth.py'))
	synth = __import__('importlib.util').util.module_from_spec(spec)
	spec.loader.exec_module(synth)
	P, S, R, V, mgs, gen = synth.synthetic_field(n, seed)
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


def run_reference(args):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	from gaussian_fluids_sizes import SIZES
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
	print(json.dumps({
		'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
		'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': workload_name(args, n), 'note': 'reference Taichi kernels cannot run (taichi not installed): CPU restatement oracle/ (C, OpenMP), bounded sample'},
		'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
		'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
	}))


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
	from gaussian_fluids_code_b200 import _lib, gsr3d, timestep3d
	from gaussian_fluids_code_b200.synth import SIZES
	gsr3d.device = torch.device('cuda', local)
	lib = _lib.lib()
	n = SIZES[args.size]
	dev = gsr3d.device

	def barrier():
		torch.cuda.synchronize()
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	ts = timestep3d.LeapfrogTimestep(n=n, iters=args.iters, test_res=args.test_res, rank=rank, world=world, use_graph=not args.no_graph)
	N = ts.N
	# host-resident copies of the parameters (pinned) for the e2e path, and pinned landing buffers for its outputs
	host_params = [torch.as_tensor(a).clone().pin_memory() for a in ts.params0]
	host_out = [torch.empty_like(p).pin_memory() for p in host_params]
	host_fields = [torch.empty(ts.lattice.shape[0], dtype=torch.float32).pin_memory() for _ in range(2)]

	# ---- warm-up; the work census (candidate visits, accepted pairs) is taken on the last warm-up step --------------
	census = timestep3d.Census(dev)
	probe = []
	for w in range(max(args.warmup, 3)):
		ts.reset()
		ts.step(census if w == max(args.warmup, 3) - 1 else None)
	C_step, P_step = census.value()
	if world > 1:
		tot = torch.tensor([C_step, P_step], dtype=torch.int64, device=dev)
		dist.all_reduce(tot)
		C_job, P_job = [int(v) for v in tot.tolist()]
	else:
		C_job, P_job = C_step, P_step

	# ---- timed region: device-resident inputs -------------------------------------------------------------------
	sampler = ClockSampler(local)
	sampler.start()
	launches0 = lib.gsr_launch_count() + ts.graph_launches
	ts.probe = probe
	barrier()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	for _ in range(args.steps):
		ts.reset()
		ts.step()
	e1.record()
	barrier()
	ms = e0.elapsed_time(e1)
	ts.probe = None
	launches = lib.gsr_launch_count() + ts.graph_launches - launches0
	sampler.stop_flag = True
	sampler.join()
	t = torch.tensor([ms], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	ms = float(t.item())
	value = C_job * args.steps / (ms * 1e-3)

	# ---- end-to-end: host buffers in, host buffers out, copies inside the timed region -----------------------------
	barrier()
	e0.record()
	for _ in range(args.steps):
		ts.reset([p.to(dev, non_blocking=True) for p in host_params])
		vor, div = ts.step()
		for dst, src in zip(host_out, (ts.cur.positions, ts.cur.scalings, ts.cur.rotations, ts.cur.values)):
			dst.copy_(src.detach(), non_blocking=True)
		host_fields[0].copy_(vor, non_blocking=True)
		host_fields[1].copy_(div, non_blocking=True)
		torch.cuda.synchronize()	# the caller owns the results only once they are on the host
	e1.record()
	barrier()
	t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	ms_e2e = float(t.item())
	h2d = sum(p.numel() * 4 for p in host_params)
	d2h = sum(p.numel() * 4 for p in host_out) + sum(f.numel() * 4 for f in host_fields)

	def finish():
		"""leave without tearing NCCL down: the captured iteration graphs hold the communicator's kernels, and destroying the
		process group under them can block; every rank has synchronised and rank 0 has printed by the time this runs"""
		sys.stdout.flush()
		sys.stderr.flush()
		if world > 1:
			torch.cuda.synchronize()
			dist.barrier()
			torch.cuda.synchronize()
			os._exit(0)

	if rank != 0:
		finish()
		return

	# ---- roofline of the dominant kernel: the RK4 pull-back (rk4_tiled3s_kernel<2>) on the test lattice --------------------
	fma, mufu = C.c_double(0.), C.c_double(0.)
	lib.gsr_peak_fma(C.c_int(20000), C.byref(fma), _lib.stream())
	lib.gsr_peak_mufu(C.c_int(20000), C.byref(mufu), _lib.stream())
	cl = torch.zeros(2, dtype=torch.int64, device=dev)
	ts.cur._engine.ensure_packed(ts.cur._params())
	ts.cur._engine.count_pairs(ts.lattice, cl, 1, True)
	C_lat, P_lat = [int(v) for v in cl.tolist()]
	kernel_ms = [a.elapsed_time(b) for a, b in probe]
	k_ms = sum(kernel_ms) / max(len(kernel_ms), 1)
	flop_per_launch = 5 * (24 * C_lat + 28 * P_lat)
	achieved = flop_per_launch / (k_ms * 1e-3) / 1e12 if kernel_ms else None
	roofline = {'bound': 'fp32', 'kernel': 'rk4_tiled3s_kernel<2> (RK4 pull-back of the previous field on the test lattice, 5 field evaluations per point)',
				'achieved': achieved, 'peak': fma.value, 'unit': 'TFLOP/s', 'frac': (achieved / fma.value) if achieved else None,
				'traffic': NCU_TRAFFIC.get((args.size, args.test_res, world)), 'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel on this workload (profiles/ncu_full_rk4_tiled3s_r01.csv); null for workloads that were not captured',
				'peak_source': 'FP32 FFMA peak measured live by gsr_peak_fma on this GPU (MEASURED_PEAKS.json holds only HBM and bf16 peaks); nominal 74.4 at 1965 MHz',
				'mufu_peak_Tops': mufu.value, 'mufu_achieved_Tops': (5 * P_lat / (k_ms * 1e-3) / 1e12) if kernel_ms else None,
				'pair_evals_per_s': (5 * C_lat / (k_ms * 1e-3)) if kernel_ms else None,
				'algorithmic_flop_per_launch': flop_per_launch, 'launches_timed': len(kernel_ms), 'avg_launch_ms': k_ms,
				'share_of_step': (sum(kernel_ms) / ms) if kernel_ms else None}
	peaks = {}
	try:
		peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
	except Exception:
		pass

	out = {
		'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
		'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
		'config': {'workload': workload_name(args, n), 'sharding': f'samples sharded over {world} rank(s): per rank N training + 8192 boundary samples per iteration (global Q = {world}*N, global normalisers) and {args.test_res}^3 of a {args.test_res}x{args.test_res}x{world * args.test_res} test lattice; parameters replicated; per iteration one exchange of the 3*N*12 gradient accumulators + loss sums: ' + ('a single kernel summing all ranks\' buffers over NVLink peer memory (csrc/xrank.cu)' if os.environ.get('GSR_EXCHANGE', 'p2p') == 'p2p' else 'NCCL all-reduce') if world > 1 else 'single GPU',
				   'l2_policy': 'every step sweeps > 126 MB (test lattice passes write 2.1M x 12 floats; the iterations rewrite all buffers), inputs regenerated each iteration',
				   'work_census': 'candidate visits counted on the last warm-up step (gsr_count_pairs), RK4 counted as 4 or 5 evaluations of its start points'},
		'timesteps_per_s': args.steps / (ms * 1e-3), 'project_iters_per_s': args.steps * args.iters / (ms * 1e-3),
		'pair_evals_per_step': C_job, 'accepted_pairs_per_step': P_job,
		'e2e': {'value': C_job * args.steps / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
				'ms_per_step': ms_e2e / args.steps, 'timesteps_per_s': args.steps / (ms_e2e * 1e-3)},
		'gpu_launches': int(launches),
		'clocks': sampler.summary(),
		'roofline': roofline,
		'measured_peaks': {'hbm_gbs': peaks.get('hbm_gbs'), 'fp32_tflops_live': fma.value, 'mufu_tops_live': mufu.value},
	}
	if world == 1 and not args.no_cpu_baseline:
		cores = os.cpu_count() or 1
		lattice_points = 131072 if n <= 40 else 16384
		cpu_sample(n, cores, lattice_points=2048)
		c, tsec = cpu_sample(n, cores, lattice_points=lattice_points, reps=2)
		out['cpu_baseline'] = {'value': c / tsec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
							   'sample': f'2 x (RK4 pull-back + fwd + bwd on Q=N={N}, boundary fwd+bwd on 8192, RK4 pull-back + fwd on {lattice_points} lattice points), oracle/ f32 OpenMP'}
	print(json.dumps(out))
	finish()


if __name__ == '__main__':
	args = parse()
	if args.impl == 'reference':
		# sizes without importing the CUDA package
		sys.modules['gaussian_fluids_sizes'] = type(sys)('gaussian_fluids_sizes')
		sys.modules['gaussian_fluids_sizes'].SIZES = {'S1': 10, 'S2': 40, 'S3': 100, 'S4': 160, 'S5': 256}
		run_reference(args)
	else:
		run_ours(args)
