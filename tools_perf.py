"""quick kernel timing probe (development tool): python tools_perf.py n [Q]"""
import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from test_gpu_3d_kernels import synthetic, make_fast3d

def timeit(fn, reps=10, warm=3):
	for _ in range(warm): fn()
	torch.cuda.synchronize()
	evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
	for a, b in evs:
		a.record(); fn(); b.record()
	torch.cuda.synchronize()
	return float(np.median([a.elapsed_time(b) for a, b in evs]))

n = int(sys.argv[1]); N = n ** 3
Q = int(sys.argv[2]) if len(sys.argv) > 2 else N
P, S, R, V, mgs, gen = synthetic(n)
o = make_fast3d(P, S, R, V, 5e-3, mgs)
e = o._engine
x = torch.rand((Q, 3), generator=gen).cuda()
e.ensure_packed(o._params())
perm, scs = e.bin_samples(x, True)
cs = e.cell_start.cpu().numpy().astype(np.int64)
# candidates per sample via cell counts (3x3x3 box sums)
dims = o.grid_size
cnt = np.diff(cs).reshape(dims)
pad = np.pad(cnt, 1)
box = sum(pad[i:i+dims[0], j:j+dims[1], k:k+dims[2]] for i in range(3) for j in range(3) for k in range(3))
xc = np.floor((x.cpu().numpy() - np.float32(o.x_min)) / np.float32(o.grid_scale)).astype(int)
C = int(box[xc[:, 0], xc[:, 1], xc[:, 2]].sum())
res = {'n': n, 'N': N, 'Q': Q, 'C_per_Q': C / Q}
val = torch.zeros((Q, 3), device='cuda'); grad = torch.zeros((Q, 3, 3), device='cuda')
t = timeit(lambda: e.forward(x, val, grad, True, perm)); res['fwd_ms'] = t; res['fwd_Gcand_s'] = C / t / 1e6
t = timeit(lambda: e.forward(x, val, None, True, perm)); res['fwd_valonly_ms'] = t
gp = torch.zeros((Q, 3), device='cuda'); df = torch.zeros((Q, 3, 3), device='cuda')
def rk4_full():
	perm_, _ = e.bin_samples(x, False)
	e.rk4(x, -.02, gp, df, val, grad)
t = timeit(rk4_full)
res['rk4_full_ms'] = t; res['rk4_Gcand_s'] = 5 * C / t / 1e6
ref_vor = torch.randn((Q, 3), device='cuda') * .1; ref_hel = torch.randn((Q,), device='cuda') * .1
val.zero_(); grad.zero_(); e.forward(x, val, grad, True, perm)
t = timeit(lambda: e.backward_gather(x, perm, scs, val, grad, (0, 0, 0, 1., 1., 1.), {'ref_vor': ref_vor, 'ref_hel': ref_hel}, None)); res['bwd_gather_ms'] = t; res['bwd_Gcand_s'] = C / t / 1e6
t = timeit(lambda: e.bin_samples(x, True)); res['bin_samples_ms'] = t
t = timeit(lambda: e.build(o.positions.detach())); res['build_grid_ms'] = t
t = timeit(lambda: (setattr(e, '_packed_key', None), e.ensure_packed(o._params()))); res['pack_ms'] = t
t = timeit(lambda: o.get_losses(x)); res['get_losses_fwd_api_ms'] = t
print(json.dumps(res))
