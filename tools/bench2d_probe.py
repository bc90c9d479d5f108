"""development probe: where a 2D frame's time goes (clone / advect / project; graph vs eager).  python tools/bench2d_probe.py [scene] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_more
from gaussian_fluids_code_b200 import advance2d, gsr2d, init_cond2d, graphloop
gsr2d.device = torch.device('cuda', 0)
name = sys.argv[1] if len(sys.argv) > 1 else 'taylor_green'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 600
scene = init_cond2d.Scene2D(name)
cur, new = bench_more.synthetic_field2d(scene), bench_more.synthetic_field2d(scene)
params0 = [p.detach().clone() for p in cur._params()]
gen = lambda n_, gs, restrict=None: scene.data_generator(gs)
gen.graph_safe = True
test = lambda gs: scene.test_generator()
b1, b2 = scene.boundary_samplers
dt = .001


def T(fn):
	torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r


for use_graph in (False, True, True, True):
	with torch.no_grad():
		for f in (cur, new):
			for p, q in zip(f._params(), params0):
				p.copy_(q)
			f.zero_grad()
	t_clone, _ = T(lambda: advance2d.clone_velocity_field(new, cur, gen, test, max_epoch=iters, verbose=0))
	t_adv, _ = T(lambda: advance2d.advect_covector_field(new, cur, dt))
	ref = advance2d.AdvectedCovectorField(cur, cur, dt, domain=scene.scaled(scene.advance_domain))
	g0 = graphloop.GRAPH_LAUNCHES
	t_proj, _ = T(lambda: advance2d.project(new, ref, gen, test, boundary_generator_1=b1, boundary_generator_2=b2, boundary_lambda=1., max_epoch=iters, patience=10 ** 9, verbose=0, use_graph=use_graph))
	print(f'{name} N={cur.N} graph={use_graph}: clone {t_clone:.1f} ms, advect {t_adv:.1f} ms, project {t_proj:.1f} ms ({t_proj / iters * 1e3:.0f} us/iter), graph launches {graphloop.GRAPH_LAUNCHES - g0}', flush=True)
