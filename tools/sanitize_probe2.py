"""development probe for compute-sanitizer: the round-2 kernels on small inputs — four-lane cluster step with the in-kernel hash (3D and 2D), device
split, density slab, 2D project with both boundary samplers.  compute-sanitizer --tool memcheck python tools/sanitize_probe2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import advance2d, advance3d, advance_density, gsr2d, gsr3d, init_cond2d, init_cond3d, reseed
from gaussian_fluids_code_b200.synth import make_fast3d, synthetic_field
gsr3d.device = gsr2d.device = torch.device('cuda', 0)
# 3D frame through the API, eager (no graph under the sanitizer), 6 iterations
P, S, R, V, mgs, gen = synthetic_field(6)
a, b = make_fast3d(P, S, R, V, 5e-3, mgs), make_fast3d(P, S, R, V, 5e-3, mgs)
box = (0., 1.) * 3
a, b, ep, _ = advance3d.advance_frame(a, b, *box, .02, advance3d.BoxSampler(*box), advance3d.LatticeGenerator(*box, 16, 16, 16), boundary_generator=advance3d.BoxSurfaceSampler(*box),
									 boundary_lambda=10., max_epoch=6, patience=10 ** 9, verbose=0, batch_size=1024, check_iter=2, use_graph=False)
print('3d frame', ep, flush=True)
# split
S2 = S.copy(); S2[::7, 0] -= 1.
f = make_fast3d(P, S2, R, V, 5e-3, mgs)
print('split', reseed.split_once(f, 3, clamp_box=(f.x_min, f.x_max, f.y_min, f.y_max, f.z_min, f.z_max), seed=3)[0], flush=True)
# density slab
adv = advance_density.DensityAdvector(0., 1., 0., 1., 0., 1., res=(24, 24, 24))
d1 = adv.set_ring(init_cond3d.other_info['ring_collide']['ring1'])
o = torch.zeros_like(d1)
adv.advect(a, .02, d1, None, x_range=(5, 17), out=[o])
print('density', float(o.sum()), flush=True)
# 2D karman projection, eager
scene = init_cond2d.Scene2D('karman')
scene.particle_count, scene.visualize_res = (40, 8), (32, 16)
gv = advance2d.simulation_initialize(scene, max_epoch=3, verbose=0, project_epochs=3, use_graph=False)
print('2d', gv.N, bool(torch.isfinite(gv.positions).all()), flush=True)
torch.cuda.synchronize()
print('ok')
