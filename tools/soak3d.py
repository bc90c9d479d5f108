"""end-to-end soak of the reference's 3D driver sequence on its own leapfrog scene (development tool): initialize.py's fit of the two
vortex rings (N = 10^3), then advance.py's time loop with the reference's early stop (patience 500, max 20000 iterations) — every
frame: iterations run, test losses, kinetic energy proxy, Gaussian count.  python tools/soak3d.py [frames] [fit_epochs] [scene] [density_res]
(ring_collide: 40^3 Gaussians, and its two smoke rings advected on a density_res^3 lattice after every frame, as 3D/advance_density.py does)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussian_fluids_code_b200 import advance3d, gsr3d, init_cond3d, initialize3d
gsr3d.device = torch.device('cuda', 0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10
fit_epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 500
name = sys.argv[3] if len(sys.argv) > 3 else 'leapfrog'
box = init_cond3d.domain[name]
torch.manual_seed(42)
t0 = time.time()
gv = initialize3d.simulation_initialize(name, max_epoch=fit_epochs, verbose=0)
new = initialize3d.simulation_initialize(name, max_epoch=0, verbose=0)
torch.cuda.synchronize()
print(f'init: N = {gv.N}, {time.time() - t0:.2f} s', flush=True)
log = []
dens = None
if name == 'ring_collide':
	from gaussian_fluids_code_b200 import advance_density
	res = int(sys.argv[4]) if len(sys.argv) > 4 else 256
	adv = advance_density.DensityAdvector(*box, res=(res,) * 3)
	info = init_cond3d.other_info[name]
	dens = [adv.set_ring(info['ring1']), adv.set_ring(info['ring2'])]
	mass0 = [float(d.sum()) for d in dens]


def on_frame(k, field, vor, div):
	global dens
	if dens is not None:
		dens = list(adv.advect(field, .02, dens[0], dens[1]))
	torch.cuda.synchronize()
	log.append((k, field.N, float(vor.mean()), float(vor.max()), float(div.abs().mean()), bool(all(torch.isfinite(p).all() for p in field._params()))))
	print('frame %d: N = %d, mean |vorticity| %.4f, max %.3f, mean |div| %.5f, finite %s, %.2f s' % (log[-1] + (time.time() - t0,)), flush=True)


advance3d.advance(gv, new, *box, .02, frames * .02 - 1e-9, boundary_generator=advance3d.BoxSurfaceSampler(*box), boundary_lambda=10., visualize_res=(64, 64, 64),
				  max_epoch=20000, patience=500, verbose=0, on_frame=on_frame, batch_size=8192)
assert len(log) == frames and all(r[-1] for r in log)
if dens is not None:
	print('smoke mass (voxels) start', mass0, 'end', [float(d.sum()) for d in dens])
print('ok')
