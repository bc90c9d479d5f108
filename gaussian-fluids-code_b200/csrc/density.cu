// density.cu — passive density advection on a regular lattice (reference 3D/advance_density.py:13-63).
//
// The reference materialises the 512^3 lattice (1.6 GB of coordinates), back-traces it with advection_rk4 (another 1.6 GB),
// clamps, and resamples the old density with a trilinear Taichi kernel (ti_get_interp_val, :25-50).  Here one kernel does
// all of it per voxel: coordinates from three small axis tables, the RK4 back-trace through the warp-collective evaluator of
// the tiled kernels (eval_tiled.cuh: 4 voxels per thread on packed FP32, lane-parallel bounding-sphere culling — a warp owns a
// 4 x 4 x 8 block of voxels, so its candidate set is as compact as it gets without any sorting), clamp, trilinear gather,
// store.  Nothing but the two density fields touches HBM: 4 B read (8 cached taps) + 4 B written per voxel.
#include "eval_tiled.cuh"

namespace gsr {

struct LatticeArgs {
	const float *xs, *ys, *zs;	// axis coordinates (NX), (NY), (NZ) — torch.linspace of the reference's get_grid_points
	int nx, ny, nz;
	int x_begin, x_end;		// the slab of x planes this launch computes (a process's share of the lattice); the whole lattice: 0, nx
	float lo[3], hi[3];		// the domain [x_min, x_max] x ... the back-traced points are clamped to, and the lattice spans
};

// ti_get_interp_val (3D/advance_density.py:25-50), the reference's arithmetic in f32
__device__ __forceinline__ float trilinear(const float *__restrict__ f, const LatticeArgs &L, float px, float py, float pz)
{
	const float dx = (L.hi[0] - L.lo[0]) / (float)(L.nx - 1), dy = (L.hi[1] - L.lo[1]) / (float)(L.ny - 1), dz = (L.hi[2] - L.lo[2]) / (float)(L.nz - 1);
	const float qx = px - L.lo[0], qy = py - L.lo[1], qz = pz - L.lo[2];
	int i = (int)floorf(__fdiv_rn(qx, dx)), j = (int)floorf(__fdiv_rn(qy, dy)), k = (int)floorf(__fdiv_rn(qz, dz));
	i = min(max(i, 0), L.nx - 1); j = min(max(j, 0), L.ny - 1); k = min(max(k, 0), L.nz - 1);	// clamped points: i <= n - 1 already; guards rounding
	const int i1 = min(i + 1, L.nx - 1), j1 = min(j + 1, L.ny - 1), k1 = min(k + 1, L.nz - 1);
	// corner_min = ti_get_coord(i, j, k) - zero_p = extent / (n - 1) * index
	const float wx = __fdiv_rn(qx - dx * (float)i, dx), wy = __fdiv_rn(qy - dy * (float)j, dy), wz = __fdiv_rn(qz - dz * (float)k, dz);
	const size_t sy = (size_t)L.nz, sx = (size_t)L.ny * L.nz;
	const float f000 = __ldg(f + i * sx + j * sy + k), f100 = __ldg(f + i1 * sx + j * sy + k), f010 = __ldg(f + i * sx + j1 * sy + k), f110 = __ldg(f + i1 * sx + j1 * sy + k);
	const float f001 = __ldg(f + i * sx + j * sy + k1), f101 = __ldg(f + i1 * sx + j * sy + k1), f011 = __ldg(f + i * sx + j1 * sy + k1), f111 = __ldg(f + i1 * sx + j1 * sy + k1);
	return f000 * (1.f - wx) * (1.f - wy) * (1.f - wz) + f100 * wx * (1.f - wy) * (1.f - wz) + f010 * (1.f - wx) * wy * (1.f - wz) + f110 * wx * wy * (1.f - wz)
	       + f001 * (1.f - wx) * (1.f - wy) * wz + f101 * wx * (1.f - wy) * wz + f011 * (1.f - wx) * wy * wz + f111 * wx * wy * wz;
}

constexpr int DN_P = 4;

// CTA = 8 x 8 x 8 voxels, 4 warps; warp w owns the 4 x 4 x 8 block at (4 (w / 2), 4 (w % 2), 0); slot p of lane l is voxel
// v = 32 p + l of that block, (v / 32, (v / 8) % 4, v % 8).  NF density fields are advected through the same back-trace.
// 80 registers -> six CTAs (24 warps) per SM: the evaluator's dependent chains want warps to hide behind (measured at 512^3:
// 53.1 ms with four CTAs / 128 registers, 49.6 with five, 46.2 with six, 45.7 with eight / 64 registers and spills)
template <int NF, bool COUNT = false, int OCC = 6>
__global__ void __launch_bounds__(128, OCC) advect_density_kernel(TiledArgs a, LatticeArgs L, float dt, const float *__restrict__ f0, const float *__restrict__ f1,
								  float *__restrict__ o0, float *__restrict__ o1)
{
	__shared__ TileSh sh;
	if (threadIdx.x == 0) sh.staged = 0;	// candidates come straight from global memory / L1 (compact warps, no ordering pass)
	__syncthreads();
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int bi = L.x_begin + blockIdx.x * 8 + 4 * (w >> 1), bj = blockIdx.y * 8 + 4 * (w & 1), bk = blockIdx.z * 8;
	constexpr int P = DN_P;
	float x0[P], x1[P], x2[P], px[P], py[P], pz[P], v[P][3], vs[P][3], dummy[P][9];
	bool ok[P];
	int vi[P], vj[P], vk[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int q = 32 * p + lane;
		vi[p] = bi + q / 32; vj[p] = bj + (q / 8) % 4; vk[p] = bk + q % 8;
		ok[p] = vi[p] < L.x_end && vj[p] < L.ny && vk[p] < L.nz;
		px[p] = x0[p] = ok[p] ? __ldg(L.xs + vi[p]) : 0.f;
		py[p] = x1[p] = ok[p] ? __ldg(L.ys + vj[p]) : 0.f;
		pz[p] = x2[p] = ok[p] ? __ldg(L.zs + vk[p]) : 0.f;
		vs[p][0] = vs[p][1] = vs[p][2] = 0.f;
	}
	const float hdt = dt * .5f, dt6 = dt / 6.f;
#pragma unroll 1
	for (int st = 0; st < 4; st++) {	// RK4 of x' = u(x), positions only (3D/GSR.py:639-658)
		warp_eval3<P, false, true, COUNT>(a, sh, nullptr, px, py, pz, ok, v, dummy);
		const float wgt = (st == 0 || st == 3) ? 1.f : 2.f, step = (st < 2) ? hdt : dt;
#pragma unroll
		for (int p = 0; p < P; p++) {
			vs[p][0] += wgt * v[p][0]; vs[p][1] += wgt * v[p][1]; vs[p][2] += wgt * v[p][2];
			px[p] = x0[p] + step * v[p][0]; py[p] = x1[p] + step * v[p][1]; pz[p] = x2[p] + step * v[p][2];
		}
	}
#pragma unroll
	for (int p = 0; p < P; p++) {
		if (!ok[p]) continue;
		const float gx = fminf(fmaxf(x0[p] + dt6 * vs[p][0], L.lo[0]), L.hi[0]), gy = fminf(fmaxf(x1[p] + dt6 * vs[p][1], L.lo[1]), L.hi[1]),
			    gz = fminf(fmaxf(x2[p] + dt6 * vs[p][2], L.lo[2]), L.hi[2]);
		const size_t o = ((size_t)vi[p] * L.ny + vj[p]) * L.nz + vk[p];
		o0[o] = trilinear(f0, L, gx, gy, gz);
		if (NF > 1) o1[o] = trilinear(f1, L, gx, gy, gz);
	}
}

}  // namespace gsr

using namespace gsr;

static int advect_density_launch(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull, const float *xs, const float *ys,
				 const float *zs, int nx, int ny, int nz, int x_begin, int x_end, const float *domain, float dt, const float *density_a,
				 const float *density_b, float *out_a, float *out_b, unsigned long long *executed, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || g.D != 3 || !cell_start || !packed || !xs || !ys || !zs || nx < 2 || ny < 2 || nz < 2 || !domain || !density_a || !out_a) return GSR_EINVAL;
	if (x_begin < 0 || x_end > nx || x_begin > x_end) return GSR_EINVAL;
	if (x_begin == x_end) return GSR_OK;
	if ((density_b != nullptr) != (out_b != nullptr) || density_a == out_a || (density_b && density_b == out_b)) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	TiledArgs a;
	a.P = make_params(g);
	a.cell_start = cell_start; a.packed = (const float4 *)packed; a.cull = cull;
	a.x = nullptr; a.Q = 0; a.perm = nullptr; a.scs = nullptr; a.tile_row = nullptr; a.cap = 0;
	a.exec_count = executed;
	LatticeArgs L;
	L.xs = xs; L.ys = ys; L.zs = zs; L.nx = nx; L.ny = ny; L.nz = nz; L.x_begin = x_begin; L.x_end = x_end;
	for (int k = 0; k < 3; k++) { L.lo[k] = domain[2 * k]; L.hi[k] = domain[2 * k + 1]; }
	dim3 grid((x_end - x_begin + 7) / 8, (ny + 7) / 8, (nz + 7) / 8);
	g_launches += 1;
	if (executed) {	// census variant: same arithmetic, plus the count of pair tests the culling lets through
		if (density_b) advect_density_kernel<2, true><<<grid, 128, 0, st>>>(a, L, dt, density_a, density_b, out_a, out_b);
		else advect_density_kernel<1, true><<<grid, 128, 0, st>>>(a, L, dt, density_a, nullptr, out_a, nullptr);
	} else if (density_b) advect_density_kernel<2><<<grid, 128, 0, st>>>(a, L, dt, density_a, density_b, out_a, out_b);
	else advect_density_kernel<1><<<grid, 128, 0, st>>>(a, L, dt, density_a, nullptr, out_a, nullptr);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_advect_density_slab(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull,
				       const float *xs, const float *ys, const float *zs, int nx, int ny, int nz, int x_begin, int x_end, const float *domain, float dt,
				       const float *density_a, const float *density_b, float *out_a, float *out_b, void *stream)
{
	return advect_density_launch(d, cell_start, packed, cull, xs, ys, zs, nx, ny, nz, x_begin, x_end, domain, dt, density_a, density_b, out_a, out_b, nullptr, stream);
}

extern "C" int gsr_advect_density_census(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull,
					 const float *xs, const float *ys, const float *zs, int nx, int ny, int nz, int x_begin, int x_end, const float *domain, float dt,
					 const float *density_a, const float *density_b, float *out_a, float *out_b, unsigned long long *executed_pair_tests, void *stream)
{
	if (!executed_pair_tests) return GSR_EINVAL;
	return advect_density_launch(d, cell_start, packed, cull, xs, ys, zs, nx, ny, nz, x_begin, x_end, domain, dt, density_a, density_b, out_a, out_b,
				     executed_pair_tests, stream);
}

extern "C" int gsr_advect_density(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull,
				  const float *xs, const float *ys, const float *zs, int nx, int ny, int nz, const float *domain, float dt,
				  const float *density_a, const float *density_b, float *out_a, float *out_b, void *stream)
{
	return gsr_advect_density_slab(d, cell_start, packed, cull, xs, ys, zs, nx, ny, nz, 0, nx, domain, dt, density_a, density_b, out_a, out_b, stream);
}
