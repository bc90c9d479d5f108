// eval.cu — forward evaluation, RK4 advection / pull-back, advected-covector reference, neighbour marking and
// the work census.  Points are visited in cell-sorted order (perm) so that the
// lanes of a warp walk (nearly) the same runs of packed Gaussians: loads are warp-broadcast float4 and the loop
// trip counts are (nearly) uniform.  FP32-issue bound at large Q; at small Q the LANES = 8 / 32 variants spread
// one point over several lanes to fill the machine (see eval.cuh).
#include "eval.cuh"
#include <math.h>
#include <string.h>

namespace gsr {

// h_thr = min{ h (f32) : expf(h) >= tau }, found by bisection over the f32 bit patterns with libm's expf —
// the same expf the CPU oracle's `gaussian >= tau` test goes through.
float host_h_threshold(float tau)
{
	if (!(tau > 0.f)) return -INFINITY;
	if (tau > 1.f) return INFINITY;	// expf(h) of h <= 0 never reaches it; a pair can never be accepted
	// h in [-128, 0]; as bit patterns of negative floats, larger magnitude = larger integer
	uint32_t lo = 0x80000000u /* -0 */, hi = 0xC3000000u /* -128 */;
	// invariant: expf(lo) >= tau, expf(hi) < tau
	while (hi - lo > 1) {
		uint32_t mid = lo + (hi - lo) / 2;
		float h;
		memcpy(&h, &mid, 4);
		if (expf(h) >= tau) lo = mid; else hi = mid;
	}
	float h;
	memcpy(&h, &lo, 4);
	return h;
}

constexpr int EV_THREADS = 128;

template <int D, bool NEED_VAL, bool NEED_GRAD, bool ACCUM, int LANES>
__global__ void __launch_bounds__(EV_THREADS) forward_kernel(EvalParams P, const int32_t *__restrict__ cell_start, const float4 *__restrict__ packed,
							     const float *__restrict__ x, int Q, const int32_t *__restrict__ perm,
							     float *__restrict__ val, float *__restrict__ grad)
{
	const int gt = blockIdx.x * EV_THREADS + threadIdx.x;
	const int t = gt / LANES, lane = gt % LANES;
	if (LANES == 1 && t >= Q) return;
	const bool valid = t < Q;
	const int tt = valid ? t : Q - 1;
	const int j = perm ? perm[tt] : tt;
	float u[D], G[D * D];
	if (D == 3) eval_point3<NEED_GRAD, LANES>(P, cell_start, packed, x[3 * (size_t)j], x[3 * (size_t)j + 1], x[3 * (size_t)j + 2], lane, u, G);
	else eval_point2<NEED_GRAD, LANES>(P, cell_start, packed, x[2 * (size_t)j], x[2 * (size_t)j + 1], lane, u, G);
	if (!valid || lane != 0) return;
	if (NEED_VAL) {
#pragma unroll
		for (int k = 0; k < D; k++) {
			float *o = val + (size_t)D * j + k;
			*o = ACCUM ? *o + u[k] : u[k];
		}
	}
	if (NEED_GRAD) {
#pragma unroll
		for (int k = 0; k < D * D; k++) {
			float *o = grad + (size_t)D * D * j + k;
			*o = ACCUM ? *o + G[k] : G[k];
		}
	}
}

// RK4 of x' = u(x) with the chained stage Jacobians (3D/GSR.py:639-665; 2D/GSR.py:554-580).
// MODE 0: position only; 1: + deformation, value and gradient at the end point (5 evaluations);
// MODE 2: the advected-covector reference (3D/advance.py:35-47): omega_ref = Dpsi^-1 curl, hel_ref = u.curl.
template <int MODE, int LANES>
__global__ void __launch_bounds__(EV_THREADS) rk4_3d_kernel(EvalParams P, const int32_t *__restrict__ cell_start, const float4 *__restrict__ packed,
							    const float *__restrict__ start, int Q, const int32_t *__restrict__ perm, float dt,
							    float *__restrict__ goal_pos, float *__restrict__ deformation, float *__restrict__ goal_val, float *__restrict__ goal_grad,
							    float *__restrict__ ref_vor, float *__restrict__ ref_hel)
{
	constexpr bool FULL = MODE != 0;
	const int gt = blockIdx.x * EV_THREADS + threadIdx.x;
	const int t = gt / LANES, lane = gt % LANES;
	if (LANES == 1 && t >= Q) return;
	const bool valid = t < Q;
	const int tt = valid ? t : Q - 1;
	const size_t j = perm ? perm[tt] : tt;
	const float x0 = start[3 * j], x1 = start[3 * j + 1], x2 = start[3 * j + 2];
	const float hdt = dt * .5f, dt6 = dt / 6.f;
	float v[3], dv[9], vs[3], A[9], B[9], S[9];
	// stage 0
	eval_point3<FULL, LANES>(P, cell_start, packed, x0, x1, x2, lane, v, dv);
	vs[0] = v[0]; vs[1] = v[1]; vs[2] = v[2];
	if (FULL) {
#pragma unroll
		for (int k = 0; k < 9; k++) { S[k] = dv[k]; A[k] = ((k % 4 == 0) ? 1.f : 0.f) + hdt * dv[k]; }	// A = dphi1
	}
	// stage 1
	eval_point3<FULL, LANES>(P, cell_start, packed, x0 + hdt * v[0], x1 + hdt * v[1], x2 + hdt * v[2], lane, v, dv);
	vs[0] += 2.f * v[0]; vs[1] += 2.f * v[1]; vs[2] += 2.f * v[2];
	if (FULL) {
		mm3(dv, A, B);	// dv1 @ dphi1
#pragma unroll
		for (int k = 0; k < 9; k++) { S[k] += 2.f * B[k]; A[k] = ((k % 4 == 0) ? 1.f : 0.f) + hdt * B[k]; }	// A = dphi2
	}
	// stage 2
	eval_point3<FULL, LANES>(P, cell_start, packed, x0 + hdt * v[0], x1 + hdt * v[1], x2 + hdt * v[2], lane, v, dv);
	vs[0] += 2.f * v[0]; vs[1] += 2.f * v[1]; vs[2] += 2.f * v[2];
	if (FULL) {
		mm3(dv, A, B);	// dv2 @ dphi2
#pragma unroll
		for (int k = 0; k < 9; k++) { S[k] += 2.f * B[k]; A[k] = ((k % 4 == 0) ? 1.f : 0.f) + dt * B[k]; }	// A = dphi3
	}
	// stage 3
	eval_point3<FULL, LANES>(P, cell_start, packed, x0 + dt * v[0], x1 + dt * v[1], x2 + dt * v[2], lane, v, dv);
	vs[0] += v[0]; vs[1] += v[1]; vs[2] += v[2];
	const float p0 = x0 + dt6 * vs[0], p1 = x1 + dt6 * vs[1], p2 = x2 + dt6 * vs[2];
	const bool writer = valid && lane == 0;
	if (MODE != 2 && writer) {
		goal_pos[3 * j] = p0; goal_pos[3 * j + 1] = p1; goal_pos[3 * j + 2] = p2;
	}
	if (FULL) {
		mm3(dv, A, B);	// dv3 @ dphi3
#pragma unroll
		for (int k = 0; k < 9; k++) S[k] = ((k % 4 == 0) ? 1.f : 0.f) + dt6 * (S[k] + B[k]);	// S = dphi
		eval_point3<true, LANES>(P, cell_start, packed, p0, p1, p2, lane, v, dv);
		if (!writer) return;
		if (MODE == 1) {
#pragma unroll
			for (int k = 0; k < 9; k++) { deformation[9 * j + k] = S[k]; goal_grad[9 * j + k] = dv[k]; }
			goal_val[3 * j] = v[0]; goal_val[3 * j + 1] = v[1]; goal_val[3 * j + 2] = v[2];
		} else {
			const float w0 = dv[7] - dv[5], w1 = dv[2] - dv[6], w2 = dv[3] - dv[1];	// curl of the pulled-back Jacobian
			if (ref_hel) ref_hel[j] = v[0] * w0 + v[1] * w1 + v[2] * w2;
			// omega_ref = S^-1 w  via the adjugate
			const float c00 = S[4] * S[8] - S[5] * S[7], c01 = S[2] * S[7] - S[1] * S[8], c02 = S[1] * S[5] - S[2] * S[4];
			const float c10 = S[5] * S[6] - S[3] * S[8], c11 = S[0] * S[8] - S[2] * S[6], c12 = S[2] * S[3] - S[0] * S[5];
			const float c20 = S[3] * S[7] - S[4] * S[6], c21 = S[1] * S[6] - S[0] * S[7], c22 = S[0] * S[4] - S[1] * S[3];
			const float det = S[0] * c00 + S[1] * c10 + S[2] * c20;
			const float inv = 1.f / det;
			ref_vor[3 * j] = (c00 * w0 + c01 * w1 + c02 * w2) * inv;
			ref_vor[3 * j + 1] = (c10 * w0 + c11 * w1 + c12 * w2) * inv;
			ref_vor[3 * j + 2] = (c20 * w0 + c21 * w1 + c22 * w2) * inv;
		}
	}
}

// 2D: MODE 2 = 2D/advance.py:46-54 (omega_ref = curl at the back-traced point, zero where it left the domain).
template <int MODE, int LANES>
__global__ void __launch_bounds__(EV_THREADS) rk4_2d_kernel(EvalParams P, const int32_t *__restrict__ cell_start, const float4 *__restrict__ packed,
							    const float *__restrict__ start, int Q, const int32_t *__restrict__ perm, float dt,
							    float *__restrict__ goal_pos, float *__restrict__ deformation, float *__restrict__ goal_val, float *__restrict__ goal_grad,
							    float *__restrict__ ref_vor, float4 dom, int use_dom)
{
	constexpr bool CHAIN = MODE == 1;
	const int gt = blockIdx.x * EV_THREADS + threadIdx.x;
	const int t = gt / LANES, lane = gt % LANES;
	if (LANES == 1 && t >= Q) return;
	const bool valid = t < Q;
	const int tt = valid ? t : Q - 1;
	const size_t j = perm ? perm[tt] : tt;
	const float x0 = start[2 * j], x1 = start[2 * j + 1];
	const float hdt = dt * .5f, dt6 = dt / 6.f;
	float v[2], dv[4], vs[2], A[4], B[4], S[4];
	const float I[4] = {1.f, 0.f, 0.f, 1.f};
	eval_point2<CHAIN, LANES>(P, cell_start, packed, x0, x1, lane, v, dv);
	vs[0] = v[0]; vs[1] = v[1];
	if (CHAIN) {
#pragma unroll
		for (int k = 0; k < 4; k++) { S[k] = dv[k]; A[k] = I[k] + hdt * dv[k]; }
	}
	eval_point2<CHAIN, LANES>(P, cell_start, packed, x0 + hdt * v[0], x1 + hdt * v[1], lane, v, dv);
	vs[0] += 2.f * v[0]; vs[1] += 2.f * v[1];
	if (CHAIN) {
		mm2(dv, A, B);
#pragma unroll
		for (int k = 0; k < 4; k++) { S[k] += 2.f * B[k]; A[k] = I[k] + hdt * B[k]; }
	}
	eval_point2<CHAIN, LANES>(P, cell_start, packed, x0 + hdt * v[0], x1 + hdt * v[1], lane, v, dv);
	vs[0] += 2.f * v[0]; vs[1] += 2.f * v[1];
	if (CHAIN) {
		mm2(dv, A, B);
#pragma unroll
		for (int k = 0; k < 4; k++) { S[k] += 2.f * B[k]; A[k] = I[k] + dt * B[k]; }
	}
	eval_point2<CHAIN, LANES>(P, cell_start, packed, x0 + dt * v[0], x1 + dt * v[1], lane, v, dv);
	vs[0] += v[0]; vs[1] += v[1];
	const float p0 = x0 + dt6 * vs[0], p1 = x1 + dt6 * vs[1];
	const bool writer = valid && lane == 0;
	if (MODE != 2 && writer) { goal_pos[2 * j] = p0; goal_pos[2 * j + 1] = p1; }
	if (MODE == 1) {
		mm2(dv, A, B);
		if (writer) {
#pragma unroll
			for (int k = 0; k < 4; k++) deformation[4 * j + k] = I[k] + dt6 * (S[k] + B[k]);
		}
	}
	if (MODE != 0) {
		eval_point2<true, LANES>(P, cell_start, packed, p0, p1, lane, v, dv);
		if (!writer) return;
		if (MODE == 1) {
			goal_val[2 * j] = v[0]; goal_val[2 * j + 1] = v[1];
#pragma unroll
			for (int k = 0; k < 4; k++) goal_grad[4 * j + k] = dv[k];
		} else {
			float w = dv[2] - dv[1];
			if (use_dom && (p0 < dom.x || p0 > dom.y || p1 < dom.z || p1 > dom.w)) w = 0.f;
			ref_vor[j] = w;
		}
	}
}

// get_all_neighbors_ti (3D/GSR.py:679-690; 2D/GSR.py:620-630): mark[i] = 1 if some query point lies within
// grid_scale of mu_i.  The distance is formed unfused in the reference's order so the integer output is exact.
template <int D>
__global__ void __launch_bounds__(EV_THREADS) mark_kernel(Grid g, const int32_t *__restrict__ cell_start, const int32_t *__restrict__ sorted_id,
							  const float4 *__restrict__ packed, const float *__restrict__ x, int Q, int32_t *__restrict__ mark)
{
	int j = blockIdx.x * EV_THREADS + threadIdx.x;
	if (j >= Q) return;
	float px = x[(size_t)D * j], py = x[(size_t)D * j + 1], pz = (D == 3) ? x[(size_t)D * j + 2] : 0.f;
	const float gs = grid_gs(g);
	int cx = cell_coord(px, g.lo[0], gs), cy = cell_coord(py, g.lo[1], gs), cz = (D == 3) ? cell_coord(pz, g.lo[2], gs) : 0;
	constexpr int REC = (D == 3) ? 3 : 2;
	for (int gi = max(cx - 1, 0); gi <= min(cx + 1, g.dims[0] - 1); gi++)
		for (int gj = max(cy - 1, 0); gj <= min(cy + 1, g.dims[1] - 1); gj++)
			for (int gk = (D == 3 ? max(cz - 1, 0) : 0); gk <= (D == 3 ? min(cz + 1, g.dims[2] - 1) : 0); gk++) {
				int c = (gi * g.dims[1] + gj) * g.dims[2] + gk;
				for (int t = cell_start[c]; t < cell_start[c + 1]; t++) {
					float4 p0 = packed[REC * (size_t)t];
					float dx = __fsub_rn(px, p0.x), dy = __fsub_rn(py, p0.y);
					float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
					if (D == 3) {
						float dz = __fsub_rn(pz, p0.z);
						d2 = __fadd_rn(d2, __fmul_rn(dz, dz));
					}
					if (sqrtf(d2) <= gs) mark[sorted_id[t]] = 1;
				}
			}
}

// ---- work census (bench.py): C = candidate visits (occupancy of each sample's stencil), P = accepted pairs ----
template <int D, bool WITH_P>
__global__ void __launch_bounds__(EV_THREADS) count_kernel(EvalParams P_, const int32_t *__restrict__ cell_start, const float4 *__restrict__ packed,
							   const float *__restrict__ x, int Q, unsigned long long *out)
{
	const Grid &g = P_.g;
	int j = blockIdx.x * EV_THREADS + threadIdx.x;
	unsigned long long c = 0, p = 0;
	if (j < Q) {
		const float gs = grid_gs(g);
		const float px = x[(size_t)D * j], py = x[(size_t)D * j + 1], pz = (D == 3) ? x[(size_t)D * j + 2] : 0.f;
		int cx = cell_coord(px, g.lo[0], gs), cy = cell_coord(py, g.lo[1], gs), cz = (D == 3) ? cell_coord(pz, g.lo[2], gs) : 0;
		const int last = (D == 3) ? 2 : 1;	// the fastest axis: runs along it are contiguous
		const int cl = (D == 3) ? cz : cy;
		const int lo = max(cl - 1, 0), hi = min(cl + 1, g.dims[last] - 1);
		if (lo <= hi)
			for (int gi = max(cx - 1, 0); gi <= min(cx + 1, g.dims[0] - 1); gi++)
				for (int gj = (D == 3 ? max(cy - 1, 0) : 0); gj <= (D == 3 ? min(cy + 1, g.dims[1] - 1) : 0); gj++) {
					const int base = (D == 3) ? (gi * g.dims[1] + gj) * g.dims[2] : gi * g.dims[1];
					const int s = cell_start[base + lo], e = cell_start[base + hi + 1];
					c += (unsigned long long)(e - s);
					if (WITH_P)
						for (int t = s; t < e; t++) {
							float q;
							if (D == 3) {
								const float4 p0 = packed[3 * t], p1 = packed[3 * t + 1], p2 = packed[3 * t + 2];
								const float dx = px - p0.x, dy = py - p0.y, dz = pz - p0.z;
								const float wx = p1.x * dx + p1.y * dy + p1.z * dz, wy = p1.y * dx + p2.x * dy + p2.y * dz, wz = p1.z * dx + p2.y * dy + p2.z * dz;
								q = dx * wx + dy * wy + dz * wz;
							} else {
								const float4 p0 = packed[2 * t], p1 = packed[2 * t + 1];
								const float dx = px - p0.x, dy = py - p0.y;
								q = dx * (p1.x * dx + p1.y * dy) + dy * (p1.y * dx + p1.z * dy);
							}
							p += (q <= P_.q_thr) ? 1 : 0;
						}
				}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		c += __shfl_xor_sync(0xffffffffu, c, o);
		p += __shfl_xor_sync(0xffffffffu, p, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (c) atomicAdd(out, c);
		if (WITH_P && p) atomicAdd(out + 1, p);
	}
}

int g_lanes8_min_n = 1 << 14;	// items (points or Gaussians) from which 8 lanes share one item instead of a warp

static inline int blocks_for(int64_t Q, int lanes) { return (int)((Q * lanes + EV_THREADS - 1) / EV_THREADS); }

// eval_tiled.cu
bool use_tiled(const Grid &g, int64_t Q, const int32_t *perm, const int32_t *scs, const int32_t *tile_row);
int launch_forward_tiled3(const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
			  const int32_t *scs, const int32_t *tile_row, float *val, float *grad, bool accumulate, cudaStream_t st);
int launch_rk4_tiled3(int mode, const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
		      const int32_t *scs, const int32_t *tile_row, float dt, float *goal_pos, float *deformation, float *goal_val, float *goal_grad,
		      float *ref_vor, float *ref_hel, cudaStream_t st);

}  // namespace gsr

using namespace gsr;

#define LANES_SWITCH(L, ...)                                                \
	do {                                                                \
		if ((L) == 1) { constexpr int LN = 1; __VA_ARGS__; }        \
		else if ((L) == 8) { constexpr int LN = 8; __VA_ARGS__; }   \
		else { constexpr int LN = 32; __VA_ARGS__; }                \
	} while (0)

template <int D, bool ACC, int LN>
static void launch_forward(const EvalParams &P, const int32_t *cs, const float *packed, const float *x, int Q, const int32_t *perm,
			   float *val, float *grad, cudaStream_t st)
{
	int blocks = blocks_for(Q, LN);
	const float4 *pk = (const float4 *)packed;
	if (val && grad) forward_kernel<D, true, true, ACC, LN><<<blocks, EV_THREADS, 0, st>>>(P, cs, pk, x, Q, perm, val, grad);
	else if (val) forward_kernel<D, true, false, ACC, LN><<<blocks, EV_THREADS, 0, st>>>(P, cs, pk, x, Q, perm, val, grad);
	else forward_kernel<D, false, true, ACC, LN><<<blocks, EV_THREADS, 0, st>>>(P, cs, pk, x, Q, perm, val, grad);
}

extern "C" int gsr_forward(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull,
			   const float *x, int64_t Q, const int32_t *perm, const int32_t *scs, const int32_t *tile_row,
			   float *val, float *grad, int accumulate, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || Q >= ((int64_t)1 << 26) || (!val && !grad) || !cell_start || !packed) return GSR_EINVAL;
	if (Q == 0) return GSR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	EvalParams P = make_params(g);
	g_launches += 1;
	if (use_tiled(g, Q, perm, scs, tile_row)) return launch_forward_tiled3(P, cell_start, packed, cull, x, Q, perm, scs, tile_row, val, grad, accumulate != 0, st);
	const int L = pick_lanes(Q);
	if (g.D == 3) {
		if (accumulate) LANES_SWITCH(L, (launch_forward<3, true, LN>(P, cell_start, packed, x, (int)Q, perm, val, grad, st)));
		else LANES_SWITCH(L, (launch_forward<3, false, LN>(P, cell_start, packed, x, (int)Q, perm, val, grad, st)));
	} else {
		if (accumulate) LANES_SWITCH(L, (launch_forward<2, true, LN>(P, cell_start, packed, x, (int)Q, perm, val, grad, st)));
		else LANES_SWITCH(L, (launch_forward<2, false, LN>(P, cell_start, packed, x, (int)Q, perm, val, grad, st)));
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_rk4(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull,
		       const float *start, int64_t Q, const int32_t *perm, const int32_t *scs, const int32_t *tile_row, float dt,
		       float *goal_pos, float *deformation, float *goal_val, float *goal_grad, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || Q >= ((int64_t)1 << 26) || !goal_pos || !cell_start || !packed) return GSR_EINVAL;
	if (Q == 0) return GSR_OK;
	// the reference computes the end-point value/gradient only when BOTH outputs are requested, and the
	// deformation independently (3D/GSR.py:650, :660); its callers ask for all three or none.
	bool full = deformation && goal_val && goal_grad;
	if (!full && (deformation || goal_val || goal_grad)) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	EvalParams P = make_params(g);
	const float4 *pk = (const float4 *)packed;
	g_launches += 1;
	if (use_tiled(g, Q, perm, scs, tile_row))
		return launch_rk4_tiled3(full ? 1 : 0, P, cell_start, packed, cull, start, Q, perm, scs, tile_row, dt, goal_pos, deformation, goal_val, goal_grad, nullptr, nullptr, st);
	const int L = pick_lanes(Q);
	const float4 dom = make_float4(0, 0, 0, 0);
	if (g.D == 3) {
		if (full) LANES_SWITCH(L, (rk4_3d_kernel<1, LN><<<blocks_for(Q, LN), EV_THREADS, 0, st>>>(P, cell_start, pk, start, (int)Q, perm, dt, goal_pos, deformation, goal_val, goal_grad, nullptr, nullptr)));
		else LANES_SWITCH(L, (rk4_3d_kernel<0, LN><<<blocks_for(Q, LN), EV_THREADS, 0, st>>>(P, cell_start, pk, start, (int)Q, perm, dt, goal_pos, nullptr, nullptr, nullptr, nullptr, nullptr)));
	} else {
		if (full) LANES_SWITCH(L, (rk4_2d_kernel<1, LN><<<blocks_for(Q, LN), EV_THREADS, 0, st>>>(P, cell_start, pk, start, (int)Q, perm, dt, goal_pos, deformation, goal_val, goal_grad, nullptr, dom, 0)));
		else LANES_SWITCH(L, (rk4_2d_kernel<0, LN><<<blocks_for(Q, LN), EV_THREADS, 0, st>>>(P, cell_start, pk, start, (int)Q, perm, dt, goal_pos, nullptr, nullptr, nullptr, nullptr, dom, 0)));
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_advected_vorticity(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *cull,
				      const float *x, int64_t Q, const int32_t *perm, const int32_t *scs, const int32_t *tile_row, float dt, const float *domain,
				      float *ref_vor, float *ref_hel, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || Q >= ((int64_t)1 << 26) || !ref_vor || !cell_start || !packed) return GSR_EINVAL;
	if (Q == 0) return GSR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	EvalParams P = make_params(g);
	const float4 *pk = (const float4 *)packed;
	g_launches += 1;
	if (use_tiled(g, Q, perm, scs, tile_row))
		return launch_rk4_tiled3(2, P, cell_start, packed, cull, x, Q, perm, scs, tile_row, dt, nullptr, nullptr, nullptr, nullptr, ref_vor, ref_hel, st);
	const int L = pick_lanes(Q);
	if (g.D == 3) {
		LANES_SWITCH(L, (rk4_3d_kernel<2, LN><<<blocks_for(Q, LN), EV_THREADS, 0, st>>>(P, cell_start, pk, x, (int)Q, perm, dt, nullptr, nullptr, nullptr, nullptr, ref_vor, ref_hel)));
	} else {
		const float4 dom = domain ? make_float4(domain[0], domain[1], domain[2], domain[3]) : make_float4(0, 0, 0, 0);
		const int use = domain ? 1 : 0;
		LANES_SWITCH(L, (rk4_2d_kernel<2, LN><<<blocks_for(Q, LN), EV_THREADS, 0, st>>>(P, cell_start, pk, x, (int)Q, perm, dt, nullptr, nullptr, nullptr, nullptr, ref_vor, dom, use)));
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_mark_neighbors(const gsr_grid_desc *d, const int32_t *cell_start, const int32_t *sorted_id, const float *packed,
				  const float *x, int64_t Q, int32_t *mark, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || !mark || !cell_start || !sorted_id || !packed) return GSR_EINVAL;
	if (Q == 0) return GSR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	int blocks = (int)((Q + EV_THREADS - 1) / EV_THREADS);
	g_launches += 1;
	if (g.D == 3) mark_kernel<3><<<blocks, EV_THREADS, 0, st>>>(g, cell_start, sorted_id, (const float4 *)packed, x, (int)Q, mark);
	else mark_kernel<2><<<blocks, EV_THREADS, 0, st>>>(g, cell_start, sorted_id, (const float4 *)packed, x, (int)Q, mark);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_count_pairs(const gsr_grid_desc *d, const int32_t *cell_start, const float *packed, const float *x, int64_t Q, uint64_t *counts, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || !counts || !cell_start) return GSR_EINVAL;
	if (Q == 0) return GSR_OK;
	cudaStream_t st = (cudaStream_t)stream;
	EvalParams P = make_params(g);
	int blocks = (int)((Q + EV_THREADS - 1) / EV_THREADS);
	unsigned long long *o = (unsigned long long *)counts;
	const float4 *pk = (const float4 *)packed;
	g_launches += 1;
	if (g.D == 3) {
		if (packed) count_kernel<3, true><<<blocks, EV_THREADS, 0, st>>>(P, cell_start, pk, x, (int)Q, o);
		else count_kernel<3, false><<<blocks, EV_THREADS, 0, st>>>(P, cell_start, pk, x, (int)Q, o);
	} else {
		if (packed) count_kernel<2, true><<<blocks, EV_THREADS, 0, st>>>(P, cell_start, pk, x, (int)Q, o);
		else count_kernel<2, false><<<blocks, EV_THREADS, 0, st>>>(P, cell_start, pk, x, (int)Q, o);
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
