// hash_small.cuh — device pieces of the spatial hash shared by grid.cu and the fused small-problem step kernel (step.cu):
// cell keys, the per-Gaussian packed record, and the single-CTA hash build.
#pragma once
#include "common.cuh"
#include <math.h>

namespace gsr {

// Gaussian keys: row-major cell index, or ncell for Gaussians outside the extended domain
// (the reference silently drops those from the hash: 3D/GSR.py:212).
template <int D>
__device__ __forceinline__ uint32_t gauss_key_vals(const float *pt, const Grid &g, float gs)
{
	bool in = true;
	int c[3] = {0, 0, 0};
#pragma unroll
	for (int k = 0; k < D; k++) {
		float p = pt[k];
		in = in && (g.lo[k] <= p) && (p <= g.hi[k]);
		c[k] = cell_coord(p, g.lo[k], gs);
	}
	// quirk B.8 of the survey: an index == dims is possible at the upper face in the reference (unchecked
	// write there).  We treat any out-of-grid index as "not in the hash" instead of writing out of bounds.
#pragma unroll
	for (int k = 0; k < D; k++) in = in && c[k] >= 0 && c[k] < g.dims[k];
	uint32_t key = (uint32_t)g.ncell;
	if (in) key = (uint32_t)((c[0] * g.dims[1] + c[1]) * g.dims[2] + c[2]);
	return key;
}

template <int D>
__device__ __forceinline__ uint32_t gauss_key(const float *__restrict__ pos, int i, const Grid &g)
{
	float pt[3] = {0.f, 0.f, 0.f};
#pragma unroll
	for (int k = 0; k < D; k++) pt[k] = pos[(size_t)D * i + k];
	return gauss_key_vals<D>(pt, g, grid_gs(g));
}


// Sample keys on the padded grid (dims+2): a sample whose cell index is -1 or dims still sees the border
// cells through the reference's clamped stencil (3D/GSR.py:272-274), anything further out sees nothing.
// FINE: the key is extended by 2 bits per axis of sub-cell position (a 4^D raster inside the cell), so that consecutive
// sorted samples are spatially compact — the warps of the tiled evaluation kernels then reject most candidates as a whole.
template <int D, bool FINE>
__device__ __forceinline__ uint32_t sample_key_of(const float (&pt)[3], const Grid &g)
{
	bool ok = true;
	int c[3] = {-1, -1, -1};
	uint32_t sub = 0;
	const float gs = grid_gs(g);
#pragma unroll
	for (int k = 0; k < D; k++) {
		const float xv = pt[k];
		c[k] = cell_coord(xv, g.lo[k], gs);
		ok = ok && c[k] >= -1 && c[k] <= g.dims[k];
		if (FINE) {
			const float f = (xv - g.lo[k]) / gs - (float)c[k];	// ordering only: any rounding here is harmless
			sub = sub * 4u + (uint32_t)min(max((int)(f * 4.f), 0), 3);
		}
	}
	uint32_t key = (uint32_t)g.pcell;
	if (ok) key = (uint32_t)(((c[0] + 1) * g.pdims[1] + (c[1] + 1)) * g.pdims[2] + (c[2] + 1));
	if (FINE) key = (key << (2 * D)) | (ok ? sub : 0u);
	return key;
}

template <int D, bool FINE>
__device__ __forceinline__ uint32_t sample_key(const float *__restrict__ x, int i, const Grid &g)
{
	float pt[3] = {0.f, 0.f, 0.f};
#pragma unroll
	for (int k = 0; k < D; k++) pt[k] = x[(size_t)D * i + k];
	return sample_key_of<D, FINE>(pt, g);
}


// exp(2 s) rounded once from double: matches a correctly-rounded expf (glibc) bit-for-bit in practice
__device__ __forceinline__ float exp2s(float s) { return (float)exp(2.0 * (double)s); }

// 3D record (3 x float4): {mu.x, mu.y, mu.z, v.x} {A00, A01, A02, v.y} {A11, A12, A22, v.z},  A = Sigma^-1.
// R(q), S^2 and R S^2 R^T are formed in the reference's operation order (3D/GSR.py:278-289), unfused.
// cull[t] = (1 + margin) / lambda_min(Sigma^-1) = (1 + margin) exp(-2 min_k s_k): a point farther than sqrt(q_thr * cull) from
// mu has q = d^T Sigma^-1 d >= lambda_min |d|^2 > q_thr, i.e. is certainly rejected (the tiled kernels' warp-level culling).
// The margin covers the rounding of Sigma^-1, of q and of the box distance (a few ulp times the condition number).
__device__ __forceinline__ float cull_coef(float smin, float smax)
{
	const float kappa = expf(2.f * (smax - smin));
	return expf(-2.f * smin) * (1.f + 1e-4f + 8e-6f * kappa);
}

// the record from values: p, s = the Gaussian's position and log inverse radii, S[k] = exp2s(s[k]), r = quaternion, v = weight
__device__ __forceinline__ void pack3d_core(const float *p, const float *s, const float *S, float4 r, const float *v, int t, float4 *__restrict__ packed,
					    float *__restrict__ cull)
{
	float len = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.x, r.x), __fmul_rn(r.y, r.y)), __fmul_rn(r.z, r.z)), __fmul_rn(r.w, r.w)));
	float q0 = __fdiv_rn(r.x, len), q1 = __fdiv_rn(r.y, len), q2 = __fdiv_rn(r.z, len), q3 = __fdiv_rn(r.w, len);
#define MUL __fmul_rn
#define ADD __fadd_rn
#define SUB __fsub_rn
	float R[3][3];
	R[0][0] = SUB(1.f, MUL(2.f, ADD(MUL(q2, q2), MUL(q3, q3))));
	R[0][1] = MUL(2.f, SUB(MUL(q1, q2), MUL(q0, q3)));
	R[0][2] = MUL(2.f, ADD(MUL(q1, q3), MUL(q0, q2)));
	R[1][0] = MUL(2.f, ADD(MUL(q1, q2), MUL(q0, q3)));
	R[1][1] = SUB(1.f, MUL(2.f, ADD(MUL(q1, q1), MUL(q3, q3))));
	R[1][2] = MUL(2.f, SUB(MUL(q2, q3), MUL(q0, q1)));
	R[2][0] = MUL(2.f, SUB(MUL(q1, q3), MUL(q0, q2)));
	R[2][1] = MUL(2.f, ADD(MUL(q2, q3), MUL(q0, q1)));
	R[2][2] = SUB(1.f, MUL(2.f, ADD(MUL(q1, q1), MUL(q2, q2))));
	float A[3][3];
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int b = a; b < 3; b++)
			A[a][b] = ADD(ADD(MUL(MUL(R[a][0], S[0]), R[b][0]), MUL(MUL(R[a][1], S[1]), R[b][1])), MUL(MUL(R[a][2], S[2]), R[b][2]));
#undef MUL
#undef ADD
#undef SUB
	packed[3 * (size_t)t + 0] = make_float4(p[0], p[1], p[2], v[0]);
	packed[3 * (size_t)t + 1] = make_float4(A[0][0], A[0][1], A[0][2], v[1]);
	packed[3 * (size_t)t + 2] = make_float4(A[1][1], A[1][2], A[2][2], v[2]);
	if (cull) cull[t] = cull_coef(fminf(s[0], fminf(s[1], s[2])), fmaxf(s[0], fmaxf(s[1], s[2])));
}

__device__ __forceinline__ void pack3d_one(const float *__restrict__ pos, const float *__restrict__ scal, const float *__restrict__ rot,
					   const float *__restrict__ vals, int t, int i, float4 *__restrict__ packed, float *__restrict__ cull)
{
	const float4 r = reinterpret_cast<const float4 *>(rot)[i];
	const float s[3] = {scal[3 * (size_t)i], scal[3 * (size_t)i + 1], scal[3 * (size_t)i + 2]};
	const float S[3] = {exp2s(s[0]), exp2s(s[1]), exp2s(s[2])};
	pack3d_core(pos + 3 * (size_t)i, s, S, r, vals + 3 * (size_t)i, t, packed, cull);
}


// 2D record (2 x float4): {mu.x, mu.y, v.x, v.y} {A00, A01, A11, 0},  A = R(theta) diag(e^{2s}) R^T (2D/GSR.py:275-277)
// c, s = (float)cos / sin of the angle in double, S0, S1 = exp2s of the two log inverse radii sc[0], sc[1]
__device__ __forceinline__ void pack2d_core(const float *p, const float *sc, float S0, float S1, float c, float s, const float *v, int t,
					    float4 *__restrict__ packed, float *__restrict__ cull)
{
	float R[2][2] = {{c, -s}, {s, c}};
	float A00 = __fadd_rn(__fmul_rn(__fmul_rn(R[0][0], S0), R[0][0]), __fmul_rn(__fmul_rn(R[0][1], S1), R[0][1]));
	float A01 = __fadd_rn(__fmul_rn(__fmul_rn(R[0][0], S0), R[1][0]), __fmul_rn(__fmul_rn(R[0][1], S1), R[1][1]));
	float A11 = __fadd_rn(__fmul_rn(__fmul_rn(R[1][0], S0), R[1][0]), __fmul_rn(__fmul_rn(R[1][1], S1), R[1][1]));
	packed[2 * (size_t)t + 0] = make_float4(p[0], p[1], v[0], v[1]);
	packed[2 * (size_t)t + 1] = make_float4(A00, A01, A11, 0.f);
	if (cull) cull[t] = cull_coef(fminf(sc[0], sc[1]), fmaxf(sc[0], sc[1]));
}

__device__ __forceinline__ void pack2d_one(const float *__restrict__ pos, const float *__restrict__ scal, const float *__restrict__ rot,
					   const float *__restrict__ vals, int t, int i, float4 *__restrict__ packed, float *__restrict__ cull)
{
	const double th = (double)rot[i];
	const float sc[2] = {scal[2 * (size_t)i], scal[2 * (size_t)i + 1]};
	pack2d_core(pos + 2 * (size_t)i, sc, exp2s(sc[0]), exp2s(sc[1]), (float)cos(th), (float)sin(th), vals + 2 * (size_t)i, t, packed, cull);
}


constexpr int SH_THREADS = 1024;
constexpr int SH_MAX_N = 16384;
constexpr int SH_MAX_CELLS = 26000;	// 2 x (cells + 1) x 4 B of shared memory

// One CTA of SH_THREADS threads; sh_mem: 2 x (ncell + 1) uint32 of shared memory, warp_sums: 32 uint32 of shared memory.
template <int D, bool GAUSS, bool RANK>
__device__ __forceinline__ void small_hash_body(const float *__restrict__ pts, int n, const Grid &g, int ncell, int32_t *__restrict__ cell_start,
						int32_t *__restrict__ ids_out, uint32_t *__restrict__ keys_tmp, uint32_t *__restrict__ ids_tmp,
						const float *__restrict__ scal, const float *__restrict__ rot, const float *__restrict__ vals,
						float4 *__restrict__ packed, float *__restrict__ cull, uint32_t *sh_mem, uint32_t *warp_sums)
{
	uint32_t *start = sh_mem;		// [ncell + 1] histogram, then exclusive prefix
	uint32_t *fill = sh_mem + ncell + 1;	// [ncell + 1] slot counters
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	for (int c = tid; c <= ncell; c += SH_THREADS) { start[c] = 0; fill[c] = 0; }
	__syncthreads();
	for (int i = tid; i < n; i += SH_THREADS) {
		const uint32_t key = GAUSS ? gauss_key<D>(pts, i, g) : sample_key<D, false>(pts, i, g);
		keys_tmp[i] = key;
		atomicAdd(&start[key], 1u);
	}
	__syncthreads();
	{	// exclusive scan of start[0..ncell]
		const int m = ncell + 1, chunk = (m + SH_THREADS - 1) / SH_THREADS;
		const int b = tid * chunk, e = min(b + chunk, m);
		uint32_t s = 0;
		for (int c = b; c < e; c++) s += start[c];
		uint32_t v = s;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
			if (lane >= o) v += t;
		}
		if (lane == 31) warp_sums[w] = v;
		__syncthreads();
		if (w == 0) {
			const uint32_t ws = warp_sums[lane];
			uint32_t t2 = ws;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t t = __shfl_up_sync(0xffffffffu, t2, o);
				if (lane >= o) t2 += t;
			}
			warp_sums[lane] = t2 - ws;
		}
		__syncthreads();
		uint32_t run = warp_sums[w] + (v - s);
		for (int c = b; c < e; c++) {
			const uint32_t t = start[c];
			start[c] = run;
			cell_start[c] = (int32_t)run;
			run += t;
		}
	}
	__syncthreads();
	for (int i = tid; i < n; i += SH_THREADS) {
		const uint32_t key = keys_tmp[i];
		ids_tmp[start[key] + atomicAdd(&fill[key], 1u)] = (uint32_t)i;
	}
	if (!RANK) return;	// crowded cells: small_rank_kernel finishes on the whole machine
	__syncthreads();
	for (int t = tid; t < n; t += SH_THREADS) {
		const uint32_t id = ids_tmp[t];
		const uint32_t key = keys_tmp[id];
		uint32_t pos = (uint32_t)t;
		if (key != (uint32_t)ncell) {
			const uint32_t s = start[key], cnt = fill[key];
			uint32_t rank = 0;
			for (uint32_t k = 0; k < cnt; k++) rank += ids_tmp[s + k] < id;
			pos = s + rank;
		}
		ids_out[pos] = (int32_t)id;
		if (GAUSS && packed) {
			if (D == 3) pack3d_one(pts, scal, rot, vals, (int)pos, (int)id, packed, cull);
			else pack2d_one(pts, scal, rot, vals, (int)pos, (int)id, packed, cull);
		}
	}
}


}  // namespace gsr
