// eval_tiled.cu — the throughput shape of the forward / RK4 kernels for large Q.
//
// Measured on B200 (profiles/): with one thread per point and one 48-byte record load per (point, candidate) pair
// the kernels of eval.cu are bound by the L1/shared-memory return path, not by FP32 issue — a warp-wide 128-bit load
// occupies the SM's load pipe for 4 cycles even when all lanes read the same address, i.e. 12 cycles per candidate per
// warp against ~5 cycles of arithmetic.  This file removes that bound and the divergence on the truncation test:
//
//   * REGISTER TILING: every thread owns P points (P = 4 forward, 2 in the RK4 pull-back), so one candidate record,
//     loaded once per warp, is tested against 32 P points;
//   * WARP-UNIFORM CANDIDATE STREAM: a warp walks the hull of its points' (clamped) 27-cell stencils cell by cell; a
//     point that does not have the current cell in its own stencil gets the threshold -1 for that cell (q >= 0 is never
//     accepted), so each point still sees exactly the reference's candidate list, in the reference's order;
//   * COHERENT WARPS: gsr_bin_samples sorts the samples by (cell, 4x4x4 sub-cell), so the 32 P points of a warp occupy
//     a small part of one cell and most candidates are rejected by the whole warp — the accepted branch
//     (ex2 + 28 flop) is then skipped by a uniform branch instead of running under a mostly idle warp;
//   * TMA STAGING: a CTA owns a TILE of up to 512 sorted samples of one (x, y) row of cells.  The Gaussians of the
//     3 x 3 neighbouring rows over the tile's z range (+-1 cell) are 9 CONTIGUOUS runs of packed records (the hash is
//     cell-sorted, z fastest): one lane per run issues a TMA bulk copy (cp.async.bulk global -> shared, completion on
//     an mbarrier).  The staged box serves every point of the tile and, in the RK4 kernels, all 4-5 evaluations.
//     Cells outside the staged box (an RK4 stage point that drifted into another row, a box above the staging
//     capacity) are read from global memory through the same loop.
#include "eval_tiled.cuh"

namespace gsr {

extern __shared__ __align__(16) unsigned char tl_smem[];

// sorted position of slot p of this thread: a warp's slot p is a run of 32 consecutive sorted samples
template <int P>
__device__ __forceinline__ int slot_pos(int t0, int p)
{
	return t0 + (int)(threadIdx.x >> 5) * (32 * P) + p * 32 + (int)(threadIdx.x & 31);
}

// CTAs per SM (P = 4): measured on the 128^3 lattice at S1 — with the Jacobian 4: 0.235 ms, 5: 0.224, 6: 0.252 (spills); value only 5: 0.164, 8: 0.154
template <int P, bool NEED_VAL, bool NEED_GRAD, bool ACCUM>
__global__ void __launch_bounds__(TL_TILE / P, P == 4 ? (NEED_GRAD ? 5 : 8) : 2) forward_tiled3_kernel(TiledArgs a, float *__restrict__ val, float *__restrict__ grad)
{
	__shared__ TileSh sh;
	__shared__ __align__(8) uint64_t mbar;
	float4 *srec = reinterpret_cast<float4 *>(tl_smem);
	int r, t0, t1;
	if (!tile_locate3(a, r, t0, t1)) return;
	float x[P], y[P], z[P], u[P][3], G[P][9];
	bool ok[P];
	size_t j[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int t = slot_pos<P>(t0, p);
		ok[p] = t < t1;
		j[p] = ok[p] ? (size_t)a.perm[t] : 0;
		x[p] = ok[p] ? a.x[3 * j[p]] : 0.f;
		y[p] = ok[p] ? a.x[3 * j[p] + 1] : 0.f;
		z[p] = ok[p] ? a.x[3 * j[p] + 2] : 0.f;
	}
	tile_stage3(a, r, t0, t1, sh, srec, &mbar);
	warp_eval3<P, NEED_GRAD, NEED_GRAD>(a, sh, srec, x, y, z, ok, u, G);
#pragma unroll
	for (int p = 0; p < P; p++) {
		if (!ok[p]) continue;
		if (NEED_VAL) {
#pragma unroll
			for (int k = 0; k < 3; k++) {
				float *o = val + 3 * j[p] + k;
				*o = ACCUM ? *o + u[p][k] : u[p][k];
			}
		}
		if (NEED_GRAD) {
#pragma unroll
			for (int k = 0; k < 9; k++) {
				float *o = grad + 9 * j[p] + k;
				*o = ACCUM ? *o + G[p][k] : G[p][k];
			}
		}
	}
}

// RK4 (modes as rk4_3d_kernel in eval.cu): one staged box serves all 4-5 evaluations of the tile's points.
template <int MODE, int P>
__global__ void __launch_bounds__(TL_TILE / P) rk4_tiled3_kernel(TiledArgs a, float dt, float *__restrict__ goal_pos, float *__restrict__ deformation,
								 float *__restrict__ goal_val, float *__restrict__ goal_grad, float *__restrict__ ref_vor,
								 float *__restrict__ ref_hel)
{
	constexpr bool FULLM = MODE != 0;
	__shared__ TileSh sh;
	__shared__ __align__(8) uint64_t mbar;
	float4 *srec = reinterpret_cast<float4 *>(tl_smem);
	int r, t0, t1;
	if (!tile_locate3(a, r, t0, t1)) return;
	float x0[P], x1[P], x2[P], px[P], py[P], pz[P], v[P][3], dv[P][9], vs[P][3], A[P][9], S[P][9];
	bool ok[P];
	size_t j[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int t = slot_pos<P>(t0, p);
		ok[p] = t < t1;
		j[p] = ok[p] ? (size_t)a.perm[t] : 0;
		px[p] = x0[p] = ok[p] ? a.x[3 * j[p]] : 0.f;
		py[p] = x1[p] = ok[p] ? a.x[3 * j[p] + 1] : 0.f;
		pz[p] = x2[p] = ok[p] ? a.x[3 * j[p] + 2] : 0.f;
		vs[p][0] = vs[p][1] = vs[p][2] = 0.f;
	}
	tile_stage3(a, r, t0, t1, sh, srec, &mbar);
	const float hdt = dt * .5f, dt6 = dt / 6.f;
#pragma unroll 1
	for (int st = 0; st < 4; st++) {
		warp_eval3<P, FULLM, false>(a, sh, srec, px, py, pz, ok, v, dv);
		const float wgt = (st == 0 || st == 3) ? 1.f : 2.f;	// RK4 weights 1 2 2 1
		const float step = (st < 2) ? hdt : dt;			// the NEXT stage point: x + step * v
#pragma unroll
		for (int p = 0; p < P; p++) {
			vs[p][0] += wgt * v[p][0]; vs[p][1] += wgt * v[p][1]; vs[p][2] += wgt * v[p][2];
			if (FULLM) {
				float B[9];
				if (st == 0) {
#pragma unroll
					for (int k = 0; k < 9; k++) B[k] = dv[p][k];
				} else {
					mm3(dv[p], A[p], B);	// dv_st @ dphi_st
				}
#pragma unroll
				for (int k = 0; k < 9; k++) {
					S[p][k] = (st == 0) ? B[k] : S[p][k] + wgt * B[k];
					A[p][k] = ((k % 4 == 0) ? 1.f : 0.f) + step * B[k];	// dphi_{st+1}
				}
			}
			px[p] = x0[p] + step * v[p][0]; py[p] = x1[p] + step * v[p][1]; pz[p] = x2[p] + step * v[p][2];
		}
	}
#pragma unroll
	for (int p = 0; p < P; p++) {
		px[p] = x0[p] + dt6 * vs[p][0]; py[p] = x1[p] + dt6 * vs[p][1]; pz[p] = x2[p] + dt6 * vs[p][2];
		if (MODE != 2 && ok[p]) {
			goal_pos[3 * j[p]] = px[p]; goal_pos[3 * j[p] + 1] = py[p]; goal_pos[3 * j[p] + 2] = pz[p];
		}
	}
	if (FULLM) {
		warp_eval3<P, true, false>(a, sh, srec, px, py, pz, ok, v, dv);
#pragma unroll
		for (int p = 0; p < P; p++) {
			if (!ok[p]) continue;
			float D[9];
#pragma unroll
			for (int k = 0; k < 9; k++) D[k] = ((k % 4 == 0) ? 1.f : 0.f) + dt6 * S[p][k];	// dphi
			const size_t jj = j[p];
			if (MODE == 1) {
#pragma unroll
				for (int k = 0; k < 9; k++) { deformation[9 * jj + k] = D[k]; goal_grad[9 * jj + k] = dv[p][k]; }
				goal_val[3 * jj] = v[p][0]; goal_val[3 * jj + 1] = v[p][1]; goal_val[3 * jj + 2] = v[p][2];
			} else {
				const float *d = dv[p];
				const float w0 = d[7] - d[5], w1 = d[2] - d[6], w2 = d[3] - d[1];
				if (ref_hel) ref_hel[jj] = v[p][0] * w0 + v[p][1] * w1 + v[p][2] * w2;
				const float c00 = D[4] * D[8] - D[5] * D[7], c01 = D[2] * D[7] - D[1] * D[8], c02 = D[1] * D[5] - D[2] * D[4];
				const float c10 = D[5] * D[6] - D[3] * D[8], c11 = D[0] * D[8] - D[2] * D[6], c12 = D[2] * D[3] - D[0] * D[5];
				const float c20 = D[3] * D[7] - D[4] * D[6], c21 = D[1] * D[6] - D[0] * D[7], c22 = D[0] * D[4] - D[1] * D[3];
				const float det = D[0] * c00 + D[1] * c10 + D[2] * c20;
				const float inv = 1.f / det;
				ref_vor[3 * jj] = (c00 * w0 + c01 * w1 + c02 * w2) * inv;
				ref_vor[3 * jj + 1] = (c10 * w0 + c11 * w1 + c12 * w2) * inv;
				ref_vor[3 * jj + 2] = (c20 * w0 + c21 * w1 + c22 * w2) * inv;
			}
		}
	}
}

// RK4 with the deformation chain, 4 points per thread: the per-point integrator state (start point, velocity sum, S and
// dphi: 24 floats) lives in SHARED memory between the evaluations, [slot][thread] so that lanes touch consecutive banks;
// only the evaluation itself (the forward kernel's register footprint) is in registers.
constexpr int RKS_P = 4;
constexpr int RKS_STATE = 24;	// x0 3 | vs 3 | S 9 | A 9

template <int MODE>
__global__ void __launch_bounds__(TL_TILE / RKS_P, 4) rk4_tiled3s_kernel(TiledArgs a, float dt, float *__restrict__ goal_pos, float *__restrict__ deformation,
									  float *__restrict__ goal_val, float *__restrict__ goal_grad, float *__restrict__ ref_vor,
									  float *__restrict__ ref_hel)
{
	constexpr int P = RKS_P, T = TL_TILE / RKS_P;
	__shared__ TileSh sh;
	__shared__ __align__(8) uint64_t mbar;
	float4 *srec = reinterpret_cast<float4 *>(tl_smem);
	float *state = reinterpret_cast<float *>(tl_smem + (size_t)a.cap * 48) + threadIdx.x;	// + (p * RKS_STATE + k) * T
	int r, t0, t1;
	if (!tile_locate3(a, r, t0, t1)) return;
	float px[P], py[P], pz[P], v[P][3], dv[P][9];
	bool ok[P];
#pragma unroll
	for (int p = 0; p < P; p++) {
		const int t = slot_pos<P>(t0, p);
		ok[p] = t < t1;
		const size_t j = ok[p] ? (size_t)a.perm[t] : 0;
		px[p] = ok[p] ? a.x[3 * j] : 0.f;
		py[p] = ok[p] ? a.x[3 * j + 1] : 0.f;
		pz[p] = ok[p] ? a.x[3 * j + 2] : 0.f;
		float *sp = state + (p * RKS_STATE) * T;
		sp[0] = px[p]; sp[T] = py[p]; sp[2 * T] = pz[p];
		sp[3 * T] = 0.f; sp[4 * T] = 0.f; sp[5 * T] = 0.f;
	}
	tile_stage3(a, r, t0, t1, sh, srec, &mbar);
	const float hdt = dt * .5f, dt6 = dt / 6.f;
#pragma unroll 1
	for (int st = 0; st < 4; st++) {
		warp_eval3<P, true, true>(a, sh, srec, px, py, pz, ok, v, dv);
		const float wgt = (st == 0 || st == 3) ? 1.f : 2.f;
		const float step = (st < 2) ? hdt : dt;
#pragma unroll
		for (int p = 0; p < P; p++) {
			float *sp = state + (p * RKS_STATE) * T;
#pragma unroll
			for (int k = 0; k < 3; k++) sp[(3 + k) * T] += wgt * v[p][k];
			float B[9];
			if (st == 0) {
#pragma unroll
				for (int k = 0; k < 9; k++) B[k] = dv[p][k];
			} else {
				float A[9];
#pragma unroll
				for (int k = 0; k < 9; k++) A[k] = sp[(15 + k) * T];
				mm3(dv[p], A, B);	// dv_st @ dphi_st
			}
#pragma unroll
			for (int k = 0; k < 9; k++) {
				sp[(6 + k) * T] = (st == 0) ? B[k] : sp[(6 + k) * T] + wgt * B[k];
				sp[(15 + k) * T] = ((k % 4 == 0) ? 1.f : 0.f) + step * B[k];	// dphi_{st+1}
			}
			px[p] = sp[0] + step * v[p][0]; py[p] = sp[T] + step * v[p][1]; pz[p] = sp[2 * T] + step * v[p][2];
		}
	}
#pragma unroll
	for (int p = 0; p < P; p++) {
		const float *sp = state + (p * RKS_STATE) * T;
		px[p] = sp[0] + dt6 * sp[3 * T]; py[p] = sp[T] + dt6 * sp[4 * T]; pz[p] = sp[2 * T] + dt6 * sp[5 * T];
	}
	warp_eval3<P, true, true>(a, sh, srec, px, py, pz, ok, v, dv);
#pragma unroll
	for (int p = 0; p < P; p++) {
		if (!ok[p]) continue;
		const float *sp = state + (p * RKS_STATE) * T;
		const size_t jj = (size_t)a.perm[slot_pos<P>(t0, p)];
		float D[9];
#pragma unroll
		for (int k = 0; k < 9; k++) D[k] = ((k % 4 == 0) ? 1.f : 0.f) + dt6 * sp[(6 + k) * T];	// dphi
		if (MODE == 1) {
			goal_pos[3 * jj] = px[p]; goal_pos[3 * jj + 1] = py[p]; goal_pos[3 * jj + 2] = pz[p];
#pragma unroll
			for (int k = 0; k < 9; k++) { deformation[9 * jj + k] = D[k]; goal_grad[9 * jj + k] = dv[p][k]; }
			goal_val[3 * jj] = v[p][0]; goal_val[3 * jj + 1] = v[p][1]; goal_val[3 * jj + 2] = v[p][2];
		} else {
			const float *d = dv[p];
			const float w0 = d[7] - d[5], w1 = d[2] - d[6], w2 = d[3] - d[1];
			if (ref_hel) ref_hel[jj] = v[p][0] * w0 + v[p][1] * w1 + v[p][2] * w2;
			const float c00 = D[4] * D[8] - D[5] * D[7], c01 = D[2] * D[7] - D[1] * D[8], c02 = D[1] * D[5] - D[2] * D[4];
			const float c10 = D[5] * D[6] - D[3] * D[8], c11 = D[0] * D[8] - D[2] * D[6], c12 = D[2] * D[3] - D[0] * D[5];
			const float c20 = D[3] * D[7] - D[4] * D[6], c21 = D[1] * D[6] - D[0] * D[7], c22 = D[0] * D[4] - D[1] * D[3];
			const float det = D[0] * c00 + D[1] * c10 + D[2] * c20;
			const float inv = 1.f / det;
			ref_vor[3 * jj] = (c00 * w0 + c01 * w1 + c02 * w2) * inv;
			ref_vor[3 * jj + 1] = (c10 * w0 + c11 * w1 + c12 * w2) * inv;
			ref_vor[3 * jj + 2] = (c20 * w0 + c21 * w1 + c22 * w2) * inv;
		}
	}
}

// ---- tile -> row table ----------------------------------------------------------------------------
// Row r (a fixed (x, y) pair of the padded sample grid; r == nrows: the samples outside it) owns the tile ids
// [s_r / T + r, s_r / T + r + ceil(n_r / T)), s_r = first sorted sample of the row.  The ranges of different rows
// never overlap, so no prefix sum over rows is needed; unused ids stay -1.
__global__ void tile_rows_kernel(const int32_t *__restrict__ scs, int Q, int rowlen, int nrows, int pcell, int32_t *__restrict__ tile_row)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r > nrows) return;
	const int s = scs[r < nrows ? r * rowlen : pcell];
	const int e = (r < nrows) ? scs[(r + 1) * rowlen] : Q;
	const int base = s / TL_TILE + r;
	const int n = (e - s + TL_TILE - 1) / TL_TILE;
	for (int c = 0; c < n; c++) tile_row[base + c] = r;
}

static int64_t tile_slots(const Grid &g, int64_t Q)
{
	const int rowlen = g.pdims[g.D - 1];
	return Q / TL_TILE + g.pcell / rowlen + 2;
}

extern int g_force_radix;	// grid.cu
extern int g_step_small_n;	// step.cu
extern int g_step_lanes4;	// step.cu
extern int g_step_fused_hash;	// step.cu
extern int g_gather_cta_max_n;	// backward.cu

// tunables (gsr_set_tuning)
int g_tiled_min_q = 1 << 17;
int g_tiled_cap = 512;
int g_fw_p4_min_spc = 0;	// samples per cell above which the forward kernel takes 4 points per thread (else 2)
int g_rk4s_min_spc = 1024;	// samples per hash cell from which the shared-memory-state RK4 kernel is used (GSR_TUNE_RK4S_MIN_SPC)
int g_rk4s_cap = 128;	// staging capacity of rk4_tiled3s_kernel (GSR_TUNE_RK4S_CAP)
int g_rk4_smem_state = 1;	// 1: RK4 with the deformation chain keeps its state in shared memory (4 points / thread)

static size_t tiled_smem(int cap) { return (size_t)cap * 48; }

template <typename K>
static int prep(K kernel, int cap, bool max_shared)
{
	cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem(cap));
	// forward: 4-5 CTAs per SM are register-feasible, so ask for the large shared-memory carve-out; the RK4 kernels are
	// register-limited to 2 CTAs and keep the default split (their few spilled values live in L1)
	if (e == cudaSuccess && max_shared) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace gsr

using namespace gsr;

extern "C" int64_t gsr_tile_slots(const gsr_grid_desc *d, int64_t Q)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0) return GSR_EINVAL;
	return tile_slots(g, Q);
}

extern "C" int gsr_build_tiles(const gsr_grid_desc *d, const int32_t *sample_cell_start, int64_t Q, int32_t *tile_row, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || Q >= ((int64_t)1 << 30) || !sample_cell_start || !tile_row) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	const int rowlen = g.pdims[g.D - 1], nrows = g.pcell / rowlen;
	cudaError_t e = cudaMemsetAsync(tile_row, 0xff, sizeof(int32_t) * (size_t)tile_slots(g, Q), st);
	if (e != cudaSuccess) return (int)e;
	g_launches += 1;
	tile_rows_kernel<<<(nrows + 1 + 127) / 128, 128, 0, st>>>(sample_cell_start, (int)Q, rowlen, nrows, g.pcell, tile_row);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_set_tuning(int key, int value)
{
	switch (key) {
	case GSR_TUNE_TILED_MIN_Q: g_tiled_min_q = value; return GSR_OK;
	case GSR_TUNE_RK4_SMEM_STATE: g_rk4_smem_state = value; return GSR_OK;
	case GSR_TUNE_FORCE_RADIX: g_force_radix = value; return GSR_OK;
	case GSR_TUNE_STEP_SMALL_N: g_step_small_n = value; return GSR_OK;
	case GSR_TUNE_RK4S_CAP: g_rk4s_cap = value; return GSR_OK;
	case GSR_TUNE_RK4S_MIN_SPC: g_rk4s_min_spc = value; return GSR_OK;
	case GSR_TUNE_STEP_LANES4: g_step_lanes4 = value; return GSR_OK;
	case GSR_TUNE_STEP_FUSED_HASH: g_step_fused_hash = value; return GSR_OK;
	case GSR_TUNE_GATHER_CTA_MAX_N: g_gather_cta_max_n = value; return GSR_OK;
	case GSR_TUNE_LANES8_MIN_N: g_lanes8_min_n = value; return GSR_OK;
	case GSR_TUNE_FW_P4_MIN_SPC: g_fw_p4_min_spc = value; return GSR_OK;
	case GSR_TUNE_TILED_CAP:
		if (value < 0 || tiled_smem(value) > 200 * 1024) return GSR_EINVAL;
		g_tiled_cap = value;
		return GSR_OK;
	default: return GSR_EINVAL;
	}
}

namespace gsr {

bool use_tiled(const Grid &g, int64_t Q, const int32_t *perm, const int32_t *scs, const int32_t *tile_row)
{
	return g.D == 3 && perm && scs && tile_row && Q >= g_tiled_min_q;
}

static TiledArgs make_targs(const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
			    const int32_t *scs, const int32_t *tile_row)
{
	TiledArgs a;
	a.exec_count = nullptr;
	a.P = P; a.cell_start = cell_start; a.packed = (const float4 *)packed; a.cull = cull; a.x = x; a.Q = (int)Q; a.perm = perm; a.scs = scs; a.tile_row = tile_row;
	a.cap = g_tiled_cap;
	return a;
}

#define TL_LAUNCH(PP, KERNEL, ...)                                                              \
	do {                                                                                    \
		int rc__ = prep(KERNEL, a.cap, (PP) == FW_P);                                   \
		if (rc__) return rc__;                                                          \
		KERNEL<<<blocks, TL_TILE / (PP), tiled_smem(a.cap), st>>>(__VA_ARGS__);         \
	} while (0)

constexpr int FW_P = 4;	// points per thread, forward
constexpr int RK_P = 2;	// points per thread, RK4 with the deformation chain (register budget)

int launch_forward_tiled3(const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
			  const int32_t *scs, const int32_t *tile_row, float *val, float *grad, bool accumulate, cudaStream_t st)
{
	TiledArgs a = make_targs(P, cell_start, packed, cull, x, Q, perm, scs, tile_row);
	const int blocks = (int)tile_slots(P.g, Q);
	const bool dense = (double)Q >= (double)g_fw_p4_min_spc * (double)P.g.ncell;	// samples per cell: see launch_rk4_tiled3
#define FW_DISPATCH(PP)                                                                                                 \
	do {                                                                                                            \
		if (accumulate) {                                                                                       \
			if (val && grad) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, true, true>), a, val, grad);    \
			else if (val) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, false, true>), a, val, grad);      \
			else TL_LAUNCH(PP, (forward_tiled3_kernel<PP, false, true, true>), a, val, grad);               \
		} else {                                                                                                \
			if (val && grad) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, true, false>), a, val, grad);   \
			else if (val) TL_LAUNCH(PP, (forward_tiled3_kernel<PP, true, false, false>), a, val, grad);     \
			else TL_LAUNCH(PP, (forward_tiled3_kernel<PP, false, true, false>), a, val, grad);              \
		}                                                                                                       \
	} while (0)
	if (dense) FW_DISPATCH(4);
	else FW_DISPATCH(2);
#undef FW_DISPATCH
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

int launch_rk4_tiled3(int mode, const EvalParams &P, const int32_t *cell_start, const float *packed, const float *cull, const float *x, int64_t Q, const int32_t *perm,
		      const int32_t *scs, const int32_t *tile_row, float dt, float *goal_pos, float *deformation, float *goal_val, float *goal_grad,
		      float *ref_vor, float *ref_hel, cudaStream_t st)
{
	TiledArgs a = make_targs(P, cell_start, packed, cull, x, Q, perm, scs, tile_row);
	const int blocks = (int)tile_slots(P.g, Q);
	// many samples per cell: warps of 128 points are still compact, and 4 points per thread halve the per-candidate overhead;
	// fewer: 64-point warps cull better
	if (mode != 0 && g_rk4_smem_state && (double)Q >= (double)g_rk4s_min_spc * (double)P.g.ncell) {
		// four CTAs per SM (launch bounds: 128 registers) instead of three: the integrator state takes 48 KB of shared memory per CTA,
		// so the staging area shrinks to 128 records (what does not fit is read through L1) — measured on the 128^3 lattice at S1:
		// 1.216 -> 1.124 ms; with 128 registers and the full staging area (three CTAs) 1.280 ms
		a.cap = min(a.cap, g_rk4s_cap);
		const size_t sm = tiled_smem(a.cap) + sizeof(float) * RKS_STATE * RKS_P * (TL_TILE / RKS_P);
		cudaError_t e = (mode == 1) ? cudaFuncSetAttribute(rk4_tiled3s_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
					    : cudaFuncSetAttribute(rk4_tiled3s_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e == cudaSuccess)
			e = (mode == 1) ? cudaFuncSetAttribute(rk4_tiled3s_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
					: cudaFuncSetAttribute(rk4_tiled3s_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		if (e != cudaSuccess) return (int)e;
		if (mode == 1) rk4_tiled3s_kernel<1><<<blocks, TL_TILE / RKS_P, sm, st>>>(a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
		else rk4_tiled3s_kernel<2><<<blocks, TL_TILE / RKS_P, sm, st>>>(a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
		GSR_CHECK_LAUNCH();
		return GSR_OK;
	}
	if (mode == 0) TL_LAUNCH(FW_P, (rk4_tiled3_kernel<0, FW_P>), a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
	else if (mode == 1) TL_LAUNCH(RK_P, (rk4_tiled3_kernel<1, RK_P>), a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
	else TL_LAUNCH(RK_P, (rk4_tiled3_kernel<2, RK_P>), a, dt, goal_pos, deformation, goal_val, goal_grad, ref_vor, ref_hel);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

}  // namespace gsr
