// grid.cu — the cell-binned spatial hash.
//
// Replaces reinitialize_grid_ti (reference 3D/GSR.py:205-245, 2D/GSR.py:194-222): instead of a
// histogram + serial prefix + atomic-slot scatter, Gaussian cell keys are radix-sorted (stable, LSD,
// 8-bit digits, hand-written) which yields the reference's grid_cnt / grid_offset bit-exactly and
// sorted_id in canonical (ascending id inside a cell) order, deterministically.
//
// HBM-bound integer work: 1 key pass + P radix passes (P = ceil(log2(ncell+1)/8)) over 8 B/item,
// then a gather that writes the 48 B (3D) / 32 B (2D) packed Gaussian record in cell order.
#include "common.cuh"
#include "hash_small.cuh"
#include "sampling.cuh"
#include <math.h>
#include <cooperative_groups.h>

namespace gsr {

// ------------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------------

template <int D>
__global__ void gauss_keys_kernel(const float *__restrict__ pos, int n, Grid g, uint32_t *__restrict__ keys)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	keys[i] = gauss_key<D>(pos, i, g);
}

template <int D, bool FINE>
__global__ void sample_keys_kernel(const float *__restrict__ x, int n, Grid g, uint32_t *__restrict__ keys)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	keys[i] = sample_key<D, FINE>(x, i, g);
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, index) pairs, 8 bits per pass
// ------------------------------------------------------------------------------------------------

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;	// 2048 keys per block

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t *__restrict__ keys, int n, int shift, uint32_t *__restrict__ hist, int nblocks)
{
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	int base = blockIdx.x * RS_TILE;
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		int idx = base + i * RS_THREADS + threadIdx.x;
		if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
	}
	__syncthreads();
	hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];	// digit-major
}

// Exclusive scan of the digit-major histogram hist[256][nblocks] (row-major order = the order of a stable counting sort):
// CTA d scans row d in place and publishes the row total; the last CTA to finish turns the 256 totals into row bases.
// The scatter kernel adds base[d] to its row-local offsets.  (A one-CTA scan of all 256 * nblocks entries took 0.44 ms at
// 2 M keys — more than everything else in the pass.)
__global__ void __launch_bounds__(256) rs_scan_kernel(uint32_t *__restrict__ hist, int nblocks, uint32_t *__restrict__ totals, uint32_t *__restrict__ base,
						       unsigned int *__restrict__ done)
{
	__shared__ uint32_t warp_sums[8];
	__shared__ bool last;
	const int d = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t *row = hist + (size_t)d * nblocks;
	const int chunk = (nblocks + 255) / 256, b = threadIdx.x * chunk, e = min(b + chunk, nblocks);
	uint32_t s = 0;
	for (int i = b; i < e; i++) s += row[i];
	uint32_t v = s;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
		if (lane >= o) v += t;
	}
	if (lane == 31) warp_sums[w] = v;
	__syncthreads();
	uint32_t wbase = 0, total = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) {
		if (k < w) wbase += warp_sums[k];
		total += warp_sums[k];
	}
	uint32_t run = wbase + (v - s);
	for (int i = b; i < e; i++) {
		const uint32_t t = row[i];
		row[i] = run;
		run += t;
	}
	if (threadIdx.x == 0) {
		totals[d] = total;
		__threadfence();
		last = atomicAdd(done, 1u) == 255u;
	}
	__syncthreads();
	if (!last) return;
	__threadfence();
	// 256 totals -> exclusive bases, by this one CTA (thread t owns digit t)
	const uint32_t mine = ((volatile uint32_t *)totals)[threadIdx.x];
	uint32_t v2 = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, v2, o);
		if (lane >= o) v2 += t;
	}
	__syncthreads();
	if (lane == 31) warp_sums[w] = v2;
	__syncthreads();
	uint32_t wb = 0;
#pragma unroll
	for (int k = 0; k < 8; k++)
		if (k < w) wb += warp_sums[k];
	base[threadIdx.x] = wb + v2 - mine;
	if (threadIdx.x == 0) *done = 0;	// re-arm for the next pass
}

// Stable in-warp ranking of one 32-key slice: lanes with equal digits get consecutive offsets in lane
// order; `cnt` is this warp's private digit counter row (shared memory).
__device__ __forceinline__ uint32_t warp_rank(uint32_t digit, bool valid, uint32_t *cnt, int lane)
{
	uint32_t d = valid ? digit : 256u;
	uint32_t mask = __match_any_sync(0xffffffffu, d);
	int leader = __ffs(mask) - 1;
	uint32_t old = 0;
	if (valid && lane == leader) {
		old = cnt[d];
		cnt[d] = old + __popc(mask);
	}
	old = __shfl_sync(0xffffffffu, old, leader);
	__syncwarp();
	return old + __popc(mask & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
								uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
								int n, int shift, const uint32_t *__restrict__ hist, int nblocks, const uint32_t *__restrict__ digit_base)
{
	__shared__ uint32_t cnt[RS_WARPS][256];
	__shared__ uint32_t gbase[256];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
	gbase[threadIdx.x] = hist[(size_t)threadIdx.x * nblocks + blockIdx.x] + digit_base[threadIdx.x];
	__syncthreads();
	int base = blockIdx.x * RS_TILE + w * (32 * RS_ITEMS);
	uint32_t key[RS_ITEMS], off[RS_ITEMS];
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		int idx = base + i * 32 + lane;
		bool valid = idx < n;
		key[i] = valid ? keys_in[idx] : 0u;
		off[i] = warp_rank((key[i] >> shift) & 255u, valid, cnt[w], lane);
	}
	__syncthreads();
	{	// exclusive prefix over warps for digit = threadIdx.x
		uint32_t run = 0;
#pragma unroll
		for (int ww = 0; ww < RS_WARPS; ww++) {
			uint32_t t = cnt[ww][threadIdx.x];
			cnt[ww][threadIdx.x] = run;
			run += t;
		}
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		int idx = base + i * 32 + lane;
		if (idx < n) {
			uint32_t d = (key[i] >> shift) & 255u;
			uint32_t p = gbase[d] + cnt[w][d] + off[i];
			keys_out[p] = key[i];
			vals_out[p] = vals_in ? vals_in[idx] : (uint32_t)idx;
		}
	}
}

// Whole sort in ONE block for small inputs (n <= 16384): the latency-bound regime of the reference's own
// problem sizes (N = 1000 .. 64000, rebuilt every optimiser iteration).
constexpr int SB_THREADS = 1024;
constexpr int SB_WARPS = 32;
constexpr int SB_MAX_ITERS = 16;
constexpr int SB_MAX_N = SB_WARPS * 32 * SB_MAX_ITERS;

__global__ void __launch_bounds__(SB_THREADS) rs_single_block_kernel(const uint32_t *__restrict__ keys0, int n, int passes,
								     uint32_t *kA, uint32_t *vA, uint32_t *kB, uint32_t *vB)
{
	extern __shared__ uint32_t smem[];
	uint32_t(*cnt)[256] = reinterpret_cast<uint32_t(*)[256]>(smem);	// [SB_WARPS][256]
	uint32_t *dbase = smem + SB_WARPS * 256;				// [256]
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	int per_warp = ((n + SB_WARPS - 1) / SB_WARPS + 31) & ~31;
	int iters = per_warp / 32;
	const uint32_t *kin = keys0, *vin = nullptr;
	for (int p = 0; p < passes; p++) {
		uint32_t *kout = (p & 1) ? kB : kA, *vout = (p & 1) ? vB : vA;
		int shift = 8 * p;
		for (int i = threadIdx.x; i < SB_WARPS * 256; i += SB_THREADS) (&cnt[0][0])[i] = 0;
		__syncthreads();
		uint32_t key[SB_MAX_ITERS], off[SB_MAX_ITERS];
#pragma unroll
		for (int i = 0; i < SB_MAX_ITERS; i++) {
			if (i < iters) {
				int idx = w * per_warp + i * 32 + lane;
				bool valid = idx < n;
				key[i] = valid ? kin[idx] : 0u;
				off[i] = warp_rank((key[i] >> shift) & 255u, valid, cnt[w], lane);
			}
		}
		__syncthreads();
		if (threadIdx.x < 256) {
			uint32_t run = 0;
			for (int ww = 0; ww < SB_WARPS; ww++) {
				uint32_t t = cnt[ww][threadIdx.x];
				cnt[ww][threadIdx.x] = run;
				run += t;
			}
			dbase[threadIdx.x] = run;	// digit total
		}
		__syncthreads();
		if (w == 0) {	// exclusive scan of the 256 digit totals by one warp (8 per lane)
			uint32_t loc[8], s = 0;
#pragma unroll
			for (int k = 0; k < 8; k++) { loc[k] = dbase[lane * 8 + k]; s += loc[k]; }
			uint32_t v = s;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
				if (lane >= o) v += t;
			}
			uint32_t run = v - s;
#pragma unroll
			for (int k = 0; k < 8; k++) { dbase[lane * 8 + k] = run; run += loc[k]; }
		}
		__syncthreads();
#pragma unroll
		for (int i = 0; i < SB_MAX_ITERS; i++) {
			if (i < iters) {
				int idx = w * per_warp + i * 32 + lane;
				if (idx < n) {
					uint32_t d = (key[i] >> shift) & 255u;
					uint32_t q = dbase[d] + cnt[w][d] + off[i];
					kout[q] = key[i];
					vout[q] = vin ? vin[idx] : (uint32_t)idx;
				}
			}
		}
		__syncthreads();	// global writes of this block are visible to it after the barrier
		kin = kout;
		vin = vout;
	}
}

static inline int key_passes(uint32_t max_key)
{
	int bits = 1;
	while (bits < 32 && (max_key >> bits)) bits++;
	return (bits + 7) / 8;
}

struct SortWs {
	uint32_t *keys0, *kA, *kB, *vTmp, *hist, *cells;
	size_t cells_cap;
	int nblocks;
};

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static size_t sort_ws_bytes(int64_t n, int64_t cells = 0)
{
	int64_t nb = (n + RS_TILE - 1) / RS_TILE;
	if (nb < 1) nb = 1;
	// 4 arrays of n, the radix histograms, and one per-cell counter array (the counting-sort path)
	return 4 * align256(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1)) + align256(sizeof(uint32_t) * (256 * (size_t)nb + 1024)) + align256(sizeof(uint32_t) * (size_t)(cells + 2));
}

static bool carve_sort_ws(void *ws, size_t ws_bytes, int64_t n, SortWs &s)
{
	if (ws_bytes < sort_ws_bytes(n)) return false;
	char *p = (char *)ws;
	size_t a = align256(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
	s.keys0 = (uint32_t *)p; p += a;
	s.kA = (uint32_t *)p; p += a;
	s.kB = (uint32_t *)p; p += a;
	s.vTmp = (uint32_t *)p; p += a;
	s.hist = (uint32_t *)p;
	s.nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
	if (s.nblocks < 1) s.nblocks = 1;
	s.cells = (uint32_t *)(p + align256(sizeof(uint32_t) * (256 * (size_t)s.nblocks + 1024)));
	s.cells_cap = (ws_bytes - (size_t)((char *)s.cells - (char *)ws)) / sizeof(uint32_t);
	return true;
}

// Sort (s.keys0[i], i) by key; sorted indices land in out_vals, sorted keys in *keys_sorted.
static int radix_sort_index(SortWs &s, int n, uint32_t max_key, uint32_t *out_vals, const uint32_t **keys_sorted, cudaStream_t st)
{
	int passes = key_passes(max_key);
	// ping-pong so that the LAST pass writes its values into out_vals
	uint32_t *vA = (passes & 1) ? out_vals : s.vTmp;
	uint32_t *vB = (passes & 1) ? s.vTmp : out_vals;
	if (n <= SB_MAX_N) {
		size_t sm = sizeof(uint32_t) * (SB_WARPS * 256 + 256);
		rs_single_block_kernel<<<1, SB_THREADS, sm, st>>>(s.keys0, n, passes, s.kA, vA, s.kB, vB);
		g_launches += 1;
		GSR_CHECK_LAUNCH();
	} else {
		const uint32_t *kin = s.keys0, *vin = nullptr;
		for (int p = 0; p < passes; p++) {
			uint32_t *kout = (p & 1) ? s.kB : s.kA, *vout = (p & 1) ? vB : vA;
			uint32_t *totals = s.hist + (size_t)256 * s.nblocks, *base = totals + 256;
			unsigned int *done = base + 256;
			if (p == 0) {
				cudaError_t e = cudaMemsetAsync(done, 0, sizeof(unsigned int), st);
				if (e != cudaSuccess) return (int)e;
			}
			rs_hist_kernel<<<s.nblocks, RS_THREADS, 0, st>>>(kin, n, 8 * p, s.hist, s.nblocks);
			rs_scan_kernel<<<256, 256, 0, st>>>(s.hist, s.nblocks, totals, base, done);
			rs_scatter_kernel<<<s.nblocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, 8 * p, s.hist, s.nblocks, base);
			g_launches += 3;
			GSR_CHECK_LAUNCH();
			kin = kout;
			vin = vout;
		}
	}
	*keys_sorted = ((passes - 1) & 1) ? s.kB : s.kA;
	return 0;
}

// cell_start[c] = first sorted position whose key >= c, for c in [0, ncell].  Items whose key is ncell (not in
// the hash) stay at the tail of the sorted order, past cell_start[ncell].
__global__ void cell_start_kernel(const uint32_t *__restrict__ keys_sorted, int n, int ncell, int32_t *__restrict__ cell_start, int shift)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	int prev = (i == 0) ? -1 : min((int)(keys_sorted[i - 1] >> shift), ncell);
	int cur = (i == n) ? ncell : min((int)(keys_sorted[i] >> shift), ncell);
	for (int c = prev + 1; c <= cur; c++) cell_start[c] = i;
}

__global__ void ref_format_kernel(const int32_t *__restrict__ cell_start, int ncell, int32_t *__restrict__ cnt, int32_t *__restrict__ offset)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ncell) return;
	int s = cell_start[c];
	if (cnt) cnt[c] = cell_start[c + 1] - s;
	if (offset) offset[c] = s;
}

// ------------------------------------------------------------------------------------------------
// per-Gaussian precompute, gathered into cell order
// ------------------------------------------------------------------------------------------------

__global__ void pack3d_kernel(const float *__restrict__ pos, const float *__restrict__ scal, const float *__restrict__ rot, const float *__restrict__ vals,
			      int n, const int32_t *__restrict__ sorted_id, float4 *__restrict__ packed, float *__restrict__ cull)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	pack3d_one(pos, scal, rot, vals, t, sorted_id[t], packed, cull);
}

__global__ void pack2d_kernel(const float *__restrict__ pos, const float *__restrict__ scal, const float *__restrict__ rot, const float *__restrict__ vals,
			      int n, const int32_t *__restrict__ sorted_id, float4 *__restrict__ packed, float *__restrict__ cull)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	pack2d_one(pos, scal, rot, vals, t, sorted_id[t], packed, cull);
}

// ------------------------------------------------------------------------------------------------
// Whole hash in ONE single-CTA kernel for small inputs (the latency-bound regime of the reference's own sizes, where
// the hash of N = 1000..16384 Gaussians and of every sample batch is rebuilt in each optimiser iteration):
//   keys -> shared-memory histogram -> exclusive scan (= cell_start) -> slot scatter -> canonical intra-cell order
//   [-> pack {mu, Sigma^-1, v} in cell order].
// Intra-cell order: an item's final rank is the number of smaller ids in its cell's (unordered) segment, so the result
// is exactly the stable sort of the radix path, deterministically, without sorting.  The tail bucket (items outside the
// hash) keeps its arrival order: nothing reads it positionally.
// ------------------------------------------------------------------------------------------------
template <int D, bool GAUSS, bool RANK>
__global__ void __launch_bounds__(SH_THREADS) small_hash_kernel(const float *__restrict__ pts, int n, Grid g, int ncell, int32_t *__restrict__ cell_start,
								 int32_t *__restrict__ ids_out, uint32_t *__restrict__ keys_tmp, uint32_t *__restrict__ ids_tmp,
								 const float *__restrict__ scal, const float *__restrict__ rot, const float *__restrict__ vals,
								 float4 *__restrict__ packed, float *__restrict__ cull)
{
	extern __shared__ uint32_t sh_mem[];
	__shared__ uint32_t warp_sums[32];
	small_hash_body<D, GAUSS, RANK>(pts, n, g, ncell, cell_start, ids_out, keys_tmp, ids_tmp, scal, rot, vals, packed, cull, sh_mem, warp_sums);
}

// second half of the small hash when cells are crowded: one thread per item, whole machine
template <int D, bool GAUSS>
__global__ void __launch_bounds__(128) small_rank_kernel(const float *__restrict__ pts, int n, int ncell, const int32_t *__restrict__ cell_start,
							 int32_t *__restrict__ ids_out, const uint32_t *__restrict__ keys_tmp, const uint32_t *__restrict__ ids_tmp,
							 const float *__restrict__ scal, const float *__restrict__ rot, const float *__restrict__ vals,
							 float4 *__restrict__ packed, float *__restrict__ cull)
{
	const int t = blockIdx.x * 128 + threadIdx.x;
	if (t >= n) return;
	const uint32_t id = ids_tmp[t];
	const uint32_t key = keys_tmp[id];
	uint32_t pos = (uint32_t)t;
	if (key != (uint32_t)ncell) {
		const uint32_t s = (uint32_t)cell_start[key], cnt = (uint32_t)cell_start[key + 1] - s;
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

template <int D, bool GAUSS>
static int launch_small_hash(const float *pts, int n, const Grid &g, int ncell, int32_t *cell_start, int32_t *ids_out, uint32_t *keys_tmp, uint32_t *ids_tmp,
			     const float *scal, const float *rot, const float *vals, float4 *packed, float *cull, cudaStream_t st)
{
	const size_t sm = sizeof(uint32_t) * 2 * (size_t)(ncell + 1);
	const bool fused_rank = (int64_t)n <= 4 * (int64_t)ncell;	// sparse cells: rank inside the single CTA
	if (sm > 48 * 1024) {
		cudaError_t e = fused_rank ? cudaFuncSetAttribute(small_hash_kernel<D, GAUSS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)
					   : cudaFuncSetAttribute(small_hash_kernel<D, GAUSS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e != cudaSuccess) return (int)e;
	}
	if (fused_rank) {
		g_launches += 1;
		small_hash_kernel<D, GAUSS, true><<<1, SH_THREADS, sm, st>>>(pts, n, g, ncell, cell_start, ids_out, keys_tmp, ids_tmp, scal, rot, vals, packed, cull);
	} else {
		g_launches += 2;
		small_hash_kernel<D, GAUSS, false><<<1, SH_THREADS, sm, st>>>(pts, n, g, ncell, cell_start, ids_out, keys_tmp, ids_tmp, scal, rot, vals, packed, cull);
		small_rank_kernel<D, GAUSS><<<(n + 127) / 128, 128, 0, st>>>(pts, n, ncell, cell_start, ids_out, keys_tmp, ids_tmp, scal, rot, vals, packed, cull);
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

// ------------------------------------------------------------------------------------------------
// Counting sort by cell for mid-size inputs whose cells are not crowded (n up to millions, the per-iteration hash of
// N = 64 000 .. 10^6 Gaussians and of Q = N sample batches): histogram with global atomics -> one-CTA scan (= cell_start)
// -> slot scatter -> canonical intra-cell order by rank counting (small_rank_kernel).  4 launches + 2 memsets instead of
// the 3 launches per 8-bit digit of the radix path; the result is identical (stable sort by cell).
// ------------------------------------------------------------------------------------------------
template <int D, bool GAUSS>
__global__ void __launch_bounds__(256) ch_keys_hist_kernel(const float *__restrict__ pts, int n, Grid g, uint32_t *__restrict__ keys, uint32_t *__restrict__ hist)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const uint32_t key = GAUSS ? gauss_key<D>(pts, i, g) : sample_key<D, false>(pts, i, g);
	keys[i] = key;
	atomicAdd(hist + key, 1u);
}

// exclusive scan of hist[0..m) by one CTA -> out (int32); hist may alias out
__global__ void __launch_bounds__(1024) ch_scan_kernel(const uint32_t *hist, int m, int32_t *out)
{
	__shared__ uint32_t warp_sums[32];
	const int T = 1024, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int chunk = (m + T - 1) / T, b = threadIdx.x * chunk, e = min(b + chunk, m);
	uint32_t s = 0;
	for (int i = b; i < e; i++) s += hist[i];
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
	for (int i = b; i < e; i++) {
		const uint32_t t = hist[i];
		out[i] = (int32_t)run;
		run += t;
	}
}

__global__ void __launch_bounds__(256) ch_scatter_kernel(const uint32_t *__restrict__ keys, int n, const int32_t *__restrict__ cell_start, uint32_t *__restrict__ fill,
							 uint32_t *__restrict__ ids_tmp)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const uint32_t key = keys[i];
	ids_tmp[(uint32_t)cell_start[key] + atomicAdd(fill + key, 1u)] = (uint32_t)i;
}

int g_force_radix = 0;	// GSR_TUNE_FORCE_RADIX: tests compare the hash paths

static bool count_hash_ok(int64_t n, int64_t ncell, const SortWs &s)
{
	return !g_force_radix && n > SH_MAX_N && ncell * 48 >= n && ncell <= (1 << 22) && s.cells_cap >= (size_t)(ncell + 2);
}

template <int D, bool GAUSS>
static int launch_count_hash(const float *pts, int n, const Grid &g, int ncell, int32_t *cell_start, int32_t *ids_out, SortWs &s,
			     const float *scal, const float *rot, const float *vals, float4 *packed, float *cull, cudaStream_t st)
{
	// cell_start doubles as the histogram (ncell + 1 counters: the last one collects the items outside the hash)
	cudaError_t e = cudaMemsetAsync(cell_start, 0, sizeof(int32_t) * (size_t)(ncell + 1), st);
	if (e == cudaSuccess) e = cudaMemsetAsync(s.cells, 0, sizeof(uint32_t) * (size_t)(ncell + 1), st);
	if (e != cudaSuccess) return (int)e;
	g_launches += 4;
	ch_keys_hist_kernel<D, GAUSS><<<(n + 255) / 256, 256, 0, st>>>(pts, n, g, s.kA, (uint32_t *)cell_start);
	ch_scan_kernel<<<1, 1024, 0, st>>>((const uint32_t *)cell_start, ncell + 1, cell_start);
	ch_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.kA, n, cell_start, s.cells, s.vTmp);
	small_rank_kernel<D, GAUSS><<<(n + 127) / 128, 128, 0, st>>>(pts, n, ncell, cell_start, ids_out, s.kA, s.vTmp, scal, rot, vals, packed, cull);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

// ------------------------------------------------------------------------------------------------
// Sample batches of up to 16384 points on a grid of up to 8191 padded cells (every per-iteration batch of the reference's own
// sizes — the 8192 boundary samples land 50+ to a cell there): ONE launch of ONE 8-CTA thread-block cluster that is a STABLE
// counting sort by cell, with no atomics and no separate ranking pass.  Each of the 64 warps owns a contiguous block of
// samples and a private uint16 histogram row in its CTA's shared memory; inside a round the lanes of one cell find each other
// with one ballot per key bit.  A column scan over the CTA's 8 rows, the other CTAs' per-cell totals read through distributed
// shared memory, a scan over the cells (redundantly per CTA), and every sample knows its slot:
// cell start + samples of earlier CTAs + samples of earlier warps + rank inside its own warp.
// Measured on the way here (profiles/README.md): the same sort in a single 1024-thread CTA is issue-bound on its one SM
// (69k warp instructions, 19 us for 8192 samples), and MATCH.ANY serialises over the distinct values of a warp.
// ------------------------------------------------------------------------------------------------
constexpr int CB_CTAS = 8;
constexpr int CB_THREADS = 256;
constexpr int CB_WARPS = CB_THREADS / 32;
constexpr int CB_MAX_ROUNDS = 8;
constexpr int CB_MAX_N = CB_CTAS * CB_THREADS * CB_MAX_ROUNDS;
constexpr int CB_MAX_SLOTS = 8192;	// cells + the out-of-grid bucket: 26 B of shared memory each, keys fit 13 bits

// what cell_bin_kernel<D, true> needs to draw the samples itself (the boundary batch of project(): gsr_sample_box_surface)
struct SurfaceGen {
	Box box;
	uint64_t seed;
	uint32_t stream_id;
	const float *iteration;
	float *data, *normal;
};

template <int D, bool GEN>
__global__ void __cluster_dims__(CB_CTAS, 1, 1) __launch_bounds__(CB_THREADS, 1)
cell_bin_kernel(const float *__restrict__ x, int n, Grid g, int rounds, int kbits, int32_t *__restrict__ scs, int32_t *__restrict__ perm, SurfaceGen gen)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	extern __shared__ uint32_t cb_smem[];
	__shared__ uint32_t warp_sums[CB_WARPS];
	const int m = g.pcell + 1, mp = (m + 1) & ~1;	// slots, and the (even) pitch of the uint16 rows
	uint32_t *start = cb_smem;	// [mp]  cell totals, then cell starts
	uint32_t *cnt = cb_smem + mp;	// [mp]  this CTA's samples per cell (read by the other CTAs)
	uint16_t *hist = reinterpret_cast<uint16_t *>(cb_smem + 2 * mp);	// [CB_WARPS][mp]
	uint16_t *cbase = hist + CB_WARPS * mp;	// [mp]  samples of earlier CTAs per cell
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int rank = (int)cluster.block_rank();
	const uint32_t lt = (1u << lane) - 1u;
	for (int i = tid; i < (CB_WARPS / 2) * mp; i += CB_THREADS) cb_smem[2 * mp + i] = 0u;
	uint32_t kr[CB_MAX_ROUNDS];	// key, later key | rank inside the warp << 16
	const int base = (rank * CB_WARPS + w) * rounds * 32 + lane;
#pragma unroll
	for (int j = 0; j < CB_MAX_ROUNDS; j++) {
		kr[j] = 0xffffffffu;
		if (j < rounds && base + 32 * j < n) {
			if (GEN) {	// draw sample base + 32 j exactly as sample_box_surface_kernel does, store it, and key it from registers
				const int i = base + 32 * j;
				const Philox r(gen.seed, gen.stream_id, (uint32_t)i, gen.iteration ? (uint32_t)__ldg(gen.iteration) : 0u);
				float pt[3], nm[3];
				box_surface_point(gen.box, r, pt, nm);
#pragma unroll
				for (int k = 0; k < 3; k++) {
					gen.data[3 * (size_t)i + k] = pt[k];
					gen.normal[3 * (size_t)i + k] = nm[k];
				}
				kr[j] = sample_key_of<D, false>(pt, g);
			} else {
				kr[j] = sample_key<D, false>(x, base + 32 * j, g);
			}
		}
	}
	__syncthreads();
	uint16_t *mine = hist + w * mp;
#pragma unroll
	for (int j = 0; j < CB_MAX_ROUNDS; j++) {
		if (j < rounds) {	// uniform
			const uint32_t key = kr[j];
			const bool valid = key != 0xffffffffu;
			uint32_t same = __ballot_sync(0xffffffffu, valid);
			if (!valid) same = ~same;
			for (int b = 0; b < kbits; b++) {
				const uint32_t bit = (key >> b) & 1u, bal = __ballot_sync(0xffffffffu, bit);
				same &= bit ? bal : ~bal;
			}
			uint32_t old = 0u;
			if (valid) old = mine[key];
			__syncwarp();
			if (valid) {
				kr[j] = key | ((old + (uint32_t)__popc(same & lt)) << 16);
				if ((same & lt) == 0u) mine[key] = (uint16_t)(old + (uint32_t)__popc(same));
			}
			__syncwarp();
		}
	}
	__syncthreads();
	for (int c = tid; c < m; c += CB_THREADS) {	// samples of earlier warps of this CTA, per cell
		uint32_t run = 0u;
#pragma unroll
		for (int ww = 0; ww < CB_WARPS; ww++) {
			const uint32_t t = hist[ww * mp + c];
			hist[ww * mp + c] = (uint16_t)run;
			run += t;
		}
		cnt[c] = run;
	}
	cluster.sync();
	{
		const uint32_t *peer[CB_CTAS];
#pragma unroll
		for (int r = 0; r < CB_CTAS; r++) peer[r] = cluster.map_shared_rank(cnt, r);
		for (int c = tid; c < m; c += CB_THREADS) {
			uint32_t before = 0u, total = 0u;
#pragma unroll
			for (int r = 0; r < CB_CTAS; r++) {
				const uint32_t t = peer[r][c];
				if (r < rank) before += t;
				total += t;
			}
			cbase[c] = (uint16_t)before;
			start[c] = total;
		}
	}
	cluster.sync();	// every remote read is done (no CTA may exit, or reuse cnt, before that); also the CTA barrier for start[]
	{	// exclusive scan of the cell totals -> cell starts (start[pcell] = samples inside the padded grid); same in every CTA
		const int chunk = (m + CB_THREADS - 1) / CB_THREADS, b = min(tid * chunk, m), e = min(b + chunk, m);
		uint32_t s = 0u;
		for (int c = b; c < e; c++) s += start[c];
		uint32_t v = s;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
			if (lane >= o) v += t;
		}
		if (lane == 31) warp_sums[w] = v;
		__syncthreads();
		uint32_t before = 0u;
#pragma unroll
		for (int ww = 0; ww < CB_WARPS; ww++) before += ww < w ? warp_sums[ww] : 0u;
		uint32_t run = before + v - s;
		for (int c = b; c < e; c++) {
			const uint32_t t = start[c];
			start[c] = run;
			if (rank == 0) scs[c] = (int32_t)run;
			run += t;
		}
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < CB_MAX_ROUNDS; j++) {
		if (j < rounds && kr[j] != 0xffffffffu) {
			const uint32_t key = kr[j] & 0xffffu;
			perm[start[key] + cbase[key] + mine[key] + (kr[j] >> 16)] = base + 32 * j;
		}
	}
}

static bool cell_bin_ok(int64_t n, int64_t slots)
{
	return !g_force_radix && n > 0 && n <= CB_MAX_N && slots <= CB_MAX_SLOTS;
}

template <int D, bool GEN>
static int launch_cell_bin(const float *x, int n, const Grid &g, int32_t *scs, int32_t *perm, cudaStream_t st, const SurfaceGen &gen = SurfaceGen())
{
	const int m = g.pcell + 1, mp = (m + 1) & ~1;
	const size_t sm = sizeof(uint32_t) * 2 * (size_t)mp + sizeof(uint16_t) * (CB_WARPS + 1) * (size_t)mp;
	if (sm > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(cell_bin_kernel<D, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		if (e != cudaSuccess) return (int)e;
	}
	g_launches += 1;
	int kbits = 1;
	while ((g.pcell >> kbits) != 0) kbits++;	// keys are 0 .. pcell
	const int per_warp = (n + CB_CTAS * CB_WARPS - 1) / (CB_CTAS * CB_WARPS);
	cell_bin_kernel<D, GEN><<<CB_CTAS, CB_THREADS, sm, st>>>(x, n, g, (per_warp + 31) / 32, kbits, scs, perm, gen);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

static bool small_hash_ok(int64_t n, int64_t ncell)
{
	if (g_force_radix) return false;
	// one CTA, shared-memory histogram, and an intra-cell ranking that is quadratic in the cell occupancy
	return n > 0 && n <= SH_MAX_N && ncell <= SH_MAX_CELLS && ncell * 48 >= n;
}

// ------------------------------------------------------------------------------------------------
// min over scalings (reinitialize_grid's `self.scalings.min()`)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_float(float *addr, float v)
{
	if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
	else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__global__ void min_init_kernel(float *out) { *out = __int_as_float(0x7f800000); }

__global__ void min_kernel(const float *__restrict__ a, int64_t n, float *out)
{
	float m = __int_as_float(0x7f800000);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fminf(m, a[i]);
#pragma unroll
	for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0) atomic_min_float(out, m);
}

}  // namespace gsr

using namespace gsr;

extern "C" size_t gsr_build_grid_ws_bytes(const gsr_grid_desc *d, int64_t N)
{
	Grid g;
	return sort_ws_bytes(N, make_grid(d, g) ? g.ncell : 0);
}
extern "C" size_t gsr_bin_samples_ws_bytes(const gsr_grid_desc *d, int64_t Q)
{
	Grid g;
	return sort_ws_bytes(Q, make_grid(d, g) ? g.pcell : 0);
}

extern "C" int64_t gsr_padded_cells(const gsr_grid_desc *d)
{
	Grid g;
	if (!make_grid(d, g)) return GSR_EINVAL;
	return g.pcell;
}

static int launch_pack(const Grid &g, const float *positions, const float *scalings, const float *rotations, const float *values, int n,
		       const int32_t *sorted_id, float *packed, float *cull, cudaStream_t st)
{
	g_launches += 1;
	if (g.D == 3) pack3d_kernel<<<(n + 127) / 128, 128, 0, st>>>(positions, scalings, rotations, values, n, sorted_id, (float4 *)packed, cull);
	else pack2d_kernel<<<(n + 127) / 128, 128, 0, st>>>(positions, scalings, rotations, values, n, sorted_id, (float4 *)packed, cull);
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}

extern "C" int gsr_build_grid(const gsr_grid_desc *d, const float *positions, int64_t N,
			      int32_t *cell_start, int32_t *sorted_id, int32_t *grid_cnt, int32_t *grid_offset,
			      const float *scalings, const float *rotations, const float *values, float *packed, float *cull,
			      void *ws, size_t ws_bytes, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || N < 0 || N >= ((int64_t)1 << 30) || !cell_start || !sorted_id) return GSR_EINVAL;
	if (packed && (!scalings || !rotations || !values)) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	SortWs s;
	if (!carve_sort_ws(ws, ws_bytes, N, s)) return GSR_EWS;
	int n = (int)N;
	if (small_hash_ok(N, g.ncell)) {
		// one launch: keys, histogram, scan, scatter, canonical order and (optionally) the packed records
		int rc = (g.D == 3) ? launch_small_hash<3, true>(positions, n, g, g.ncell, cell_start, sorted_id, s.kA, s.vTmp, scalings, rotations, values, (float4 *)packed, cull, st)
				    : launch_small_hash<2, true>(positions, n, g, g.ncell, cell_start, sorted_id, s.kA, s.vTmp, scalings, rotations, values, (float4 *)packed, cull, st);
		if (rc) return rc;
	} else if (count_hash_ok(N, g.ncell, s)) {
		int rc = (g.D == 3) ? launch_count_hash<3, true>(positions, n, g, g.ncell, cell_start, sorted_id, s, scalings, rotations, values, (float4 *)packed, cull, st)
				    : launch_count_hash<2, true>(positions, n, g, g.ncell, cell_start, sorted_id, s, scalings, rotations, values, (float4 *)packed, cull, st);
		if (rc) return rc;
	} else {
		const uint32_t *ks = s.keys0;
		if (n > 0) {
			if (g.D == 3) gauss_keys_kernel<3><<<(n + 255) / 256, 256, 0, st>>>(positions, n, g, s.keys0);
			else gauss_keys_kernel<2><<<(n + 255) / 256, 256, 0, st>>>(positions, n, g, s.keys0);
			GSR_CHECK_LAUNCH();
			int rc = radix_sort_index(s, n, (uint32_t)g.ncell, (uint32_t *)sorted_id, &ks, st);
			if (rc) return rc;
		}
		g_launches += (n > 0 ? 2 : 1);
		cell_start_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(ks, n, g.ncell, cell_start, 0);
		GSR_CHECK_LAUNCH();
		if (packed && n > 0) {
			int rc = launch_pack(g, positions, scalings, rotations, values, n, sorted_id, packed, cull, st);
			if (rc) return rc;
		}
	}
	if (grid_cnt || grid_offset) {
		g_launches += 1;
		ref_format_kernel<<<(g.ncell + 255) / 256, 256, 0, st>>>(cell_start, g.ncell, grid_cnt, grid_offset);
		GSR_CHECK_LAUNCH();
	}
	return GSR_OK;
}

extern "C" int gsr_bin_samples(const gsr_grid_desc *d, const float *x, int64_t Q, int32_t *perm, int32_t *sample_cell_start, int fine,
			       void *ws, size_t ws_bytes, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || Q < 0 || Q >= ((int64_t)1 << 30) || !perm) return GSR_EINVAL;
	const int shift = (fine && g.pcell < (1 << 25)) ? 2 * g.D : 0;
	cudaStream_t st = (cudaStream_t)stream;
	SortWs s;
	if (!carve_sort_ws(ws, ws_bytes, Q, s)) return GSR_EWS;
	int n = (int)Q;
	if (!shift && cell_bin_ok(Q, (int64_t)g.pcell + 1)) {
		int32_t *scs = sample_cell_start ? sample_cell_start : ((size_t)(g.pcell + 1) <= 256 * (size_t)s.nblocks + 1024 ? (int32_t *)s.hist : nullptr);
		if (scs) return (g.D == 3) ? launch_cell_bin<3, false>(x, n, g, scs, perm, st) : launch_cell_bin<2, false>(x, n, g, scs, perm, st);
	}
	if (!shift && small_hash_ok(Q, g.pcell)) {
		// sample_cell_start is produced as a by-product; when the caller does not want it, it lands in scratch
		int32_t *scs = sample_cell_start ? sample_cell_start : (int32_t *)s.hist;
		if (!sample_cell_start && (size_t)(g.pcell + 1) > 256 * (size_t)s.nblocks) scs = nullptr;
		if (scs) {
			return (g.D == 3) ? launch_small_hash<3, false>(x, n, g, g.pcell, scs, perm, s.kA, s.vTmp, nullptr, nullptr, nullptr, nullptr, nullptr, st)
					  : launch_small_hash<2, false>(x, n, g, g.pcell, scs, perm, s.kA, s.vTmp, nullptr, nullptr, nullptr, nullptr, nullptr, st);
		}
	}
	if (!shift && count_hash_ok(Q, g.pcell, s)) {
		// the cell table is a by-product; without a caller buffer it lands in the radix histogram scratch when that is large enough
		int32_t *scs = sample_cell_start ? sample_cell_start : ((size_t)(g.pcell + 1) <= 256 * (size_t)s.nblocks ? (int32_t *)s.hist : nullptr);
		if (scs)
			return (g.D == 3) ? launch_count_hash<3, false>(x, n, g, g.pcell, scs, perm, s, nullptr, nullptr, nullptr, nullptr, nullptr, st)
					  : launch_count_hash<2, false>(x, n, g, g.pcell, scs, perm, s, nullptr, nullptr, nullptr, nullptr, nullptr, st);
	}
	const uint32_t *ks = s.keys0;
	if (n > 0) {
		if (g.D == 3) {
			if (shift) sample_keys_kernel<3, true><<<(n + 255) / 256, 256, 0, st>>>(x, n, g, s.keys0);
			else sample_keys_kernel<3, false><<<(n + 255) / 256, 256, 0, st>>>(x, n, g, s.keys0);
		} else {
			if (shift) sample_keys_kernel<2, true><<<(n + 255) / 256, 256, 0, st>>>(x, n, g, s.keys0);
			else sample_keys_kernel<2, false><<<(n + 255) / 256, 256, 0, st>>>(x, n, g, s.keys0);
		}
		GSR_CHECK_LAUNCH();
		int rc = radix_sort_index(s, n, ((uint32_t)g.pcell << shift) | ((1u << shift) - 1u), (uint32_t *)perm, &ks, st);
		if (rc) return rc;
	}
	g_launches += (n > 0 ? 1 : 0) + (sample_cell_start ? 1 : 0);
	if (sample_cell_start) {
		cell_start_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(ks, n, g.pcell, sample_cell_start, shift);
		GSR_CHECK_LAUNCH();
	}
	return GSR_OK;
}

extern "C" int gsr_sample_box_surface(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev, float *data, float *normal,
				      void *stream);

extern "C" int gsr_sample_box_surface_binned(const float *box, int64_t n, uint64_t seed, uint32_t stream_id, const float *iteration_dev,
					     float *data, float *normal, const gsr_grid_desc *d, int32_t *perm, int32_t *sample_cell_start,
					     void *ws, size_t ws_bytes, void *stream)
{
	Grid g;
	if (!box || !data || !normal || !perm || !sample_cell_start || !make_grid(d, g) || g.D != 3 || n < 0 || n >= ((int64_t)1 << 30)) return GSR_EINVAL;
	if (n > 0 && cell_bin_ok(n, (int64_t)g.pcell + 1)) {	// one launch draws the samples and sorts them by cell
		SurfaceGen gen = {make_box(box), seed, stream_id, iteration_dev, data, normal};
		return launch_cell_bin<3, true>(nullptr, (int)n, g, sample_cell_start, perm, (cudaStream_t)stream, gen);
	}
	int rc = gsr_sample_box_surface(box, n, seed, stream_id, iteration_dev, data, normal, stream);
	if (rc) return rc;
	return gsr_bin_samples(d, data, n, perm, sample_cell_start, 0, ws, ws_bytes, stream);
}

extern "C" int gsr_pack_gaussians(const gsr_grid_desc *d, const float *positions, const float *scalings, const float *rotations,
				  const float *values, int64_t N, const int32_t *, const int32_t *sorted_id, float *packed, float *cull, void *stream)
{
	Grid g;
	if (!make_grid(d, g) || N < 0 || !packed || !sorted_id) return GSR_EINVAL;
	if (N == 0) return GSR_OK;
	return launch_pack(g, positions, scalings, rotations, values, (int)N, sorted_id, packed, cull, (cudaStream_t)stream);
}

extern "C" int gsr_min_scaling(const float *scalings, int64_t count, float *out_min, void *stream)
{
	if (!out_min || count < 0) return GSR_EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	g_launches += count > 0 ? 2 : 1;
	min_init_kernel<<<1, 1, 0, st>>>(out_min);
	if (count > 0) {
		int blocks = (int)((count + 1023) / 1024);
		if (blocks > kSMs * 8) blocks = kSMs * 8;
		min_kernel<<<blocks, 256, 0, st>>>(scalings, count, out_min);
	}
	GSR_CHECK_LAUNCH();
	return GSR_OK;
}
