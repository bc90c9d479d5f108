/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU oracle for the Gaussian-Fluids hot path: a restatement in C of the reference's
 * Taichi kernels (3D/GSR.py, 2D/GSR.py).  Build: `make -C oracle` → oracle/_build/libgsr_oracle.so.
 * Loaded through ctypes by oracle/oracle.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it (as the checker / the CPU
 * baseline), never the product path.
 *
 * Parity status: pinned (see gsr3d_oracle_impl.h header).
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>

typedef struct {
	float lo[3], hi[3];	/* extended domain, rounded to f32 like the kernel constants */
	float grid_scale;
	int dims[3];
	const int *cnt, *offset, *sorted_id;
} gsr_grid3;

typedef struct {
	float lo[2], hi[2];
	float grid_scale;
	int dims[2];
	const int *cnt, *offset, *sorted_id;
} gsr_grid2;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* `int((p - x_min) // grid_scale)` on f32 operands — 3D/GSR.py:213, :271 */
static inline int cell1(float p, float lo, float gs) { return (int)floorf((p - lo) / gs); }

static inline void gsr_cell_of_f32(const gsr_grid3 *g, float x, float y, float z, int c[3])
{
	c[0] = cell1(x, g->lo[0], g->grid_scale);
	c[1] = cell1(y, g->lo[1], g->grid_scale);
	c[2] = cell1(z, g->lo[2], g->grid_scale);
}
static inline void gsr_cell_of2_f32(const gsr_grid2 *g, float x, float y, int c[2])
{
	c[0] = cell1(x, g->lo[0], g->grid_scale);
	c[1] = cell1(y, g->lo[1], g->grid_scale);
}

/*
 * reinitialize_grid_ti, 3D/GSR.py:205-245.  Executed serially, the reference's scatter
 * loop visits ids in ascending order, so each cell's segment of sorted_id comes out
 * ascending — the canonical form the CUDA path is compared against bit-for-bit.
 * Returns the number of Gaussians inside the (extended) domain.
 */
int o3_build_grid(const float *pos, long N, const float *lo, const float *hi, float grid_scale, const int *dims,
		  int *cnt, int *offset, int *sorted_id)
{
	long ncell = (long)dims[0] * dims[1] * dims[2];
	memset(cnt, 0, sizeof(int) * ncell);
	for (long i = 0; i < N; i++) {
		const float *p = pos + 3 * i;
		if (lo[0] <= p[0] && p[0] <= hi[0] && lo[1] <= p[1] && p[1] <= hi[1] && lo[2] <= p[2] && p[2] <= hi[2]) {
			int ix = cell1(p[0], lo[0], grid_scale), iy = cell1(p[1], lo[1], grid_scale), iz = cell1(p[2], lo[2], grid_scale);
			cnt[((long)ix * dims[1] + iy) * dims[2] + iz] += 1;
		}
	}
	/* per-x prefix then row-major scan inside each x slab == exclusive scan in row-major cell order */
	int run = 0;
	for (long c = 0; c < ncell; c++) { offset[c] = run; run += cnt[c]; }
	memset(cnt, 0, sizeof(int) * ncell);
	for (long i = 0; i < N; i++) {
		const float *p = pos + 3 * i;
		if (lo[0] <= p[0] && p[0] <= hi[0] && lo[1] <= p[1] && p[1] <= hi[1] && lo[2] <= p[2] && p[2] <= hi[2]) {
			int ix = cell1(p[0], lo[0], grid_scale), iy = cell1(p[1], lo[1], grid_scale), iz = cell1(p[2], lo[2], grid_scale);
			long c = ((long)ix * dims[1] + iy) * dims[2] + iz;
			sorted_id[offset[c] + cnt[c]++] = (int)i;
		}
	}
	return run;
}

/* reinitialize_grid_ti, 2D/GSR.py:194-222 */
int o2_build_grid(const float *pos, long N, const float *lo, const float *hi, float grid_scale, const int *dims,
		  int *cnt, int *offset, int *sorted_id)
{
	long ncell = (long)dims[0] * dims[1];
	memset(cnt, 0, sizeof(int) * ncell);
	for (long i = 0; i < N; i++) {
		const float *p = pos + 2 * i;
		if (lo[0] <= p[0] && p[0] <= hi[0] && lo[1] <= p[1] && p[1] <= hi[1])
			cnt[(long)cell1(p[0], lo[0], grid_scale) * dims[1] + cell1(p[1], lo[1], grid_scale)] += 1;
	}
	int run = 0;
	for (long c = 0; c < ncell; c++) { offset[c] = run; run += cnt[c]; }
	memset(cnt, 0, sizeof(int) * ncell);
	for (long i = 0; i < N; i++) {
		const float *p = pos + 2 * i;
		if (lo[0] <= p[0] && p[0] <= hi[0] && lo[1] <= p[1] && p[1] <= hi[1]) {
			long c = (long)cell1(p[0], lo[0], grid_scale) * dims[1] + cell1(p[1], lo[1], grid_scale);
			sorted_id[offset[c] + cnt[c]++] = (int)i;
		}
	}
	return run;
}

/* get_all_neighbors_ti, 3D/GSR.py:679-690 */
void o3_mark_neighbors(const gsr_grid3 *g, const float *pos, const float *x, long Q, int *mark)
{
	for (long j = 0; j < Q; j++) {
		int c[3];
		gsr_cell_of_f32(g, x[3 * j], x[3 * j + 1], x[3 * j + 2], c);
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++)
		for (int gk = imax(c[2] - 1, 0); gk <= imin(c[2] + 1, g->dims[2] - 1); gk++) {
			long cell = ((long)gi * g->dims[1] + gj) * g->dims[2] + gk;
			for (int t = g->offset[cell]; t < g->offset[cell] + g->cnt[cell]; t++) {
				int i = g->sorted_id[t];
				float dx = x[3 * j] - pos[3 * i], dy = x[3 * j + 1] - pos[3 * i + 1], dz = x[3 * j + 2] - pos[3 * i + 2];
				if (sqrtf(dx * dx + dy * dy + dz * dz) <= g->grid_scale) mark[i] = 1;
			}
		}
	}
}

/* get_all_neighbors_ti, 2D/GSR.py:620-630 */
void o2_mark_neighbors(const gsr_grid2 *g, const float *pos, const float *x, long Q, int *mark)
{
	for (long j = 0; j < Q; j++) {
		int c[2];
		gsr_cell_of2_f32(g, x[2 * j], x[2 * j + 1], c);
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++) {
			long cell = (long)gi * g->dims[1] + gj;
			for (int t = 0; t < g->cnt[cell]; t++) {
				int i = g->sorted_id[g->offset[cell] + t];
				float dx = x[2 * j] - pos[2 * i], dy = x[2 * j + 1] - pos[2 * i + 1];
				if (sqrtf(dx * dx + dy * dy) <= g->grid_scale) mark[i] = 1;
			}
		}
	}
}

/*
 * Work counters for the benchmark's unit of work: C = candidate visits,
 * P = accepted pairs (g >= tau), for one forward evaluation of the Q points.
 */
void o3_count_pairs(const gsr_grid3 *g, const float *x, long Q, long long *C_out)
{
	long long C = 0;
	#pragma omp parallel for reduction(+:C)
	for (long j = 0; j < Q; j++) {
		int c[3];
		gsr_cell_of_f32(g, x[3 * j], x[3 * j + 1], x[3 * j + 2], c);
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++)
		for (int gk = imax(c[2] - 1, 0); gk <= imin(c[2] + 1, g->dims[2] - 1); gk++)
			C += g->cnt[((long)gi * g->dims[1] + gj) * g->dims[2] + gk];
	}
	*C_out = C;
}

#define REAL float
#define SFX _f32
#include "gsr3d_oracle_impl.h"
#include "gsr2d_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL double
#define SFX _f64
#include "gsr3d_oracle_impl.h"
#include "gsr2d_oracle_impl.h"
#undef REAL
#undef SFX

/*
 * Pair classification for the borderline-band policy: for every sample,
 * the number of accepted pairs and the number of pairs whose quadratic form lies within
 * rel_band of q_max = -2 ln(tau) (computed in double).  n_acc/n_band are (Q,) int arrays.
 */
void o3_classify_pairs(const gsr_grid3 *g, const float *pos, const float *scal, const float *rot,
		       double tau, double rel_band, const float *x, long Q, int *n_acc, int *n_band)
{
	const double qmax = -2.0 * log((double)(float)tau);
	#pragma omp parallel for schedule(dynamic, 64)
	for (long j = 0; j < Q; j++) {
		int c[3], na = 0, nb = 0;
		gsr_cell_of_f32(g, x[3 * j], x[3 * j + 1], x[3 * j + 2], c);
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++)
		for (int gk = imax(c[2] - 1, 0); gk <= imin(c[2] + 1, g->dims[2] - 1); gk++) {
			long cell = ((long)gi * g->dims[1] + gj) * g->dims[2] + gk;
			for (int t = g->offset[cell]; t < g->offset[cell] + g->cnt[cell]; t++) {
				int i = g->sorted_id[t];
				V3_f64 d = {{(double)x[3 * j] - pos[3 * i], (double)x[3 * j + 1] - pos[3 * i + 1], (double)x[3 * j + 2] - pos[3 * i + 2]}};
				double q[4]; M3_f64 R, S2, C;
				gauss_geom_f64(rot + 4 * i, scal + 3 * i, q, &R, &S2, &C);
				double quad = v_dot_f64(v_mulm_f64(d, C), d);
				if (quad <= qmax) na++;
				if (fabs(quad - qmax) <= rel_band * qmax) nb++;
			}
		}
		n_acc[j] = na; n_band[j] = nb;
	}
}

void o2_classify_pairs(const gsr_grid2 *g, const float *pos, const float *scal, const float *rot,
		       double tau, double rel_band, const float *x, long Q, int *n_acc, int *n_band)
{
	const double qmax = -2.0 * log((double)(float)tau);
	#pragma omp parallel for schedule(dynamic, 64)
	for (long j = 0; j < Q; j++) {
		int c[2], na = 0, nb = 0;
		gsr_cell_of2_f32(g, x[2 * j], x[2 * j + 1], c);
		for (int gi = imax(c[0] - 1, 0); gi <= imin(c[0] + 1, g->dims[0] - 1); gi++)
		for (int gj = imax(c[1] - 1, 0); gj <= imin(c[1] + 1, g->dims[1] - 1); gj++) {
			long cell = (long)gi * g->dims[1] + gj;
			for (int t = 0; t < g->cnt[cell]; t++) {
				int i = g->sorted_id[g->offset[cell] + t];
				double d[2] = {(double)x[2 * j] - pos[2 * i], (double)x[2 * j + 1] - pos[2 * i + 1]};
				double C[2][2];
				geom2_f64(rot[i], scal + 2 * i, C);
				double quad = (d[0] * C[0][0] + d[1] * C[1][0]) * d[0] + (d[0] * C[0][1] + d[1] * C[1][1]) * d[1];
				if (quad <= qmax) na++;
				if (fabs(quad - qmax) <= rel_band * qmax) nb++;
			}
		}
		n_acc[j] = na; n_band[j] = nb;
	}
}

/* ---- N3: mesh boundary sampler (3D/mesh_sampler.py:12-21, :60-88), float32 as in the reference --------------------------
 * o3_mesh_area_presum: per-face area, then the SERIAL inclusive prefix sum of ti_get_tri_area.
 * o3_mesh_sample: the map (three uniforms per sample) -> (point, normal) of ti_lower_bound + ti_sample. */
void o3_mesh_area_presum(const float *vertices, const int *faces, long F, float *presum)
{
	for (long i = 0; i < F; i++) {
		const float *a = vertices + 3 * faces[3 * i], *b = vertices + 3 * faces[3 * i + 1], *c = vertices + 3 * faces[3 * i + 2];
		const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
		const float x = e1[1] * e2[2] - e1[2] * e2[1], y = e1[2] * e2[0] - e1[0] * e2[2], z = e1[0] * e2[1] - e1[1] * e2[0];
		presum[i] = sqrtf(x * x + y * y + z * z) * .5f;
	}
	for (long i = 1; i < F; i++) presum[i] += presum[i - 1];
}

void o3_mesh_sample(long n, const float *uniforms, const float *vertices, const float *normals, const int *faces, const int *facenormals,
		    const float *presum, long F, float *data, float *normal)
{
	const float total = presum[F - 1];
	for (long i = 0; i < n; i++) {
		const float t = uniforms[3 * i] * total;
		long l = 0, r = F;
		while (l < r) {
			const long m = (l + r) / 2;
			if (presum[m] < t) l = m + 1;
			else r = m;
		}
		const long f = l < F - 1 ? l : F - 1;
		const float u = 1.f - sqrtf(uniforms[3 * i + 1]), v = uniforms[3 * i + 2] * (1.f - u), w = 1.f - u - v;
		float nn[3];
		for (int k = 0; k < 3; k++) {
			data[3 * i + k] = u * vertices[3 * faces[3 * f] + k] + v * vertices[3 * faces[3 * f + 1] + k] + w * vertices[3 * faces[3 * f + 2] + k];
			nn[k] = u * normals[3 * facenormals[3 * f] + k] + v * normals[3 * facenormals[3 * f + 1] + k] + w * normals[3 * facenormals[3 * f + 2] + k];
		}
		const float len = sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
		for (int k = 0; k < 3; k++) normal[3 * i + k] = nn[k] / len;
	}
}
