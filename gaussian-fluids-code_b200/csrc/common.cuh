// common.cuh — shared device helpers of the B200-native GSR engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gsr_b200.h"

#define GSR_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

#include <atomic>

namespace gsr {

extern std::atomic<unsigned long long> g_launches;	// kernels launched by this library (misc.cu)

constexpr int kSMs = 148;	// B200

struct Grid {
	int D;
	int dims[3];
	float lo[3], hi[3];
	float gs;
	const float *gs_dev;	// optional device-resident grid_scale (overrides gs)
	float tau;
	int ncell;	// prod(dims)
	int pdims[3];	// dims + 2 (padded sample grid: one virtual layer on each side)
	int pcell;	// prod(pdims)
};

inline bool make_grid(const gsr_grid_desc *d, Grid &g)
{
	if (!d || (d->D != 2 && d->D != 3)) return false;
	g.D = d->D;
	int64_t nc = 1, pc = 1;
	for (int k = 0; k < 3; k++) {
		g.dims[k] = (k < d->D) ? d->dims[k] : 1;
		if (g.dims[k] < 1) return false;
		g.pdims[k] = (k < d->D) ? g.dims[k] + 2 : 1;
		g.lo[k] = (k < d->D) ? d->lo[k] : 0.f;
		g.hi[k] = (k < d->D) ? d->hi[k] : 0.f;
		nc *= g.dims[k];
		pc *= g.pdims[k];
	}
	if (pc >= (int64_t)1 << 30) return false;
	g.ncell = (int)nc;
	g.pcell = (int)pc;
	g.gs = d->grid_scale;
	g.gs_dev = d->grid_scale_dev;
	g.tau = d->tau;
	return true;
}

// `int((p - x_min) // grid_scale)` of the reference (3D/GSR.py:213, :271) — IEEE f32 subtract, divide, floor.
// Written with the _rn intrinsics so that no compiler flag (fast-math, fmad) can change a cell index.
__device__ __forceinline__ int cell_coord(float p, float lo, float gs)
{
	return (int)floorf(__fdiv_rn(__fsub_rn(p, lo), gs));
}

__device__ __forceinline__ float grid_gs(const Grid &g) { return g.gs_dev ? __ldg(g.gs_dev) : g.gs; }

__device__ __forceinline__ int iclamp(int v, int a, int b) { return v < a ? a : (v > b ? b : v); }

}  // namespace gsr
