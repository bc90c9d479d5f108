// chain.cuh — per-Gaussian chain rule from dL/dSigma^-1 to the reference's parameters (scalings, rotation).
// Shared by the backward epilogue (backward.cu) and the fused optimiser step (step.cu).
#pragma once
#include "common.cuh"

namespace gsr {

// 3D:  Sigma^-1 = R(q) diag(e^{2s}) R(q)^T,  q = r/|r|.
// dL/ds_k = 2 e^{2 s_k} r_k^T G r_k ;  dL/dr_m = 2 <G R S^2, dR/dr_m>,
// dR/dr_m = -r_m/|r|^3 sum_n r_n dR/dq_n + (dR/dq_m)/|r|     (3D/GSR.py:328-368, contracted analytically)
__device__ __forceinline__ void chain3d(const float G6[6], const float s[3], const float r[4], float ds[3], float dr[4])
{
	const float len2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
	const float len = sqrtf(len2), inv = 1.f / len;
	const float q0 = r[0] * inv, q1 = r[1] * inv, q2 = r[2] * inv, q3 = r[3] * inv;
	const float R[3][3] = {{1.f - 2.f * (q2 * q2 + q3 * q3), 2.f * (q1 * q2 - q0 * q3), 2.f * (q1 * q3 + q0 * q2)},
			       {2.f * (q1 * q2 + q0 * q3), 1.f - 2.f * (q1 * q1 + q3 * q3), 2.f * (q2 * q3 - q0 * q1)},
			       {2.f * (q1 * q3 - q0 * q2), 2.f * (q2 * q3 + q0 * q1), 1.f - 2.f * (q1 * q1 + q2 * q2)}};
	const float G[3][3] = {{G6[0], G6[1], G6[2]}, {G6[1], G6[3], G6[4]}, {G6[2], G6[4], G6[5]}};
	const float e[3] = {expf(2.f * s[0]), expf(2.f * s[1]), expf(2.f * s[2])};
	float H[3][3];	// G R S^2
#pragma unroll
	for (int a = 0; a < 3; a++)
#pragma unroll
		for (int k = 0; k < 3; k++) H[a][k] = (G[a][0] * R[0][k] + G[a][1] * R[1][k] + G[a][2] * R[2][k]) * e[k];
#pragma unroll
	for (int k = 0; k < 3; k++) ds[k] = 2.f * (R[0][k] * H[0][k] + R[1][k] * H[1][k] + R[2][k] * H[2][k]);
	// P_n = <H, dR/dq_n>
	const float P0 = 2.f * (q3 * (H[1][0] - H[0][1]) + q2 * (H[0][2] - H[2][0]) + q1 * (H[2][1] - H[1][2]));
	const float P1 = 2.f * (q2 * (H[0][1] + H[1][0]) + q3 * (H[0][2] + H[2][0]) + q0 * (H[2][1] - H[1][2])) - 4.f * q1 * (H[1][1] + H[2][2]);
	const float P2 = 2.f * (q1 * (H[0][1] + H[1][0]) + q0 * (H[0][2] - H[2][0]) + q3 * (H[1][2] + H[2][1])) - 4.f * q2 * (H[0][0] + H[2][2]);
	const float P3 = 2.f * (q0 * (H[1][0] - H[0][1]) + q1 * (H[0][2] + H[2][0]) + q2 * (H[1][2] + H[2][1])) - 4.f * q3 * (H[0][0] + H[1][1]);
	const float rp = (r[0] * P0 + r[1] * P1 + r[2] * P2 + r[3] * P3) * inv * inv * inv;
	dr[0] = 2.f * (P0 * inv - r[0] * rp);
	dr[1] = 2.f * (P1 * inv - r[1] * rp);
	dr[2] = 2.f * (P2 * inv - r[2] * rp);
	dr[3] = 2.f * (P3 * inv - r[3] * rp);
}

// 2D:  Sigma^-1 = R(theta) diag(e^{2s}) R(theta)^T;  dSigma^-1/dtheta = (e0 - e1) [[-sin 2t, cos 2t],[cos 2t, sin 2t]]   (2D/GSR.py:468)
__device__ __forceinline__ void chain2d(const float G3[3], const float s[2], float theta, float ds[2], float *dth)
{
	float sn, cs;
	sincosf(theta, &sn, &cs);
	const float e0 = expf(2.f * s[0]), e1 = expf(2.f * s[1]);
	const float a0 = cs * cs * G3[0] + 2.f * cs * sn * G3[1] + sn * sn * G3[2];	// r0^T G r0, r0 = (cos, sin)
	const float a1 = sn * sn * G3[0] - 2.f * cs * sn * G3[1] + cs * cs * G3[2];	// r1 = (-sin, cos)
	ds[0] = 2.f * e0 * a0;
	ds[1] = 2.f * e1 * a1;
	const float s2 = 2.f * sn * cs, c2 = cs * cs - sn * sn;
	*dth = (e0 - e1) * (-s2 * G3[0] + 2.f * c2 * G3[1] + s2 * G3[2]);
}

}  // namespace gsr
