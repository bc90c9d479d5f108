// f32x2.cuh — Blackwell packed-FP32 arithmetic (PTX fma/add/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2).
//
// One instruction performs two IEEE round-to-nearest FP32 operations on a 64-bit register pair, i.e. half the issue
// slots and half the register-bank reads of the scalar form for the same flops.  The evaluation kernels put TWO POINTS
// of a thread in the two halves and feed the candidate Gaussian's scalars through bc(): ptxas folds pack(s, s) into the
// scalar-broadcast operand form (`FFMA2 R, R.F32x2.HI_LO, R.F32, R.F32x2.HI_LO`), so the broadcast costs no instruction.
// Every half is computed exactly as the scalar code would (same operation order, fused where the scalar code fuses),
// so results are bit-identical to the scalar kernels.
#pragma once
#include <cuda_runtime.h>

namespace gsr {

typedef unsigned long long f2;	// {lo, hi} = two floats

__device__ __forceinline__ f2 pack2(float lo, float hi)
{
	f2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ f2 bc(float s) { return pack2(s, s); }
__device__ __forceinline__ void unpack2(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float lo2(f2 v) { float a, b; unpack2(v, a, b); return a; }
__device__ __forceinline__ float hi2(f2 v) { float a, b; unpack2(v, a, b); return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
	f2 r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
	f2 r;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
	f2 r;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}

}  // namespace gsr
