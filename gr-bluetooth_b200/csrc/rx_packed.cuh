// rx_packed.cuh -- Blackwell packed-fp32 helpers (FMUL2 / FADD2 / FFMA2), each half rounding exactly like
// the scalar instruction.  Shared by the tiled FIR (rx_kernels.cu) and the delay-line FIR (rx_firdl.cu).
#pragma once
namespace btb200 {
typedef unsigned long long u64;
// NOTE on ptxas 12.9: it contracts mul.rn.f32x2 feeding add/sub.rn.f32x2 into one FFMA2 even
// though the operations carry an explicit .rn (it does not do that for scalar mul.rn/add.rn).
// That would change the rounding.  A product must therefore never be the direct operand of a
// packed add/sub: "x - p" is written fma(p, -1, x) (exact: p*(-1) is exact, one rounding),
// which ptxas keeps as FFMA2 with an immediate and cannot merge with the FMUL2 that made p.
// The SASS is checked for this in tests/test_build.py.
__device__ __forceinline__ u64 pk_mul(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pk_add(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// x - p, p a product
__device__ __forceinline__ u64 pk_xsubp(u64 x, u64 p)
{ u64 d; asm("{.reg .b64 m1; mov.b64 m1, 0xbf800000bf800000; fma.rn.f32x2 %0, %2, m1, %1;}" : "=l"(d) : "l"(x), "l"(p)); return d; }
// (x + p cannot be written fma(p, +1, x): ptxas folds the multiply by one and then contracts p's FMUL2 into a
// real FFMA2.  Sums of products are therefore written x - (-p) with one factor negated: pk_neg + pk_xsubp.)
// a * b + c, one rounding per half (tolerance-mode kernels only: the exact path never fuses)
__device__ __forceinline__ u64 pk_fma(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk_neg(u64 a) { return a ^ 0x8000000080000000ull; }
__device__ __forceinline__ float pk_lo(u64 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float pk_hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ u64 pk_pack(float lo, float hi)
{ u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi))); return d; }
}  // namespace btb200
