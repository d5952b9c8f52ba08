// 3-vector arithmetic with a fixed operation order, plus the deterministic sin/cos/log2 used by the
// shading kernels.
//
// Arithmetic contract: binary32, one IEEE-754 operation per written operator, strict
// left-to-right association, no implicit FMA (nvcc -fmad=false), correctly rounded sqrt and
// division (no --use_fast_math). Where a fused multiply-add is part of the spec it is written as
// fma_rn(). With that contract the GPU and the CPU oracle agree bit for bit.
#pragma once
#include "compat.cuh"

namespace brt {

struct f3 {
  float x, y, z;
};
struct f2 {
  float x, y;
};
BRT_HD f3 F3(float x, float y, float z) { return f3{x, y, z}; }
BRT_HD f3 F3(float s) { return f3{s, s, s}; }
BRT_HD f3 operator+(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
BRT_HD f3 operator-(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
BRT_HD f3 operator*(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
BRT_HD f3 operator*(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
BRT_HD f3 operator*(float s, f3 a) { return F3(s * a.x, s * a.y, s * a.z); }
BRT_HD f3 operator-(f3 a) { return F3(-a.x, -a.y, -a.z); }
BRT_HD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
BRT_HD f3 cross(f3 a, f3 b) { return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
BRT_HD float length(f3 a) { return sqrtf(dot(a, a)); }
// normalize(v) := v * (1 / sqrt(dot(v, v)))
BRT_HD f3 normalize(f3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
BRT_HD float square(float f) { return f * f; }  // SH/shadermath.slang:3
BRT_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// lerp(x, y, s) := x*(1-s) + y*s  (FMix of the shipped SPIR-V)
BRT_HD float lerpf(float x, float y, float s) { return x * (1.0f - s) + y * s; }
BRT_HD f3 lerp3(f3 x, f3 y, float s) { return F3(lerpf(x.x, y.x, s), lerpf(x.y, y.y, s), lerpf(x.z, y.z, s)); }
BRT_HD f3 reflect(f3 i, f3 n) { return i - n * (2.0f * dot(n, i)); }
BRT_HD float comp(f3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
BRT_HD f3 fmin3(f3 a, f3 b) { return F3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
BRT_HD f3 fmax3(f3 a, f3 b) { return F3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
BRT_HD f3 xyz(float4 v) { return F3(v.x, v.y, v.z); }

// row-major 3x4 affine transforms (VkTransformMatrixKHR, RT/Scene.cpp:183-190)
BRT_HD f3 xform_point(const float4 m[3], f3 p) {
  return F3(((m[0].x * p.x + m[0].y * p.y) + m[0].z * p.z) + m[0].w, ((m[1].x * p.x + m[1].y * p.y) + m[1].z * p.z) + m[1].w,
            ((m[2].x * p.x + m[2].y * p.y) + m[2].z * p.z) + m[2].w);
}
BRT_HD f3 xform_dir(const float4 m[3], f3 p) {
  return F3((m[0].x * p.x + m[0].y * p.y) + m[0].z * p.z, (m[1].x * p.x + m[1].y * p.y) + m[1].z * p.z,
            (m[2].x * p.x + m[2].y * p.y) + m[2].z * p.z);
}
// mul(n, WorldToObject4x3): n'_j = sum_i W2O(i,j) n_i  — the inverse-transpose normal transform of
// SH/raytracing.slang:150
BRT_HD f3 xform_normal(const float4 w2o[3], f3 n) {
  return F3((w2o[0].x * n.x + w2o[1].x * n.y) + w2o[2].x * n.z, (w2o[0].y * n.x + w2o[1].y * n.y) + w2o[2].y * n.z,
            (w2o[0].z * n.x + w2o[1].z * n.y) + w2o[2].z * n.z);
}

// world box of an instance = |M| applied to the mesh box in centre/extent form (shared by the TLAS
// builder and Smart Culling; the oracle uses the same formula so culling decisions agree exactly)
BRT_HD void instance_world_box(const float4 o2w[3], f3 mlo, f3 mhi, f3& wlo, f3& whi) {
  const f3 c = (mlo + mhi) * 0.5f, e = (mhi - mlo) * 0.5f;
  const f3 wc = xform_point(o2w, c);
  const f3 we = F3((fabsf(o2w[0].x) * e.x + fabsf(o2w[0].y) * e.y) + fabsf(o2w[0].z) * e.z,
                   (fabsf(o2w[1].x) * e.x + fabsf(o2w[1].y) * e.y) + fabsf(o2w[1].z) * e.z,
                   (fabsf(o2w[2].x) * e.x + fabsf(o2w[2].y) * e.y) + fabsf(o2w[2].z) * e.z);
  wlo = wc - we;
  whi = wc + we;
}

// ---- SH/constants.slang ---------------------------------------------------------------------------
#define BRT_INFINITE 1e32f                // :3-5
#define BRT_PI 3.1415926535897f           // :11-13
#define BRT_TWO_PI 6.2831853071795f       // :15-17
#define BRT_ONE_OVER_PI 0.3183098861837f  // :19-21
#define BRT_LIGHT_TRESHOLD 0.0001f        // :27-29

// ---- deterministic transcendentals ----------------------------------------------------------------
// The shaders call sin/cos (SH/sampler.slang:60-62,77-78) and log2 (SH/disney.slang:18) whose
// precision SPIR-V leaves to the driver. The spec here fixes them to explicit polynomials so that
// every implementation of it computes the same bits.
// sin/cos for x in [0, ~2*pi]: k = trunc(x * 2/pi + 0.5), two-term Cody-Waite remainder, Taylor
// polynomials of degree 9 / 8 in Horner form, all with explicit fma.
BRT_HD void det_sincos(float x, float* s, float* c) {
  const float two_over_pi = 0.636619772f;
  const float pio2_hi = 1.57079625f;
  const float pio2_lo = 7.54978942e-08f;
  int k = (int)(x * two_over_pi + 0.5f);
  float kf = (float)k;
  float r = fma_rn(-kf, pio2_hi, x);
  r = fma_rn(-kf, pio2_lo, r);
  float r2 = r * r;
  float sp = fma_rn(r2, 2.75573192e-06f, -1.98412698e-04f);
  sp = fma_rn(sp, r2, 8.33333377e-03f);
  sp = fma_rn(sp, r2, -1.66666672e-01f);
  sp = fma_rn(sp * r2, r, r);
  float cp = fma_rn(r2, 2.48015876e-05f, -1.38888892e-03f);
  cp = fma_rn(cp, r2, 4.16666679e-02f);
  cp = fma_rn(cp, r2, -0.5f);
  cp = fma_rn(cp, r2, 1.0f);
  switch (k & 3) {
    case 0: *s = sp; *c = cp; break;
    case 1: *s = cp; *c = -sp; break;
    case 2: *s = -sp; *c = -cp; break;
    default: *s = -cp; *c = sp; break;
  }
}
// log2 for normal x > 0: x = m 2^e with m in [sqrt(1/2), sqrt(2)), ln m = 2 atanh((m-1)/(m+1)),
// odd series to t^9.
BRT_HD float det_log2(float x) {
  uint32_t bits = f2u(x);
  int e = (int)((bits >> 23) & 0xffu) - 127;
  float m = u2f((bits & 0x007fffffu) | 0x3f800000u);
  if (m > 1.41421354f) {
    m = m * 0.5f;
    e = e + 1;
  }
  float t = (m - 1.0f) / (m + 1.0f);
  float t2 = t * t;
  float p = fma_rn(t2, 0.111111112f, 0.142857149f);
  p = fma_rn(p, t2, 0.2f);
  p = fma_rn(p, t2, 0.333333343f);
  p = fma_rn(p, t2, 1.0f);
  float ln_m = (2.0f * t) * p;
  return (float)e + ln_m * 1.44269502f;
}
// exp2 for |y| <= 120: y = k + f with k = floor(y + 1/2), f in [-1/2, 1/2); e^(f ln 2) by its series to z^7, scaled by 2^k
// through the exponent field. (exp2 / pow are implementation-defined in SPIR-V; the spec here fixes them)
BRT_HD float det_exp2(float y) {
  const float kf = floorf(y + 0.5f);
  const float z = (y - kf) * 0.693147182f;
  float p = fma_rn(z, 1.98412701e-4f, 1.38888892e-3f);
  p = fma_rn(p, z, 8.33333377e-3f);
  p = fma_rn(p, z, 4.16666679e-2f);
  p = fma_rn(p, z, 0.166666672f);
  p = fma_rn(p, z, 0.5f);
  p = fma_rn(p, z, 1.0f);
  p = fma_rn(p, z, 1.0f);
  return u2f(f2u(p) + ((uint32_t)(int)kf << 23));
}
// Storage-image format conversion of the present path (Pipeline::rebuildRenderOutput(format, extent), RT/RTPipeline.cpp:49-55;
// copyImageToSwapchain, RT/RTApp.cpp:87-152). UNORM8: clamp to [0, 1] (NaN -> 0), scale by 255, round to nearest even.
// SRGB: the sRGB transfer function first (alpha stays linear).
BRT_HD uint32_t float_to_unorm8(float f) {
  const float c = f > 0.0f ? (f < 1.0f ? f : 1.0f) : 0.0f;  // NaN fails f > 0
  return (uint32_t)rintf(c * 255.0f);
}
BRT_HD float linear_to_srgb(float c) {
  if (!(c > 0.0031308f)) return c > 0.0f ? c * 12.92f : 0.0f;
  if (c >= 1.0f) return 1.0f;
  return 1.055f * det_exp2(det_log2(c) * 0.416666657f) - 0.055f;
}
// pow(x, 5) := (x*x)*(x*x)*x   (SH/disney.slang:11)
BRT_HD float pow5(float x) {
  float x2 = x * x;
  return (x2 * x2) * x;
}

}  // namespace brt
