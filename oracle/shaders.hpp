// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.
//
// PINNED TO THE REFERENCE: the reference (CodingBloon/Hardware-Ray-Tracer) has no tests, golden vectors or
// CPU path, but its shipped shader binary can be executed: this scalar C++ restatement of the Slang
// shaders reproduces the frames that oracle/ref/spv_interp.py computes from shaders/raytracing.slang.spv
// (tests/golden/spv_kat.json, tests/test_spv_kat.py: BRDF, processLight, calculateColor, geometry fetch, ray generation). hash / pcg /
// rand are dead code in that binary: pinned by the integer KATs of SURVEY.md Appendix C (tests/test_oracle_kat.py); the samplers by
// analytic cases.
//
// Shorthand: SH/ = /root/reference/Hardware Ray Tracer/shaders/
//
// Arithmetic contract: binary32, one IEEE operation per written operator, no FMA
// contraction (built with -ffp-contract=off), left-to-right association as written, sqrt and
// division correctly rounded. The CUDA kernels are compiled with -fmad=false and follow the same
// association so both sides agree bit for bit wherever no transcendental is involved; sin, cos and
// log2 are replaced by the explicit polynomials below (identical on both sides by construction).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct vec3 {
  float x, y, z;
};
struct vec2 {
  float x, y;
};
static inline vec3 V3(float x, float y, float z) { return vec3{x, y, z}; }
static inline vec3 V3(float s) { return vec3{s, s, s}; }
static inline vec3 operator+(vec3 a, vec3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline vec3 operator-(vec3 a, vec3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline vec3 operator*(vec3 a, vec3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline vec3 operator*(vec3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline vec3 operator*(float s, vec3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline vec3 operator+(vec3 a, float s) { return V3(a.x + s, a.y + s, a.z + s); }
static inline vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline vec3 cross(vec3 a, vec3 b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
// normalize := v * (1 / sqrt(dot(v, v)))   (GLSL.std.450 Normalize leaves the precision open)
static inline vec3 normalize(vec3 a) {
  float inv = 1.0f / std::sqrt(dot(a, a));
  return a * inv;
}
static inline float square(float f) { return f * f; }  // SH/shadermath.slang:3
static inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
// lerp := x*(1-s) + y*s  (GLSL.std.450 FMix, what the shipped .spv uses)
static inline float lerpf(float x, float y, float s) { return x * (1.0f - s) + y * s; }
static inline vec3 lerp3(vec3 x, vec3 y, float s) { return V3(lerpf(x.x, y.x, s), lerpf(x.y, y.y, s), lerpf(x.z, y.z, s)); }
static inline vec3 reflect(vec3 i, vec3 n) { return i - n * (2.0f * dot(n, i)); }

// ---- SH/constants.slang ------------------------------------------------------------------------
static const float kINFINITE = 1e32f;         // :3-5
static const int kMISS_DEPTH = 1000;          // :7-9
static const float kPI = 3.1415926535897f;    // :11-13
static const float kTWO_PI = 6.2831853071795f;  // :15-17
static const float kONE_OVER_PI = 0.3183098861837f;  // :19-21
static const float kLIGHT_TRESHOLD = 0.0001f;  // :27-29

// ---- deterministic transcendental replacements (spec) -----------------------------
// sin/cos for x in [0, ~2*pi]: quadrant reduction k = trunc(x*2/pi + 0.5), two-term Cody-Waite
// remainder with explicit fma, degree-9 / degree-8 Taylor polynomials in Horner form with fma.
static inline void det_sincos(float x, float* s, float* c) {
  const float two_over_pi = 0.636619772f;
  const float pio2_hi = 1.57079625f;       // 0x3FC90FDA
  const float pio2_lo = 7.54978942e-08f;   // pi/2 - pio2_hi
  int k = (int)(x * two_over_pi + 0.5f);
  float kf = (float)k;
  float r = std::fma(-kf, pio2_hi, x);
  r = std::fma(-kf, pio2_lo, r);
  float r2 = r * r;
  float sp = std::fma(r2, 2.75573192e-06f, -1.98412698e-04f);
  sp = std::fma(sp, r2, 8.33333377e-03f);
  sp = std::fma(sp, r2, -1.66666672e-01f);
  sp = std::fma(sp * r2, r, r);
  float cp = std::fma(r2, 2.48015876e-05f, -1.38888892e-03f);
  cp = std::fma(cp, r2, 4.16666679e-02f);
  cp = std::fma(cp, r2, -0.5f);
  cp = std::fma(cp, r2, 1.0f);
  switch (k & 3) {
    case 0: *s = sp; *c = cp; break;
    case 1: *s = cp; *c = -sp; break;
    case 2: *s = -sp; *c = -cp; break;
    default: *s = -cp; *c = sp; break;
  }
}
// log2 for finite x > 0 (normal floats): x = m * 2^e with m in [sqrt(1/2), sqrt(2)),
// ln(m) = 2*atanh(t), t = (m-1)/(m+1), odd series to t^9.
static inline float det_log2(float x) {
  uint32_t bits;
  std::memcpy(&bits, &x, 4);
  int e = (int)((bits >> 23) & 0xffu) - 127;
  uint32_t mb = (bits & 0x007fffffu) | 0x3f800000u;
  float m;
  std::memcpy(&m, &mb, 4);
  if (m > 1.41421354f) {
    m = m * 0.5f;
    e = e + 1;
  }
  float t = (m - 1.0f) / (m + 1.0f);
  float t2 = t * t;
  float p = std::fma(t2, 0.111111112f, 0.142857149f);
  p = std::fma(p, t2, 0.2f);
  p = std::fma(p, t2, 0.333333343f);
  p = std::fma(p, t2, 1.0f);
  float ln_m = (2.0f * t) * p;
  return (float)e + ln_m * 1.44269502f;
}
// exp2 for |y| <= 120 (implementation-defined in SPIR-V, fixed by DESIGN.md §3): y = k + f, k = floor(y + 1/2);
// e^(f ln 2) by the exponential series to the 7th power, times 2^k through the exponent bits.
static inline float det_exp2(float y) {
  float kf = std::floor(y + 0.5f);
  float z = (y - kf) * 0.693147182f;
  float p = std::fma(z, 1.98412701e-4f, 1.38888892e-3f);
  p = std::fma(p, z, 8.33333377e-3f);
  p = std::fma(p, z, 4.16666679e-2f);
  p = std::fma(p, z, 0.166666672f);
  p = std::fma(p, z, 0.5f);
  p = std::fma(p, z, 1.0f);
  p = std::fma(p, z, 1.0f);
  uint32_t bits;
  std::memcpy(&bits, &p, 4);
  bits += (uint32_t)((int)kf) << 23;
  std::memcpy(&p, &bits, 4);
  return p;
}
// Present path: what storing a float4 into a storage image of the swapchain's format does (Pipeline::rebuildRenderOutput(format,
// extent), RT/RTPipeline.cpp:49-55; the image is then copied texel by texel, RT/RTApp.cpp:87-152). Vulkan's float -> UNORM8
// conversion: clamp to [0, 1] (NaN -> 0), x 255, round to nearest even; SRGB formats encode R, G, B with the sRGB curve first.
static inline uint32_t float_to_unorm8(float f) {
  float c = f > 0.0f ? (f < 1.0f ? f : 1.0f) : 0.0f;
  return (uint32_t)std::nearbyint(c * 255.0f);
}
static inline float linear_to_srgb(float c) {
  if (!(c > 0.0031308f)) return c > 0.0f ? c * 12.92f : 0.0f;
  if (c >= 1.0f) return 1.0f;
  return 1.055f * det_exp2(det_log2(c) * 0.416666657f) - 0.055f;
}
// pow(x, 5) := (x*x)*(x*x)*x
static inline float pow5(float x) {
  float x2 = x * x;
  return (x2 * x2) * x;
}

// ---- SH/random.slang -----------------------------------------------------------------------------
static inline uint32_t hash3(uint32_t px, uint32_t py, uint32_t pz) {  // :2-12
  const uint32_t p0 = 2246822519u, p1 = 3266489917u, p2 = 668265263u, p3 = 374761393u;
  uint32_t h32 = pz + p3 + px * p1;
  h32 = p2 * ((h32 << 17) | (h32 >> (32 - 17)));
  h32 += py * p1;
  h32 = p2 * ((h32 << 17) | (h32 >> (32 - 17)));
  h32 = p0 * (h32 ^ (h32 >> 15));
  h32 = p1 * (h32 ^ (h32 >> 13));
  return h32 ^ (h32 >> 16);
}
static inline uint32_t pcg(uint32_t& state) {  // :14-19
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
static inline float rnd(uint32_t& seed) {  // :21-24  (can return exactly 1.0)
  uint32_t r = pcg(seed);
  return (float)r * (1.0f / (float)0xffffffffu);
}

// ---- SH/material.slang:3-15 ---------------------------------------------------------------------
struct Material {
  vec3 color;
  float subsurface, metallic, roughness, specular, specularTint, anisotropic, sheen, sheenTint, clearCoat, clearCoatGloss;
};

// ---- SH/shadermath.slang --------------------------------------------------------------------------
static inline void orthonormalBasis(vec3 n, vec3& tangent, vec3& bitangent) {  // :5-16
  if (n.z < -0.99998796f) {
    tangent = V3(0.0f, -1.0f, 0.0f);
    bitangent = V3(-1.0f, 0.0f, 0.0f);
    return;
  }
  float a = 1.0f / (1.0f + n.z);
  float b = -n.x * n.y * a;
  tangent = V3(1.0f - n.x * n.x * a, b, -n.x);
  bitangent = V3(b, 1.0f - n.y * n.y * a, -n.y);
}
static inline vec3 toLocal(vec3 v, vec3 n) {  // :18-23
  vec3 t, b;
  orthonormalBasis(n, t, b);
  return V3(dot(v, t), dot(v, b), dot(v, n));
}
static inline vec3 toWorld(vec3 v, vec3 n) {  // :25-30
  vec3 t, b;
  orthonormalBasis(n, t, b);
  return (v.x * t + v.y * b) + v.z * n;
}

// ---- SH/disney.slang ------------------------------------------------------------------------------
static inline float schlickFresnel(float F0, float VdotH) { return F0 + (1.0f - F0) * pow5(1.0f - VdotH); }  // :11
static inline float schlickWeight(float f) {  // :12
  float m = clampf(1.0f - f, 0.0f, 1.0f);
  return m * m * m * m * m;
}
static inline float GTR1(float NdotH, float a) {  // :15-19 (log2, not ln — kept)
  if (a >= 1.0f) return kONE_OVER_PI;
  float a2 = a * a;
  return (a2 - 1.0f) / (kPI * det_log2(a2) * (1.0f + (a2 - 1.0f) * NdotH * NdotH));
}
static inline float GTR2_anisotropic(float NdotH, float HdotX, float HdotY, vec2 a) {  // :26-28
  return 1.0f / (kPI * a.x * a.y * square(square(HdotX / a.x) + square(HdotY / a.y) + NdotH * NdotH));
}
static inline float GGX(float NdotV, float a) {  // :30-33
  float a2 = a * a;
  return 2.0f / (1.0f + std::sqrt(a2 + (1.0f - a2) * NdotV * NdotV));
}
static inline float GGX_anisotropic(float NdotV, float VdotX, float VdotY, vec2 a) {  // :35-37
  return 1.0f / (NdotV + std::sqrt(square(VdotX * a.x) + square(VdotY * a.y) * NdotV * NdotV));
}
static inline vec3 calculateTint(vec3 color) {  // :39-42
  float l = dot(V3(0.3f, 0.6f, 1.0f), color);
  return l > 0.0f ? color * (1.0f / l) : V3(1.0f);
}
static inline vec3 evalSheen(const Material& m, float HdotL) {  // :44-47 (ignores material.sheen — kept)
  vec3 tint = calculateTint(m.color);
  return lerp3(V3(1.0f), tint, m.sheenTint) * schlickWeight(HdotL);
}
static inline float evalClearcoat(const Material& m, float NdotH, float NdotL, float NdotV, float LdotH) {  // :49-55
  float d = GTR1(NdotH, lerpf(0.1f, 0.001f, m.clearCoatGloss));
  float f = schlickFresnel(0.04f, LdotH);
  float g = GGX(NdotL, 0.25f) * GGX(NdotV, 0.25f);
  return 0.25f * m.clearCoat * d * f * g;
}
static inline float evalDiffuse(const Material& m, vec3 L, vec3 V, vec3 H) {  // :57-69 (local-frame vectors)
  float FL = schlickWeight(L.z);
  float FV = schlickWeight(V.z);
  float FD90 = 0.5f + 2.0f * m.roughness * square(dot(H, L));
  float FD = lerpf(1.0f, FD90, FL) * lerpf(1.0f, FD90, FV);
  float Fss90 = square(dot(L, H)) * m.roughness;
  float Fss = lerpf(1.0f, Fss90, FL) * lerpf(1.0f, Fss90, FV);
  float ss = 1.25f * (Fss * (1.0f / (L.z + V.z) - 0.5f) + 0.5f);
  return lerpf(FD, ss, m.subsurface);
}
static inline vec2 calculateAnisotropicParameters(float anisotropic, float roughness) {  // :71-77
  float aspect = std::sqrt(1.0f - anisotropic * 0.9f);
  float r2 = roughness * roughness;
  return vec2{std::fmax(0.001f, r2 / aspect), std::fmax(0.001f, r2 * aspect)};
}
static inline vec3 evalSpecular(const Material& m, float NdotH, vec3 H, vec3 V, vec3 L) {  // :79-92
  vec2 ap = calculateAnisotropicParameters(m.anisotropic, m.roughness);
  vec3 tint = calculateTint(m.color);
  vec3 color = lerp3(m.specular * 0.08f * lerp3(V3(1.0f), tint, m.specularTint), m.color, m.metallic);
  float d = GTR2_anisotropic(NdotH, H.x, H.y, ap);
  float fresnel = schlickWeight(dot(L, H));
  vec3 f = lerp3(color, V3(1.0f), fresnel);
  float g = GGX_anisotropic(L.z, L.x, L.y, ap) * GGX_anisotropic(V.z, V.x, V.y, ap);
  return d * f * g;
}
static inline vec3 BRDF(const Material& m, vec3 N, vec3 V, vec3 L) {  // :95-116 (no N.L cosine — kept)
  float NdotL = dot(N, L);
  float NdotV = dot(N, V);
  if (NdotL <= 0.0f || NdotV <= 0.0f) return V3(0.0f);
  vec3 H = normalize(V + L);
  float NdotH = dot(N, H);
  float HdotL = dot(H, L);
  vec3 localH = toLocal(H, N);
  vec3 localV = toLocal(V, N);
  vec3 localL = toLocal(L, N);
  vec3 sheen = evalSheen(m, HdotL);
  float clearCoat = evalClearcoat(m, NdotH, NdotL, NdotV, HdotL);
  vec3 specular = evalSpecular(m, NdotH, localH, localV, localL);
  float diffuse = evalDiffuse(m, localL, localV, localH);
  return (kONE_OVER_PI * diffuse * m.color + sheen) * (1.0f - m.metallic) + specular + clearCoat;
}

// ---- SH/sampler.slang -----------------------------------------------------------------------------
static inline float GGXVNDFPDF(const Material& m, vec3 wo, vec3 wi) {  // :23-33
  float a2 = square(m.roughness);
  float NdotL = wi.z;
  float NdotV = wo.z;
  float f1 = std::sqrt(a2 + (1.0f - a2) * NdotL * NdotL);
  float f2 = std::sqrt(a2 + (1.0f - a2) * NdotV * NdotV);
  float G1 = 2.0f * NdotV / std::sqrt(a2 + (1.0f - a2) * NdotV * NdotV) + NdotV;
  float G2 = 2.0f * NdotL * NdotV / (f1 + f2);
  return G2 / G1;
}
static inline vec2 anisotropicFromMaterial(const Material& m) {  // :35-42
  float aspect = std::sqrt(1.0f - m.anisotropic * 0.9f);
  float r2 = square(m.roughness);
  return vec2{std::fmax(0.001f, r2 / aspect), std::fmax(0.001f, r2 * aspect)};
}
// :53-65 — returns the tangent-space direction; `pdf` is really 1/pdf; unused by the path loop
static inline vec3 sampleCosineWeightedHemisphere(vec2 randoms, float& pdf) {
  float phi = kTWO_PI * randoms.y;
  float cosTheta = std::sqrt(randoms.x);
  float sinTheta = std::sqrt(std::fmax(0.0f, 1.0f - cosTheta * cosTheta));
  pdf = 1.0f / (cosTheta * kONE_OVER_PI);
  float s, c;
  det_sincos(phi, &s, &c);
  return V3(sinTheta * c, sinTheta * s, cosTheta);
}
// :67-93 — V is the incoming ray direction (the call site passes WorldRayDirection(), SH/raytracing.slang:166)
static inline vec3 sampleGGXVNDFSphericalCap(const Material& m, vec3 V, vec3 N, vec2 randoms, float& pdf) {
  vec3 wo = toLocal(V, N);
  vec2 an = anisotropicFromMaterial(m);
  vec3 v = normalize(V3(an.x * -wo.x, an.y * -wo.y, -wo.z));
  float lensq = square(v.x) + square(v.y);
  vec3 t1 = lensq > 0.0f ? V3(-v.y, v.x, 0.0f) * (1.0f / std::sqrt(lensq)) : V3(1.0f, 0.0f, 0.0f);
  vec3 t2 = cross(v, t1);
  float r = std::sqrt(randoms.x);
  float phi = kTWO_PI * randoms.y;
  float sn, cs;
  det_sincos(phi, &sn, &cs);
  float p1 = r * cs;
  float p2 = r * sn;
  float s = 0.5f * (1.0f + v.z);
  p2 = (1.0f - s) * std::sqrt(1.0f - square(p1)) + s * p2;
  vec3 n = (t1 * p1 + t2 * p2) + std::sqrt(std::fmax(0.0f, 1.0f - square(p1) - square(p2))) * v;
  vec3 wm = normalize(V3(an.x * n.x, an.y * n.y, std::fmax(0.0f, n.z)));
  vec3 wi = reflect(wo, wm);
  if (wi.z < 0.0f)
    pdf = 0.0f;
  else
    pdf = GGXVNDFPDF(m, wo, wi) * 4.0f;
  return toWorld(wi, N);
}

}  // namespace orc
