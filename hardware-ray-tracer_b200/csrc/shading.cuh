// Device-side restatement of the reference's shading functions (SH/ = reference shaders/):
// random.slang, material.slang, shadermath.slang, disney.slang, sampler.slang. Every deviation of
// the reference from the textbook Disney BRDF is kept on purpose.
#pragma once
#include "vecmath.cuh"

namespace brt {

// ---- SH/random.slang --------------------------------------------------------------------------------
BRT_HD uint32_t hash3(uint32_t px, uint32_t py, uint32_t pz) {  // :2-12, xxHash32-style avalanche
  const uint32_t p0 = 2246822519u, p1 = 3266489917u, p2 = 668265263u, p3 = 374761393u;
  uint32_t h = pz + p3 + px * p1;
  h = p2 * ((h << 17) | (h >> 15));
  h += py * p1;
  h = p2 * ((h << 17) | (h >> 15));
  h = p0 * (h ^ (h >> 15));
  h = p1 * (h ^ (h >> 13));
  return h ^ (h >> 16);
}
BRT_HD uint32_t pcg(uint32_t& state) {  // :14-19, PCG-RXS-M-XS 32
  uint32_t prev = state * 747796405u + 2891336453u;
  uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
  state = prev;
  return (word >> 22u) ^ word;
}
BRT_HD float rnd(uint32_t& seed) {  // :21-24 — float(0xffffffff) rounds to 2^32, so 1.0 is reachable
  uint32_t r = pcg(seed);
  return (float)r * (1.0f / 4294967296.0f);
}

// ---- SH/material.slang:3-15 (== brt_material, 52 bytes) ------------------------------------------------
struct Material {
  f3 color;
  float subsurface, metallic, roughness, specular, specularTint, anisotropic, sheen, sheenTint, clearCoat, clearCoatGloss;
};

// ---- SH/shadermath.slang ---------------------------------------------------------------------------
BRT_HD void orthonormalBasis(f3 n, f3& tangent, f3& bitangent) {  // :5-16
  if (n.z < -0.99998796f) {
    tangent = F3(0.0f, -1.0f, 0.0f);
    bitangent = F3(-1.0f, 0.0f, 0.0f);
    return;
  }
  float a = 1.0f / (1.0f + n.z);
  float b = -n.x * n.y * a;
  tangent = F3(1.0f - n.x * n.x * a, b, -n.x);
  bitangent = F3(b, 1.0f - n.y * n.y * a, -n.y);
}
struct Frame {  // the tangent frame is built once per shading point and reused (same values as recomputing it)
  f3 t, b, n;
};
BRT_HD Frame make_frame(f3 n) {
  Frame f;
  f.n = n;
  orthonormalBasis(n, f.t, f.b);
  return f;
}
BRT_HD f3 toLocal(f3 v, const Frame& f) { return F3(dot(v, f.t), dot(v, f.b), dot(v, f.n)); }  // :18-23
BRT_HD f3 toWorld(f3 v, const Frame& f) { return (v.x * f.t + v.y * f.b) + v.z * f.n; }        // :25-30

// ---- SH/disney.slang -----------------------------------------------------------------------------
BRT_HD float schlickFresnel(float F0, float VdotH) { return F0 + (1.0f - F0) * pow5(1.0f - VdotH); }  // :11
BRT_HD float schlickWeight(float f) {                                                                 // :12
  float m = clampf(1.0f - f, 0.0f, 1.0f);
  return m * m * m * m * m;
}
BRT_HD float GTR1(float NdotH, float a) {  // :15-19 (log2 where Disney has ln: kept)
  if (a >= 1.0f) return BRT_ONE_OVER_PI;
  float a2 = a * a;
  return (a2 - 1.0f) / (BRT_PI * det_log2(a2) * (1.0f + (a2 - 1.0f) * NdotH * NdotH));
}
BRT_HD float GTR2_anisotropic(float NdotH, float HdotX, float HdotY, f2 a) {  // :26-28
  return 1.0f / (BRT_PI * a.x * a.y * square(square(HdotX / a.x) + square(HdotY / a.y) + NdotH * NdotH));
}
BRT_HD float GGX(float NdotV, float a) {  // :30-33
  float a2 = a * a;
  return 2.0f / (1.0f + sqrtf(a2 + (1.0f - a2) * NdotV * NdotV));
}
BRT_HD float GGX_anisotropic(float NdotV, float VdotX, float VdotY, f2 a) {  // :35-37 (NdotV^2 on one term only: kept)
  return 1.0f / (NdotV + sqrtf(square(VdotX * a.x) + square(VdotY * a.y) * NdotV * NdotV));
}
BRT_HD f3 calculateTint(f3 color) {  // :39-42
  float l = dot(F3(0.3f, 0.6f, 1.0f), color);
  return l > 0.0f ? color * (1.0f / l) : F3(1.0f);
}
BRT_HD f2 calculateAnisotropicParameters(float anisotropic, float roughness) {  // :71-77 == sampler.slang:35-42
  float aspect = sqrtf(1.0f - anisotropic * 0.9f);
  float r2 = roughness * roughness;
  return f2{fmaxf(0.001f, r2 / aspect), fmaxf(0.001f, r2 * aspect)};
}

// Material-only terms of the BRDF, hoisted out of the per-light loop (they do not depend on L).
struct BrdfSetup {
  f3 tint;
  f2 ap;
  f3 spec_color;
  float cc_a;
};
BRT_HD BrdfSetup brdf_setup(const Material& m) {
  BrdfSetup s;
  s.tint = calculateTint(m.color);
  s.ap = calculateAnisotropicParameters(m.anisotropic, m.roughness);
  s.spec_color = lerp3(m.specular * 0.08f * lerp3(F3(1.0f), s.tint, m.specularTint), m.color, m.metallic);  // :84
  s.cc_a = lerpf(0.1f, 0.001f, m.clearCoatGloss);                                                          // :50
  return s;
}

// BRDF (:95-116). N, V, L are world-space unit vectors; V points away from the surface.
BRT_HD f3 BRDF(const Material& m, const BrdfSetup& s, const Frame& fr, f3 V, f3 L) {
  const f3 N = fr.n;
  float NdotL = dot(N, L);
  float NdotV = dot(N, V);
  if (NdotL <= 0.0f || NdotV <= 0.0f) return F3(0.0f);  // :99-100
  f3 H = normalize(V + L);
  float NdotH = dot(N, H);
  float HdotL = dot(H, L);
  f3 lH = toLocal(H, fr), lV = toLocal(V, fr), lL = toLocal(L, fr);
  // evalSheen :44-47 (material.sheen is never read by the reference: kept)
  f3 sheen = lerp3(F3(1.0f), s.tint, m.sheenTint) * schlickWeight(HdotL);
  // evalClearcoat :49-55
  float cc_d = GTR1(NdotH, s.cc_a);
  float cc_f = schlickFresnel(0.04f, HdotL);
  float cc_g = GGX(NdotL, 0.25f) * GGX(NdotV, 0.25f);
  float clearCoat = 0.25f * m.clearCoat * cc_d * cc_f * cc_g;
  // evalSpecular :79-92
  float d = GTR2_anisotropic(NdotH, lH.x, lH.y, s.ap);
  float fresnel = schlickWeight(dot(lL, lH));
  f3 f = lerp3(s.spec_color, F3(1.0f), fresnel);
  float g = GGX_anisotropic(lL.z, lL.x, lL.y, s.ap) * GGX_anisotropic(lV.z, lV.x, lV.y, s.ap);
  f3 specular = d * f * g;
  // evalDiffuse :57-69 (local-frame vectors, .z as the cosines)
  float FL = schlickWeight(lL.z);
  float FV = schlickWeight(lV.z);
  float hl = dot(lH, lL);
  float FD90 = 0.5f + 2.0f * m.roughness * square(hl);
  float FD = lerpf(1.0f, FD90, FL) * lerpf(1.0f, FD90, FV);
  float Fss90 = square(dot(lL, lH)) * m.roughness;
  float Fss = lerpf(1.0f, Fss90, FL) * lerpf(1.0f, Fss90, FV);
  float ss = 1.25f * (Fss * (1.0f / (lL.z + lV.z) - 0.5f) + 0.5f);
  float diffuse = lerpf(FD, ss, m.subsurface);
  return (BRT_ONE_OVER_PI * diffuse * m.color + sheen) * (1.0f - m.metallic) + specular + F3(clearCoat);  // :115
}

// ---- SH/sampler.slang ------------------------------------------------------------------------------
BRT_HD float GGXVNDFPDF(const Material& m, f3 wo, f3 wi) {  // :23-33
  float a2 = square(m.roughness);
  float NdotL = wi.z;
  float NdotV = wo.z;
  float f1 = sqrtf(a2 + (1.0f - a2) * NdotL * NdotL);
  float f2_ = sqrtf(a2 + (1.0f - a2) * NdotV * NdotV);
  float G1 = 2.0f * NdotV / sqrtf(a2 + (1.0f - a2) * NdotV * NdotV) + NdotV;
  float G2 = 2.0f * NdotL * NdotV / (f1 + f2_);
  return G2 / G1;
}
// :53-65 — tangent-space direction; the reference's `pdf` output (really 1/pdf) is unused by the path loop
BRT_HD f3 sampleCosineWeightedHemisphere(float r1, float r2) {
  float phi = BRT_TWO_PI * r2;
  float cosTheta = sqrtf(r1);
  float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
  float s, c;
  det_sincos(phi, &s, &c);
  return F3(sinTheta * c, sinTheta * s, cosTheta);
}
// :67-93 — V is the incoming ray direction (the call site passes WorldRayDirection(), SH/raytracing.slang:166)
BRT_HD f3 sampleGGXVNDFSphericalCap(const Material& m, f2 an, f3 V, const Frame& fr, float r1, float r2, float& pdf) {
  f3 wo = toLocal(V, fr);
  f3 v = normalize(F3(an.x * -wo.x, an.y * -wo.y, -wo.z));
  float lensq = square(v.x) + square(v.y);
  f3 t1 = lensq > 0.0f ? F3(-v.y, v.x, 0.0f) * (1.0f / sqrtf(lensq)) : F3(1.0f, 0.0f, 0.0f);
  f3 t2 = cross(v, t1);
  float r = sqrtf(r1);
  float phi = BRT_TWO_PI * r2;
  float sn, cs;
  det_sincos(phi, &sn, &cs);
  float p1 = r * cs;
  float p2 = r * sn;
  float s = 0.5f * (1.0f + v.z);
  p2 = (1.0f - s) * sqrtf(1.0f - square(p1)) + s * p2;
  f3 n = (t1 * p1 + t2 * p2) + sqrtf(fmaxf(0.0f, 1.0f - square(p1) - square(p2))) * v;
  f3 wm = normalize(F3(an.x * n.x, an.y * n.y, fmaxf(0.0f, n.z)));
  f3 wi = reflect(wo, wm);
  if (wi.z < 0.0f)
    pdf = 0.0f;
  else
    pdf = GGXVNDFPDF(m, wo, wi) * 4.0f;
  return toWorld(wi, fr);
}

}  // namespace brt
