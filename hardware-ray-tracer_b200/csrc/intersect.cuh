// Primitive tests that replace the RT cores behind TraceRay (SH/raytracing.slang:67,121): the
// reference has no source for them, so the algorithm is the published watertight
// ray/triangle test of Woop, Benthin & Wald (JCGT 2013) restated, plus an analytic sphere
// (extension). Operation order is part of the spec and matched by the CPU oracle.
#pragma once
#include "vecmath.cuh"

namespace brt {

// Per-ray constants: kz = axis of the largest |d| (ties: lowest index), kx = kz+1, ky = kz+2 (mod 3);
// no winding swap because nothing is face-culled (VK_GEOMETRY_INSTANCE_TRIANGLE_CULL_DISABLE,
// RT/Scene.cpp:188).
struct RayShear {
  int kz;
  float Sx, Sy, Sz;
};
BRT_HD RayShear make_shear(f3 d) {
  RayShear s;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  s.kz = 0;
  float m = ax;
  if (ay > m) { s.kz = 1; m = ay; }
  if (az > m) { s.kz = 2; }
  float dx = s.kz == 0 ? d.y : (s.kz == 1 ? d.z : d.x);
  float dy = s.kz == 0 ? d.z : (s.kz == 1 ? d.x : d.y);
  float dz = s.kz == 0 ? d.x : (s.kz == 1 ? d.y : d.z);
  s.Sx = dx / dz;
  s.Sy = dy / dz;
  s.Sz = 1.0f / dz;
  return s;
}
// rotate components so that the result is (v[kx], v[ky], v[kz])
BRT_HD f3 permute(f3 v, int kz) { return kz == 0 ? F3(v.y, v.z, v.x) : (kz == 1 ? F3(v.z, v.x, v.y) : v); }

// true and (t, u, v) when tmin < t and (t < tmax, or t <= tmax when `inclusive`); u, v are the
// barycentric weights of v1 and v2 (BuiltInTriangleIntersectionAttributes, SH/raytracing.slang:137)
BRT_HD bool intersect_tri(f3 o, const RayShear& s, float tmin, float tmax, bool inclusive, f3 v0, f3 v1, f3 v2, float& t_out, float& u_out,
                          float& v_out) {
  f3 A = permute(v0 - o, s.kz), B = permute(v1 - o, s.kz), C = permute(v2 - o, s.kz);
  float Ax = A.x - s.Sx * A.z, Ay = A.y - s.Sy * A.z;
  float Bx = B.x - s.Sx * B.z, By = B.y - s.Sy * B.z;
  float Cx = C.x - s.Sx * C.z, Cy = C.y - s.Sy * C.z;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {  // edge case: redo the edge functions in double
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = (U + V) + W;
  if (det == 0.0f) return false;
  float Az = s.Sz * A.z, Bz = s.Sz * B.z, Cz = s.Sz * C.z;
  float T = (U * Az + V * Bz) + W * Cz;
  float rcp = 1.0f / det;
  float t = T * rcp;
  if (!(t > tmin && (inclusive ? t <= tmax : t < tmax))) return false;
  t_out = t;
  u_out = V * rcp;
  v_out = W * rcp;
  return true;
}

// extension: object-space sphere; the nearer root inside the interval wins, else the farther one
BRT_HD bool intersect_sphere(f3 o, f3 d, float tmin, float tmax, bool inclusive, f3 c, float r, float& t_out) {
  f3 oc = o - c;
  float a = dot(d, d);
  float b = dot(oc, d);
  float cc = dot(oc, oc) - r * r;
  float disc = b * b - a * cc;
  if (!(disc >= 0.0f)) return false;
  float sq = sqrtf(disc);
  float t0 = (-b - sq) / a;
  float t1 = (-b + sq) / a;
  if (t0 > tmin && (inclusive ? t0 <= tmax : t0 < tmax)) { t_out = t0; return true; }
  if (t1 > tmin && (inclusive ? t1 <= tmax : t1 < tmax)) { t_out = t1; return true; }
  return false;
}

}  // namespace brt
