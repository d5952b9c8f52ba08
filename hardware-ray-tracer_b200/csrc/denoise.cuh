// Denoiser slot of the reference (Graphics/Denoiser/Denoiser.h:5-12 is a declaration-only stub whose comment lists the stages:
// temporal accumulation with reprojection, history clamping, variance estimation, a-trous wavelet filter, bilateral pass; DLSS Ray
// Reconstruction, the README's actual denoiser, is proprietary and out of scope). The stages are specified in DESIGN.md §12 and
// restated by the oracle (oracle/denoise.hpp); both use the arithmetic contract of DESIGN.md §3, so the results are bit-identical.
//
// All kernels are one thread per pixel over row-major full-frame images (consecutive lanes = consecutive x: every tap of the
// stencils is a coalesced 16-byte-per-lane row segment). They are HBM/L2 streaming kernels: algorithmic bytes per pixel are
// stated next to each params struct.
#pragma once
#include "device_types.cuh"
#include "vecmath.cuh"

namespace brt {

BRT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
BRT_HD float luminance(float r, float g, float b) { return (0.2126f * r + 0.7152f * g) + 0.0722f * b; }
// e^(-x) for x >= 0 through the deterministic exp2
BRT_HD float exp_neg(float x) { return det_exp2(-fminf(x, 60.0f) * 1.44269502f); }

struct GBuffer {          // primary-hit attributes of sample 0 (written by k_shade with BRT_RENDER_GBUFFER)
  const float4* pos;      // world position xyz, w = 1 on a hit, 0 on a miss
  const float4* nrm;      // shading normal xyz (after the flip towards the viewer), w = hit distance; all 0 on a miss
  const uint32_t* inst;   // instance id, BRT_MISS on a miss
  const float* t;         // hit distance
};

// ---- temporal accumulation with reprojection + history clamping + variance estimation --------------------------------------
// per pixel: 16 colour + 9 neighbours (cached) + 40 G-buffer read, on a valid reprojection 16 + 8 + 16 + 4 + 4 history read;
// 16 + 16 + 8 written  ->  ~130 B
struct DnTemporalParams {
  uint32_t count;  // width * height
  const uint32_t* count_ptr;
  uint32_t width, height;
  const float4* color;     // this frame, linear RGBA32F
  GBuffer g;               // this frame
  GBuffer hg;              // previous frame
  const float4* h_color;   // previous accumulated colour, a = history length
  const float2* h_moments; // previous luminance moments
  float prev_vp[16];       // previous frame's projection * view, row-major: world -> clip
  float prev_eye[3];
  float pixel_offset;      // 0.5: frames sample pixel corners (the reference's ray-gen); 0: jittered frames sample [id, id + 1)
  uint32_t have_history;
  float clamp_gamma;       // history clamped to mean +- gamma * sigma of the 3x3 neighbourhood; <= 0: off
  float max_history;
  float4* out;             // accumulated colour, w = variance of the luminance   (input of the wavelet filter)
  float4* out_h_color;     // next frame's history
  float2* out_h_moments;
};
BRT_HD void dn_temporal_body(const DnTemporalParams& p, uint32_t i) {
  const int W = (int)p.width, H = (int)p.height;
  const int x = (int)(i % p.width), y = (int)(i / p.width);
  const float4 c4 = p.color[i];
  const float L = luminance(c4.x, c4.y, c4.z);
  // 3x3 statistics of this frame (clamped addressing)
  float m[3] = {0.0f, 0.0f, 0.0f}, s[3] = {0.0f, 0.0f, 0.0f}, lm = 0.0f, ls = 0.0f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int qx = clampi(x + dx, 0, W - 1), qy = clampi(y + dy, 0, H - 1);
      const float4 q = p.color[(size_t)qy * p.width + qx];
      m[0] = m[0] + q.x; m[1] = m[1] + q.y; m[2] = m[2] + q.z;
      s[0] = s[0] + q.x * q.x; s[1] = s[1] + q.y * q.y; s[2] = s[2] + q.z * q.z;
      const float ql = luminance(q.x, q.y, q.z);
      lm = lm + ql;
      ls = ls + ql * ql;
    }
  const float inv9 = 0.111111112f;
  for (int k = 0; k < 3; ++k) { m[k] = m[k] * inv9; s[k] = s[k] * inv9; }
  lm = lm * inv9;
  ls = ls * inv9;
  const float spatial_var = fmaxf(ls - lm * lm, 0.0f);
  const uint32_t id = p.g.inst[i];
  if (id == BRT_MISS) {  // background: nothing to accumulate
    p.out[i] = make_float4(c4.x, c4.y, c4.z, 0.0f);
    p.out_h_color[i] = make_float4(c4.x, c4.y, c4.z, 0.0f);
    p.out_h_moments[i] = make_float2(L, L * L);
    return;
  }
  const float4 P = p.g.pos[i], N = p.g.nrm[i];
  bool valid = false;
  size_t q = 0;
  if (p.have_history) {
    const float* M = p.prev_vp;
    const float cx = ((M[0] * P.x + M[1] * P.y) + M[2] * P.z) + M[3];
    const float cy = ((M[4] * P.x + M[5] * P.y) + M[6] * P.z) + M[7];
    const float cw = ((M[12] * P.x + M[13] * P.y) + M[14] * P.z) + M[15];
    if (cw > 0.0f) {
      const float sx = ((cx / cw) * 0.5f + 0.5f) * (float)W, sy = ((cy / cw) * 0.5f + 0.5f) * (float)H;  // raygen: clip = id / size * 2 - 1
      const float fx = floorf(sx + p.pixel_offset), fy = floorf(sy + p.pixel_offset);
      if (fx >= 0.0f && fy >= 0.0f && fx < (float)W && fy < (float)H) {
        q = (size_t)(int)fy * p.width + (size_t)(int)fx;
        const float4 Nq = p.hg.nrm[q];
        const float ex = P.x - p.prev_eye[0], ey = P.y - p.prev_eye[1], ez = P.z - p.prev_eye[2];
        const float dprev = sqrtf((ex * ex + ey * ey) + ez * ez);
        valid = p.hg.inst[q] == id && ((N.x * Nq.x + N.y * Nq.y) + N.z * Nq.z) >= 0.9f && fabsf(dprev - p.hg.t[q]) <= 0.05f * dprev;
      }
    }
  }
  float oc[3] = {c4.x, c4.y, c4.z}, m1 = L, m2 = L * L, n = 1.0f;
  if (valid) {
    const float4 h = p.h_color[q];
    const float2 hm = p.h_moments[q];
    float hc[3] = {h.x, h.y, h.z};
    if (p.clamp_gamma > 0.0f)
      for (int k = 0; k < 3; ++k) {
        const float sd = sqrtf(fmaxf(s[k] - m[k] * m[k], 0.0f)) * p.clamp_gamma;
        hc[k] = fminf(fmaxf(hc[k], m[k] - sd), m[k] + sd);
      }
    n = fminf(h.w + 1.0f, p.max_history);
    const float a = 1.0f / n;
    for (int k = 0; k < 3; ++k) oc[k] = hc[k] + (oc[k] - hc[k]) * a;
    m1 = hm.x + (L - hm.x) * a;
    m2 = hm.y + (L * L - hm.y) * a;
  }
  const float var = n >= 4.0f ? fmaxf(m2 - m1 * m1, 0.0f) : spatial_var;
  p.out[i] = make_float4(oc[0], oc[1], oc[2], var);
  p.out_h_color[i] = make_float4(oc[0], oc[1], oc[2], n);
  p.out_h_moments[i] = make_float2(m1, m2);
}

// ---- edge-avoiding a-trous wavelet iteration (5x5 B3 spline, holes of `step` pixels) ------------------------------------------
// per pixel: 25 taps x (16 colour/variance + 16 normal/depth) read (neighbouring lanes share them through L1/L2: 36 B of DRAM
// traffic per pixel with the 4-byte id of the centre), 16 written
struct DnAtrousParams {
  uint32_t count;
  const uint32_t* count_ptr;
  uint32_t width, height;
  int step;
  int radius;           // 2: the 5x5 wavelet; 1: the final 3x3 bilateral pass (depth term off); 0: copy
  float sigma_z, sigma_l;
  uint32_t sigma_n_log2;  // normal weight = max(0, N.N')^(2^sigma_n_log2)
  uint32_t final_pass;    // write alpha = 1 instead of the filtered variance
  const float4* in;     // rgb + variance
  GBuffer g;
  float4* out;
};
BRT_HD void dn_atrous_body(const DnAtrousParams& p, uint32_t i) {
  const int W = (int)p.width, H = (int)p.height;
  const int x = (int)(i % p.width), y = (int)(i / p.width);
  const float4 c = p.in[i];
  const uint32_t id = p.g.inst[i];
  if (id == BRT_MISS || p.radius == 0) {  // background, or the pass-through that only sets alpha (no filtering asked for)
    p.out[i] = p.final_pass ? make_float4(c.x, c.y, c.z, 1.0f) : c;
    return;
  }
  const float4 N = p.g.nrm[i];  // w = hit distance
  const float T = N.w;
  const float Lp = luminance(c.x, c.y, c.z);
  // both exponential edge-stopping terms share one exponential: exp(-dz / (sigma_z step T)) exp(-dl / (sigma_l sqrt(var)))
  const float rcp_l = 1.0f / (p.sigma_l * sqrtf(fmaxf(c.w, 0.0f)) + 1e-6f);
  const float rcp_z = p.radius == 2 ? 1.0f / ((p.sigma_z * (float)p.step) * T + 1e-6f) : 0.0f;  // the bilateral pass has no depth term
  const float kern[3] = {0.375f, 0.25f, 0.0625f};
  const float w0 = kern[0] * kern[0];
  float sw = w0, sc[3] = {c.x * w0, c.y * w0, c.z * w0}, sv = c.w * (w0 * w0);
  for (int dy = -p.radius; dy <= p.radius; ++dy)
    for (int dx = -p.radius; dx <= p.radius; ++dx) {
      if (dx == 0 && dy == 0) continue;
      const int qx = x + dx * p.step, qy = y + dy * p.step;
      if (qx < 0 || qy < 0 || qx >= W || qy >= H) continue;
      const size_t q = (size_t)qy * p.width + qx;
      const float4 Nq = p.g.nrm[q];  // a miss has a zero normal: weight 0
      float wn = fmaxf(((N.x * Nq.x + N.y * Nq.y) + N.z * Nq.z), 0.0f);
      if (!(wn > 0.0f)) continue;
      for (uint32_t k = 0; k < p.sigma_n_log2; ++k) wn = wn * wn;
      const float4 cq = p.in[q];
      const float e = fabsf(Nq.w - T) * rcp_z + fabsf(luminance(cq.x, cq.y, cq.z) - Lp) * rcp_l;
      const float w = ((kern[dx < 0 ? -dx : dx] * kern[dy < 0 ? -dy : dy]) * wn) * exp_neg(e);
      sc[0] = sc[0] + cq.x * w; sc[1] = sc[1] + cq.y * w; sc[2] = sc[2] + cq.z * w;
      sv = sv + cq.w * (w * w);
      sw = sw + w;
    }
  p.out[i] = make_float4(sc[0] / sw, sc[1] / sw, sc[2] / sw, p.final_pass ? 1.0f : sv / (sw * sw));
}

}  // namespace brt
