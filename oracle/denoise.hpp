// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED by the reference (it has no denoiser code and no tests).
//
// CPU restatement of the denoiser slot: the reference only DECLARES Extensions::Denoiser::denoise()
// (Graphics/Denoiser/Denoiser.h:14-20); the comment at :5-12 lists its stages — temporal accumulation (with reprojection), history
// clamping (to prevent ghosting), variance estimation, a-trous wavelet denoiser, bilateral pass. DESIGN.md §12 specifies each
// stage; this file states that specification as plain scalar loops over full images, one function per stage.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/brt.h"
#include "shaders.hpp"

namespace orc {

struct Px4 { float x, y, z, w; };

static inline float dn_luminance(float r, float g, float b) { return (0.2126f * r + 0.7152f * g) + 0.0722f * b; }
static inline float dn_exp_neg(float x) { return det_exp2(-std::fmin(x, 60.0f) * 1.44269502f); }
static inline int dn_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct DenoiseState {                 // what one call leaves for the next
  uint32_t w = 0, h = 0;
  bool have = false;
  std::vector<Px4> color;             // accumulated colour, w = history length
  std::vector<float> m1, m2;          // luminance moments
  std::vector<Px4> nrm;               // previous G-buffer
  std::vector<uint32_t> inst;
  std::vector<float> t;
  float vp[16] = {0}, eye[3] = {0, 0, 0};
};

// stage 1-3: temporal accumulation with reprojection, history clamping, variance estimation
static void dn_temporal(const Px4* color, const Px4* pos, const Px4* nrm, const uint32_t* inst, const DenoiseState& hs, uint32_t W, uint32_t H,
                        float clamp_gamma, float max_history, float pixel_offset, std::vector<Px4>& out, DenoiseState& next) {
  const size_t npx = (size_t)W * H;
  out.resize(npx);
  next.color.resize(npx);
  next.m1.resize(npx);
  next.m2.resize(npx);
  for (uint32_t y = 0; y < H; ++y)
    for (uint32_t x = 0; x < W; ++x) {
      const size_t i = (size_t)y * W + x;
      const Px4 c = color[i];
      const float L = dn_luminance(c.x, c.y, c.z);
      float m[3] = {0, 0, 0}, s[3] = {0, 0, 0}, lm = 0.0f, ls = 0.0f;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const Px4 q = color[(size_t)dn_clamp((int)y + dy, 0, (int)H - 1) * W + dn_clamp((int)x + dx, 0, (int)W - 1)];
          m[0] = m[0] + q.x; m[1] = m[1] + q.y; m[2] = m[2] + q.z;
          s[0] = s[0] + q.x * q.x; s[1] = s[1] + q.y * q.y; s[2] = s[2] + q.z * q.z;
          const float ql = dn_luminance(q.x, q.y, q.z);
          lm = lm + ql;
          ls = ls + ql * ql;
        }
      const float inv9 = 0.111111112f;
      for (int k = 0; k < 3; ++k) { m[k] = m[k] * inv9; s[k] = s[k] * inv9; }
      lm = lm * inv9;
      ls = ls * inv9;
      const float spatial_var = std::fmax(ls - lm * lm, 0.0f);
      if (inst[i] == BRT_AOV_MISS) {
        out[i] = Px4{c.x, c.y, c.z, 0.0f};
        next.color[i] = Px4{c.x, c.y, c.z, 0.0f};
        next.m1[i] = L;
        next.m2[i] = L * L;
        continue;
      }
      const Px4 P = pos[i], N = nrm[i];
      bool valid = false;
      size_t q = 0;
      if (hs.have) {
        const float* M = hs.vp;
        const float cx = ((M[0] * P.x + M[1] * P.y) + M[2] * P.z) + M[3];
        const float cy = ((M[4] * P.x + M[5] * P.y) + M[6] * P.z) + M[7];
        const float cw = ((M[12] * P.x + M[13] * P.y) + M[14] * P.z) + M[15];
        if (cw > 0.0f) {
          const float sx = ((cx / cw) * 0.5f + 0.5f) * (float)W, sy = ((cy / cw) * 0.5f + 0.5f) * (float)H;
          const float fx = std::floor(sx + pixel_offset), fy = std::floor(sy + pixel_offset);  // ray-gen samples pixel corners, or [id, id+1) when jittered
          if (fx >= 0.0f && fy >= 0.0f && fx < (float)W && fy < (float)H) {
            q = (size_t)(int)fy * W + (size_t)(int)fx;
            const Px4 Nq = hs.nrm[q];
            const float ex = P.x - hs.eye[0], ey = P.y - hs.eye[1], ez = P.z - hs.eye[2];
            const float dprev = std::sqrt((ex * ex + ey * ey) + ez * ez);
            valid = hs.inst[q] == inst[i] && ((N.x * Nq.x + N.y * Nq.y) + N.z * Nq.z) >= 0.9f && std::fabs(dprev - hs.t[q]) <= 0.05f * dprev;
          }
        }
      }
      float oc[3] = {c.x, c.y, c.z}, m1 = L, m2 = L * L, n = 1.0f;
      if (valid) {
        const Px4 h = hs.color[q];
        float hc[3] = {h.x, h.y, h.z};
        if (clamp_gamma > 0.0f)
          for (int k = 0; k < 3; ++k) {
            const float sd = std::sqrt(std::fmax(s[k] - m[k] * m[k], 0.0f)) * clamp_gamma;
            hc[k] = std::fmin(std::fmax(hc[k], m[k] - sd), m[k] + sd);
          }
        n = std::fmin(h.w + 1.0f, max_history);
        const float a = 1.0f / n;
        for (int k = 0; k < 3; ++k) oc[k] = hc[k] + (oc[k] - hc[k]) * a;
        m1 = hs.m1[q] + (L - hs.m1[q]) * a;
        m2 = hs.m2[q] + (L * L - hs.m2[q]) * a;
      }
      const float var = n >= 4.0f ? std::fmax(m2 - m1 * m1, 0.0f) : spatial_var;
      out[i] = Px4{oc[0], oc[1], oc[2], var};
      next.color[i] = Px4{oc[0], oc[1], oc[2], n};
      next.m1[i] = m1;
      next.m2[i] = m2;
    }
}

// stage 4 / 5: one edge-avoiding a-trous iteration (radius 2, holes of `step`), or the final 3x3 bilateral pass (radius 1) (§12.2-3)
static void dn_atrous(const std::vector<Px4>& in, const Px4* nrm, const uint32_t* inst, uint32_t W, uint32_t H, int step, int radius,
                      float sigma_z, float sigma_l, uint32_t sigma_n_log2, bool final_pass, std::vector<Px4>& out) {
  out.resize(in.size());
  const float kern[3] = {0.375f, 0.25f, 0.0625f};
  for (int y = 0; y < (int)H; ++y)
    for (int x = 0; x < (int)W; ++x) {
      const size_t i = (size_t)y * W + x;
      const Px4 c = in[i];
      if (inst[i] == BRT_AOV_MISS || radius == 0) {
        out[i] = final_pass ? Px4{c.x, c.y, c.z, 1.0f} : c;
        continue;
      }
      const Px4 N = nrm[i];  // w = hit distance
      const float T = N.w;
      const float Lp = dn_luminance(c.x, c.y, c.z);
      const float rcp_l = 1.0f / (sigma_l * std::sqrt(std::fmax(c.w, 0.0f)) + 1e-6f);
      const float rcp_z = radius == 2 ? 1.0f / ((sigma_z * (float)step) * T + 1e-6f) : 0.0f;
      const float w0 = kern[0] * kern[0];
      float sw = w0, sc[3] = {c.x * w0, c.y * w0, c.z * w0}, sv = c.w * (w0 * w0);
      for (int dy = -radius; dy <= radius; ++dy)
        for (int dx = -radius; dx <= radius; ++dx) {
          if (dx == 0 && dy == 0) continue;
          const int qx = x + dx * step, qy = y + dy * step;
          if (qx < 0 || qy < 0 || qx >= (int)W || qy >= (int)H) continue;
          const size_t q = (size_t)qy * W + qx;
          const Px4 Nq = nrm[q];
          float wn = std::fmax(((N.x * Nq.x + N.y * Nq.y) + N.z * Nq.z), 0.0f);
          if (!(wn > 0.0f)) continue;  // includes misses: their normal is zero
          for (uint32_t k = 0; k < sigma_n_log2; ++k) wn = wn * wn;
          const Px4 cq = in[q];
          const float e = std::fabs(Nq.w - T) * rcp_z + std::fabs(dn_luminance(cq.x, cq.y, cq.z) - Lp) * rcp_l;
          const float w = ((kern[dx < 0 ? -dx : dx] * kern[dy < 0 ? -dy : dy]) * wn) * dn_exp_neg(e);
          sc[0] = sc[0] + cq.x * w; sc[1] = sc[1] + cq.y * w; sc[2] = sc[2] + cq.z * w;
          sv = sv + cq.w * (w * w);
          sw = sw + w;
        }
      out[i] = Px4{sc[0] / sw, sc[1] / sw, sc[2] / sw, final_pass ? 1.0f : sv / (sw * sw)};
    }
}

}  // namespace orc
