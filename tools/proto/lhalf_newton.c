// Prototype (host) of the RootNormLhalf stationary-point Newton used by the CUDA kernels:
//   val = s^2, s the largest root of s^3 - az s + a = 0  (a = σλ/2),
// started from a Float32 evaluation of the closed form.  Built by tools/proto/check_lhalf_newton.py
// to measure the error against mpmath before the device version is trusted.
#include <math.h>
#include <stdint.h>

static float perturb(float x, uint32_t* st, float rel) {
  *st = *st * 1664525u + 1013904223u;
  float u = ((*st >> 8) * (1.0f / 16777216.0f)) * 2.0f - 1.0f;
  return x * (1.0f + rel * u);
}

// variant 0: two plain fused steps; 1: second step compensated; bit 2 (4): third step when t32 > 0.8
double lhalf_root_fast(double az, double c4, int variant, uint32_t seed, float* t32_out) {
  uint32_t st = seed;
  const float zf = (float)az;
  const float w = zf * 0.33333334f;
  const float r = perturb((float)(1.0 / sqrt((double)w)), &st, 2.4e-7f);  // rsqrt.approx.ftz.f32
  const float t32 = ((float)c4 * r) * (r * r);
  float om = 1.0f - t32;
  om = om > 0.0f ? om : 0.0f;
  const float sf = perturb(sqrtf(om), &st, 1.2e-7f);  // sqrt.approx.ftz.f32
  float pf = 0.0011198767460882664f;
  pf = fmaf(pf, sf, -0.005838877987116575f);
  pf = fmaf(pf, sf, 0.01784452795982361f);
  pf = fmaf(pf, sf, -0.05533028766512871f);
  pf = fmaf(pf, sf, 0.40823012590408325f);
  pf = fmaf(pf, sf, 0.5000002384185791f);
  const float s0 = (2.0f * (w * r)) * pf;
  const float slope = fmaf(3.0f * s0, s0, -zf);
  const float invf = perturb(1.0f / slope, &st, 1.2e-7f);  // rcp.approx.ftz.f32
  if (t32_out) *t32_out = t32;
  const double a = 2.0 * c4;
  const double inv = (double)invf;
  double s = (double)s0;
  // step 1
  double u = fma(s, s, -az);
  double f = fma(s, u, a);
  s = fma(-f, inv, s);
  if ((variant & 4) && t32 > 0.8f) {
    u = fma(s, s, -az);
    f = fma(s, u, a);
    s = fma(-f, inv, s);
  }
  if ((variant & 3) == 0) {
    u = fma(s, s, -az);
    f = fma(s, u, a);
    s = fma(-f, inv, s);
  } else {
    const double p = s * s;
    const double e = fma(s, s, -p);
    u = p - az;
    const double ue = (p - (u + az));  // Fast2Sum tail, az >= p
    f = fma(s, u, a);
    f = fma(s, e + ue, f);
    s = fma(-f, inv, s);
  }
  return s * s;
}

void lhalf_root_fast_v(const double* az, double c4, int variant, int n, double* out, float* t32) {
  for (int i = 0; i < n; ++i) out[i] = lhalf_root_fast(az[i], c4, variant, 12345u + (uint32_t)i, t32 + i);
}
