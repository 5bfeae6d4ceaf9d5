"""CPU check of the two-term Float32 form of candidate 4 in ShiftedRootNormLhalfBox's Float32 kernel
(csrc/spx_ops.cuh, ProxLhalfBox::apply_f32): the stationary point's magnitude |val| = s² is carried as an unevaluated
sum of two Float32 numbers after ONE Newton step on s³ - |z| s + σλ/2 = 0 from the Float32 closed-form start.  The
kernel's claims, restated here in numpy Float32 arithmetic (FMA emulated through Float64, exact for Float32 operands):

  * p - |z| is exact (Sterbenz) for t <= 0.9, so the residual of the cubic loses nothing to cancellation;
  * mh + ml is within 1e-10 |val| of the Float64 value of the reference's closed form
    (2/3)|z| (1 + cos(2π/3 - (2/3) acos t))  (shiftedRootNormLhalf.jl:48,57),
    i.e. far below half a Float32 ulp: rounding `val - xs` to Float32 afterwards gives the reference's stored value
    except when the exact result sits within ~1e-3 ulp of a rounding boundary.
"""
import numpy as np

f32 = np.float32


def fma32(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def start(zf, c4f):
    """lhalf_start of spx_ops.cuh with exact Float32 sqrt/rsqrt in place of the SFU approximations (rel. 1e-7)."""
    w = f32(zf * f32(0.33333334))
    r = f32(1.0) / np.sqrt(w, dtype=f32)
    t32 = f32(f32(c4f * r) * f32(r * r))
    sf = np.sqrt(max(f32(1.0) - t32, f32(0.0)), dtype=f32)
    pf = f32(0.0011198767460882664)
    for c in (-0.005838877987116575, 0.01784452795982361, -0.05533028766512871, 0.40823012590408325, 0.5000002384185791):
        pf = fma32(pf, sf, f32(c))
    s0 = f32(f32(f32(2.0) * f32(w * r)) * pf)
    slope = fma32(f32(f32(3.0) * s0), s0, -zf)
    return t32, s0, f32(1.0) / slope


def two_term(zf, a2f):
    t32, s0, inv = start(zf, f32(a2f * f32(0.5)))
    p = f32(s0 * s0)
    e = fma32(s0, s0, -p)
    uh = f32(p - zf)
    exact_sub = np.float64(p) - np.float64(zf) == np.float64(uh)
    t1 = f32(s0 * uh)
    t1e = fma32(s0, uh, -t1)
    f = f32(f32(t1 + a2f) + fma32(s0, e, t1e))
    d = f32(-f * inv)
    m1 = fma32(d, d, fma32(f32(s0 + s0), d, e))
    mh = f32(p + m1)
    ml = f32(f32(p - mh) + m1)
    return t32, np.float64(mh) + np.float64(ml), exact_sub


def closed_form(z, a2):
    t = (a2 / 2.0) * (z / 3.0) ** -1.5
    return (2.0 / 3.0) * z * (1.0 + np.cos(2.0 * np.pi / 3.0 - (2.0 / 3.0) * np.arccos(t)))


def test_two_term_float32_magnitude_matches_the_float64_closed_form():
    rng = np.random.default_rng(5)
    worst = 0.0
    checked = 0
    for sigma_lambda in (0.1, 0.013, 2.5):
        a2f = f32(f32(sigma_lambda) * f32(0.5))
        # |z| from just above the t = 0.9 boundary up to large values (t -> 0)
        zmin = 3.0 * (float(a2f) / 2.0 / 0.9) ** (2.0 / 3.0)
        for z in np.exp(rng.uniform(np.log(zmin * 1.001), np.log(zmin * 1e4), 4000)):
            zf = f32(z)
            t32, val, exact_sub = two_term(zf, a2f)
            if t32 > f32(0.9):
                continue
            assert exact_sub, (zf, a2f)
            ref = closed_form(np.float64(zf), np.float64(a2f))
            worst = max(worst, abs(val - ref) / ref)
            checked += 1
    assert checked > 10_000
    assert worst < 1e-10, worst
