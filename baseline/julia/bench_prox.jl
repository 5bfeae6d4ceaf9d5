# baseline/julia/bench_prox.jl -- the REAL reference (ShiftedProximalOperators.jl) on the host cores, for the
# configurations of BASELINE.json.  Not runnable in the build image (no `julia`, no network): run it wherever the
# package resolves, e.g.
#     julia --project=/path/to/ShiftedProximalOperators.jl baseline/julia/bench_prox.jl [log2n]
# It prints one JSON line per configuration (elements/s, single-threaded like the package itself), with the same
# synthetic inputs as bench.py (splitmix64 hash uniforms, seed 20261018) so the numbers replace the oracle-port
# column of BASELINE.md §4 one for one.
using ShiftedProximalOperators, ProximalOperators, Printf

const SEED = UInt64(20261018)
function splitmix64(x::UInt64)
  x += 0x9E3779B97F4A7C15
  x = (x ⊻ (x >> 30)) * 0xBF58476D1CE4E5B9
  x = (x ⊻ (x >> 27)) * 0x94D049BB133111EB
  x ⊻ (x >> 31)
end
uniform(n, stream; scale = 1.0, shift = 0.0, i0 = 0) =
  [scale * (Float64(splitmix64((SEED ⊻ (UInt64(stream) << 40)) + UInt64(i0 + i - 1)) >> 11) * 2.0^-53) + shift for i = 1:n]

function timeit(f; reps = 5)
  f()
  ts = [(@elapsed f()) for _ = 1:reps]
  sort(ts)[(reps + 1) ÷ 2]
end
report(name, n, t) = @printf("{\"config\": \"%s\", \"n\": %d, \"seconds\": %.6f, \"elements_per_s\": %.4e, \"threads\": 1}\n", name, n, t, n / t)

function main()
  log2n = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 24
  n = 1 << log2n
  λ, σ = 1.0, 0.1
  xk, sj, q = uniform(n, 0, scale = 4.0, shift = -2.0), uniform(n, 1, shift = -0.5), uniform(n, 2, scale = 4.0, shift = -2.0)
  l, u = -(0.25 .+ uniform(n, 3)), 0.25 .+ uniform(n, 4)
  b = uniform(n, 6)
  d = 0.5 .+ uniform(n, 5)
  d = ifelse.(b .< 0.1, -d, ifelse.(b .< 0.2, 0.0, d))
  y = similar(q)
  # C1
  ψ = shifted(shifted(NormL1(λ), xk), sj)
  report("C1 ShiftedNormL1 prox!", n, timeit(() -> prox!(y, ψ, q, σ)))
  # C2
  ψ0 = shifted(shifted(NormL0(λ), xk, l, u), sj)
  ψh = shifted(shifted(RootNormLhalf(λ), xk, l, u), sj)
  report("C2 ShiftedNormL0Box prox!", n, timeit(() -> prox!(y, ψ0, q, σ)))
  report("C2 ShiftedRootNormLhalfBox prox!", n, timeit(() -> prox!(y, ψh, q, σ)))
  report("C2 ShiftedNormL0Box iprox!", n, timeit(() -> iprox!(y, ψ0, q, d)))
  # C3
  χ = NormLinf(1.0)
  ψb = shifted(shifted(NormL1(λ), xk, 0.75, χ), sj)
  report("C3 ShiftedNormL1 BInf prox!", n, timeit(() -> prox!(y, ψb, q, σ)))
  ψ2 = shifted(shifted(NormL1(λ), xk, 1.0e9, NormL2(1.0)), sj)
  prox!(y, ψ2, q, σ)
  Δ = 0.5 * sqrt(sum(abs2, sj .+ y))
  ψ2 = shifted(shifted(NormL1(λ), xk, Δ, NormL2(1.0)), sj)
  report("C3 ShiftedNormL1B2 prox!", n, timeit(() -> prox!(y, ψ2, q, σ); reps = 3))
  # C4
  ng = n ÷ 64
  idx = [((g - 1) * 64 + 1):(g * 64) for g = 1:ng]
  lam = 0.5 .+ uniform(ng, 12)
  hg = GroupNormL2(lam, idx)
  ψg = shifted(shifted(hg, xk), sj)
  report("C4 ShiftedGroupNormL2 prox! (groups of 64)", n, timeit(() -> prox!(y, ψg, q, 0.3); reps = 3))
  ψgb = shifted(shifted(hg, xk, 0.5, χ), sj)
  report("C4 ShiftedGroupNormL2Binf prox! (groups of 64)", n, timeit(() -> prox!(y, ψgb, q, 0.3); reps = 3))
  # C5 (one problem of 65536 at a time, as a batch would run it)
  pn = 65536
  np = max(1, n ÷ pn)
  t = 0.0
  for p = 1:min(np, 16)
    r = ((p - 1) * pn + 1):(p * pn)
    ψt = shifted(shifted(IndBallL0(1024), xk[r], 1.0, χ), sj[r])
    yy = similar(q[r])
    qq = q[r]
    t += timeit(() -> prox!(yy, ψt, qq, 1.0); reps = 3)
  end
  report("C5 ShiftedIndBallL0BInf prox! (problems of 65536, r = 1024)", min(np, 16) * pn, t)
end

main()
