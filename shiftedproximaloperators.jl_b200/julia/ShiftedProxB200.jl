# ShiftedProxB200.jl -- Julia glue over libshiftedprox.so (B200, sm_100a).
#
# Drop-in for the shifted prox path of ShiftedProximalOperators.jl v0.2.2: same exported
# verbs (src/ShiftedProximalOperators.jl:11-12), same struct field names, methods added to
# ProximalOperators.prox / prox! exactly as the reference does (:15).  The Julia side only
# OWNS DEVICE BUFFERS (DeviceVector) and `ccall`s the C ABI of include/shiftedprox.h; there
# is no CUDA.jl kernel, no multi-backend dispatch and no CPU fallback on this path.
#
# NOTE: `julia` is not installed in the build image, so this file has been reviewed but not
# executed; the same C ABI is exercised by the Python mirror (shiftedprox/) in tests/.
module ShiftedProxB200

using ProximalOperators
import ProximalOperators: prox, prox!

export DeviceVector, ShiftedProximableFunction
export prox, prox!, iprox, iprox!, set_radius!, shift!, shifted, set_bounds!, step!

const libshiftedprox = get(ENV, "LIBSHIFTEDPROX", joinpath(@__DIR__, "..", "libshiftedprox.so"))

# ---------------------------------------------------------------- status check ---
# style of `chklapackerror(info[])` at src/psvd.jl:137
struct SpxError <: Exception
  status::Int32
  msg::String
end
function chkspx(status::Int32)
  status == 0 && return nothing
  msg = unsafe_string(ccall((:spx_last_error, libshiftedprox), Cstring, ()))
  status == -2 && throw(AssertionError(msg))          # `@assert d[i] > 0`, shiftedNormL1.jl:70
  status == -3 && error("Error: at least one lower bound is greater than the upper bound.")
  throw(SpxError(status, msg))
end

# --------------------------------------------------------------------- context ---
mutable struct Context
  handle::Ptr{Cvoid}
  function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    chkspx(ccall((:spx_ctx_create, libshiftedprox), Int32, (Ref{Ptr{Cvoid}}, Int32, Ptr{Cvoid}, Int32),
                 h, device, C_NULL, 1))
    c = new(h[])
    finalizer(c -> ccall((:spx_ctx_destroy, libshiftedprox), Int32, (Ptr{Cvoid},), c.handle), c)
    c
  end
end
const CTX = Ref{Context}()
ctx() = (isassigned(CTX) || (CTX[] = Context(0)); CTX[].handle)

# --------------------------------------------------------------- device vector ---
# plays the role of Vector{R}: pointer + length, freed by a finalizer
mutable struct DeviceVector{R <: Union{Float32, Float64}} <: AbstractVector{R}
  ptr::Ptr{R}
  len::Int
  function DeviceVector{R}(::UndefInitializer, n::Integer) where {R}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    chkspx(ccall((:spx_malloc, libshiftedprox), Int32, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx(), n * sizeof(R), p))
    v = new{R}(Ptr{R}(p[]), n)
    finalizer(v -> ccall((:spx_free, libshiftedprox), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), ctx(), v.ptr), v)
    v
  end
end
Base.size(v::DeviceVector) = (v.len,)
Base.length(v::DeviceVector) = v.len
Base.similar(v::DeviceVector{R}, n::Integer = v.len) where {R} = DeviceVector{R}(undef, n)
Base.unsafe_convert(::Type{Ptr{R}}, v::DeviceVector{R}) where {R} = v.ptr
function DeviceVector(x::Vector{R}) where {R}
  v = DeviceVector{R}(undef, length(x))
  chkspx(ccall((:spx_memcpy_h2d, libshiftedprox), Int32, (Ptr{Cvoid}, Ptr{R}, Ptr{R}, Csize_t), ctx(), v, x, sizeof(x)))
  v
end
function Base.Array(v::DeviceVector{R}) where {R}
  x = Vector{R}(undef, v.len)
  chkspx(ccall((:spx_memcpy_d2h, libshiftedprox), Int32, (Ptr{Cvoid}, Ptr{R}, Ptr{R}, Csize_t), ctx(), x, v, sizeof(x)))
  x
end
Base.getindex(v::DeviceVector, i::Integer) = Array(v)[i]  # debugging only
function Base.copyto!(dst::DeviceVector{R}, src::DeviceVector{R}) where {R}  # `.=` of shift!
  chkspx(ccall((:spx_memcpy_d2d, libshiftedprox), Int32, (Ptr{Cvoid}, Ptr{R}, Ptr{R}, Csize_t), ctx(), dst, src, dst.len * sizeof(R)))
  dst
end
for (R, suf) in ((Float64, "f64"), (Float32, "f32"))
  @eval function Base.fill!(v::DeviceVector{$R}, a)
    chkspx(ccall(($(QuoteNode(Symbol("spx_fill_", suf))), libshiftedprox), Int32, (Ptr{Cvoid}, Ptr{$R}, Int64, $R), ctx(), v, v.len, a))
    v
  end
end
Base.zero(v::DeviceVector{R}) where {R} = fill!(similar(v), zero(R))

# C structs of the ABI
struct SpxBound
  vec::Ptr{Cvoid}
  val::Cdouble
end
bound(b::Real) = SpxBound(C_NULL, Cdouble(b))
bound(b::DeviceVector) = SpxBound(Ptr{Cvoid}(b.ptr), 0.0)
struct SpxSel
  kind::Int32
  start::Int64
  step::Int64
  stop::Int64
  mask::Ptr{UInt32}
  list::Ptr{Int64}
  nlist::Int64
end
# `selected` (1-based on the Julia side) -> 0-based device form
sel(r::AbstractUnitRange, n) = SpxSel(first(r) == 1 && last(r) == n ? 0 : 1, first(r) - 1, 1, last(r) - 1, C_NULL, C_NULL, 0)
sel(r::StepRange, n) = SpxSel(1, first(r) - 1, step(r), last(r) - 1, C_NULL, C_NULL, 0)
# general AbstractArray{<:Integer}: upload the list, build the bitmask with spx_build_mask (kept in ψ)

# ------------------------------------------------------------------ the types ---
abstract type ShiftedProximableFunction end

sfx(::Type{Float64}) = "f64"
sfx(::Type{Float32}) = "f32"

mutable struct ShiftedNormL1{R, V0, V1, V2} <: ShiftedProximableFunction  # shiftedNormL1.jl:3-26
  h::NormL1{R}
  xk::V0
  sj::V1
  sol::V2
  shifted_twice::Bool
  xsy::V2
end
mutable struct ShiftedNormL0Box{R, T, V0, V1, V2, V3, V4} <: ShiftedProximableFunction  # shiftedNormL0Box.jl:3-48
  h::NormL0{R}
  xk::V0
  sj::V1
  sol::V2
  l::V3
  u::V4
  shifted_twice::Bool
  selected::T
  xsy::V2
end
# ... ShiftedNormL0, ShiftedRootNormLhalf, ShiftedNormL1Box, ShiftedRootNormLhalfBox, ShiftedNormL1B2,
# ShiftedIndBallL0(BInf), ShiftedGroupNormL2(Binf) follow the same pattern, field for field.

shifted(h::NormL1{R}, xk::DeviceVector{R}) where {R} =
  ShiftedNormL1{R, typeof(xk), typeof(xk), typeof(xk)}(h, xk, zero(xk), similar(xk), false, similar(xk))
shifted(ψ::ShiftedNormL1{R}, sj::DeviceVector{R}) where {R} =
  ShiftedNormL1{R, typeof(ψ.xk), typeof(sj), typeof(ψ.sol)}(ψ.h, ψ.xk, sj, similar(ψ.xk), true, similar(ψ.xk))
function shifted(h::NormL0{R}, xk::DeviceVector{R}, l, u, selected::AbstractArray{T} = 1:length(xk)) where {R, T <: Integer}
  flag = Ref{Int32}(0)  # any(l .> u)  shiftedNormL0Box.jl:33
  R == Float64 ?
    chkspx(ccall((:spx_any_gt_f64, libshiftedprox), Int32, (Ptr{Cvoid}, Int64, Ref{SpxBound}, Ref{SpxBound}, Ref{Int32}), ctx(), length(xk), bound(l), bound(u), flag)) :
    chkspx(ccall((:spx_any_gt_f32, libshiftedprox), Int32, (Ptr{Cvoid}, Int64, Ref{SpxBound}, Ref{SpxBound}, Ref{Int32}), ctx(), length(xk), bound(l), bound(u), flag))
  flag[] != 0 && error("Error: at least one lower bound is greater than the upper bound.")
  ShiftedNormL0Box{R, T, typeof(xk), typeof(xk), typeof(xk), typeof(l), typeof(u)}(h, xk, zero(xk), similar(xk), l, u, false, selected, similar(xk))
end

# ---------------------------------------------------------------- generic verbs ---
function shift!(ψ::ShiftedProximableFunction, shift::DeviceVector)  # ShiftedProximalOperators.jl:72-79
  copyto!(ψ.shifted_twice ? ψ.sj : ψ.xk, shift)
  ψ
end
function set_bounds!(ψ::ShiftedNormL0Box, l, u)  # :107-111
  isa(l, Real) ? (ψ.l = l) : copyto!(ψ.l, l)
  isa(u, Real) ? (ψ.u = u) : copyto!(ψ.u, u)
  ψ
end
set_radius!(ψ::ShiftedNormL0Box, Δ) = set_bounds!(ψ, -Δ, Δ)  # :97-99
prox(ψ::ShiftedProximableFunction, q, σ) = prox!(ψ.sol, ψ, q, σ)   # :189-190
iprox(ψ::ShiftedProximableFunction, g, d) = iprox!(ψ.sol, ψ, g, d) # :180
function Base.getproperty(ψ::ShiftedProximableFunction, p::Symbol)  # :113-121
  p == :λ ? getfield(ψ, :h).lambda : p == :r ? getfield(ψ, :h).r : getfield(ψ, p)
end

# ------------------------------------------------------------- per-type methods ---
for R in (Float64, Float32)
  s = sfx(R)
  @eval begin
    # shiftedNormL1.jl:40-54
    function prox!(y::DeviceVector{$R}, ψ::ShiftedNormL1{$R}, q::DeviceVector{$R}, σ::$R)
      chkspx(ccall(($(QuoteNode(Symbol("spx_prox_l1_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Cdouble, Cdouble, Ptr{Cdouble}),
                   ctx(), length(y), y, ψ.xk, ψ.sj, q, ψ.λ, σ, C_NULL))
      y
    end
    # shiftedNormL1.jl:60-75 (AssertionError if some d[i] <= 0)
    function iprox!(y::DeviceVector{$R}, ψ::ShiftedNormL1{$R}, g::DeviceVector{$R}, d::DeviceVector{$R})
      bad = Ref{Int64}(-1)
      chkspx(ccall(($(QuoteNode(Symbol("spx_iprox_l1_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Cdouble, Ref{Int64}, Ptr{Cdouble}),
                   ctx(), length(y), y, ψ.xk, ψ.sj, g, d, ψ.λ, bad, C_NULL))
      y
    end
    # ψ(y)  ShiftedProximalOperators.jl:51-54
    function (ψ::ShiftedNormL1{$R})(y::DeviceVector{$R})
      out = Ref{Cdouble}(0)
      chkspx(ccall(($(QuoteNode(Symbol("spx_value_sep_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int32, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Cdouble, Int64, Ref{Cdouble}),
                   ctx(), 0, length(y), ψ.xk, ψ.sj, y, ψ.λ, 0, out))
      $R(out[])
    end
    # shiftedNormL0Box.jl:89-131
    function prox!(y::DeviceVector{$R}, ψ::ShiftedNormL0Box{$R}, q::DeviceVector{$R}, σ::$R)
      chkspx(ccall(($(QuoteNode(Symbol("spx_prox_l0box_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ref{SpxBound}, Ref{SpxBound}, Ref{SpxSel}, Cdouble, Cdouble, Ptr{Cdouble}),
                   ctx(), length(y), y, ψ.xk, ψ.sj, q, bound(ψ.l), bound(ψ.u), sel(ψ.selected, length(y)), ψ.λ, σ, C_NULL))
      y
    end
    # shiftedNormL0Box.jl:137-231
    function iprox!(y::DeviceVector{$R}, ψ::ShiftedNormL0Box{$R}, g::DeviceVector{$R}, d::DeviceVector{$R})
      chkspx(ccall(($(QuoteNode(Symbol("spx_iprox_l0box_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ref{SpxBound}, Ref{SpxBound}, Ref{SpxSel}, Cdouble, Ptr{Cdouble}),
                   ctx(), length(y), y, ψ.xk, ψ.sj, g, d, bound(ψ.l), bound(ψ.u), sel(ψ.selected, length(y)), ψ.λ, C_NULL))
      y
    end
    # shiftedNormL0Box.jl:70-82
    function (ψ::ShiftedNormL0Box{$R})(y::DeviceVector{$R})
      out = Ref{Cdouble}(0)
      chkspx(ccall(($(QuoteNode(Symbol("spx_value_box_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int32, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ref{SpxBound}, Ref{SpxBound}, Ref{SpxSel}, Cdouble, Ref{Cdouble}),
                   ctx(), 1, length(y), ψ.xk, ψ.sj, y, bound(ψ.l), bound(ψ.u), sel(ψ.selected, length(y)), ψ.λ, out))
      $R(out[])
    end
    # Fused solver step (extension; SURVEY.md §8f rank 1): the sweeps an R2 / TR iteration of
    # RegularizedOptimization.jl wraps around its prox! (reference README.md:17), in the pass of the prox!:
    #   s .= prox(ψ, -ν .* ∇f, ν);  xsy .= ψ.xk .+ ψ.sj .+ s;  returns (ψ(s), ‖s‖₂, ∇f's)
    function step!(s::DeviceVector{$R}, xsy::Union{DeviceVector{$R}, Nothing}, ψ::ShiftedNormL1{$R},
                   ∇f::DeviceVector{$R}, ν::$R)
      out = zeros(Cdouble, 3)
      chkspx(ccall(($(QuoteNode(Symbol("spx_step_sep_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int32, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Cdouble, Cdouble, Ptr{Cdouble}),
                   ctx(), 0, length(s), s, xsy === nothing ? C_NULL : xsy, ψ.xk,
                   ψ.shifted_twice ? ψ.sj : C_NULL,   # shifted once: sj is the constructor's zero vector
                   ∇f, ψ.λ, ν, out))
      ($R(out[1]), sqrt(out[2]), out[3])
    end
    function step!(s::DeviceVector{$R}, xsy::Union{DeviceVector{$R}, Nothing}, ψ::ShiftedNormL0Box{$R},
                   ∇f::DeviceVector{$R}, ν::$R)
      out = zeros(Cdouble, 3)
      chkspx(ccall(($(QuoteNode(Symbol("spx_step_box_", s))), libshiftedprox), Int32,
                   (Ptr{Cvoid}, Int32, Int64, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ptr{$R}, Ref{SpxBound}, Ref{SpxBound},
                    Ref{SpxSel}, Cdouble, Cdouble, Ptr{Cdouble}),
                   ctx(), 1, length(s), s, xsy === nothing ? C_NULL : xsy, ψ.xk, ψ.sj, ∇f, bound(ψ.l), bound(ψ.u),
                   sel(ψ.selected, length(s)), ψ.λ, ν, out))
      ($R(out[1]), sqrt(out[2]), out[3])
    end
  end
end

end # module
