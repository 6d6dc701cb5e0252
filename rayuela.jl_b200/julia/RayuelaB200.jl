# RayuelaB200.jl -- drop-in Julia shim over librayuela_b200.so for Rayuela.jl's two hot paths.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The Python mirror
# (rayuela_b200/julia_api.py) passes the library exactly the buffers these ccalls pass and is what the parity
# tests drive.  Usage:
#     using Rayuela
#     include("RayuelaB200.jl")
#     RayuelaB200.install!(Rayuela)            # Rayuela's OWN functions now run on the GPU
#     RayuelaB200.init_devices!(0:7)           # optional: one process drives 8 GPUs
# install! defines, inside module Rayuela, methods of encoding_icm / encode_icm_cuda / veccost / qerror /
# quantize_pq / quantize_opq / linscan_* / quantize_norms / quantize_chainq / update_codebooks_fast_bin with the
# reference's own signatures that forward to this module, so train_lsq, train_lsq_cuda, train_sr_cuda,
# experiment_lsq_cuda, experiment_sr_cuda (src/LSQ_GPU.jl:267-368, src/SR.jl:88-306) and
# demos/demos_train_query_base.jl run UNMODIFIED on top of librayuela_b200.so.  The functions below can also be
# called directly (same positional signatures and return values; citations: file:line in the Rayuela.jl tree).
module RayuelaB200

export encoding_icm, encode_icm_cuda, veccost, qerror, quantize_pq, quantize_opq,
       linscan_pq, linscan_opq, linscan_lsq, linscan_cq, seed_b200!,
       quantize_norms, quantize_chainq, fast_bin_matmul, update_codebooks_fast_bin,
       install!, init_devices!, shutdown_devices!, fast_unaries!, encode_icm_timings

using Printf, Statistics, LinearAlgebra

const librayuela_b200 = get(ENV, "RAYUELA_B200_LIB",
                            joinpath(@__DIR__, "..", "lib", "librayuela_b200.so"))

# flags passed to rayuela_encode_icm: 0 = exact (bit-identical to the CPU path), 2 = RAYUELA_FAST_UNARIES (opt-in tcgen05
# bf16x3 unaries; codes may differ on near-ties).  For the scans set ENV["RAYUELA_B200_FAST_LUT"] = "1" instead (their
# exact-signature symbols have no flags argument).
const ENCODE_FLAGS = Ref{Cuint}(0)
fast_unaries!(on::Bool) = (ENCODE_FLAGS[] = on ? Cuint(2) : Cuint(0); nothing)

const SEED = Ref{UInt64}(0)
"Reproducible stream for the ILS perturbations / visiting orders (each encode call consumes one seed)."
seed_b200!(s::Integer) = (SEED[] = UInt64(s); nothing)
function next_seed()
  s = SEED[]; SEED[] = s + 0x9E3779B97F4A7C15; s
end

function check(rc::Cint)
  rc == 0 && return
  msg = unsafe_string(ccall((:rayuela_last_error, librayuela_b200), Cstring, ()))
  error("librayuela_b200 error $rc: $msg")
end

"""
    init_devices!(devices)

One Julia process, several GPUs (the reference hard-codes device 0, src/LSQ_GPU.jl:41,45, and splits the base in
time with `nsplits`): after this call every encode splits the base over `devices` and every linscan shards it, with
bit-identical results.  The environment variable RAYUELA_B200_DEVICES="0,1,..." does the same without code.
"""
function init_devices!(devices)
  devs = convert(Vector{Cint}, collect(devices))
  check(ccall((:rayuela_init, librayuela_b200), Cint, (Ptr{Cint}, Cint), devs, length(devs)))
end
shutdown_devices!() = check(ccall((:rayuela_shutdown, librayuela_b200), Cint, ()))

codes0(B::Matrix{<:Integer}) = convert(Matrix{UInt8}, B .- one(eltype(B)))   # src/LSQ.jl:228
codes1(B::Matrix{UInt8})     = convert(Matrix{Int16}, B) .+ one(Int16)        # src/LSQ.jl:232

# --- path (1) ---------------------------------------------------------------------------------------------
"encoding_icm(X, oldB, C, ilsiter, icmiter, randord, npert, cpp=true, V=true) -> B   (src/LSQ.jl:272-294)"
function encoding_icm(X::Matrix{Float32}, oldB::Matrix{Int16}, C::Vector{Matrix{Float32}},
                      ilsiter::Integer, icmiter::Integer, randord::Bool, npert::Integer,
                      cpp::Bool=true, V::Bool=true)
  d, n = size(X); m = length(C); _, h = size(C[1])
  h <= 256 || error("The B200 implementation of ICM encoding stores codes in one byte: h must be <= 256")
  B     = codes0(oldB)
  stats = zeros(Cint, 2, max(ilsiter, 1))
  check(ccall((:rayuela_encode_icm, librayuela_b200), Cint,
    (Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cuchar}, Int64, Cint, Cint, Cint, Cint, Cint, Cint, Cint, UInt64, Int64,
     Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cint}, Cuint, Ptr{Cvoid}),
    X, hcat(C...), B, n, d, m, h, ilsiter, icmiter, npert, randord, next_seed(), 0,
    C_NULL, C_NULL, 0, C_NULL, C_NULL, C_NULL, stats, ENCODE_FLAGS[], C_NULL))
  if V
    for i = 1:ilsiter                                                         # src/LSQ.jl:243-245
      @printf(" ILS iteration %d/%d done. %5.2f%% new codes are equal. %5.2f%% new codes are better.\n",
              i, ilsiter, 100*stats[1,i]/n, 100*stats[2,i]/n)
    end
  end
  newB = codes1(B)
  copyto!(oldB, newB)                                                         # src/LSQ.jl:248
  return newB
end

"encode_icm_cuda(RX, B, C, ilsiters, icmiter, npert, randord, nsplits=2, V=false) -> Bs, objs   (src/LSQ_GPU.jl:218-264)"
function encode_icm_cuda(RX::Matrix{Float32}, B::Matrix{Int16}, C::Vector{Matrix{Float32}},
                         ilsiters::Vector{Int64}, icmiter::Integer, npert::Integer, randord::Bool,
                         nsplits::Integer=2, V::Bool=false)
  d, n = size(RX); m = length(C); _, h = size(C[1]); nr = length(ilsiters)
  B0    = codes0(B)
  snaps = zeros(UInt8, m, n, nr)
  objs  = zeros(Cfloat, nr)
  check(ccall((:rayuela_encode_icm, librayuela_b200), Cint,
    (Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cuchar}, Int64, Cint, Cint, Cint, Cint, Cint, Cint, Cint, UInt64, Int64,
     Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cint}, Cuint, Ptr{Cvoid}),
    RX, hcat(C...), B0, n, d, m, h, maximum(ilsiters), icmiter, npert, randord, next_seed(), 0,
    C_NULL, convert(Vector{Cint}, ilsiters), nr, snaps, objs, C_NULL, C_NULL, ENCODE_FLAGS[], C_NULL))
  Bs = [codes1(snaps[:, :, i]) for i = 1:nr]
  return Bs, objs
end

"Phase times (ms) of the last verbose encode: (setup, unaries, icm, total) -- the reference prints time_ns() deltas, src/LSQ_GPU.jl:50-58,213"
function encode_icm_timings()
  t = zeros(Cfloat, 4)
  check(ccall((:rayuela_encode_icm_timings, librayuela_b200), Cint, (Ptr{Cfloat},), t))
  return (setup=t[1], unaries=t[2], icm=t[3], total=t[4])
end

"veccost(X, B, C)   (src/qerrors.jl:36-66)"
function veccost(X::Matrix{Float32}, B::Matrix{<:Integer}, C::Vector{Matrix{Float32}})
  d, n = size(X); m = length(C)
  cost = zeros(Cfloat, n)
  check(ccall((:rayuela_veccost, librayuela_b200), Cint,
    (Ptr{Cfloat}, Ptr{Cuchar}, Ptr{Cfloat}, Int64, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cdouble}, Cuint, Ptr{Cvoid}),
    X, codes0(B), hcat(C...), n, d, m, 256, cost, C_NULL, 0, C_NULL))
  return cost
end
qerror(X, B, C) = mean(veccost(X, B, C))                                     # src/qerrors.jl:69-74

# --- PQ / OPQ encode -----------------------------------------------------------------------------------------
"quantize_pq(X, C, V=false) -> B   (src/PQ.jl:18-48)"
function quantize_pq(X::Matrix{Float32}, C::Vector{Matrix{Float32}}, V::Bool=false)
  d, n = size(X); m = length(C); h = size(C[1], 2)
  B = zeros(UInt8, m, n)
  check(ccall((:rayuela_quantize_pq, librayuela_b200), Cint,
    (Ptr{Cfloat}, Ptr{Cfloat}, Int64, Cint, Cint, Cint, Ptr{Cuchar}, Cuint, Ptr{Cvoid}),
    X, cat(C..., dims=3), n, d, m, h, B, 0, C_NULL))
  return codes1(B)
end
quantize_opq(X, R, C, V::Bool=false) = quantize_pq(R' * X, C, V)             # src/OPQ.jl:19-27

# --- path (2): the three reference symbols, exact signatures (only the library path changes) ---------------
"linscan_pq(B, X, C, b, k)   (src/Linscan.jl:5-37)"
function linscan_pq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, b::Int, k::Int=10000)
  m, n = size(B); d, nq = size(X)
  dists = zeros(Cfloat, k, nq); res = zeros(Cuint, k, nq)
  ccall((:linscan_aqd_query, librayuela_b200), Nothing,
    (Ptr{Cfloat}, Ptr{Cuint}, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cuint, Cint, Cint, Cint, Cint, Cint),
    dists, res, B, cat(C..., dims=3), X, Cint(n), Cuint(nq), Cint(b), Cint(k), Cint(m), Cint(d), Cint(d/m))
  return dists, (res .+= 1)
end
linscan_pq(B::Matrix{<:Integer}, X, C, b::Int, k::Int=10000) = linscan_pq(codes0(B), X, C, b, k)
linscan_opq(B, X, C, b::Int, R::Matrix{Cfloat}, k::Int=10000) = linscan_pq(B, R' * X, C, b, k)  # src/Linscan.jl:93-115

"linscan_lsq(B, X, C, dbnorms, R, k)   (src/Linscan.jl:118-157)"
function linscan_lsq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, dbnorms::Vector{Cfloat},
                     R::Matrix{Cfloat}, k::Int=10000)
  RX = R' * X
  m, n = size(B); d, nq = size(RX); _, h = size(C[1])
  dists = zeros(Cfloat, k, nq); res = zeros(Cuint, k, nq)
  ccall((:linscan_aqd_query_extra_byte, librayuela_b200), Nothing,
    (Ptr{Cfloat}, Ptr{Cint}, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cuint, Cint, Cint, Cint, Cint, Cint),
    dists, res, B, RX, hcat(C...), dbnorms, Cint(nq), Cint(n), Cint(m), Cint(h), Cint(d), Cint(k))
  return dists, res
end
linscan_lsq(B::Matrix{<:Integer}, X, C, dbnorms, R, k::Int=10000) = linscan_lsq(codes0(B), X, C, dbnorms, R, k)

"linscan_cq(B, X, C, k)   (src/Linscan.jl:160-193)"
function linscan_cq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, k::Int=10000)
  m, n = size(B); d, nq = size(X); _, h = size(C[1])
  dists = zeros(Cfloat, k, nq); res = zeros(Cuint, k, nq)
  ccall((:linscan_aqd_cq_query_extra_byte, librayuela_b200), Nothing,
    (Ptr{Cfloat}, Ptr{Cint}, Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Cuint, Cint, Cint, Cint, Cint, Cint),
    dists, res, B, X, hcat(C...), Cint(nq), Cint(n), Cint(m), Cint(h), Cint(d), Cint(k))
  return dists, res
end
linscan_cq(B::Matrix{<:Integer}, X, C, k::Int=10000) = linscan_cq(codes0(B), X, C, k)

# --- "next" rows (SURVEY 8f): norm quantization, ChainQ Viterbi encode, codebook update ------------------------
"quantize_norms(B, C, cbnorms) -> dbnormsB, dbnormsX   (src/utils.jl:29-59)"
function quantize_norms(B::Matrix{T1}, C::Vector{Matrix{Float32}}, cbnorms::Vector{Float32}) where T1<:Integer
  m, n = size(B); d, h = size(C[1])
  codes = zeros(UInt8, n); norms = zeros(Cfloat, n)
  check(ccall((:rayuela_quantize_norms, librayuela_b200), Cint,
    (Ptr{Cuchar}, Ptr{Cfloat}, Ptr{Cfloat}, Int64, Cint, Cint, Cint, Ptr{Cuchar}, Ptr{Cfloat}, Cuint, Ptr{Cvoid}),
    codes0(B), hcat(C...), cbnorms, n, d, m, h, codes, norms, 0, C_NULL))
  return convert(Vector{T1}, codes) .+ one(T1), norms                         # findmin index is 1-based, :55
end

"quantize_chainq(X, C, use_cuda=false, use_cpp=false) -> B, ellapsed   (src/ChainQ.jl:287-348)"
function quantize_chainq(X::Matrix{Float32}, C::Vector{Matrix{Float32}}, use_cuda::Bool=false, use_cpp::Bool=false)
  d, n = size(X); m = length(C); _, h = size(C[1])
  B = zeros(UInt8, m, n)
  st = time()
  check(ccall((:rayuela_quantize_chainq, librayuela_b200), Cint,
    (Ptr{Cfloat}, Ptr{Cfloat}, Int64, Cint, Cint, Cint, Ptr{Cuchar}, Cuint, Ptr{Cvoid}),
    X, hcat(C...), n, d, m, h, B, 0, C_NULL))
  return codes1(B), time() - st
end

"fast_bin_matmul(X, B, h, V=false, rho=1e-4) -> A, b   (src/codebook_update.jl:96-171)"
function fast_bin_matmul(X::Matrix{Float32}, B::Matrix{Int16}, h::Integer, V::Bool=false, rho::Float64=1e-4)
  d, n = size(X); m, _ = size(B)
  A = zeros(Cdouble, m*h, m*h); b = zeros(Cdouble, m*h, d)
  check(ccall((:rayuela_fast_bin_matmul, librayuela_b200), Cint,
    (Ptr{Cfloat}, Ptr{Cuchar}, Int64, Cint, Cint, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Cuint, Ptr{Cvoid}),
    X, codes0(B), n, d, m, h, rho, A, b, 0, C_NULL))
  return A, b
end

"update_codebooks_fast_bin(X, B, h, V=false, rho=1e-4) -> C   (src/codebook_update.jl:175-204; the solve stays LAPACK)"
function update_codebooks_fast_bin(X::Matrix{Float32}, B::Matrix{Int16}, h::Integer, V::Bool=false, rho::Float64=1e-4)
  m, n = size(B)
  A, b = fast_bin_matmul(X, B, h, V, rho)
  lpt = LAPACK.getrf!(A)
  Cm  = convert(Matrix{Float32}, LAPACK.getrs!('N', lpt[1], lpt[2], b))       # (m*h)-by-d
  Ct  = collect(Cm')                                                          # d-by-(m*h)
  return [Ct[:, (i-1)*h+1:i*h] for i = 1:m]                                   # K2vec, src/utils.jl
end

# --- make Rayuela's own functions call the GPU ------------------------------------------------------------------
"""
    install!(mod)          # mod = Rayuela

Defines in `mod` methods with the reference's signatures (same or more specific: Float32 data) that forward to this
module.  Julia dispatches `Rayuela.train_lsq_cuda`'s internal call of `encode_icm_cuda(...)` etc. to them, so the
trainers and experiment_* drivers need no source change.  `nsplits` arguments are accepted and ignored.
"""
function install!(mod::Module)
  G = @__MODULE__
  Core.eval(mod, quote
    # path (1)   src/LSQ.jl:272-281, src/LSQ_GPU.jl:218-227, src/qerrors.jl:36-39,69-72
    encoding_icm(X::Matrix{Float32}, oldB::Matrix{Int16}, C::Vector{Matrix{Float32}}, ilsiter::Integer,
                 icmiter::Integer, randord::Bool, npert::Integer, cpp::Bool=true, V::Bool=true) =
      $G.encoding_icm(X, oldB, C, ilsiter, icmiter, randord, npert, cpp, V)
    encode_icm_cuda(RX::Matrix{Float32}, B::Matrix{Int16}, C::Vector{Matrix{Float32}}, ilsiters::Vector{Int64},
                    icmiter::Integer, npert::Integer, randord::Bool, nsplits::Integer=2, V::Bool=false) =
      $G.encode_icm_cuda(RX, B, C, ilsiters, icmiter, npert, randord, nsplits, V)
    veccost(X::Matrix{Float32}, B::Matrix{T2}, C::Vector{Matrix{Float32}}) where {T2 <: Integer} = $G.veccost(X, B, C)
    qerror(X::Matrix{Float32}, B::Matrix{T2}, C::Vector{Matrix{Float32}}) where {T2 <: Integer} = $G.qerror(X, B, C)
    # PQ / OPQ encode   src/PQ.jl:18-21, src/OPQ.jl:19-23
    quantize_pq(X::Matrix{Float32}, C::Vector{Matrix{Float32}}, V::Bool=false) = $G.quantize_pq(X, C, V)
    quantize_opq(X::Matrix{Float32}, R::Matrix{Float32}, C::Vector{Matrix{Float32}}, V::Bool=false) =
      $G.quantize_opq(X, R, C, V)
    # path (2)   src/Linscan.jl:5-10,93-99,118-124,160-164 (the Integer-code methods :28-37 etc. call these)
    linscan_pq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, b::Int, k::Int=10000) =
      $G.linscan_pq(B, X, C, b, k)
    linscan_opq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, b::Int, R::Matrix{Cfloat},
                k::Int=10000) = $G.linscan_opq(B, X, C, b, R, k)
    linscan_lsq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, dbnorms::Vector{Cfloat},
                R::Matrix{Cfloat}, k::Int=10000) = $G.linscan_lsq(B, X, C, dbnorms, R, k)
    linscan_cq(B::Matrix{UInt8}, X::Matrix{Cfloat}, C::Vector{Matrix{Cfloat}}, k::Int=10000) =
      $G.linscan_cq(B, X, C, k)
    # "next" rows   src/utils.jl:29-32, src/ChainQ.jl:305-309, src/codebook_update.jl:175-180
    quantize_norms(B::Matrix{T1}, C::Vector{Matrix{Float32}}, cbnorms::Vector{Float32}) where {T1 <: Integer} =
      $G.quantize_norms(B, C, cbnorms)
    quantize_chainq(X::Matrix{Float32}, C::Vector{Matrix{Float32}}, use_cuda::Bool=false, use_cpp::Bool=false) =
      $G.quantize_chainq(X, C, use_cuda, use_cpp)
    update_codebooks_fast_bin(X::Matrix{Float32}, B::Matrix{Int16}, h::Integer, V::Bool=false, rho::Float64=1e-4) =
      $G.update_codebooks_fast_bin(X, B, h, V, rho)
  end)
  return nothing
end

end # module
