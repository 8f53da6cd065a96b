# GATB200.jl -- the reference-side binding of libgat (include/gat.h).
#
# This is what a maintainer of coezmaden/GPUAcceleratedTracking adds to use the B200 engine.  The file is meant to be
# `include`d INSIDE the reference's module, after src/algorithms.jl (src/GPUAcceleratedTracking.jl:92-98), where
# `kernel_algorithm`, `KernelAlgorithm`, `NumAnts`, `SVector`, `Hz`, `ustrip` are already in scope:
#
#     include("GATB200.jl")                         # this file
#     ALGODICT["b200"] = :b200                      # next to src/GPUAcceleratedTracking.jl:44-61
#     GATB200.init!(0; systems = [(GATB200.GAT_GPSL1, GPSL1()), (GATB200.GAT_GPSL5, GPSL5())])
#
# It adds ONE method to each of the reference's two entry points, with the reference's own positional signatures:
#     kernel_algorithm(...25 arguments..., ::KernelAlgorithm{:b200})            (src/algorithms.jl:1485-1512)
#     downconvert_and_correlate!(::B200, system, signal, correlator, ...)        (call site src/benchmarks.jl:63-79)
# INTEGRATION.md quotes these two methods verbatim.  Julia is not installed in the build image, so this file is kept
# minimal and mechanically checkable against include/gat.h (tests/test_abi_and_host.py checks every ccall symbol and
# its argument count); the same calls are exercised from Python (gpuacceleratedtracking_b200/api.py) in the test suite.
module GATB200

using CUDA, StaticArrays, StructArrays
import Unitful: Hz, ustrip

const libgat = get(ENV, "LIBGAT", "libgat.so")

const GAT_ACCUMULATE     = Cuint(1)
const GAT_CODE_PHASE_F64 = Cuint(2)
const GAT_TENSOR_TF32    = Cuint(8)      # tensor-core path for blocks shared by many channels (include/gat.h)
const GAT_GPSL1 = Cint(0)
const GAT_GPSL5 = Cint(1)

# mirrors `struct gat_channel` (include/gat.h)
struct GatChannel
    system_id::Int32
    prn::Int32
    code_phase_chips::Float64
    code_freq_hz::Float64
    carrier_phase_cycles::Float64
    carrier_freq_hz::Float64
end

mutable struct Context
    handle::Ptr{Cvoid}
end

function check(ctx::Context, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:gat_last_error, libgat), Cstring, (Ptr{Cvoid},), ctx.handle))
    error("libgat status $rc: $msg")
end

function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:gat_create, libgat), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
    rc == 0 || error("gat_create failed with status $rc (libgat needs an sm_100 GPU; there is no CPU fallback)")
    ctx = Context(h[])
    finalizer(c -> ccall((:gat_destroy, libgat), Cint, (Ptr{Cvoid},), c.handle), ctx)
    # queue libgat's work on CUDA.jl's task-local stream so it is ordered with the caller's kernels
    check(ctx, ccall((:gat_set_stream, libgat), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, CUDA.stream().handle))
    ctx
end

"Upload `system.codes` (Int8 +-1, column-major [code_length x n_prn]) once per system."
function set_codes!(ctx::Context, system_id::Integer, codes::AbstractMatrix, code_frequency = nothing)
    tab = Matrix{Int8}(codes[1:size(codes, 1), :])
    check(ctx, ccall((:gat_set_codes, libgat), Cint, (Ptr{Cvoid}, Cint, Ptr{Int8}, Cint, Cint),
                     ctx.handle, system_id, tab, size(tab, 1), size(tab, 2)))
    code_frequency === nothing ||
        check(ctx, ccall((:gat_set_code_frequency, libgat), Cint, (Ptr{Cvoid}, Cint, Cdouble), ctx.handle, system_id, hz(code_frequency)))
end

hz(x) = Float64(ustrip(Hz, x))
hz(x::Real) = Float64(x)

# ---- process-wide state: one context, the system -> id table ---------------------------------------------------
const CTX = Ref{Context}()
const SYSTEM_IDS = IdDict{Any, Cint}()          # GNSS system type -> libgat system id

"One-time setup: create the context and upload the chip tables (replaces the CuTexture construction at src/benchmarks.jl:829-837)."
function init!(device::Integer = 0; systems = ())
    CTX[] = Context(device)
    for (id, system) in systems
        set_codes!(CTX[], id, system.codes, system.code_frequency)      # GNSSSignals: `codes`, `code_frequency` fields
        SYSTEM_IDS[typeof(system)] = Cint(id)
    end
    CTX[]
end

system_id(system) = get(SYSTEM_IDS, typeof(system), GAT_GPSL1)

"""
K satellite channels over ONE bound signal block in one launch -- the receiver case the reference's 3-D kernels index
with `sat = blockIdx.z` (src/algorithms.jl:656).  `out_re`/`out_im` are `CuArray{Float32}` [num_ants x NCOR x K]
(the 3d_4431 layout, src/algorithms.jl:712).  Asynchronous on the bound stream.
"""
function correlate_channels!(ctx::Context, out_re::CuArray{Float32}, out_im::CuArray{Float32},
        channels::Vector{GatChannel}, signal_re::CuArray{Float32}, signal_im::CuArray{Float32},
        correlator_sample_shifts::SVector{NCOR, Int64}, sampling_frequency, num_samples::Integer;
        start_sample::Integer = 0, tensor::Bool = false, accumulate::Bool = false) where {NCOR}
    M = size(signal_re, 2)
    ld = size(signal_re, 1)
    check(ctx, ccall((:gat_bind_signal, libgat), Cint,
                     (Ptr{Cvoid}, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cint, Cint, Cint),
                     ctx.handle, 0, pointer(signal_re), pointer(signal_im), start_sample + num_samples, M, ld))
    shifts = Int32.(collect(correlator_sample_shifts))
    flags = (tensor ? GAT_TENSOR_TF32 : Cuint(0)) | (accumulate ? GAT_ACCUMULATE : Cuint(0))
    check(ctx, ccall((:gat_correlate, libgat), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{GatChannel}, Cdouble, Ptr{Int32}, Cint, Cint, Cint,
                      CuPtr{Cfloat}, CuPtr{Cfloat}, Cint, Cuint),
                     ctx.handle, 0, length(channels), channels, hz(sampling_frequency), shifts, NCOR, start_sample,
                     num_samples, pointer(out_re), pointer(out_im), 1, flags))
    return nothing
end

"""
A host loop of `downconvert_and_correlate!` calls over P consecutive blocks in ONE call (gat_ingest_correlate): `re`, `im`
are host `Array{Float32,3}` [ld x M x P]; `channels` is [K x P]; the library pipelines the H2D copies under the kernels.
Pin the arrays once with `host_register!` to get asynchronous copies at full PCIe rate.  Returns `Array{ComplexF32,4}`
[M x NCOR x K x P].
"""
function ingest_correlate(ctx::Context, re::Array{Float32, 3}, im::Array{Float32, 3}, channels::Matrix{GatChannel},
        correlator_sample_shifts::SVector{NCOR, Int64}, sampling_frequency, num_samples::Integer; start_sample::Integer = 0) where {NCOR}
    ld, M, P = size(re)
    K = size(channels, 1)
    pre = [pointer(re, 1 + (p - 1) * ld * M) for p in 1:P]
    pim = [pointer(im, 1 + (p - 1) * ld * M) for p in 1:P]
    out_re = Array{Float32}(undef, M, NCOR, K, P)
    out_im = similar(out_re)
    shifts = Int32.(collect(correlator_sample_shifts))
    GC.@preserve re im check(ctx, ccall((:gat_ingest_correlate, libgat), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Ptr{Cfloat}}, Ptr{Ptr{Cfloat}}, Cint, Cint, Cint, Ptr{GatChannel}, Cdouble, Ptr{Int32}, Cint, Cint, Cint,
         Ptr{Cfloat}, Ptr{Cfloat}, Cuint),
        ctx.handle, P, pre, pim, ld, M, K, channels, hz(sampling_frequency), shifts, NCOR, start_sample, num_samples, out_re, out_im, 0))
    return complex.(out_re, out_im)
end

host_register!(a::Array) = ccall((:gat_host_register, libgat), Cint, (Ptr{Cvoid}, UInt64), a, sizeof(a))

# ---- resident sessions (gat_resident_*): the per-millisecond call of a tracking loop without a kernel launch ------------
"""
`resident_begin!(ctx, slots, channels, shifts, fs, start_sample, num_samples)` launches the correlate kernel once; it stays on
the device and `resident_correlate!(ctx, slot_index, channels, out_re, out_im)` then costs a PCIe round trip plus the
correlation itself instead of a kernel launch and a stream synchronisation (the reference's measurement is exactly one such
call per 1 ms block: src/benchmarks.jl:872).  `slots` hold blocks of one geometry (bind or upload them first); `channels`
is a `Vector{GatChannel}` of at most 32 channels; `out_re`, `out_im` are host `Array{Float32,3}` [M x NCOR x K].
`resident_end!(ctx)` frees the device again.
"""
function resident_begin!(ctx::Context, slots::Vector{<:Integer}, channels::Vector{GatChannel}, correlator_sample_shifts::SVector{NCOR, Int64},
        sampling_frequency, start_sample::Integer, num_samples::Integer) where {NCOR}
    shifts = Int32.(collect(correlator_sample_shifts))
    check(ctx, ccall((:gat_resident_begin, libgat), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Cint, Cint, Ptr{GatChannel}, Cdouble, Ptr{Int32}, Cint, Cint, Cint),
        ctx.handle, Int32.(slots), length(slots), length(channels), channels, hz(sampling_frequency), shifts, NCOR, start_sample, num_samples))
end

function resident_correlate!(ctx::Context, slot_index::Integer, channels::Vector{GatChannel}, out_re::Array{Float32, 3}, out_im::Array{Float32, 3})
    check(ctx, ccall((:gat_resident_correlate, libgat), Cint, (Ptr{Cvoid}, Cint, Ptr{GatChannel}, Ptr{Cfloat}, Ptr{Cfloat}),
        ctx.handle, slot_index, channels, out_re, out_im))
    return out_re, out_im
end

resident_end!(ctx::Context) = check(ctx, ccall((:gat_resident_end, libgat), Cint, (Ptr{Cvoid},), ctx.handle))

# ---- sample sharding: a ctx that holds samples [origin, ...) of every period (gat_set_sample_origin / gat_gather_sum) ----
set_sample_origin!(ctx::Context, origin::Integer) = check(ctx, ccall((:gat_set_sample_origin, libgat), Cint, (Ptr{Cvoid}, Cint), ctx.handle, origin))
gather_sum!(ctx::Context, out_re::CuArray{Float32}, out_im::CuArray{Float32}) =
    check(ctx, ccall((:gat_gather_sum, libgat), Cint, (Ptr{Cvoid}, UInt64, CuPtr{Cfloat}, CuPtr{Cfloat}), ctx.handle, length(out_re), out_re, out_im))

# ---- all GPUs of the box from this one process (gat_mg_*) -------------------------------------------------------
mutable struct MultiContext
    handle::Ptr{Cvoid}
end

function mgcheck(mg::MultiContext, rc::Cint)
    rc == 0 && return nothing
    error("libgat status $rc: " * unsafe_string(ccall((:gat_mg_last_error, libgat), Cstring, (Ptr{Cvoid},), mg.handle)))
end

function MultiContext(devices::Vector{<:Integer}, n_slots::Integer, num_samples::Integer, num_ants::Integer; systems = ())
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:gat_mg_create, libgat), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), h, length(devices), Cint.(devices))
    rc == 0 || error("gat_mg_create failed with status $rc")
    mg = MultiContext(h[])
    finalizer(m -> ccall((:gat_mg_destroy, libgat), Cint, (Ptr{Cvoid},), m.handle), mg)
    for (id, system) in systems
        tab = Matrix{Int8}(system.codes)
        mgcheck(mg, ccall((:gat_mg_set_codes, libgat), Cint, (Ptr{Cvoid}, Cint, Ptr{Int8}, Cint, Cint), mg.handle, id, tab, size(tab, 1), size(tab, 2)))
    end
    mgcheck(mg, ccall((:gat_mg_configure, libgat), Cint, (Ptr{Cvoid}, Cint, Cint, Cint), mg.handle, n_slots, num_samples, num_ants))
    mg
end

"Scatter one host block ([ld x M] planes) over the devices: each gets ITS sample range through its own PCIe link (asynchronous)."
upload_signal!(mg::MultiContext, slot::Integer, re::Matrix{Float32}, im::Matrix{Float32}) =
    mgcheck(mg, ccall((:gat_mg_upload_signal, libgat), Cint, (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint), mg.handle, slot, re, im, size(re, 1)))

"Channels [K x P] sharded over the devices, blocks gathered over NVLink inside the kernels; `Array{ComplexF32,4}` [M x NCOR x K x P] back."
# :samples (default) or :satellites -- how a call is split over the devices (gat_mg_set_sharding)
set_sharding!(mg::MultiContext, mode::Symbol) =
    mgcheck(mg, ccall((:gat_mg_set_sharding, libgat), Cint, (Ptr{Cvoid}, Cint), mg.handle, mode === :satellites ? 1 : 0))

function correlate(mg::MultiContext, slots::Vector{<:Integer}, channels::Matrix{GatChannel}, correlator_sample_shifts::SVector{NCOR, Int64},
        sampling_frequency, num_samples::Integer, num_ants::Integer; start_sample::Integer = 0) where {NCOR}
    K, P = size(channels)
    out_re = Array{Float32}(undef, num_ants, NCOR, K, P)
    out_im = similar(out_re)
    mgcheck(mg, ccall((:gat_mg_correlate, libgat), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Int32}, Cint, Ptr{GatChannel}, Cdouble, Ptr{Int32}, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cuint),
        mg.handle, P, Int32.(slots), K, channels, hz(sampling_frequency), Int32.(collect(correlator_sample_shifts)), NCOR,
        start_sample, num_samples, out_re, out_im, 0))
    return complex.(out_re, out_im)
end

end # module GATB200

# =====================================================================================================================
# The two methods the reference gains.  Evaluated in the reference's module (see the header of this file).
# =====================================================================================================================

"""
GPU-style entry: the reference's own 26 positional arguments (src/algorithms.jl:1485-1512) dispatching on
`KernelAlgorithm{:b200}`.  Launch geometry, replica and scratch arguments are ignored (one fused launch plans itself);
the result ACCUMULATES into `accum_re`/`accum_im` (`CuMatrix{Float32}` [num_ants x num_corrs], the 4431 semantics of
src/algorithms.jl:625-632) and the call is asynchronous -- wrap it in `CUDA.@sync` as src/benchmarks.jl:872 does.
"""
function kernel_algorithm(
    threads_per_block, blocks_per_grid, shmem_size, code_replica, codes, code_frequency, sampling_frequency,
    start_code_phase, prn, num_samples, num_of_shifts, code_length, accum_re, accum_im,
    carrier_replica_re, carrier_replica_im, downconverted_signal_re, downconverted_signal_im, signal_re, signal_im,
    correlator_sample_shifts::SVector{NCOR, Int64}, carrier_frequency, carrier_phase, num_ants::NumAnts{NANT}, num_corrs,
    algorithm::KernelAlgorithm{:b200}
) where {NANT, NCOR}
    ctx = GATB200.CTX[]
    sid = code_length == 10230 ? GATB200.GAT_GPSL5 : GATB200.GAT_GPSL1       # the reference passes no system object on this path
    GATB200.check(ctx, ccall((:gat_bind_signal, GATB200.libgat), Cint,
                             (Ptr{Cvoid}, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cint, Cint, Cint),
                             ctx.handle, 0, pointer(signal_re), pointer(signal_im), num_samples, NANT, size(signal_re, 1)))
    ch = Ref(GATB200.GatChannel(sid, prn, start_code_phase, GATB200.hz(code_frequency), carrier_phase, GATB200.hz(carrier_frequency)))
    GATB200.check(ctx, ccall((:gat_correlate, GATB200.libgat), Cint,
                             (Ptr{Cvoid}, Cint, Cint, Ref{GATB200.GatChannel}, Cdouble, Ptr{Int32}, Cint, Cint, Cint,
                              CuPtr{Cfloat}, CuPtr{Cfloat}, Cint, Cuint),
                             ctx.handle, 0, 1, ch, GATB200.hz(sampling_frequency), Int32.(collect(correlator_sample_shifts)), NCOR,
                             0, num_samples, pointer(accum_re), pointer(accum_im), 1, GATB200.GAT_ACCUMULATE))
    return nothing
end

"Backend tag for the CPU-style entry: `downconvert_and_correlate!(B200(), system, signal, ...)`."
struct B200 end

"""
CPU-style entry: the 15 positional arguments of `Tracking.downconvert_and_correlate!` (call site src/benchmarks.jl:63-79)
behind a `B200()` tag.  `signal` is a host `StructArray{ComplexF32}` ([N] or [N, M]); the three scratch buffers are accepted
and left untouched.  Returns a NEW correlator whose accumulators are the old ones plus this block's sums, like upstream.
"""
function downconvert_and_correlate!(::B200, system, signal::StructArray, correlator, code_replica, code_phase, carrier_replica,
        carrier_phase, downconverted_signal, code_frequency, correlator_sample_shifts::SVector{NCOR, Int64}, carrier_frequency,
        sampling_frequency, signal_start_sample::Integer, num_samples_left::Integer, prn::Integer) where {NCOR}
    ctx = GATB200.CTX[]
    re, im = signal.re, signal.im
    M = ndims(re) == 1 ? 1 : size(re, 2)
    ch = Ref(GATB200.GatChannel(GATB200.system_id(system), prn, code_phase, GATB200.hz(code_frequency), carrier_phase,
                                GATB200.hz(carrier_frequency)))
    out_re = Matrix{Float32}(undef, M, NCOR)
    out_im = Matrix{Float32}(undef, M, NCOR)
    GATB200.check(ctx, ccall((:gat_downconvert_and_correlate, GATB200.libgat), Cint,
                             (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Ref{GATB200.GatChannel}, Cdouble, Ptr{Int32},
                              Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cuint),
                             ctx.handle, re, im, size(re, 1), M, 1, ch, GATB200.hz(sampling_frequency),
                             Int32.(collect(correlator_sample_shifts)), NCOR, signal_start_sample - 1, num_samples_left, out_re, out_im, 0))
    acc = complex.(out_re, out_im)                                    # [M, L]
    new = M == 1 ? SVector{NCOR}(vec(acc)) : SVector{NCOR}(ntuple(l -> SVector{M}(acc[:, l]), NCOR))
    return typeof(correlator)(Tracking.get_accumulators(correlator) .+ new)
end
