# GATB200.jl -- the reference-side binding of libgat (include/gat.h).
#
# This is what a maintainer of coezmaden/GPUAcceleratedTracking adds to use the B200 engine:
# a `kernel_algorithm(..., ::KernelAlgorithm{:b200})` method next to the existing ones
# (src/algorithms.jl:869-1545) and a `downconvert_and_correlate!` method for a `B200Correlator`
# backend (call site src/benchmarks.jl:63-79).  Julia is not installed in the build image, so
# this file is kept minimal and mechanically checkable against include/gat.h; the same calls are
# exercised from Python (gpuacceleratedtracking_b200/api.py) in the test suite.
module GATB200

using CUDA, StaticArrays, StructArrays
import Unitful: Hz, ustrip

const libgat = get(ENV, "LIBGAT", "libgat.so")

const GAT_ACCUMULATE     = Cuint(1)
const GAT_CODE_PHASE_F64 = Cuint(2)
const GAT_TENSOR_TF32    = Cuint(8)      # opt-in tensor-core path for blocks shared by many channels (include/gat.h)
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
function set_codes!(ctx::Context, system_id::Integer, codes::AbstractMatrix)
    tab = Matrix{Int8}(codes[1:size(codes, 1), :])
    check(ctx, ccall((:gat_set_codes, libgat), Cint, (Ptr{Cvoid}, Cint, Ptr{Int8}, Cint, Cint),
                     ctx.handle, system_id, tab, size(tab, 1), size(tab, 2)))
end

hz(x) = Float64(ustrip(Hz, x))
hz(x::Real) = Float64(x)

"""
GPU-style entry: same 26 positional arguments as `kernel_algorithm(..., ::KernelAlgorithm{4431})`
(src/algorithms.jl:1485-1512).  Launch geometry, replica and scratch arguments are ignored; the
result ACCUMULATES into `accum_re`/`accum_im` (CuMatrix{Float32} [num_ants x num_corrs]) and the
call is asynchronous -- wrap in `CUDA.@sync` as src/benchmarks.jl:872 does.
"""
function kernel_algorithm(ctx::Context, system_id::Integer,
        threads_per_block, blocks_per_grid, shmem_size, code_replica, codes, code_frequency,
        sampling_frequency, start_code_phase, prn, num_samples, num_of_shifts, code_length,
        accum_re::CuArray{Float32}, accum_im::CuArray{Float32},
        carrier_replica_re, carrier_replica_im, downconverted_signal_re, downconverted_signal_im,
        signal_re::CuArray{Float32}, signal_im::CuArray{Float32},
        correlator_sample_shifts::SVector{NCOR, Int64}, carrier_frequency, carrier_phase,
        num_ants, num_corrs) where {NCOR}
    M = size(signal_re, 2)
    ld = size(signal_re, 1)
    check(ctx, ccall((:gat_bind_signal, libgat), Cint,
                     (Ptr{Cvoid}, Cint, CuPtr{Cfloat}, CuPtr{Cfloat}, Cint, Cint, Cint),
                     ctx.handle, 0, pointer(signal_re), pointer(signal_im), num_samples, M, ld))
    ch = Ref(GatChannel(system_id, prn, start_code_phase, hz(code_frequency), carrier_phase, hz(carrier_frequency)))
    shifts = Int32.(collect(correlator_sample_shifts))
    check(ctx, ccall((:gat_correlate, libgat), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ref{GatChannel}, Cdouble, Ptr{Int32}, Cint, Cint, Cint,
                      CuPtr{Cfloat}, CuPtr{Cfloat}, Cint, Cuint),
                     ctx.handle, 0, 1, ch, hz(sampling_frequency), shifts, NCOR, 0, num_samples,
                     pointer(accum_re), pointer(accum_im), 1, GAT_ACCUMULATE))
    return nothing
end

"""
The receiver case the reference's 3-D kernels index with `sat = blockIdx.z` (src/algorithms.jl:656): K satellite
channels over ONE bound signal block in one launch.  `out_re`/`out_im` are `CuArray{Float32}` [num_ants x NCOR x K]
(the 3d_4431 layout, src/algorithms.jl:712).  `tensor = true` allows the `tcgen05` path (TF32 operands, FP32 sums;
<= 4 taps, <= 16 antennas) -- it pays from 32 channels per block upwards; the call is asynchronous.
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
CPU-style entry: the 15 positional arguments of `Tracking.downconvert_and_correlate!`
(src/benchmarks.jl:63-79).  `signal` is a host `StructArray{ComplexF32}` ([N] or [N, M]); the
three scratch buffers are accepted and left untouched.  Returns the accumulators of this block as
a `Matrix{ComplexF32}` [M, L]; the caller adds them to its immutable correlator
(`Tracking` returns `typeof(correlator)(get_accumulators(correlator) .+ ...)`).
"""
function downconvert_and_correlate!(ctx::Context, system_id::Integer, signal::StructArray, correlator, code_replica,
        code_phase, carrier_replica, carrier_phase, downconverted_signal, code_frequency,
        correlator_sample_shifts::SVector{NCOR, Int64}, carrier_frequency, sampling_frequency,
        signal_start_sample::Integer, num_samples_left::Integer, prn::Integer) where {NCOR}
    re, im = signal.re, signal.im
    M = ndims(re) == 1 ? 1 : size(re, 2)
    ld = size(re, 1)
    ch = Ref(GatChannel(system_id, prn, code_phase, hz(code_frequency), carrier_phase, hz(carrier_frequency)))
    shifts = Int32.(collect(correlator_sample_shifts))
    out_re = Matrix{Float32}(undef, M, NCOR)
    out_im = Matrix{Float32}(undef, M, NCOR)
    check(ctx, ccall((:gat_downconvert_and_correlate, libgat), Cint,
                     (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Ref{GatChannel}, Cdouble, Ptr{Int32},
                      Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cuint),
                     ctx.handle, re, im, ld, M, 1, ch, hz(sampling_frequency), shifts, NCOR,
                     signal_start_sample - 1, num_samples_left, out_re, out_im, 0))
    return complex.(out_re, out_im)
end

end # module
