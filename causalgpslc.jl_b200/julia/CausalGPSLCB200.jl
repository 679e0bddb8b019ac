# CausalGPSLCB200.jl — `ccall` glue that re-points the hot path of CausalGPSLC.jl at libgpslc_b200.so.
#
# NOT EXECUTED in the build image (no Julia there): kept declarative and mechanical on purpose. Every function below
# replaces the body of the reference function named in its docstring; signatures are the reference's.
#
# Usage inside the reference package (see INTEGRATION.md):
#     include("CausalGPSLCB200.jl"); using .CausalGPSLCB200
#     CausalGPSLCB200.init!("/path/to/libgpslc_b200.so")        # once per process, binds one GPU
module CausalGPSLCB200

using Gen
using LinearAlgebra      # PosDefException
using Random             # default seeds come from Julia's global RNG, so Random.seed!(...) keeps working as in the reference

const LIB = Ref{String}("libgpslc_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

struct GpslcData            # include/gpslc.h: gpslc_data
    n::Cint; nX::Cint; nU::Cint; binary::Cint
    X::Ptr{Cdouble}; T::Ptr{Cdouble}; Y::Ptr{Cdouble}
    n_obj::Cint; obj_counts::Ptr{Cint}
    sigma_u_eps::Cdouble; sigma_u_cov::Cdouble
    per_chain_data::Cint
    sigma_u_dense::Ptr{Cdouble}
end
struct GpslcPrior           # gpslc_prior
    shape::NTuple{13,Cdouble}; scale::NTuple{13,Cdouble}; drift::Cdouble
end
struct GpslcOpts            # gpslc_opts
    nOuter::Cint; nMHInner::Cint; nESInner::Cint; n_chains::Cint
    seed::UInt64; chain_offset::Cint; u_layout_mode::Cint; ess_rule::Cint; observe_x::Cint
end

const FAMILIES = ["uNoise", "xNoise", "tNoise", "yNoise", "xScale", "tScale", "yScale", "uxLS", "utLS", "xtLS", "uyLS", "xyLS", "tyLS"]

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:gpslc_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[]))
    rc == 3 && throw(LinearAlgebra.PosDefException(0))      # GPSLC_ERR_NOT_PD: what `cholesky` throws in the reference
    error("gpslc error $rc: $msg")
end

"per-task LAPACK-style info of gpslc_ite / gpslc_sate: the reference aborts with PosDefException when Kp or CovITE is not PD"
function check_info(info::Vector{Cint})
    k = findfirst(!=(0), info)
    k === nothing || throw(LinearAlgebra.PosDefException(Int(info[k])))
end

function init!(libpath::String=LIB[]; device::Int=0)
    LIB[] = libpath
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:gpslc_create, LIB[]), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
    rc == 0 || error("gpslc_create failed ($rc): no sm_100 GPU; there is no CPU fallback")
    CTX[] = h[]
end

"""
Block structure of a SigmaU built by generateSigmaU (src/utils.jl:17-33): (counts, eps, cov) with 1 + eps == the common diagonal
value and `cov` the common within-object entry, or `nothing` for any other matrix (which is then passed densely: the library factors
it once, as generateU's `mvnormal` would on every update, src/model_prior.jl:27-30).
"""
function sigmaUstructure(SigmaU::Matrix{Float64})
    n = size(SigmaU, 1); counts = Cint[]; i = 1
    dg = SigmaU[1, 1]; cov = nothing
    all(SigmaU[k, k] == dg for k in 1:n) || return nothing
    while i <= n
        j = i + 1
        while j <= n && SigmaU[i, j] != 0.0; j += 1; end
        if j - i > 1
            c = SigmaU[i, i+1]
            cov === nothing ? (cov = c) : (c == cov || return nothing)
        end
        push!(counts, j - i); i = j
    end
    cov = cov === nothing ? 0.0 : cov
    R = zeros(n, n); i = 1
    for m in counts; R[i:i+m-1, i:i+m-1] .= cov; i += m; end
    for k in 1:n; R[k, k] = dg; end
    (R == SigmaU && dg - cov > 0 && cov >= 0) || return nothing
    counts, dg - 1.0, cov
end

prior_struct(pp) = GpslcPrior(ntuple(k -> Float64(pp[FAMILIES[k] * "Shape"]), 13),
                              ntuple(k -> Float64(pp[FAMILIES[k] * "Scale"]), 13), Float64(pp["drift"]))

"packed record (SURVEY.md App. A7) -> Gen choicemap with the reference's addresses (src/proposal.jl:8-22, src/inference.jl:9-15)"
function to_choicemap(rec::AbstractVector{Float64}, n, nU, nX, X, T, Y, binary)
    cm = Gen.choicemap()
    names = (:uNoise, :tNoise, :yNoise, :tyLS, :tScale, :yScale)
    for (k, s) in enumerate(names); isnan(rec[k]) || (cm[s] = rec[k]); end
    np = 6 + 4nX + 2nU + nU * nX
    # no-U models never observe X (SURVEY.md App. B3): the record then carries the model's own X after U / logitT (include/gpslc.h, stride)
    xoff = np + nU * n + (binary ? n : 0)
    has_xmodel = nU == 0 && nX > 0 && length(rec) >= xoff + n * nX
    for k in 1:nX
        isnan(rec[6+k]) || (cm[:xNoise=>k=>:Noise] = rec[6+k])
        isnan(rec[6+nX+k]) || (cm[:xScale=>k=>:Scale] = rec[6+nX+k])
        cm[:xtLS=>k=>:LS] = rec[6+2nX+k]; cm[:xyLS=>k=>:LS] = rec[6+3nX+k]
        cm[:X=>k=>:X] = has_xmodel ? rec[xoff+(k-1)*n+1:xoff+k*n] : X[:, k]
    end
    for i in 1:nU
        cm[:utLS=>i=>:LS] = rec[6+4nX+i]; cm[:uyLS=>i=>:LS] = rec[6+4nX+nU+i]
        for j in 1:nX; cm[:uxLS=>i=>j=>:LS] = rec[6+4nX+2nU+(i-1)*nX+j]; end
        cm[:U=>i=>:U] = rec[np+(i-1)*n+1:np+i*n]
    end
    cm[:Y] = Y
    if binary
        cm[:logitT] = rec[np+nU*n+1:np+nU*n+n]
        for r in 1:n; cm[:T=>r=>:T] = T[r]; end
    else
        cm[:T] = T
    end
    cm
end

"""
    Posterior(priorparams, X, T, Y, nU, nOuter, nMHInner, nESInner)
Replaces all eight methods of src/inference.jl:4-379 (called from samplePosterior, src/driver.jl:59-69).
Returns (posteriorSamples::Vector{Any} of choicemaps, packed::Array{Float64,3}).
DEVIATION from the reference, on purpose: the reference returns `(samples, trace)` (src/inference.jl:58); no Gen trace exists on
the GPU path, so the second value is the packed sample buffer [stride, n_chains, nOuter] instead. Its only caller,
samplePosterior (src/driver.jl:62), discards it.
`seed` defaults to a draw from Julia's global RNG: like the reference, repeated calls give different chains and
`Random.seed!(1234)` (test/runtests.jl:18) makes a run reproducible.
"""
function Posterior(priorparams, X, T, Y, nU, nOuter, nMHInner, nESInner; n_chains=1, seed=rand(UInt64), chain_offset=0,
                   u_layout_mode=0, ess_rule=0, observe_x=0)
    n = length(Y); nX = X === nothing ? 0 : size(X, 2); nu = nU === nothing ? 0 : nU
    binary = eltype(T) == Bool
    Tf = Float64.(T); Xf = X === nothing ? Float64[] : Matrix{Float64}(X); Yf = Float64.(Y)
    st = nu > 0 ? sigmaUstructure(Matrix{Float64}(priorparams["SigmaU"])) : nothing
    counts = st === nothing ? Cint[] : st[1]
    eps_u = st === nothing ? Float64(priorparams["sigmaUNoise"]) : st[2]
    cov_u = st === nothing ? Float64(priorparams["sigmaUCov"]) : st[3]
    Sd = (nu > 0 && st === nothing) ? Matrix{Float64}(priorparams["SigmaU"]) : zeros(0, 0)     # unstructured SigmaU: dense path
    np = 6 + 4nX + 2nu + nu * nX
    stride = np + nu * n + (binary ? n : 0) + ((nu == 0 && nX > 0 && observe_x == 0) ? n * nX : 0)
    out = Array{Float64}(undef, stride, n_chains, nOuter)          # column-major == C [nOuter][n_chains][stride]
    GC.@preserve Tf Xf Yf counts Sd out begin
        d = GpslcData(n, nX, nu, binary, nX > 0 ? pointer(Xf) : C_NULL, pointer(Tf), pointer(Yf), length(counts),
                      length(counts) > 0 ? pointer(counts) : C_NULL, eps_u, cov_u, 0, length(Sd) > 0 ? pointer(Sd) : C_NULL)
        o = GpslcOpts(nOuter, something(nMHInner, 0), something(nESInner, 0), n_chains, seed, chain_offset, u_layout_mode, ess_rule, observe_x)
        check(ccall((:gpslc_posterior, LIB[]), Cint,
                    (Ptr{Cvoid}, Ref{GpslcData}, Ref{GpslcPrior}, Ref{GpslcOpts}, Ptr{Cdouble}, Ptr{Culonglong}, Ptr{Culonglong}),
                    CTX[], d, prior_struct(priorparams), o, out, C_NULL, C_NULL))
    end
    samples = Any[to_choicemap(view(out, :, 1, i), n, nu, nX, X, T, Y, binary) for i in 1:nOuter]
    samples, out
end

retained(h) = Cint.(collect(h.nBurnIn:h.stepSize:h.nOuter) .- 1)      # 0-based; src/estimation.jl:72,78

"pack g.posteriorSamples (choicemaps) back into records for the estimation calls"
function pack(g)
    n = length(g.Y); nX = g.X === nothing ? 0 : size(g.X, 2); nU = g.hyperparams.nU === nothing ? 0 : g.hyperparams.nU
    np = 6 + 4nX + 2nU + nU * nX
    out = fill(NaN, np + nU * n, 1, length(g.posteriorSamples))
    for (i, cm) in enumerate(g.posteriorSamples)
        out[3, 1, i] = cm[:yNoise]; out[4, 1, i] = cm[:tyLS]; out[6, 1, i] = cm[:yScale]
        for k in 1:nX; out[6+3nX+k, 1, i] = cm[:xyLS=>k=>:LS]; end
        for u in 1:nU
            out[6+4nX+nU+u, 1, i] = cm[:uyLS=>u=>:LS]
            out[np+(u-1)*n+1:np+u*n, 1, i] = cm[:U=>u=>:U]
        end
    end
    out
end

function data_struct(g, Tf, Xf, Yf)
    nX = g.X === nothing ? 0 : size(g.X, 2); nU = g.hyperparams.nU === nothing ? 0 : g.hyperparams.nU
    GpslcData(length(Yf), nX, nU, eltype(g.T) == Bool, nX > 0 ? pointer(Xf) : C_NULL, pointer(Tf), pointer(Yf), 0, C_NULL, 0.0, 0.0, 0, C_NULL)
end

"""
    sampleITE(g, doT; samplesPerPosterior=10)   — replaces src/driver.jl:86-89 (ITEDistributions + ITEsamples,
src/estimation.jl:66-109, likelihoodDistribution src/likelihood.jl:8-174). Returns n × (R*samplesPerPosterior).
"""
function sampleITE(g, doT; samplesPerPosterior::Int=10, seed=rand(UInt64))
    packed = pack(g); ret = retained(g.hyperparams); R = length(ret); n = length(g.Y)
    Tf = Float64.(g.T); Xf = g.X === nothing ? Float64[] : Matrix{Float64}(g.X); Yf = Float64.(g.Y)
    out = Array{Float64}(undef, n, R * samplesPerPosterior)
    dts = Float64[doT]; info = zeros(Cint, R)
    GC.@preserve packed ret Tf Xf Yf out dts info begin
        check(ccall((:gpslc_ite, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ref{GpslcData}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Cint, Cdouble,
                     Cint, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                    CTX[], 0, data_struct(g, Tf, Xf, Yf), packed, size(packed, 3), 1, size(packed, 1), ret, R, dts, 1,
                    g.hyperparams.predictionCovarianceNoise, samplesPerPosterior, seed, 0, C_NULL, C_NULL, out, info))
    end
    check_info(info)
    out
end

"""
    sampleSATE(g, doT; samplesPerPosterior=10)  — replaces src/driver.jl:108-111 (SATEDistributions + SATEsamples,
src/estimation.jl:116-163). var_as_std=1 keeps the reference's `normal(mean, var)` behaviour.
"""
function sampleSATE(g, doT; samplesPerPosterior::Int=10, seed=rand(UInt64), var_as_std::Int=1)
    packed = pack(g); ret = retained(g.hyperparams); R = length(ret)
    Tf = Float64.(g.T); Xf = g.X === nothing ? Float64[] : Matrix{Float64}(g.X); Yf = Float64.(g.Y)
    out = Vector{Float64}(undef, R * samplesPerPosterior)
    dts = Float64[doT]; info = zeros(Cint, R)
    GC.@preserve packed ret Tf Xf Yf out dts info begin
        check(ccall((:gpslc_sate, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ref{GpslcData}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Cint, Cdouble,
                     Cint, UInt64, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                    CTX[], 0, data_struct(g, Tf, Xf, Yf), packed, size(packed, 3), 1, size(packed, 1), ret, R, dts, 1,
                    g.hyperparams.predictionCovarianceNoise, samplesPerPosterior, seed, 0, var_as_std, C_NULL, C_NULL, out, info))
    end
    check_info(info)
    out
end

"""
    predictCounterfactualEffects(g, nSamplesPerMixture; fidelity=100, minDoT, maxDoT) — replaces src/prediction.jl:23-36:
one gpslc_ite call for all doT values instead of one sampleITE per doT.
"""
function predictCounterfactualEffects(g, nSamplesPerMixture::Int; fidelity::Int=100, minDoT=min(g.T...), maxDoT=max(g.T...), seed=rand(UInt64))
    doTrange = minDoT:(abs(maxDoT - minDoT) / fidelity):maxDoT
    dts = collect(Float64, doTrange); D = length(dts)
    packed = pack(g); ret = retained(g.hyperparams); R = length(ret); n = length(g.Y)
    Tf = Float64.(g.T); Xf = g.X === nothing ? Float64[] : Matrix{Float64}(g.X); Yf = Float64.(g.Y)
    out = Array{Float64}(undef, n, R * nSamplesPerMixture, D)      # C layout [D][1][R*spp][n]
    info = zeros(Cint, R * D)
    GC.@preserve packed ret Tf Xf Yf out dts info begin
        check(ccall((:gpslc_ite, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ref{GpslcData}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Cint, Cdouble,
                     Cint, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                    CTX[], 0, data_struct(g, Tf, Xf, Yf), packed, size(packed, 3), 1, size(packed, 1), ret, R, dts, D,
                    g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, seed, 0, C_NULL, C_NULL, out, info))
    end
    check_info(info)
    permutedims(out, (3, 1, 2)), doTrange                           # ite[d, n, R*spp] as the reference returns
end

"""
    predictCounterfactualEffectsShard(g, nSamplesPerMixture, world_size, rank; fidelity=100, minDoT, maxDoT)
One rank's block of the doT range of predictCounterfactualEffects (src/prediction.jl:23-36) through gpslc_ite_slice: the doT values
are the independent units of the sweep; the draws are keyed by the global doT index, so the blocks of all ranks concatenate to
exactly the unsharded result. Returns (ite[d_local, n, R*spp], the full doT range, offset of this block).
"""
function predictCounterfactualEffectsShard(g, nSamplesPerMixture::Int, world_size::Int, rank::Int; fidelity::Int=100,
                                           minDoT=min(g.T...), maxDoT=max(g.T...), seed=UInt64(0))   # all ranks must pass the SAME seed
    doTrange = minDoT:(abs(maxDoT - minDoT) / fidelity):maxDoT
    alldts = collect(Float64, doTrange); D = length(alldts)
    base, rem = divrem(D, world_size)
    cnt = base + (rank < rem ? 1 : 0); off = rank * base + min(rank, rem)
    dts = alldts[off+1:off+cnt]
    packed = pack(g); ret = retained(g.hyperparams); R = length(ret); n = length(g.Y)
    Tf = Float64.(g.T); Xf = g.X === nothing ? Float64[] : Matrix{Float64}(g.X); Yf = Float64.(g.Y)
    out = Array{Float64}(undef, n, R * nSamplesPerMixture, cnt)
    info = zeros(Cint, R * cnt)
    GC.@preserve packed ret Tf Xf Yf out dts info begin
        check(ccall((:gpslc_ite_slice, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ref{GpslcData}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Cint, Cint, Cdouble,
                     Cint, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                    CTX[], 0, data_struct(g, Tf, Xf, Yf), packed, size(packed, 3), 1, size(packed, 1), ret, R, dts, cnt, off,
                    g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, seed, 0, C_NULL, C_NULL, out, info))
    end
    check_info(info)
    permutedims(out, (3, 1, 2)), doTrange, off
end

"""
    predictCounterfactualSummaryShard(g, nSamplesPerMixture, world_size, rank; fidelity=100, minDoT, maxDoT, credible_interval=0.90)
predictCounterfactualEffects (src/prediction.jl:23-36) followed by summarizeEstimates (src/driver.jl:129-149) per doT, fused on the
device through gpslc_ite_summary: the ITE draws of this rank's block of doT values never leave HBM (BASELINE c5: 168 MB of draws), only
`summary[3, n, d_local]` = (Mean, LowerBound, UpperBound) per individual and doT comes back. Returns (summary, doTrange, offset).
"""
function predictCounterfactualSummaryShard(g, nSamplesPerMixture::Int, world_size::Int, rank::Int; fidelity::Int=100,
                                           minDoT=min(g.T...), maxDoT=max(g.T...), seed=UInt64(0), credible_interval::Float64=0.90)
    doTrange = minDoT:(abs(maxDoT - minDoT) / fidelity):maxDoT
    alldts = collect(Float64, doTrange); D = length(alldts)
    base, rem = divrem(D, world_size)
    cnt = base + (rank < rem ? 1 : 0); off = rank * base + min(rank, rem)
    dts = alldts[off+1:off+cnt]
    packed = pack(g); ret = retained(g.hyperparams); R = length(ret); n = length(g.Y)
    Tf = Float64.(g.T); Xf = g.X === nothing ? Float64[] : Matrix{Float64}(g.X); Yf = Float64.(g.Y)
    out = Array{Float64}(undef, 3, n, cnt)                           # C layout [d_local][1][n][3]
    info = zeros(Cint, R * cnt)
    GC.@preserve packed ret Tf Xf Yf out dts info begin
        check(ccall((:gpslc_ite_summary, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ref{GpslcData}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Cint, Cint, Cdouble,
                     Cint, UInt64, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cint}),
                    CTX[], 0, data_struct(g, Tf, Xf, Yf), packed, size(packed, 3), 1, size(packed, 1), ret, R, dts, cnt, off,
                    g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, seed, 0, credible_interval, out, info))
    end
    check_info(info)
    out, doTrange, off
end

"""
    subgroupEffectCurveShard(g, idx, nSamplesPerMixture, world_size, rank; fidelity=100, minDoT, maxDoT, credible_interval=0.90)
The documented subgroup workflow (docs/src/index.md:101-114) fused on the device through gpslc_ite_subset_summary:
`ite, doT = predictCounterfactualEffects(g, nSamples); sate = mean(ite[:, idx, :], dims=2)[:, 1, :]; summarizeEstimates(sate)`.
`idx` is a Bool vector over the individuals (e.g. `vec(g.obj .== "MA")`). Returns (summary, sate, doTrange, offset) for this rank's
block of doT values: `summary[3, d_local]` = (Mean, LowerBound, UpperBound) per doT, `sate[d_local, R*nSamplesPerMixture]`.
"""
function subgroupEffectCurveShard(g, idx::AbstractVector{Bool}, nSamplesPerMixture::Int, world_size::Int, rank::Int; fidelity::Int=100,
                                  minDoT=min(g.T...), maxDoT=max(g.T...), seed=UInt64(0), credible_interval::Float64=0.90)
    doTrange = minDoT:(abs(maxDoT - minDoT) / fidelity):maxDoT
    alldts = collect(Float64, doTrange); D = length(alldts)
    base, rem = divrem(D, world_size)
    cnt = base + (rank < rem ? 1 : 0); off = rank * base + min(rank, rem)
    dts = alldts[off+1:off+cnt]
    packed = pack(g); ret = retained(g.hyperparams); R = length(ret); n = length(g.Y)
    length(idx) == n || throw(ArgumentError("idx must have one entry per individual"))
    mask = UInt8.(idx)
    Tf = Float64.(g.T); Xf = g.X === nothing ? Float64[] : Matrix{Float64}(g.X); Yf = Float64.(g.Y)
    out = Array{Float64}(undef, 3, cnt)                              # C layout [1][d_local][3]
    sate = Array{Float64}(undef, cnt, R * nSamplesPerMixture)        # C layout [1][R*spp][d_local]
    info = zeros(Cint, R * cnt)
    GC.@preserve packed ret Tf Xf Yf out sate dts info mask begin
        check(ccall((:gpslc_ite_subset_summary, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ref{GpslcData}, Ptr{Cdouble}, Cint, Cint, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Cint, Cint, Cdouble,
                     Cint, UInt64, Cint, Ptr{UInt8}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cint}),
                    CTX[], 0, data_struct(g, Tf, Xf, Yf), packed, size(packed, 3), 1, size(packed, 1), ret, R, dts, cnt, off,
                    g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, seed, 0, mask, credible_interval, sate, out, info))
    end
    check_info(info)
    out, sate, doTrange, off
end

"""
    summarizeEstimates(samples; credible_interval=0.90) — the statistics of src/driver.jl:129-149 (row mean and the two
quantiles, Julia `quantile` default) through gpslc_summarize. `samples` is the n × m matrix sampleITE returns, whose memory
is exactly the C layout [m][n]. Returns (Mean, LowerBound, UpperBound) vectors; the DataFrame/CSV part stays in driver.jl.
"""
function summarizeEstimates(samples::Matrix{Float64}; credible_interval::Float64=0.90)
    n, m = size(samples)
    out = Matrix{Float64}(undef, 3, n)                               # C layout [n][3]
    GC.@preserve samples out begin
        check(ccall((:gpslc_summarize, LIB[]), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Cint, Cint, Cdouble, Ptr{Cdouble}),
                    CTX[], 0, samples, 1, m, n, credible_interval, out))
    end
    out[1, :], out[2, :], out[3, :]
end

"rbfKernelLog / processCov (src/kernel.jl:24-59) through gpslc_cov_build — parity layer"
function covBuild(X1::Matrix{Float64}, X2::Matrix{Float64}, LS::Vector{Float64}, scale::Float64, noise::Union{Float64,Nothing}=nothing)
    n, D = size(X1); K = Matrix{Float64}(undef, n, n)
    sc = [scale]; nz = noise === nothing ? Float64[] : [noise]
    GC.@preserve X1 X2 LS sc nz K begin
        check(ccall((:gpslc_cov_build, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                    CTX[], 0, n, 1, D, X1, X2, 1, LS, sc, noise === nothing ? C_NULL : pointer(nz), K))
    end
    K
end

# ---- persistence: the language-neutral `.gpslc` container of gpslc_b200/io.py (the reference's src/io.jl serialises the Julia
# object, which only Julia can read). Layout: 8-byte magic "GPSLCB2\0", little-endian UInt64 header length, JSON header
# {"version","hyperparams","priorparams","seed","arrays":[{"name","dtype","shape"}...]}, then the arrays raw, C order, each padded
# to 8 bytes. Only the packed posterior samples are needed to rebuild g.posteriorSamples (to_choicemap).
json_num(x) = x === nothing ? "null" : (x isa Bool ? (x ? "true" : "false") : (x isa Integer ? string(x) : repr(Float64(x))))
np_dtype(::Type{Float64}) = "<f8"; np_dtype(::Type{Bool}) = "|b1"; np_dtype(::Type{Int64}) = "<i8"

"""
    savePacked(filename, g, packed; seed=0)
Write `g` (a GPSLCObject) with the packed posterior samples `packed` [stride, n_chains, nOuter] (second return value of `Posterior`)
as a `.gpslc` container that `gpslc_b200.loadGPSLCObject` (Python mirror) reads: data arrays in C order, SigmaU as object counts when
it has the generateSigmaU structure. The `.gpslc` extension is optional (src/io.jl:15-17).
"""
function savePacked(filename::String, g, packed::Array{Float64,3}; seed::Integer=0)
    if length(filename) > 6 && filename[end-5:end] == ".gpslc"; filename = filename[1:end-6]; end
    h = g.hyperparams
    arrays = Pair{String,Array}["packed" => packed, "T" => (eltype(g.T) == Bool ? Vector{Bool}(g.T) : Float64.(g.T)), "Y" => Float64.(g.Y)]
    shapes = Dict("packed" => reverse(size(packed)), "T" => (length(g.T),), "Y" => (length(g.Y),))
    if g.X !== nothing
        push!(arrays, "X" => permutedims(Matrix{Float64}(g.X)))             # C order [n][nX] == column-major [nX, n]
        shapes["X"] = size(g.X)
    end
    if g.SigmaU !== nothing
        st = sigmaUstructure(Matrix{Float64}(g.SigmaU))
        if st === nothing
            push!(arrays, "SigmaU" => Matrix{Float64}(g.SigmaU)); shapes["SigmaU"] = size(g.SigmaU)      # symmetric: order irrelevant
        else
            push!(arrays, "obj_counts" => Int64.(st[1])); shapes["obj_counts"] = (length(st[1]),)
        end
    end
    hp = join(["\"$k\": " * json_num(getfield(h, k)) for k in (:nU, :nOuter, :nMHInner, :nESInner, :nBurnIn, :stepSize, :predictionCovarianceNoise)], ", ")
    pp = join(["\"$k\": " * json_num(v) for (k, v) in g.priorparams if v isa Real], ", ")
    arr = join(["{\"name\": \"$(nm)\", \"dtype\": \"$(np_dtype(eltype(a)))\", \"shape\": [$(join(shapes[nm], ", "))]}" for (nm, a) in arrays], ", ")
    header = "{\"version\": 1, \"hyperparams\": {$hp}, \"priorparams\": {$pp}, \"seed\": $(Int(seed)), \"arrays\": [$arr]}"
    open(filename * ".gpslc", "w") do io
        write(io, b"GPSLCB2\0"); write(io, htol(UInt64(sizeof(header)))); write(io, header)
        for (_, a) in arrays
            write(io, a); write(io, zeros(UInt8, mod(-sizeof(a), 8)))
        end
    end
end

end # module
