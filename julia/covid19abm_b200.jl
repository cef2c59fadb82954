# covid19abm_b200.jl — drop-in for covid19abm.jl's `ModelParameters` / `main(ip)` / `runsim(simnum, ip)` over
# libcabm_b200.so (include/cabm.h). UNVERIFIED: Julia is not installed in the build image or on the GPU box;
# the same C ABI is exercised through Python ctypes in tests/. Field order and defaults follow
# /root/reference/src/covid19abm.jl:31-55.
module covid19abm_b200

using Parameters

const LIB = get(ENV, "CABM_LIB", joinpath(@__DIR__, "..", "covid19abm.jl_b200", "libcabm_b200.so"))

@enum HEALTH SUS LAT PRE ASYMP MILD MISO INF IISO HOS ICU REC DED UNDEF      # :5

@with_kw mutable struct ModelParameters @deftype Float64                       # :31-55
    β = 0.0
    frelasymp::Float64 = 0.11
    seasonal::Bool = false
    popsize::Int64 = 10000
    prov::Symbol = :newyork
    calibration::Bool = false
    modeltime::Int64 = 300
    initialinf::Int64 = 1
    initialhi::Int64 = 0
    asymp_props::NTuple{5, Float64} = (0.25, 0.25, 0.14, 0.07, 0.07)
    mild_props::NTuple{5, Float64} = (0.95, 0.9, 0.85, 0.6, 0.2)
    τmild::Int64 = 0
    fmild::Float64 = 0.0
    fsevere::Float64 = 0.0
    tpreiso::Int64 = 0
    fpreiso::Float64 = 0.0
    quar_grp::Int8 = 5
    quar_grp_frac::Float64 = 0.0
    ctstrat::Int8 = 0
    fctcapture::Float16 = 0.0
    fcontactst::Float16 = 0.0
    cdaysback::Int8 = 0
    strat3qdays::Int8 = 0
end

# struct cabm_params (include/cabm.h) — isbits, same layout as the C struct (192 bytes; asserted below)
struct CabmParams
    beta::Cdouble; frelasymp::Cdouble
    seasonal::Int32; popsize::Int32; prov_id::Int32; calibration::Int32; modeltime::Int32
    initialinf::Int32; initialhi::Int32; tau_mild::Int32
    asymp_props::NTuple{5, Cdouble}; mild_props::NTuple{5, Cdouble}
    fmild::Cdouble; fsevere::Cdouble; fpreiso::Cdouble; quar_grp_frac::Cdouble
    tpreiso::Int32; quar_grp::Int32; ctstrat::Int32; cdaysback::Int32; strat3qdays::Int32
    fctcapture::Cfloat; fcontactst::Cfloat; _pad::Int32
end

@assert sizeof(CabmParams) == 192 "CabmParams does not match struct cabm_params of include/cabm.h"

const PROVINCES = (:alberta, :bc, :canada, :manitoba, :newbruns, :newfdland, :nwterrito, :novasco,
                   :nunavut, :ontario, :pei, :quebec, :saskat, :yukon, :newyork)               # :319-333

function to_c(ip::ModelParameters)
    pid = findfirst(==(ip.prov), PROVINCES)
    CabmParams(ip.β, ip.frelasymp, ip.seasonal, ip.popsize, pid === nothing ? -1 : pid - 1, ip.calibration,
               ip.modeltime, ip.initialinf, ip.initialhi, ip.τmild, ip.asymp_props, ip.mild_props, ip.fmild,
               ip.fsevere, ip.fpreiso, ip.quar_grp_frac, ip.tpreiso, ip.quar_grp, ip.ctstrat, ip.cdaysback,
               ip.strat3qdays, Float32(ip.fctcapture), Float32(ip.fcontactst), 0)
end

function check(h, rc)
    rc == 0 && return
    error(unsafe_string(ccall((:cabm_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))   # the reference's error("...") text
end

# Random.seed!(simnum*726) (:78) seeds Julia's global RNG; here the replicate's draws are keyed Philox (DESIGN.md §3), so the
# "seed" is the Philox key. seed!(s) sets the key the next main(ip) uses, exactly where the reference calls Random.seed!.
const CURRENT_SEED = Ref{UInt64}(0)
seed!(s) = (CURRENT_SEED[] = UInt64(s))

# what the reference leaves in its globals after main (`humans`, `ct_data` :69-73) and runsim reads back (:84-90)
const LAST = Ref{Any}(nothing)

"main(ip) :104-142 — returns Matrix{Int16}(popsize, modeltime). In-model error()s come back as Julia errors (cabm_sync: CABM_E_MODEL)."
function main(ip::ModelParameters; device = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cabm_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint), h, device, 1, ip.popsize, ip.modeltime, 1)
    rc == 0 || error(unsafe_string(ccall((:cabm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    try
        p = Ref(to_c(ip)); seeds = UInt64[CURRENT_SEED[]]
        check(h[], ccall((:cabm_configure, LIB), Cint, (Ptr{Cvoid}, Ref{CabmParams}, Cint, Ptr{Int32}, Cint, Ptr{UInt64}, Cint),
                         h[], p, 1, C_NULL, 1, seeds, 0))
        check(h[], ccall((:cabm_run, LIB), Cint, (Ptr{Cvoid},), h[]))
        check(h[], ccall((:cabm_sync, LIB), Cint, (Ptr{Cvoid},), h[]))      # raises the reference's error("...") text if the replicate hit one
        hmatrix = Matrix{Int16}(undef, ip.popsize, ip.modeltime)          # column-major: one day per column (:111,:221)
        check(h[], ccall((:cabm_get_hmatrix, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int16}), h[], 0, hmatrix))
        inc = Array{Int32}(undef, 12, ip.modeltime, 6); prev = similar(inc)                       # C order [6][T][12]
        check(h[], ccall((:cabm_get_curves, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), h[], inc, prev))
        infectors = Vector{Int64}(undef, 4); tiso = Vector{Int64}(undef, 1)
        check(h[], ccall((:cabm_get_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt32}), h[], infectors, tiso, C_NULL, C_NULL))
        ags = Vector{Int32}(undef, ip.popsize)
        check(h[], ccall((:cabm_get_field, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}), h[], 4, 0, ags))   # CABM_F_AG
        LAST[] = (inc = inc, prev = prev, infectors = infectors, totalisolated = tiso[1], ags = ags)
        return hmatrix
    finally
        ccall((:cabm_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

const STATES = ("SUS", "LAT", "PRE", "ASYMP", "MILD", "MISO", "INF", "IISO", "HOS", "ICU", "REC", "DED")

"_collectdf :226-245 for one group, from the device-reduced curves: NamedTuple of columns (time, secondarys, <S>_INC..., <S>_PREV...)"
function _collectdf(inc::AbstractMatrix, prev::AbstractMatrix)          # [12, T]
    T = size(inc, 2)
    sec = [Float64(inc[2, t]) / Float64(sum(prev[3:8, t])) for t in 1:T]   # LAT_INC ./ sum(PRE..IISO prevalence) :237-239
    sec[isnan.(sec)] .= 0; sec[isinf.(sec)] .= 0                            # :240-241
    cols = Pair{Symbol, Any}[:time => collect(1:T), :secondarys => sec]
    for (s, n) in enumerate(STATES); push!(cols, Symbol(n, "_INC") => Int64.(inc[s, :])); end
    for (s, n) in enumerate(STATES); push!(cols, Symbol(n, "_PREV") => Int64.(prev[s, :])); end
    return (; cols...)
end

"runsim(simnum, ip) :77-101: seed simnum*726 (:78), main, and the result tuple (a, g1..g5, infectors, ct_numbers)"
function runsim(simnum, ip::ModelParameters; device = 0)
    seed!(simnum * 726)
    main(ip; device = device)
    L = LAST[]
    grp(k) = merge(_collectdf(view(L.inc, :, :, k), view(L.prev, :, :, k)), (sim = fill(simnum, ip.modeltime),))
    ct_numbers = (0, 0, Int(L.totalisolated), 0, 0, 0, 0)                   # :86-87 — only totalisolated is ever written (:649)
    return (a = grp(1), g1 = grp(2), g2 = grp(3), g3 = grp(4), g4 = grp(5), g5 = grp(6),
            infectors = Tuple(Int.(L.infectors)), ct_numbers = ct_numbers)
end

"pmap(x -> runsim(x, ip), 1:nsims) (scripts/local_run.jl:28-30) as one device run: curves [12, T, 6, nsims]"
function run_ensemble(ip::ModelParameters, nsims::Integer; device = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cabm_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint), h, device, nsims, ip.popsize, ip.modeltime, 0)
    rc == 0 || error(unsafe_string(ccall((:cabm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    try
        p = Ref(to_c(ip)); seeds = UInt64[726 * x for x in 1:nsims]                               # :78
        check(h[], ccall((:cabm_configure, LIB), Cint, (Ptr{Cvoid}, Ref{CabmParams}, Cint, Ptr{Int32}, Cint, Ptr{UInt64}, Cint),
                         h[], p, 1, C_NULL, nsims, seeds, 0))
        check(h[], ccall((:cabm_run, LIB), Cint, (Ptr{Cvoid},), h[]))
        inc = Array{Int32}(undef, 12, ip.modeltime, 6, nsims); prev = similar(inc)                # C order [R][6][T][12]
        check(h[], ccall((:cabm_get_curves, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), h[], inc, prev))
        infectors = Matrix{Int64}(undef, 4, nsims); tiso = Vector{Int64}(undef, nsims)
        check(h[], ccall((:cabm_get_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt32}), h[], infectors, tiso, C_NULL, C_NULL))
        return (inc = inc, prev = prev, infectors = infectors, totalisolated = tiso)
    finally
        ccall((:cabm_destroy, LIB), Cvoid, (Ptr{Cvoid},), h[])
    end
end

"""
Ensembles larger than one batch: `handles` handles of `batch` replicates on one GPU, used round-robin. Each handle has its own
stream, so consecutive batches overlap on the device (the short replicates of batch k+1 fill the SMs that the long epidemics of
batch k leave idle), and the kernels write every finished replicate's Int16 curves straight into one pinned host array for the
whole ensemble (cabm_set_host_curves_i16): the host copies nothing. Curves come back as [12, T, 6, nsims].
"""
function run_ensemble_pipelined(ip::ModelParameters, nsims::Integer; batch = 1000, handles = 3, device = 0)
    nsims % batch == 0 || error("nsims must be a multiple of batch")
    per = 12 * ip.modeltime * 6                                                    # Int16 elements per replicate
    bytes = UInt64(2 * per * nsims)
    pinc = ccall((:cabm_host_alloc, LIB), Ptr{Int16}, (UInt64,), bytes); pprev = ccall((:cabm_host_alloc, LIB), Ptr{Int16}, (UInt64,), bytes)
    (pinc == C_NULL || pprev == C_NULL) && error("cabm_host_alloc failed")
    hs = Ptr{Cvoid}[]
    infectors = Matrix{Int64}(undef, 4, nsims); tiso = Vector{Int64}(undef, nsims)
    try
        for _ in 1:handles
            h = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:cabm_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint), h, device, batch, ip.popsize, ip.modeltime, 0)
            rc == 0 || error(unsafe_string(ccall((:cabm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
            push!(hs, h[])
        end
        p = Ref(to_c(ip))
        inflight = fill(0, handles)                                                 # batch number running on each handle
        collect!(j) = begin                                                          # wait for the batch on handle j, read its scalars
            k = inflight[j]; k == 0 && return
            r0 = (k - 1) * batch
            check(hs[j], ccall((:cabm_sync, LIB), Cint, (Ptr{Cvoid},), hs[j]))
            check(hs[j], ccall((:cabm_get_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt32}),
                               hs[j], pointer(infectors, 4 * r0 + 1), pointer(tiso, r0 + 1), C_NULL, C_NULL))
            inflight[j] = 0
        end
        for k in 1:(nsims ÷ batch)
            j = (k - 1) % handles + 1
            collect!(j)
            r0 = (k - 1) * batch
            seeds = UInt64[726 * x for x in (r0 + 1):(r0 + batch)]                  # :78
            check(hs[j], ccall((:cabm_set_host_curves_i16, LIB), Cint, (Ptr{Cvoid}, Ptr{Int16}, Ptr{Int16}), hs[j], pinc + 2 * per * r0, pprev + 2 * per * r0))
            check(hs[j], ccall((:cabm_configure, LIB), Cint, (Ptr{Cvoid}, Ref{CabmParams}, Cint, Ptr{Int32}, Cint, Ptr{UInt64}, Cint),
                               hs[j], p, 1, C_NULL, batch, seeds, 0))
            check(hs[j], ccall((:cabm_run, LIB), Cint, (Ptr{Cvoid},), hs[j]))         # asynchronous on the handle's stream
            inflight[j] = k
        end
        foreach(collect!, 1:handles)
        inc = copy(unsafe_wrap(Array, pinc, (12, ip.modeltime, 6, nsims))); prev = copy(unsafe_wrap(Array, pprev, (12, ip.modeltime, 6, nsims)))
        return (inc = inc, prev = prev, infectors = infectors, totalisolated = tiso)
    finally
        foreach(h -> ccall((:cabm_destroy, LIB), Cvoid, (Ptr{Cvoid},), h), hs)
        ccall((:cabm_host_free, LIB), Cvoid, (Ptr{Cvoid},), pinc); ccall((:cabm_host_free, LIB), Cvoid, (Ptr{Cvoid},), pprev)
    end
end

export ModelParameters, HEALTH, main, runsim, seed!, run_ensemble, run_ensemble_pipelined
end
