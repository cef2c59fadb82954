# pmap_reference.jl — the named CPU baseline of BASELINE.json: the unmodified covid19abm.jl, one worker per core,
# `pmap(runsim)` exactly as scripts/local_run.jl:28-30 does it.
#
# UNVERIFIED: there is no `julia` binary in the build image or on the GPU box, so this script has never been run
# there; `bench.py --impl reference` times the C restatement in oracle/ instead and says so in its JSON line.
# For anyone with Julia and a checkout of affans/covid19abm.jl:
#
#     julia --project=/path/to/covid19abm.jl baseline/pmap_reference.jl [replicates] [workers]
#
# prints one JSON line with the same metric/unit/config as bench.py (agent-days/s, config B).
using Distributed

nrep = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 64
nwrk = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : Sys.CPU_THREADS
addprocs(nwrk; exeflags = "--project=$(Base.active_project())")
@everywhere using covid19abm
@everywhere const cv = covid19abm

# config B (SURVEY.md §8d): Ontario, beta = 0.0345, isolation of mild / severe / presymptomatic cases switched on
myp = cv.ModelParameters(β = 0.0345, prov = :ontario, popsize = 10000, modeltime = 500, initialinf = 1,
                         fmild = 0.5, τmild = 1, fsevere = 1.0, fpreiso = 0.3, tpreiso = 0)

pmap(x -> cv.runsim(x, myp), 1:nwrk)                     # warm-up: compile on every worker
t = @elapsed pmap(x -> cv.runsim(x, myp), 1:nrep)        # replicate x seeds with x * 726 (src/covid19abm.jl:78)
v = nrep * myp.popsize * myp.modeltime / t
println("{\"impl\": \"reference\", \"metric\": \"agent_days_per_sec\", \"value\": $v, \"unit\": \"agent-days/s\", ",
        "\"cpu_baseline\": {\"value\": $v, \"unit\": \"agent-days/s\", \"cores\": $nwrk, \"kind\": \"reference\", ",
        "\"sample\": \"$nrep replicates of config B through pmap(runsim), $(round(t, digits = 2)) s wall\"}}")
