# cpu_baseline.jl — the reference package's own CPU figure for bench.py's metric, for any machine that has Julia.
#
# SURVEY.md §8(d) "CPU baseline beside it", item (1): time GenericSchur.gschur! (src/GenericSchur.jl:350-372
# complex, :805-835 real) over the batch with Threads.@threads and BLAS.set_num_threads(1), and print one
# JSON line shaped like `bench.py --impl reference` prints.
#
# NOTE: the build environment and the GPU boxes have no Julia, so this file has never been executed; the
# numbers bench.py reports come from the C++ port under oracle/ (cpu_baseline.kind = "port").  When this
# script is run next to a bench.py run the two lines can be compared directly: same workload shapes, same
# metric and unit.  (The entries are uniform [0,1) like bench.py's, but from Julia's own generator.)
#
#   julia -t auto scripts/cpu_baseline.jl [cfg2|cfg3|f64n64] [sample] [steps]
using LinearAlgebra, Random, Printf
import GenericSchur

const WORKLOADS = Dict(
    "cfg2" => (Float64, 32, 16384),
    "cfg3" => (ComplexF64, 64, 65536),
    "f64n64" => (Float64, 64, 65536),
)

function step(A::Array{T, 3}) where {T}
    nb = size(A, 3)
    acc = zeros(Float64, Threads.nthreads() * 8)      # keeps the results alive; padded against false sharing
    t = @elapsed Threads.@threads :static for b in 1:nb
        S = GenericSchur.gschur!(A[:, :, b])              # the slice is a fresh copy; gschur! destroys it
        acc[(Threads.threadid() - 1) * 8 + 1] += abs(S.values[1])
    end
    return t, sum(acc)
end

function main()
    name = length(ARGS) >= 1 ? ARGS[1] : "cfg3"
    T, n, batch = WORKLOADS[name]
    sample = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : min(batch, 256 * Threads.nthreads())
    steps = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 3
    BLAS.set_num_threads(1)
    Random.seed!(1234)                                 # test/complex.jl:86 seeds 1234
    A = rand(T, n, n, sample)
    step(A)                                            # warm-up (compilation)
    times = [step(A)[1] for _ in 1:steps]
    tm = sum(times) / steps
    @printf("{\"impl\": \"reference\", \"metric\": \"batched n=%d Schur matrices/s\", \"value\": %.3f, ", n, sample / tm)
    @printf("\"unit\": \"matrices/s\", \"steps\": %d, \"ms_per_step\": %.3f, \"higher_is_better\": true, ", steps, 1e3 * tm)
    @printf("\"dtype\": \"f64\", \"data\": \"synthetic\", \"config\": {\"workload\": \"%s\", \"n\": %d, \"element\": \"%s\", ", name, n, string(T))
    @printf("\"per_gpu_batch\": %d, \"sample_per_step\": %d}, ", batch, sample)
    @printf("\"cpu_baseline\": {\"value\": %.3f, \"unit\": \"matrices/s\", \"cores\": %d, \"kind\": \"reference\", ", sample / tm, Threads.nthreads())
    @printf("\"sample\": \"each step = %d matrices, GenericSchur.gschur! under Threads.@threads\"}}\n", sample)
end

main()
