# reference_cpu.jl — times the UNMODIFIED reference environment on the CPU (not executed in the build image:
# Julia is not installed there; shipped so that anyone with Julia 1.10 + the reference's packages can produce the
# true Julia number that bench.py's C restatement stands in for).
#
#   cd <checkout of lucagiorgetti/Laplace-DQN-Snake-game>;  julia --threads=auto /path/to/reference_cpu.jl
#
# Workload = BASELINE config 1/2 shape: uniform random actions over the 3 available moves, a new game on loss,
# per step: step! + virtual_step + assemble_state! + Float32 cast (what one env-step of the batched kernel emits).
include("imports.jl")
using Random, Printf

function run_envs(n_envs::Int, n_steps::Int; seed = 42)
    rng = Xoshiro(seed)
    games = [SnakeGame() for _ in 1:n_envs]
    sink = 0.0f0
    for _ in 1:n_steps, i in 1:n_envs
        g = games[i]
        av = available_actions(g)
        step!(g, av[rand(rng, 1:3)])
        virtual_step(g, missing_model)            # next_is_suicidal (utils.jl:112-132)
        if !g.lost
            assemble_state!(g)
            sink += sum(Float32.(g.state))        # stack_exp's cast (utils.jl:361-362)
        else
            games[i] = SnakeGame()
        end
    end
    return sink
end

# virtual_step only uses the model argument for dispatch
const missing_model = Chain(identity)

function main()
    run_envs(64, 10)                                   # compile
    for (n, steps) in ((1, 10_000), (4096, 50))
        t = @elapsed run_envs(n, steps)
        @printf "single thread: %d envs x %d steps: %.3e env-steps/s\n" n steps n * steps / t
    end
    nt = Threads.nthreads()
    n, steps = 4096 * nt, 50
    t = @elapsed begin
        Threads.@threads for k in 1:nt
            run_envs(4096, steps; seed = 42 + k)
        end
    end
    @printf "%d threads: %d envs x %d steps: %.3e env-steps/s\n" nt n steps n * steps / t
end

main()
