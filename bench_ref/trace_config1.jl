# trace_config1.jl — BASELINE config 1 trace from the UNMODIFIED reference (not executed in the build image: no Julia).
#   cd <checkout of lucagiorgetti/Laplace-DQN-Snake-game>; julia /path/to/trace_config1.jl > trace_julia.txt
# Compare with `python tools/trace_config1.py > trace_oracle.txt` (or --cuda): the files must be identical
# (sha256 of the expected file is pinned in tests/golden/g7_config1_trace.json).
include("imports.jl")
using Printf

splitmix64(x::UInt64) = begin
    x += 0x9E3779B97F4A7C15
    x = (x ⊻ (x >> 30)) * 0xBF58476D1CE4E5B9
    x = (x ⊻ (x >> 27)) * 0x94D049BB133111EB
    x ⊻ (x >> 31)
end
action_index(t::Int; seed = UInt64(42), env = UInt64(0)) =
    Int(splitmix64(seed ⊻ splitmix64(env * 0x100000001B3 + UInt64(t))) % 3) + 1       # t is 0-based

function fnv1a(board::Matrix{Int})
    h = 0xCBF29CE484222325
    for v in vec(board)                       # column-major, like the (10,10,2,N) state
        h ⊻= UInt64((v + 1) & 0xFF)
        h *= 0x100000001B3
    end
    h
end
dircode(a) = a == CartesianIndex(-1, 0) ? 0 : a == CartesianIndex(1, 0) ? 1 : a == CartesianIndex(0, -1) ? 2 : 3

function main(steps = 10_000)
    game = SnakeGame()
    for t in 0:steps-1
        av = available_actions(game)
        a = av[action_index(t)]
        step!(game, a)
        @printf "%d %d %08x %d %d %016x\n" (t + 1) dircode(a) reinterpret(UInt32, game.reward) Int(game.lost) game.score fnv1a(game.board)
        game.lost && (game = SnakeGame())
    end
end
main()
