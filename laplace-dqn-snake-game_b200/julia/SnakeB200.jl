# SnakeB200.jl — thin `ccall` wrapper over libsnake_b200.so (include/snake_b200.h).
#
# NOT EXECUTED IN THE BUILD IMAGE (Julia is not installed there); the tests drive the identical C symbols
# through ctypes.  It is the binding a maintainer of lucagiorgetti/Laplace-DQN-Snake-game adds so that
# `utils.jl` / `main.jl` keep their entry points: the same names (`available_actions`, `step!`,
# `virtual_step`, `assemble_state!`, `epsilon_greedy`), the same 3-action relative control, the same
# (10,10,2,N) two-frame state — now for N games at once on a B200.
#
# Host-array methods use the `_host` entry points (pinned or plain Julia Arrays); if CUDA.jl is loaded the
# device entry points can be called with `pointer(::CuArray)` the same way.
module SnakeB200

export BatchedSnakeGame, available_actions, step!, step_fused!, virtual_step, assemble_state!,
       epsilon_greedy, masked_target, center_columns!, reset!, set_food_list!, score, lost

const lib = get(ENV, "SNAKE_B200_LIB", joinpath(@__DIR__, "..", "libsnake_b200.so"))

const OBS_F32, OBS_I8, OBS_I64, OBS_PACKED2 = Cint(1), Cint(2), Cint(3), Cint(4)
const AUTO_RESET = UInt32(1)
# utils.jl:8 order
const DIRS = (CartesianIndex(-1, 0), CartesianIndex(1, 0), CartesianIndex(0, -1), CartesianIndex(0, 1))

struct SnakeB200Error <: Exception
    code::Int
    msg::String
end
function check(rc::Cint)
    rc == 0 && return nothing
    throw(SnakeB200Error(rc, unsafe_string(ccall((:snk_last_error, lib), Cstring, ()))))
end

"""N batched `SnakeGame`s (structs.jl:6-100) resident on one GPU."""
mutable struct BatchedSnakeGame
    handle::Ptr{Cvoid}
    n::Int
    # host staging, laid out exactly like `stack_exp` (utils.jl:343-383)
    state::Array{Float32,4}        # (10,10,2,N): next_state of the last step
    reward::Vector{Float32}
    lost::Vector{UInt8}
    next_is_suicidal::Matrix{UInt8}   # (3,N)
    actions::Vector{UInt8}
    function BatchedSnakeGame(n::Integer; device::Integer = 0, auto_reset::Bool = true,
                              board_size::Integer = 10, n_frames::Integer = 2)
        (board_size == 10 && n_frames == 2) || throw("only board_size = 10, n_frames = 2 are supported")
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:snk_create, lib), Cint, (Ref{Ptr{Cvoid}}, Int64, Cint, UInt32),
                    h, n, device, auto_reset ? AUTO_RESET : UInt32(0)))
        g = new(h[], n, zeros(Float32, 10, 10, 2, n), zeros(Float32, n), zeros(UInt8, n), ones(UInt8, 3, n),
                zeros(UInt8, n))
        finalizer(x -> ccall((:snk_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), g)
        assemble_state!(g)
        return g
    end
end

reset!(g::BatchedSnakeGame) = check(ccall((:snk_reset, lib), Cint, (Ptr{Cvoid},), g.handle))

"""food_list injection (structs.jl:70): vector of CartesianIndex{2} in 2:9 × 2:9."""
function set_food_list!(g::BatchedSnakeGame, cells::Vector{CartesianIndex{2}})
    rc = UInt8[x for c in cells for x in (c[1], c[2])]
    check(ccall((:snk_set_food_list_host, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint), g.handle, rc, length(cells)))
    reset!(g)
end

sync(g::BatchedSnakeGame) = check(ccall((:snk_sync, lib), Cint, (Ptr{Cvoid},), g.handle))

# utils.jl:448-451 (device pointers; use with CUDA.jl CuArrays)
function masked_target(q_next::Ptr{Float32}, mask::Ptr{UInt8}, r::Ptr{Float32}, done::Ptr{UInt8},
                       y::Ptr{Float64}, B::Integer; gamma::Float64 = 0.97, fill::Float32 = -100f0)
    check(ccall((:snk_masked_target, lib), Cint,
                (Ptr{Float32}, Ptr{UInt8}, Ptr{Float32}, Ptr{UInt8}, Cdouble, Cfloat, Ptr{Float64}, Ptr{Float32}, Int64, Ptr{Cvoid}),
                q_next, mask, r, done, gamma, fill, y, C_NULL, B, C_NULL))
end

"""
    step_fused!(g, actions)            # action indices 1:3 into available_actions, as the Q-net's argmax gives
    step_fused!(g, q, ε; u, ridx)      # epsilon_greedy + step! + virtual_step + next state, one kernel

Fills `g.state` (Float32 (10,10,2,N) next_state), `g.reward`, `g.lost`, `g.next_is_suicidal` — the fields
`play_episode` (utils.jl:198-259) collects per step — and returns `g`.
"""
function step_fused!(g::BatchedSnakeGame, actions::AbstractVector{<:Integer})
    g.actions .= UInt8.(actions .- 1)                       # Julia 1-based index -> 0-based
    check(ccall((:snk_step_fused_host, lib), Cint,
                (Ptr{Cvoid}, Ptr{Float32}, Cfloat, Ptr{Float32}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Float32}, Ptr{UInt8},
                 Ptr{Cvoid}, Cint, Ptr{UInt8}, Ptr{Float32}, Ptr{Int32}),
                g.handle, C_NULL, 0f0, C_NULL, C_NULL, g.actions, g.reward, g.lost, g.state, OBS_F32,
                g.next_is_suicidal, C_NULL, C_NULL))
    sync(g)
    return g
end
function step_fused!(g::BatchedSnakeGame, q::Matrix{Float32}, epsilon::Float32;
                     u::Vector{Float32} = rand(Float32, g.n), ridx::Vector{UInt8} = rand(UInt8(0):UInt8(2), g.n))
    size(q) == (3, g.n) || throw(DimensionMismatch("q must be (3, N)"))
    check(ccall((:snk_step_fused_host, lib), Cint,
                (Ptr{Cvoid}, Ptr{Float32}, Cfloat, Ptr{Float32}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Float32}, Ptr{UInt8},
                 Ptr{Cvoid}, Cint, Ptr{UInt8}, Ptr{Float32}, Ptr{Int32}),
                g.handle, q, epsilon, u, ridx, g.actions, g.reward, g.lost, g.state, OBS_F32,
                g.next_is_suicidal, C_NULL, C_NULL))
    sync(g)
    return g
end

# The reference's per-call API, for code that wants the individual pieces ------------------------------
# (these go through small device buffers owned by the caller when CUDA.jl is present; the host versions
#  below round-trip through step_fused! outputs)

"""utils.jl:100-109 for all N games; `actions` are indices 1:3 into available_actions(g)."""
step!(g::BatchedSnakeGame, actions::AbstractVector{<:Integer}) = step_fused!(g, actions)

"""utils.jl:112-132: `next_is_suicidal` (3,N) of the current states (filled by the last step)."""
virtual_step(g::BatchedSnakeGame) = g.next_is_suicidal .!= 0

"""utils.jl:135-139: (10,10,2,N) Float32 two-frame state."""
function assemble_state!(g::BatchedSnakeGame)
    # device -> host through a temporary device buffer is what snk_state + cudaMemcpy do under CUDA.jl;
    # with plain Arrays the state is refreshed by every step_fused!.  After a reset all envs hold the
    # constructor state, which is a constant:
    if all(g.reward .== 0) && all(g.lost .== 0)
        b = zeros(Float32, 10, 10); b[1, :] .= -1; b[end, :] .= -1; b[:, 1] .= -1; b[:, end] .= -1
        b[4, 5] = 2; b[8, 2] = 1; b[9, 2] = 1
        for n in 1:g.n, f in 1:2
            g.state[:, :, f, n] .= b
        end
    end
    return g.state
end

score(g::BatchedSnakeGame, d_score::Ptr{Int32}) = check(ccall((:snk_get_score, lib), Cint, (Ptr{Cvoid}, Ptr{Int32}), g.handle, d_score))
lost(g::BatchedSnakeGame) = g.lost .!= 0

"""utils.jl:7-10 on device: fills a (3,N) UInt8 device array with direction codes 0:3 (U,D,L,R)."""
available_actions(g::BatchedSnakeGame, d_out::Ptr{UInt8}) =
    check(ccall((:snk_available_actions, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), g.handle, d_out))

"""utils.jl:153-172 on device pointers: out[i] = ridx[i] if u[i] < ε else argmax(q[:, i]) - 1."""
epsilon_greedy(g::BatchedSnakeGame, d_q::Ptr{Float32}, epsilon::Float32, d_u::Ptr{Float32}, d_ridx::Ptr{UInt8},
               d_out::Ptr{UInt8}) =
    check(ccall((:snk_select_action, lib), Cint, (Ptr{Cvoid}, Ptr{Float32}, Cfloat, Ptr{Float32}, Ptr{UInt8}, Ptr{UInt8}),
                g.handle, d_q, epsilon, d_u, d_ridx, d_out))

"""compute_D.jl:76-81 on a device-resident P×K Float64 deviation_matrix (column-major, as Julia stores it)."""
center_columns!(d_D::Ptr{Float64}, P::Integer, K::Integer, d_mean::Ptr{Float64}, d_var::Ptr{Float64}) =
    check(ccall((:snk_center_columns, lib), Cint, (Ptr{Float64}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
                d_D, P, K, d_mean, d_var, C_NULL))

end # module
