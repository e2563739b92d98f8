"""Oracle (test infrastructure): the reference's ReplayBuffer restated in plain Python.

structs.jl:104-116 (ReplayBuffer), utils.jl:265-277 (length, store!), :289-296 (isfull, isready),
:311-314 (empty_buffer!), :343-383 (stack_exp).  Transitions are dicts; positions are 1-based as in Julia.
"""
import numpy as np


class ReplayOracle:
    def __init__(self, capacity=50000, batch_size=64):
        if batch_size > capacity:
            raise ValueError("batch_size cannot be greater than the capacity of the buffer.")   # structs.jl:113
        self.capacity, self.position, self.buffer, self.batch_size = capacity, 1, [], batch_size

    def __len__(self):                                   # utils.jl:265
        return len(self.buffer)

    def store(self, exp):                                # utils.jl:267-277
        if len(self) < self.capacity:
            self.buffer.append(exp)
        else:
            self.buffer[self.position - 1] = exp
            self.position += 1
        if self.position > self.capacity:
            self.position = 1

    def isfull(self):                                    # utils.jl:289-291
        return self.position == self.capacity

    def isready(self):                                   # utils.jl:293-296
        return len(self) >= self.batch_size

    def empty(self):                                     # utils.jl:311-314
        self.buffer, self.position = [], 1

    def stack_exp(self, idx0):                           # utils.jl:343-383, idx0 = 0-based slots
        b = [self.buffer[i] for i in idx0]
        return {
            "states": np.stack([e["state"] for e in b]).astype(np.float32),          # Float32.(s)
            "next_states": np.stack([e["next_state"] for e in b]).astype(np.float32),
            "actions": np.array([e["action_idx"] for e in b], np.uint8),             # findfirst(a, av_acts) - 1
            "rewards": np.array([e["reward"] for e in b], np.float32),
            "dones": np.array([e["done"] for e in b], np.uint8),
            "mask": np.stack([e["mask"] for e in b]).astype(np.uint8),
        }
