"""Oracle (test infrastructure): per-sample gradients of the reference's DQN loss in Float64 (torch autograd on the CPU).

Reference: utils.jl:452-466 — `q_pred = q_net(states); q_sel = q_pred[a_i, i]; Flux.huber_loss(q_sel, q_target)` (delta = 1)
through the network of structs.jl:127-139.  Row i = d huber(q_net(s_i)[a_i], y_i) / d theta with theta in Flux.destructure
order (per layer weight then bias, column-major), i.e. the per-sample term of the batch loss without its 1/B.
The reference never forms per-sample gradients (Zygote returns their mean) and Flux/Zygote are unpinned third-party
dependencies: parity unpinned; this is the SURVEY 8(c) oracle ("fp64 torch per-sample grad, conv weights flipped").
The leaves are kept in Flux's own array shapes and turned into torch's cross-correlation weights by differentiable flips /
permutes, so autograd returns the gradients directly in Flux layout; the forward pass built this way is the one
tests/test_bson_qnet.py checks against the numpy restatement of Flux/NNlib (oracle/qnet_oracle.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _colmajor(t):
    return t.permute(*reversed(range(t.dim()))).reshape(-1)


def per_sample_grads(layers, states, actions, targets):
    """states (B,2,10,10) [= Julia (10,10,2,B)], actions (B) 0-based, targets (B).  Returns (J (B, P) float64, loss (B), q (B,3))."""
    leaves, plan = [], []
    for kind, p in layers:
        if kind in ("conv", "dense"):
            W = torch.tensor(np.asarray(p["W"], dtype=np.float64), requires_grad=True)
            b = torch.tensor(np.asarray(p["b"], dtype=np.float64), requires_grad=True)
            leaves += [W, b]
            plan.append((kind, W, b, int(p["pad"][0]) if kind == "conv" else None))
        else:
            plan.append((kind, None, None, None))
    n_dense = sum(1 for k, *_ in plan if k == "dense")
    states = torch.as_tensor(np.asarray(states, dtype=np.float64))
    rows, losses, qs = [], [], []
    for i in range(states.shape[0]):
        x = states[i:i + 1]
        seen = 0
        for kind, W, b, pad in plan:
            if kind == "conv":
                wt = W.flip(0, 1).permute(3, 2, 1, 0)              # wt[o,c,kh,kw] = W[K1-1-kw, K2-1-kh, c, o]
                x = F.relu(F.conv2d(x, wt, b, padding=pad))
            elif kind == "flatten":
                x = x.flatten(1)                                    # c*25 + d2*5 + d1 = Flux's x + 5 y + 25 c
            else:
                seen += 1
                x = x @ W.T + b
                if seen < n_dense:
                    x = F.relu(x)
        q = x[0]
        d = q[int(actions[i])] - float(targets[i])
        loss = 0.5 * d * d if abs(float(d.detach())) < 1.0 else d.abs() - 0.5     # Flux.huber_loss: quadratic strictly inside delta
        g = torch.autograd.grad(loss, leaves)
        rows.append(torch.cat([_colmajor(t) for t in g]).numpy())
        losses.append(float(loss.detach()))
        qs.append(q.detach().numpy())
    return np.stack(rows), np.array(losses), np.stack(qs)
