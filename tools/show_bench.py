#!/usr/bin/env python3
"""Prints the interesting numbers of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g env-steps/s  %.4f ms/step  hbm frac %.3f  n_gpus %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["n_gpus"]))
e = d["e2e"]
print("e2e %.4g (%.3f ms/step)" % (e["value"], e["ms_per_step"]), {k: "%.4g" % v["value"] for k, v in e.items() if isinstance(v, dict) and "value" in v})
print("clocks", d.get("clocks"))
for k, v in d.get("obs_format_variants", {}).items():
    if isinstance(v, dict):
        print(" variant %s: %.4g env-steps/s, hbm frac %.3f" % (k, v["value"], v["hbm_frac"]))
c2 = d.get("config2_4096_envs")
if c2:
    print("config2 %.4g env-steps/s (%.3f us/step, hbm %.3f); graph of launches %.4g" % (c2["value"], c2["ms_per_step"] * 1e3, c2["hbm_frac"], c2["per_step_launches_in_a_cuda_graph"]["value"]))
c4 = d.get("config4_65536_envs")
if c4:
    for k, v in c4.items():
        if isinstance(v, dict):
            print(" config4 %-20s %.4g env-steps/s, fwd %.3f ms, %.0f TFLOP/s useful, err %.2g" % (k, v["env_steps_per_s"], v["qnet_forward_ms"], v["qnet_useful_tflops"], v["max_err_vs_float64_of_maxQ"]))
g = d.get("gram_5a")
if g:
    print("gram5a pack %.3f ms (%.2f hbm) pack_f32 %.3f (%.2f) center %.3f (%.2f)" % (g["pack_ms"], g["pack_hbm_frac"], g["pack_f32_ms"], g["pack_f32_hbm_frac"], g["center_ms"], g["center_hbm_frac"]))
    for t in ("terms3", "terms1"):
        r = g[t]
        print("  %s: %.3f ms, mma %.0f TF (%.2f), algorithmic %.2f, err %.2g, traffic %s" % (t, r["ms"], r["mma_tflops"], r["roofline"]["frac_executed"], r["roofline"]["frac"], r["rel_fro_err_vs_fp64"], r["roofline"]["traffic"]))
gs = d.get("gram_5b_sharded")
if gs:
    print("gram5b K=%d: total %.1f ms (producer %.1f, gram %.1f), mma/gpu %.0f TF (%.2f of sustained), nccl baseline %.1f ms, same bits %s, verify %s"
          % (gs["K_total"], gs["ms"], gs["producer_ms"], gs["gram_ms"], gs["mma_tflops_per_gpu"], gs["roofline"]["frac_executed"], gs["nccl_allgather_baseline_ms"],
             gs["nccl_allgather_same_bits_rank0"], gs["verify_per_rank"]["max_err_over_sqrt_GiiGjj"]))
cb = d.get("cpu_baseline")
if cb:
    print("cpu baseline %.4g on %d cores (%s)" % (cb["value"], cb["cores"], cb["kind"]))
