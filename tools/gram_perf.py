#!/usr/bin/env python3
"""Times snk_gram variants at BASELINE config 5a size (K=1000 snapshots x P=181,395 weights)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402


def executed_flops(S, K, P, terms):
    """the tiles snk_gram computes (upper-triangle tiles of the padded problem) x products per k-step, from the library's own plan"""
    import ctypes as C
    v = C.c_double(0)
    S._check(S.lib().snk_gram_block_flops(K, K, P, terms, 1, C.byref(v)))
    return v.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=1000)
    ap.add_argument("--P", type=int, default=181395)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--variants", default="3:32:0,3:64:0,1:32:0,1:64:0")
    ap.add_argument("--cta-group", type=int, default=2)
    args = ap.parse_args()
    S = graft.load_package()
    S.lib().snk_gram_config(args.cta_group)
    dev = torch.device("cuda", 0)
    K, P = args.K, args.P
    g = torch.Generator(device=dev); g.manual_seed(0)
    A = torch.randn(K, P, device=dev, dtype=torch.float32, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    for v in args.variants.split(","):
        terms, bk, splits = (int(x) for x in v.split(":"))
        plan = S.GramPlan(K, P, dev, splits=splits).pack(A)
        G = torch.empty(K, K, dtype=torch.float32, device=dev)
        for _ in range(3):
            plan.gram(terms, bk, out=G)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()                                   # flush L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.gram(terms, bk, out=G); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        useful = 2.0 * K * K * P
        res.append({"terms": terms, "block_k": bk, "splits": splits, "ms_median": ms, "ms_min": ts[0],
                    "useful_tflops": useful / (ms * 1e-3) / 1e12,
                    "mma_tflops": executed_flops(S, K, P, terms) / (ms * 1e-3) / 1e12})
        print(json.dumps(res[-1]), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Ab = A.to(torch.bfloat16)
    for _ in range(3):
        torch.matmul(Ab, Ab.T)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.matmul(Ab, Ab.T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"cublas_bf16_matmul_ms": ms, "tflops": 2.0 * K * K * P / (ms * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
