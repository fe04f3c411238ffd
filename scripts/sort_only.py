#!/usr/bin/env python
"""The radix sort alone (profiling runs): 67.1 M Float64 keys + 32-bit payload through `ocb_sort_pairs`, CUDA-event timed.
`python scripts/sort_only.py [n] [randn|strat]`"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ocb_loader  # noqa: E402

ocb = ocb_loader.load()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512 * 512 * 256
    kind = sys.argv[2] if len(sys.argv) > 2 else "randn"
    dev = torch.device("cuda", 0)
    keys = torch.randn(n, device=dev, dtype=torch.float64)
    if kind == "strat":
        keys = 1e-5 * (torch.rand(n, device=dev, dtype=torch.float64) - 1) * 0.5 + 0.05e-5 * keys
    out = torch.empty_like(keys)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    L = ocb._lib.lib()
    st = torch.cuda.current_stream().cuda_stream

    def go():
        ocb._lib.check(L.ocb_sort_pairs(1, keys.data_ptr(), out.data_ptr(), perm.data_ptr(), n, st))
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        go()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ok = bool((out[1:] >= out[:-1]).all()) and bool((keys[perm.long()] == out).all())
    print(json.dumps({"name": f"radix sort {kind}", "n": n, "ms": round(ms, 3), "Gkeys_per_s": round(n / ms / 1e6, 2), "sorted_and_permutation_ok": ok}))


if __name__ == "__main__":
    main()
