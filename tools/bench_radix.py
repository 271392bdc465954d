"""Radix-sort pass timing on device (per-kernel CUDA events from the library's
profile hooks) on uniform random keys.  Usage: python tools/bench_radix.py [n] [bit_hi] [pairs]
(The look-back / match.any variants this tool compared during round 1 are
recorded in profiles/bench_radix_variants_r01.jsonl; the library now holds only
the per-tile offset-table version.)"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmr_b200


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
    hi = int(sys.argv[2]) if len(sys.argv) > 2 else 33
    pairs = len(sys.argv) > 3 and sys.argv[3] == "pairs"
    lib = tmr_b200.require_gpu()
    lib.tmr_b200_context.restype = ctypes.c_void_p
    ctx = ctypes.c_void_p(lib.tmr_b200_context())
    lib.tmrgpu_test_radix_sort.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    lib.tmrgpu_profile_enable.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.tmrgpu_profile_reset.argtypes = [ctypes.c_void_p]
    lib.tmrgpu_profile_json.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 1 << hi, n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32) if pairs else None
    lib.tmrgpu_profile_enable(ctx, 1)
    for rep in range(3):
        k = keys.copy()
        v = vals.copy() if pairs else None
        lib.tmrgpu_profile_reset(ctx)
        rc = lib.tmrgpu_test_radix_sort(ctx, k.ctypes.data, v.ctypes.data if pairs else None, n, 0, hi)
        assert rc == 0
        buf = ctypes.create_string_buffer(1 << 16)
        lib.tmrgpu_profile_json(ctx, buf, len(buf))
        prof = json.loads(buf.value.decode())
    ok = bool(np.all(k[1:] >= k[:-1]))
    if pairs:
        ok = ok and bool(np.array_equal(keys[v], k))
    bytes_per = 16 + (8 if pairs else 0)
    out = {"n": n, "bits": hi,
           "pairs": pairs, "sorted": ok, "kernels": prof}
    for name, st in (prof.items() if isinstance(prof, dict) else []):
        if "radix_pass" in name:
            ms = st["ms"] / st["launches"]
            out["ms_per_pass"] = ms
            out["GBps"] = n * bytes_per / ms / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
