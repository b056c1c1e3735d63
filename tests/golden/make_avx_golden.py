"""Generates tests/golden/avx_golden.json: outputs of the REFERENCE's own compiled AVX kernels
(pkg/distance/simd/cpp/avx.cpp, built unmodified into oracle/_ref/libcoltt_ref_avx.so by oracle/Makefile) on seeded
inputs — dot, norm_a*norm_b and the squared L2 distance as fp32 bit patterns.  Run in the authoring container, where
/root/reference exists:  python tests/golden/make_avx_golden.py
The inputs are regenerated from the seed by the test, so only the outputs are stored."""
import ctypes as C
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

SEED = 0xC0177
DIMS = [1, 7, 8, 9, 17, 64, 100, 128, 384, 768, 770, 1536]


def inputs(d, trial):
    g = np.random.Generator(np.random.Philox(SEED + 1000 * d + trial))
    a, b = orc.aligned_f32(d), orc.aligned_f32(d)
    if trial % 2 == 0:
        a[:] = g.random(d, dtype=np.float32)                       # the reference's own test data: rand.Float32()
        b[:] = g.random(d, dtype=np.float32)
    else:
        a[:] = g.standard_normal(d).astype(np.float32) * 3
        b[:] = g.standard_normal(d).astype(np.float32) * 3
    return a, b


def bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def main():
    R = orc.ref_lib()
    assert R is not None, "oracle/_ref is not built: run `make -C oracle` where /root/reference exists"
    out = {"source": "pkg/distance/simd/cpp/avx.cpp compiled unmodified (oracle/ref_avx_wrapper.cpp, g++ -O2 -mavx2 -ffp-contract=off)",
           "seed": SEED, "cases": []}
    for d in DIMS:
        for trial in range(4):
            a, b = inputs(d, trial)
            ap, bp = a.ctypes.data_as(orc.f32p), b.ctypes.data_as(orc.f32p)
            dot, n2, l2 = C.c_float(), C.c_float(), C.c_float()
            R.ref_cosine_similarity_dot_norm(d, ap, bp, C.byref(dot), C.byref(n2))
            R.ref_euclidean_distance_squared(d, ap, bp, C.byref(l2))
            out["cases"].append({"dim": d, "trial": trial, "dot": bits(dot.value), "norm2": bits(n2.value), "l2sq": bits(l2.value)})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "avx_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
