"""Regenerate tests/golden/mt19937.json from the REFERENCE's own Core/Math/MersenneTwister.h.

Run in the authoring container (needs /root/reference and oracle/_ref built by `make -C oracle ref`):
    python tests/golden/make_mt_golden.py
The fixture pins the oracle's (and the CUDA library's) MT19937 restatement wherever the reference tree is absent.
"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_mt.so"))

out = {"source": "Core/Math/MersenneTwister.h via oracle/ref_mt_shim.cpp", "seeds": {}}
for seed in (1, 2, 42, 124, 5489, 12345, 100 * 123 + 1, 4294967295):
    n = 1300  # crosses two reload() boundaries (624, 1248)
    u = np.empty(n, np.uint32)
    ref.ref_mt_u32(C.c_uint32(seed), n, u.ctypes.data_as(C.c_void_p))
    r = np.empty(8)
    ref.ref_mt_rand(C.c_uint32(seed), 8, r.ctypes.data_as(C.c_void_p))
    e = np.empty(8)
    ref.ref_mt_randDblExc(C.c_uint32(seed), 8, e.ctypes.data_as(C.c_void_p))
    out["seeds"][str(seed)] = {
        "first8_u32": [int(x) for x in u[:8]],
        "u32_at_623_624_625": [int(x) for x in u[623:626]],
        "u32_at_1247_1248_1299": [int(u[1247]), int(u[1248]), int(u[1299])],
        "xor_all_1300": int(np.bitwise_xor.reduce(u)),
        "rand_first8_hex": [float(x).hex() for x in r],
        "randDblExc_first8_hex": [float(x).hex() for x in e],
    }
with open(os.path.join(HERE, "mt19937.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", os.path.join(HERE, "mt19937.json"))
