"""
gen_config2_kat.py -- known answer for BASELINE configs[1]: the (k,m) = (10,1) mismatch Gram over all 9000 challenge
sequences (tests/golden/dna9000.npz: Xtr0,Xtr1,Xtr2,Xte0,Xte1,Xte2 in file order), normalised as get_mismatch_K does
(kernels.py:196-217).  *** TEST INFRASTRUCTURE ***

The reference itself cannot run this configuration (300-420 s and 8.4 MB of phi PER SEQUENCE), so the
answer comes from the plain-C oracle (oracle/kmg_oracle.c orc_mismatch_raw_block: the literal double loop over window
pairs with the neighbourhood table T, validated bit for bit against the reference for k <= 6 in tests/test_oracle_c.py),
upper triangle only, mirrored, then oracle_np.normalize_K (the reference's normalize_K, kernels.py:398-415).
~10 minutes on 8 threads.  Writes tests/golden/config2_mm10_kat.json: SHA-256 of the float64 matrix, its sum, trace and a
few entries.  tests/test_gpu_properties.py holds the GPU result to it.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_c as oc  # noqa: E402
import oracle_np as onp  # noqa: E402


def main():
    oc.build()
    codes = np.load(os.path.join(HERE, "..", "tests", "golden", "dna9000.npz"))["codes"]
    n = codes.shape[0]
    raw = np.zeros((n, n), np.int64)
    t0 = time.time()
    step = 200
    for r in range(0, n, step):
        blk = oc.mismatch_raw_block(codes[r:r + step], codes[r:], 10, 1)
        raw[r:r + step, r:] = blk
        print(f"rows {r + step}/{n}  {time.time() - t0:.0f} s", flush=True)
    iu = np.triu_indices(n, 1)
    raw.T[iu] = raw[iu]  # mirror (kernels.py:213-215)
    assert np.array_equal(raw, raw.T)
    K = onp.normalize_K(raw.astype(np.float64))
    out = {
        "what": "get_mismatch_K(X, 10, 1) on the 9000 challenge sequences (file order), float64, C order",
        "sha256": hashlib.sha256(np.ascontiguousarray(K).tobytes()).hexdigest(),
        "raw_sha256": hashlib.sha256(np.ascontiguousarray(raw).tobytes()).hexdigest(),
        "sum": float(K.sum()), "trace": float(np.trace(K)), "raw_trace": int(np.trace(raw)), "raw_sum": int(raw.sum()),
        "K_0_1": float(K[0, 1]), "K_4503_8999": float(K[4503, 8999]), "raw_0_0": int(raw[0, 0]), "raw_0_1": int(raw[0, 1]),
        "generated_by": "oracle/gen_config2_kat.py (oracle/kmg_oracle.c + oracle_np.normalize_K)",
        "seconds": time.time() - t0,
    }
    path = os.path.join(HERE, "..", "tests", "golden", "config2_mm10_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
