#!/usr/bin/env python
"""Labelled 80-channel MSR data for the config-2 objective parity test: the reference's own generator
`SUM_RATE_GEN(M=80, W=20)` (utils/dataset_generate.py:280-313, driven as datasets/sum_rate_gen.py:10-12 drives it),
run UNMODIFIED from /root/reference under a fixed numpy seed, stored as tests/golden/msr80c_data.npz.
The real `datasets/80c_20w_10000samples.csv` is one of the blobs missing from the reference repo.

    python oracle/make_golden_80c.py         # needs /root/reference; test infrastructure only
"""
import io
import sys
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "msr80c_data.npz"

if __name__ == "__main__":
    sys.path.insert(0, str(REF))
    from utils.dataset_generate import SUM_RATE_GEN
    np.random.seed(80)
    with redirect_stdout(io.StringIO()):
        gs, rates, schemes = SUM_RATE_GEN(sample_num=2048, M=80, W=20.0)
    # row layout of the reference CSV: g[M] | rate | p*[M]  (classifier_free_MSR.py:171-173 reads it back this way)
    assert gs.shape == (2048, 80) and schemes.shape == (2048, 80) and np.all(np.isfinite(schemes))
    print("label sum-rate mean", rates.mean(), "row power sum", schemes.sum(axis=1).mean())
    np.savez_compressed(OUT, g=gs.astype(np.float32), rate=rates.astype(np.float32), p=schemes.astype(np.float32),
                        W=np.float32(20.0))
    print("wrote", OUT, OUT.stat().st_size, "bytes")
