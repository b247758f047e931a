"""Coverage pre-filter of the UNMODIFIED reference on the 1 M-read synthetic set of make_golden_1m.py (same reads, same
generator): `ref_probe covfilter` = EstimateCardinalityUpperBound -> qf::cqf -> FillCoverageHistogram -> io::CoverageFilter per read.

    make -C oracle ref && python tests/golden/make_golden_1m_cov.py        (build container only)

Kept: the cardinality bound, the filter's key width, the number of surviving reads and the SHA-256 of the verdicts (one byte 0/1 per
read). tests/test_gpu_parity.py::test_million_reads_coverage_prefilter recomputes them through the C ABI on the B200."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_1m as M  # noqa: E402

CASES = [(21, 3), (55, 4)]      # (k, read_cov_threshold)


def main():
    codes = M.synthetic_reads(M.N_READS, M.READ_LEN, M.GENOME, M.ERR, seed=M.SEED, as_codes=True)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    txt = np.empty((M.N_READS, M.READ_LEN + 1), np.uint8)
    txt[:, :M.READ_LEN] = lut[codes]
    txt[:, M.READ_LEN] = 10
    out = {"reads_sha256_text": hashlib.sha256(txt.tobytes()).hexdigest(), "cases": {}}
    with tempfile.TemporaryDirectory() as d:
        rf = os.path.join(d, "reads.txt")
        txt.tofile(rf)
        for k, thr in CASES:
            od = os.path.join(d, "cov_k%d" % k)
            env = dict(os.environ); env["PROBE_COV_THR"] = str(thr)
            t0 = time.time()
            subprocess.check_call([M.PROBE, "covfilter", rf, str(k), "4", "8", od], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
            card, bits, mask, _ = [int(x) for x in open(os.path.join(od, "covfilter.txt")).read().split()]
            keep = np.array([int(x) for x in open(os.path.join(od, "keep.txt")).read().split()], dtype=np.uint8)
            out["cases"]["k%d" % k] = {"k": k, "threshold": thr, "cardinality_upper_bound": card, "key_bits": bits, "kept": int(keep.sum()),
                                       "sha256_keep": hashlib.sha256(keep.tobytes()).hexdigest(), "reference_wall_s": round(time.time() - t0, 1)}
            print(out["cases"]["k%d" % k], flush=True)
    json.dump(out, open(os.path.join(HERE, "syn1M_covfilter.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
