"""Generates tests/golden/golden_{lbd,hbd}.npz from the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The fixtures hold the reference's outputs for the seeded cases of cases.py; inputs are regenerated from the seed
by the tests (make_cases is deterministic), so only outputs + an input checksum are stored."""
import hashlib
import numpy as np
from cases import make_cases, RefBackend, golden_path


def digest(inputs):
    h = hashlib.sha256()
    for k in sorted(inputs):
        h.update(k.encode()); h.update(np.ascontiguousarray(inputs[k]).tobytes())
    return np.frombuffer(h.digest()[:8], np.uint8).copy()


if __name__ == "__main__":
    for hbd, bd in ((0, 8), (1, 10)):
        be = RefBackend(hbd, bd)
        store = {}
        cases = make_cases(hbd, bd)
        for i, (kind, p, x) in enumerate(cases):
            out = be.run(kind, p, x)
            for k, v in out.items():
                store["c%d_%s" % (i, k)] = v
            store["c%d__digest" % i] = digest(x)
        store["n"] = np.array(len(cases))
        np.savez_compressed(golden_path(hbd), **store)
        print(golden_path(hbd), len(cases), "cases")
