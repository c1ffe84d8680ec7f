"""Re-run the oracle-vs-reference pins (and any other CPU test file given) with shifted random seeds: THOR_FUZZ_OFFSET (tests/conftest.py)
shifts every np.random.default_rng(<const>) of the tests, so each round checks fresh random cases.
usage: python tools/fuzz_oracle.py [rounds] [first_offset] [pytest args...]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
extra = sys.argv[3:] or ["tests/test_oracle_vs_ref.py"]
bad, t0 = [], time.time()
for k in range(first, first + rounds):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] + extra, cwd=ROOT, env=dict(os.environ, THOR_FUZZ_OFFSET=str(k)),
                       capture_output=True, text=True)
    tail = r.stdout.strip().split("\n")[-1]
    print("offset %d: %s (%.0f s)" % (k, tail, time.time() - t0), flush=True)
    if r.returncode != 0:
        bad.append(k)
        print(r.stdout[-3000:], flush=True)
print("failing offsets:", bad)
sys.exit(1 if bad else 0)
