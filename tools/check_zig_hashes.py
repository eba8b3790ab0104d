"""Compare `name sha256` lines (stdin) printed by the Zig program of INTEGRATION.md section 2 -- sha256 of
jpeg.load(f).rgbaPixels() from the REAL zpix decoder -- with tests/golden/rgba_sha256.json, the hashes this repository's
oracle and GPU path are tested against.  Exit code 0 = absolute parity with the Zig binary on all listed files."""
import json
import os
import sys

gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rgba_sha256.json")))
bad = seen = 0
for line in sys.stdin:
    f = line.split()
    if len(f) != 2 or f[0] not in gold:
        continue
    seen += 1
    if gold[f[0]] != f[1]:
        bad += 1
        print("MISMATCH", f[0], f[1], "expected", gold[f[0]])
print(f"{seen} files compared, {bad} mismatches, {len(gold) - seen} of the golden files not seen")
sys.exit(1 if bad or seen == 0 else 0)
