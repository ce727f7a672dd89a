"""The reference's OWN unit-test files, run unmodified against this package as a drop-in
(``catalax`` -> ``catalax_b200`` through tests/golden/dropin_plugin.py; ``jax.numpy`` -> numpy).
Build container only: skipped where /root/reference does not exist (the GPU box).  Covers the
host rules SURVEY 8 calls a1 / a2 / a4 / a10 through the reference's own assertions:
``Reaction.from_scheme`` (19 tests), ``derive_stoich_matrix`` (4), ``Dataset`` / ``Dataset.pad`` (8)."""
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parents[1]
FILES = {"tests/unit/model/test_reaction.py": 19, "tests/unit/model/test_stoich.py": 4,
         "tests/unit/dataset/test_dataset.py": 8}


@pytest.mark.skipif(not (REF / "tests").exists(), reason="the reference tree is not available here")
@pytest.mark.parametrize("rel", list(FILES))
def test_reference_test_file_passes_against_this_package(rel):
    env = dict(os.environ, PYTHONPATH=f"{ROOT / 'tests' / 'golden'}:{ROOT}")
    r = subprocess.run([sys.executable, "-m", "pytest", "-p", "dropin_plugin", rel, "-q", "--no-header",
                        "-p", "no:cacheprovider"], cwd=REF, env=env, capture_output=True, text=True, timeout=600)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    m = re.search(r"(\d+) passed", tail)
    assert r.returncode == 0 and m and int(m.group(1)) == FILES[rel], r.stdout[-2000:] + r.stderr[-500:]
