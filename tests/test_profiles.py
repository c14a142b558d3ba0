"""profiles/: the summary bench.py reads must be exactly what scripts/make_profile_summary.py derives from the committed ncu
digests (no hand-typed numbers), and every digest the manifest cites must exist."""

import glob
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.mark.parametrize("manifest", sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_manifest.json"))))
def test_summary_is_derived_from_the_digests(manifest):
    import make_profile_summary as mps

    summary = mps.summarise(manifest)
    committed = json.load(open(manifest.replace("_manifest.json", "_summary.json")))
    assert json.loads(json.dumps(summary)) == committed
    for key, entry in committed.items():
        if key.startswith("_"):
            continue
        per_unit = [k for k in entry if k.startswith("warp_instructions_per_")]
        assert len(per_unit) == 1
        assert entry[per_unit[0]] == pytest.approx(entry["warp_instructions"] / entry["units"])
        assert 0.0 < entry["issue_active_frac"] <= 1.0 and entry["kernel_ms"] > 0.0


def test_bench_reads_the_current_round_summary():
    text = open(os.path.join(ROOT, "bench.py")).read()
    assert '"profiles", "r02_summary.json"' in text
    assert os.path.exists(os.path.join(ROOT, "profiles", "r02_summary.json"))
