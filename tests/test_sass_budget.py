"""The instruction budgets DESIGN.md quotes for the Monte Carlo kernels, checked against the object file of the build (no GPU):
`scripts/sass_loop.py` finds the trial / round loop of a kernel variant in its SASS and counts it.  The numbers are not contracts of
the algorithm -- they are what the latency- and issue-bound loops were tuned to, and a compiler or source change that loses them
should be noticed before it reaches a B200."""

import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJECT = os.path.join(ROOT, "pcetk_b200", "csrc", "pce_mc.o")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(OBJECT), reason="needs cuobjdump and the built object file")


@pytest.fixture(scope="module")
def sass():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sass_loop
    return sass_loop, sass_loop.disassemble(OBJECT)


@pytest.mark.parametrize("key,limit,what", [
    ("k_mc_threadILb0EtLi128ELi1ELb0E", 290, "chain per thread, short fields, quiet path (lysozyme headline): 272 when tuned, 381 before"),
    ("k_mc_threadILb0EtLi128ELi2ELb0E", 300, "the same for field vectors of up to 16 slots (defensin)"),
    ("k_mc_warpILb0ELi3ELi1ELb0ELb1E", 450, "chain per warp, field vectors of up to 64 slots (default titration curve): 424 when tuned, 689 before"),
    ("k_mc_warpILb0ELi3ELi1ELb0ELb0E", 760, "chain per warp, small fields (ifp20)"),
])
def test_loop_instruction_budget(sass, key, limit, what):
    module, text = sass
    assert len(module.loop_region(text, key)) <= limit, what
