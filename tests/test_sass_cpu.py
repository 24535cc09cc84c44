"""Static properties of the built CUDA library that the measured performance rests on (cuobjdump, no GPU needed):
the evict-first hint sits on the global accesses of exactly the launches of ptx_rewrite.STREAM, the iteration kernel
keeps its read-only state in shared memory and its frame small, and nothing on the product path is a bulk-async
experiment or a development variant."""
import collections
import pathlib
import re
import shutil
import subprocess

import pytest

LIB = pathlib.Path(__file__).resolve().parent.parent / "elmkernels_b200" / "libelmk_b200.so"
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not LIB.exists(),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    fns, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = fns.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur is not None and m:
            op = m.group(1)
            base = op.split(".")[0]
            cur[base] += 1
            if base in ("LDG", "STG") and ".EF" in op:
                cur["EF:" + base] += 1
    return fns


def _one(fns, pattern):
    hits = [k for k in fns if re.search(pattern, k)]
    assert len(hits) == 1, (pattern, hits)
    return fns[hits[0]]


def test_evict_first_hint_on_the_listed_launches_only(sass):
    it = _one(sass, r"k_canflux_iterate")
    snow = _one(sass, r"k_groups_occILj1792E")
    soil = _one(sass, r"k_groups_occILj128E")
    assert it["EF:LDG"] >= 50 and it["EF:STG"] >= 30          # loads (refill) and stores (38 values of a converged column)
    assert snow["EF:LDG"] >= 100 and snow["EF:STG"] >= 100    # loads and stores
    assert soil["EF:LDG"] == 0 and soil["EF:STG"] >= 10       # stores only: the kernel re-reads its rows
    for name, c in sass.items():
        if not re.search(r"k_canflux_iterate|k_groups_occILj1792E|k_groups_occILj128E", name):
            assert c["EF:LDG"] == 0 and c["EF:STG"] == 0, name


def test_iteration_kernel_keeps_read_only_state_in_shared_memory(sass):
    it = _one(sass, r"k_canflux_iterate")
    assert it["LDS"] >= 60 and it["STS"] >= 60      # 74 doubles per column in flight: written at refill, read in the pass
    assert it["LDL"] + it["STL"] <= 120             # what is left of the spilled frame (520 B -> 136 B)
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    m = re.search(r"Function \S*k_canflux_iterate\S*:\s*\n\s*(.*)", res)
    assert m, "resource usage of the iteration kernel not found"
    use = dict(kv.split(":") for kv in m.group(1).split() if ":" in kv)
    assert int(use["REG"]) <= 170 and int(use["STACK"]) <= 200


def test_product_library_has_one_variant_per_kernel_and_no_bulk_async_experiment(sass):
    assert sum(1 for k in sass if "k_canflux_iterate" in k) == 1
    assert sum(1 for k in sass if "k_snicar" in k) == 1
    assert not any("k_groups_classed" in k for k in sass)
    for name, c in sass.items():
        assert c["UBLKPF"] == 0 and c["UBLKCP"] == 0, name
