"""Static checks on the compiled sm_100a code (no GPU needed, cuobjdump only):
  * the fused scan kernels really are TMA pipelines (UBLKCP bulk copies + SYNCS mbarrier ops);
  * every consumer-side stage release is *predicated* — i.e. data-dependent on the values loaded
    from the stage (DESIGN.md "Release rule"): an unpredicated SYNCS.ARRIVE after ld.shared was
    measured to let the next bulk copy overtake in-flight reads (stale tiles);
  * no kernel of the library spills to local memory stack beyond a few bytes."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "velox_b200", "lib", "libvelox_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")


def sass_by_function():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return funcs


def test_tma_pipelines_release_stages_data_dependently():
    funcs = sass_by_function()
    tma = {k: v for k, v in funcs.items() if "fused_scan_agg_tma_kernel" in k or "fused_scan_compact_tma_kernel" in k}
    assert len(tma) >= 10, "fused TMA kernels missing from the library"
    for name, lines in tma.items():
        text = "\n".join(lines)
        assert "UBLKCP" in text, f"{name}: no TMA bulk copy"
        assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in text and "SYNCS.ARRIVE.TRANS64" in text, f"{name}: no mbarrier pipeline"
        # consumer releases: arrive with count 1 and no transaction bytes (A1T0 / ART0). The producer's
        # expect_tx arrive is the plain form with a byte-count register and stays unpredicated.
        releases = [l for l in lines if "SYNCS.ARRIVE.TRANS64.A1T0" in l or "SYNCS.ARRIVE.TRANS64.ART0" in l]
        assert releases, f"{name}: no consumer release found"
        for l in releases:
            instr = re.sub(r"^/\*[0-9a-f]+\*/\s*", "", l)
            assert instr.startswith("@"), f"{name}: unpredicated stage release: {l}"


def test_kernels_do_not_spill():
    """No kernel spills beyond a few bytes. The one tolerated exception: the 4-group register-
    accumulator variants of five-projection pipelines (20 double accumulators per thread at two
    blocks per SM) may keep a few accumulators on the stack; TPC-H Q1 does not run them (its six-slot
    group space takes the shared-memory accumulators)."""
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, timeout=600).stdout
    worst, worst_hot = 0, 0
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:\d+ STACK:(\d+)", out):
        name, stack = m.group(1), int(m.group(2))
        four_group_variant = "fused_" in name and re.search(r"Li4E[il]E", name) is not None
        if four_group_variant:
            worst = max(worst, stack)
        else:
            worst_hot = max(worst_hot, stack)
    assert worst_hot <= 64, f"a kernel uses {worst_hot} bytes of local-memory stack"
    assert worst <= 128, f"a 4-group register-accumulator variant uses {worst} bytes of local-memory stack"
