"""CPU: properties of the generated SASS that the measured performance depends on (no GPU needed --
cuobjdump reads the in-tree .so).  ptxas decides how many global loads a warp keeps in flight; small
source changes have silently halved that depth, so the depth is pinned here:

* the hot kernels stage the CSR tiles with TMA (UBLKCP) and have no local-memory traffic in their gather loops;
* in the gather loop of the default SpMV / SpMM kernels at least 4 dense-row loads are issued
  back to back before the first FMA that consumes one."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "sparse-matrix-linear-equations_b200" / "libsmle_b200.so"
LOG = ROOT / "sparse-matrix-linear-equations_b200" / "build.log"

# mangled-name fragments of the default configurations (smle_capi.cu: kSpmv*, kSpmm*)
SPMV_DOT = "spmv_kernelIdLi480ELi6ELi2ELb1ELi0E"
SPMM = "spmm_rows_kernelIdLi16ELi2ELi1ELi4ELi960ELi1920ELi2ELi1ELb0"
SPMM_DOT = "spmm_rows_kernelIdLi16ELi2ELi1ELi4ELi960ELi1920ELi2ELi1ELb1"


@pytest.fixture(scope="module")
def sass():
    if not LIB.exists() or shutil.which("cuobjdump") is None:
        pytest.skip("library not built or cuobjdump missing")
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line.split("*/", 1)[1].split(";")[0].strip())
    return funcs


def _one(funcs, frag):
    hits = [k for k in funcs if frag in k]
    assert len(hits) == 1, (frag, hits)
    return funcs[hits[0]]


def _deepest_load_run(code, load_pat, fma="DFMA"):
    """largest number of matching loads issued without an FMA in between"""
    best = cur = 0
    for ins in code:
        if re.search(load_pat, ins) and not ins.startswith("@"):
            cur += 1
            best = max(best, cur)
        elif fma in ins:
            cur = 0
    return best


def test_tma_staging_present(sass):
    for frag in (SPMV_DOT, SPMM, SPMM_DOT):
        code = _one(sass, frag)
        assert any(i.startswith("UBLKCP") for i in code), f"{frag}: no cp.async.bulk (UBLKCP) in SASS"
        assert any("SYNCS" in i for i in code), f"{frag}: no mbarrier traffic"


def test_gather_depth(sass):
    assert _deepest_load_run(_one(sass, SPMM), r"LDG\.E\.128\.CONSTANT") >= 4
    assert _deepest_load_run(_one(sass, SPMM_DOT), r"LDG\.E\.128\.CONSTANT") >= 4
    assert _deepest_load_run(_one(sass, SPMV_DOT), r"LDG\.E\.64\.CONSTANT") >= 4


def test_no_spills_in_hot_loops(sass):
    """ptxas may park a launch-invariant value on the stack (a few bytes, read once per kernel), but no
    local-memory instruction may sit inside the gather loops: between the first and the last gather of
    the listing there must be no LDL / STL, and the total spill stays within 16 bytes."""
    if LOG.exists():
        text = LOG.read_text()
        for frag in (SPMV_DOT, SPMM, SPMM_DOT):
            m = re.search(re.escape(frag) + r".*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", text)
            assert m, frag
            assert int(m.group(2)) <= 16 and int(m.group(3)) <= 16, (frag, m.group(0)[-80:])
    for frag, pat in ((SPMV_DOT, r"LDG\.E\.64\.CONSTANT"), (SPMM, r"LDG\.E\.128\.CONSTANT"), (SPMM_DOT, r"LDG\.E\.128\.CONSTANT")):
        code = _one(sass, frag)
        idx = [i for i, ins in enumerate(code) if re.search(pat, ins)]
        assert idx, frag
        local = [ins for ins in code[idx[0]:idx[-1] + 1] if re.match(r"(@!?U?P\d+ )?(LDL|STL)", ins)]
        assert not local, (frag, local[:4])
