"""Properties of the built sm_100a code that can be checked without a GPU, from `cuobjdump -sass` of the in-tree
library: the tensor-core kernels really are tcgen05 + TMA kernels (UTCHMMA / UTMALDG / UTMASTG / LDTM mnemonics,
B200_PROFILING.md's proof of the instruction family), they are built for sm_100a only, and two regressions that cost
measured throughput in r01 stay fixed:

* the CTA-pair conv kernel's per-tile remote mbarrier arrives carry no GPU-scope fence (`.release.cluster` compiled to
  MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR in front of every accumulator hand-back: 2 187 -> 2 225 pairs/s without it);
* the fused-GDN epilogue takes its GDN/IGDN decision once per 32-column chunk, not around every four elements
  (7.36 -> 7.23 ms per forward).
"""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hesic_b200", "lib", "libhesic_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    assert os.path.exists(LIB), "libhesic_b200.so not built (python -m hesic_b200.build)"
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"^arch = (\S+)", out, flags=re.M))
    funcs, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append(re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()))
    return archs, funcs


def _one(funcs, needle, exclude=None):
    names = [n for n in funcs if needle in n and not (exclude and exclude in n)]
    assert names, f"no kernel matching {needle}"
    return names


def _count(ins, prefix):
    return sum(1 for i in ins if i.startswith(prefix))


def test_built_for_sm_100a_only(sass):
    archs, funcs = sass
    assert archs == {"sm_100a"}, archs
    assert len(funcs) >= 40


@pytest.mark.parametrize("needle,exclude", [("conv_tc_kernel", None), ("conv_tc_pair_kernel", None), ("en_conv_kernel", None),
                                            ("conv_head_kernel", None)])
def test_tensor_core_kernels_are_tcgen05_and_tma(sass, needle, exclude):
    _, funcs = sass
    for name in _one(funcs, needle, exclude):
        ins = funcs[name]
        assert _count(ins, "UTCHMMA") > 0, f"{name}: no tcgen05.mma (UTCHMMA)"
        assert _count(ins, "UTMALDG") > 0, f"{name}: operands not loaded by TMA (UTMALDG)"
        assert _count(ins, "LDTM") > 0, f"{name}: accumulators not read from TMEM (LDTM)"
        assert _count(ins, "HMMA") == 0 and _count(ins, "HGMMA") == 0, f"{name}: legacy mma.sync / wgmma present"
    pair = funcs[_one(funcs, "conv_tc_pair_kernel")[0]]
    assert all(".2CTA" in i for i in pair if i.startswith("UTCHMMA")), "pair kernel must issue cta_group::2 MMAs only"
    assert _count(pair, "UTMASTG") > 0 and _count(funcs[_one(funcs, "conv_tc_kernelILi8E")[0]], "UTMASTG") > 0  # TMA-store epilogue
    assert _count(funcs[_one(funcs, "conv_tc_kernelILi16E")[0]], "UTMASTG") > 0                      # 16-warp epilogue (first analysis layer)


def test_pair_kernel_hands_tiles_back_without_gpu_scope_fences(sass):
    _, funcs = sass
    pair = funcs[_one(funcs, "conv_tc_pair_kernel")[0]]
    # exactly the two cluster-wide syncs at kernel start and end (barrier.cluster.arrive.release) remain
    assert _count(pair, "MEMBAR.ALL.GPU") == 2, _count(pair, "MEMBAR.ALL.GPU")
    for name in _one(funcs, "conv_tc_kernelILi"):           # the single-CTA kernel, 8- and 16-warp epilogues
        assert _count(funcs[name], "MEMBAR.ALL.GPU") == 0
    for name in _one(funcs, "en_conv_kernel"):              # the enhancement kernel
        assert _count(funcs[name], "MEMBAR.ALL.GPU") == 0


@pytest.mark.parametrize("needle", ["conv_tc_kernelILi8E", "conv_tc_pair_kernel"])
def test_gdn_scale_pass_is_one_block_per_chunk(sass, needle):
    """Between the first and the last MUFU.SQRT of the kernel (the IGDN side of epilogue pass 2: 2 chunks x 32
    columns) there are only a handful of branches; one per four elements would be >= 16."""
    _, funcs = sass
    ins = funcs[_one(funcs, needle)[0]]
    idx = [i for i, s in enumerate(ins) if s.startswith("MUFU.SQRT")]
    assert len(idx) >= 64
    # the 64 factors of the TMA-store form: the last 64 MUFU.SQRT of the function are the two 32-column chunks
    first, last = idx[-64], idx[-1]
    chunk_a = ins[idx[-64]:idx[-33] + 1]
    chunk_b = ins[idx[-32]:idx[-1] + 1]
    for chunk in (chunk_a, chunk_b):
        assert _count(chunk, "MUFU.SQRT") == 32
        assert _count(chunk, "BRA") == 0, "GDN/IGDN decision must not be taken inside a 32-column chunk"
    assert first < last


def test_register_budget_of_the_conv_kernels():
    """320-thread CTAs are capped at 168 registers (the 576-thread one at 96); spills in the epilogue were a measured cost (d9140f8).  Checked
    from the resource usage cuobjdump reports for the built cubin."""
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], check=True, capture_output=True, text=True).stdout
    seen = 0
    for m in re.finditer(r"Function (\S*conv_tc\S*kernel\S*):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        name, reg, stack = m.group(1), int(m.group(2)), int(m.group(3))
        if "conv_head" in name:
            continue
        seen += 1
        assert reg <= 168, (name, reg)
        assert stack <= 64, f"{name}: {stack} bytes of stack (spills) -- the epilogue must stay in registers"
    assert seen == 4      # conv_tc_kernel<8>, conv_tc_kernel<16>, conv_tc_pair_kernel, conv_tc_first_kernel


def test_single_thread_issue_regions_have_no_serialisation_loops(sass):
    """Producer / MMA / store-issuing threads are chosen with elect.sync, not `lane == 0`: the compiler then knows that one
    thread runs the region and feeds UTCHMMA / UTMALDG / UTMASTG from uniform registers directly.  With `lane == 0` every such
    instruction sat in an ELECT ... BRA.U.ANY loop over the active lanes (r02: 499 of them in the library; the MMA thread of the
    first-layer kernel needed ~10 instructions and a branch per MMA and was the kernel's bound)."""
    _, funcs = sass
    ins = funcs[_one(funcs, "conv_tc_first_kernel")[0]]
    assert _count(ins, "BRA.U.ANY") == 0
    # the other kernels: only the (few per tile) TMA stores / residual loads issued from inside the epilogue keep such a loop
    for needle in ("conv_tc_pair_kernel", "conv_tc_kernelILi8E", "conv_tc_kernelILi16E", "en_conv_kernel"):
        ins = funcs[_one(funcs, needle)[0]]
        loops = [i for i, x in enumerate(ins) if x.startswith("BRA.U.ANY") or " BRA.U.ANY" in x]
        assert len(loops) <= 12, (needle, len(loops))      # r02 before the change: 60-190 per kernel
        for i in loops:
            assert not any(x.startswith("UTCHMMA") for x in ins[max(0, i - 6):i]), needle


def test_first_layer_kernel_is_tcgen05_with_resident_operands(sass):
    _, funcs = sass
    ins = funcs[_one(funcs, "conv_tc_first_kernel")[0]]
    assert _count(ins, "UTCHMMA") >= 51          # 27 conv + 24 GDN MMAs per tile, unrolled
    assert _count(ins, "UTMALDG") > 0 and _count(ins, "UTMASTG") > 0 and _count(ins, "LDTM") > 0 and _count(ins, "STTM") > 0
    assert _count(ins, "BAR.SYNC") <= 2          # no block-level barrier in the tile loop (start and end only)
    assert _count(ins, "MEMBAR.ALL.GPU") == 0
