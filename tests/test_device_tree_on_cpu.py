"""Device tree code (synthesis_b200/csrc/tpg4_tree.cuh, the family-block layout behind SYN_TPG_VER=4) compiled as plain
C++ and stepped on the CPU against the oracle, node for node (tests/emu/tpg4_emu.cpp).  No GPU needed: what runs here is
the kernels' own source — descend, write_children, backprop with the memoised -q, read_root — not a restatement of it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


import pytest


def _build(tmp_path, which="tpg4"):
    exe = str(tmp_path / f"{which}_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-DSYN_HOST_EMU", os.path.join(EMU, f"{which}_emu.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("which", ["tpg4"])
def test_tree_code_matches_the_oracle_node_for_node(tmp_path, which):
    exe = _build(tmp_path, which)
    # 13 configurations x 60 positions (0..40 plies into a game) x 800 explores: every node of every tree, every bit
    out = subprocess.run([exe, "60", "800"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-4000:]
    assert "all trees identical" in out.stdout
    assert out.stdout.count(" ok ") == 15, out.stdout


@pytest.mark.parametrize("which", ["tpg4"])
def test_tree_code_at_config_sizes(tmp_path, which):
    """BASELINE.json configs[2] / configs[4] sizes: 1,600 and 10,000 explores per tree."""
    exe = _build(tmp_path, which)
    for positions, explores in ((12, 1600), (4, 10000)):
        out = subprocess.run([exe, str(positions), str(explores)], capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-4000:]


def test_a_wrong_memo_is_caught(tmp_path):
    """The harness must be able to fail: perturb the memoised -q by one part in 10^7 and the trees must differ."""
    src = open(os.path.join(ROOT, "synthesis_b200", "csrc", "tpg4_tree.cuh")).read()
    needle = "return -__fdiv_rn(__fsub_rn(o2, o0), vis); }"
    assert needle in src
    csrc = os.path.join(ROOT, "synthesis_b200", "csrc")
    mut = src.replace(needle, "return -__fdiv_rn(__fsub_rn(o2, o0), vis) * 1.0000001f; }")
    for inc in ("devport.cuh", "c4.cuh", "treedefs.cuh"):
        mut = mut.replace(f'#include "{inc}"', f'#include "{csrc}/{inc}"')
    mut = mut.replace('#include "../../include/', f'#include "{ROOT}/include/')
    (tmp_path / "tpg4_tree_mut.cuh").write_text(mut)
    emu = open(os.path.join(EMU, "tpg4_emu.cpp")).read()
    emu = emu.replace('#include "../../synthesis_b200/csrc/tpg4_tree.cuh"', f'#include "{tmp_path}/tpg4_tree_mut.cuh"')
    emu = emu.replace('#include "../../oracle/', f'#include "{ROOT}/oracle/')
    (tmp_path / "emu_mut.cpp").write_text(emu)
    exe = str(tmp_path / "emu_mut")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-DSYN_HOST_EMU", str(tmp_path / "emu_mut.cpp"), "-o", exe])
    out = subprocess.run([exe, "10", "300"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "FAIL" in out.stdout
