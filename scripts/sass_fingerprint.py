"""Fingerprints of the tuned kernels in synthesis_b200/libsynthesis_b200.so: registers, stack frame and a hash of the SASS text
(addresses and encodings stripped).  The thread-per-game kernels inline several descents and call shared noinline helpers; an
unrelated edit elsewhere in the translation unit can move their register allocation and cost percents (profiles/r2_normal_fpu.txt:
-2.5 % from one new caller of rng::chacha12_block, +25 % from one inlined descent less).  Run this before and after a change:

    python scripts/sass_fingerprint.py [path/to/lib.so]          # prints one line per kernel
    python scripts/sass_fingerprint.py > profiles/sass_fingerprint.txt

Needs cuobjdump (CUDA toolkit); no GPU."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = {
    "selfplay_nn_tpg2_kernel<5,4,false> (bench default)": "_ZN3eng23selfplay_nn_tpg2_kernelILi5ELi4ELb0ELin1EEEvNS_7KParamsE",
    "selfplay_nn_tpg2_kernel<5,4,false,NORMAL_CACHED>": "_ZN3eng23selfplay_nn_tpg2_kernelILi5ELi4ELb0ELi102EEEvNS_7KParamsE",
    "selfplay_nn_tpg2s_kernel<5,false> (split chain)": "_ZN3eng24selfplay_nn_tpg2s_kernelILi5ELb0ELin1EEEvNS_7KParamsE",
    "selfplay_rollout_tpg2_kernel<1024,3,CONST> (rollout bench)": "_ZN3eng28selfplay_rollout_tpg2_kernelILi1024ELi3ELi0EEEvNS_7KParamsE",
    "selfplay_rollout_tpg2_kernel<1024,3> (ParentQ)": "_ZN3eng28selfplay_rollout_tpg2_kernelILi1024ELi3ELin1EEEvNS_7KParamsE",
    "selfplay_nn_team_kernel<16,4,4> (configs[2])": "_ZN3eng23selfplay_nn_team_kernelILi16ELi4ELi4EEEvNS_7KParamsE",
    "selfplay_nn_team_kernel<32,4,4>": "_ZN3eng23selfplay_nn_team_kernelILi32ELi4ELi4EEEvNS_7KParamsE",
    "selfplay_rollout_kernel<32,256> (configs[0])": "_ZN3eng23selfplay_rollout_kernelILi32ELi256EEEvNS_7KParamsE",
}
LINE = re.compile(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "synthesis_b200", "libsynthesis_b200.so")
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.splitlines()
    usage = {}
    for i, l in enumerate(res):
        m = re.search(r"Function (\S+):", l)
        if m and i + 1 < len(res):
            usage[m.group(1)] = res[i + 1].strip()
    for name, sym in KERNELS.items():
        out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, lib], capture_output=True, text=True).stdout
        ins = [m.group(1) for m in map(LINE.match, out.splitlines()) if m]
        h = hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16] if ins else "(not found)"
        u = usage.get(sym, "")
        reg = re.search(r"REG:(\d+)", u)
        stack = re.search(r"STACK:(\d+)", u)
        print("%-62s %6d instructions  REG %-4s STACK %-5s sass %s" % (name, len(ins), reg.group(1) if reg else "?", stack.group(1) if stack else "?", h))


if __name__ == "__main__":
    main()
