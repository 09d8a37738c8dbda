"""Builds oracle/liboracle.so (TEST INFRASTRUCTURE) with `-O3 -march=native` for the machine it
runs on.  The library is rebuilt whenever a source is newer or when it was built on a different
CPU (the .so travels to the GPU box with the snapshot; -march=native code must not be reused there).
Only tests/, __graft_entry__ and bench.py's CPU-baseline legs call this.
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "liboracle.so")
STAMP = os.path.join(HERE, ".build_host")


def _cpu_id() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("\n".join(model).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".hpp"))]
    srcs += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    cpu = _cpu_id()
    stale = force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs)
    if not stale:
        try:
            with open(STAMP) as f:
                stale = f.read().strip() != cpu
        except OSError:
            stale = True
    if stale:
        subprocess.check_call(["make", "-B", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
        with open(STAMP, "w") as f:
            f.write(cpu)
    return SO


if __name__ == "__main__":
    print(build())
