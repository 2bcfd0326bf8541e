"""Run every `-m gpu` test of a file in its own process (a trapped kernel kills only its own CUDA context)
and write a one-line-per-test summary.  Bring-up tool; the judged run is plain `pytest -m gpu`."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    target = sys.argv[1] if len(sys.argv) > 1 else "tests/test_kernels_gpu.py"
    per_test_timeout = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    ids = subprocess.run([sys.executable, "-m", "pytest", target, "--collect-only", "-q", "-m", "gpu"],
                         capture_output=True, text=True, cwd=ROOT).stdout.splitlines()
    ids = [i for i in ids if "::" in i]
    # group by test function so one process handles all parametrisations of one kernel
    groups = {}
    for i in ids:
        groups.setdefault(i.split("[")[0], []).append(i)
    lines = []
    for name, members in groups.items():
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "--no-header", "-rfE", "--tb=short",
                                *members], capture_output=True, text=True, cwd=ROOT, timeout=per_test_timeout)
            tail = [l for l in r.stdout.splitlines() if l.strip()]
            summary = tail[-1] if tail else "?"
            fails = [l for l in tail if l.startswith(("FAILED", "ERROR"))]
            asserts = [l for l in tail if "assert " in l and l.startswith("E")][:12]
            lines.append(f"{name}: rc={r.returncode} {summary}")
            lines += ["    " + f for f in fails[:20]]
            lines += ["    " + a for a in asserts]
            if r.returncode not in (0, 1):
                lines += ["    | " + l for l in (r.stdout + r.stderr).splitlines()[-15:]]
        except subprocess.TimeoutExpired:
            lines.append(f"{name}: TIMEOUT after {per_test_timeout}s")
        print(lines[-1], flush=True)
    with open(os.path.join(out_dir, "isolated_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
