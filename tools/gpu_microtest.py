"""Runs every conv-operator case on the GPU against the torch fp32 reference, surviving CUDA faults.

python tools/gpu_microtest.py            # master: runs cases in child processes, restarts after a fault
python tools/gpu_microtest.py --from K   # child: run cases K.. in this process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(start):
    import torch
    from ctrlhair_b200 import ops
    import _conv_cases as cc
    cases = cc.make_cases()
    for i in range(start, len(cases)):
        name, fn = cases[i]
        for impl, iname in ((ops.IMPL_SIMT_DEBUG, "simt"), (ops.IMPL_TCGEN05, "tcgen05")):
            print("CASE %d %s %s BEGIN" % (i, name, iname), flush=True)
            t0 = time.time()
            got, want = fn(impl)
            torch.cuda.synchronize()
            err = cc.rel_err(got, want)
            bad = int((~torch.isfinite(got)).sum())
            print("CASE %d %s %s rel_err=%.3e nonfinite=%d max_ref=%.3f (%.2fs) %s" %
                  (i, name, iname, err, bad, float(want.abs().max()), time.time() - t0,
                   "OK" if err < 2e-3 and bad == 0 else "FAIL"), flush=True)
    print("CHILD DONE", flush=True)


def master():
    start = 0
    import _conv_cases  # noqa: F401  (import check only; needs torch)
    n = len(_conv_cases.make_cases(device="cpu"))
    while start < n:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--from", str(start)], capture_output=True,
                           text=True, timeout=900)
        out = p.stdout
        sys.stdout.write(out)
        if p.returncode != 0:
            sys.stdout.write("CHILD EXIT %d\n%s\n" % (p.returncode, p.stderr[-2000:]))
        last = -1
        for line in out.splitlines():
            if line.startswith("CASE ") and "BEGIN" in line:
                last = int(line.split()[1])
        if "CHILD DONE" in out:
            break
        start = max(last, start) + 1
    sys.stdout.flush()


if __name__ == "__main__":
    if "--from" in sys.argv:
        child(int(sys.argv[sys.argv.index("--from") + 1]))
    else:
        master()
