"""Marginal cost of each kernel family inside the replayed training graphs: the step is timed with one family's launches
left out (FVT_SKIP; numerically meaningless runs, timing only).  usage: gpu_train_ablate.py"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for skip in ("", "wgrad", "dgrad", "bnbwd", "wgrad,dgrad,bnbwd"):
    if len(sys.argv) > 1 and skip not in sys.argv[1:]:
        continue
    env = dict(os.environ, FVT_SKIP=skip)
    out = subprocess.run([sys.executable, os.path.join(here, "gpu_train_time.py"), "4", "20"], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if "ms/step" in l]
    print("skip=%-20s %s" % (skip or "-", line[0].split(":")[1].split("losses")[0].strip() if line else out.stderr[-300:]), flush=True)
