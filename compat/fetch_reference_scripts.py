"""Copies the reference's three entry scripts, UNMODIFIED, from /root/reference into the git-ignored compat/_ref/ so that
tests/test_gpu_compat_scripts.py can execute their own text on a GPU box (where /root/reference does not exist).  Nothing of
the reference is committed: compat/_ref/ is listed in .gitignore (it still travels with the working-tree snapshot)."""
import os
import shutil
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

if __name__ == "__main__":
    if not os.path.isdir(SRC):
        sys.exit("no reference tree at %s" % SRC)
    os.makedirs(DST, exist_ok=True)
    for name in ("train_simple_r3d.py", "train.py", "validation.py"):
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        print("copied", name)
