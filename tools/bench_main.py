#!/usr/bin/env python
"""Time the reference's own training script (Main.py -> Model_Trainer.train, unmodified) for one epoch on the shipped
SF data, three ways:  stock cell | B200 cell installed | installed + loop hygiene (contiguous batch slices instead of
per-item collation, per-step torch.cuda.empty_cache() neutralised; stc_gnn_b200/install.py).

  python tools/bench_main.py [--epochs 2]

Each variant runs in its own process (fresh CUDA context, fresh module state); the epoch time is the one
Model_Trainer prints ("training time: ... s/epoch", the LAST epoch of the run = warm).  One JSON line per variant.
Needs the unmodified reference on the box (probe: $STC_REF_DIR, /root/reference/framework, baseline/_ref/framework).
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand in (os.environ.get("STC_REF_DIR"), "/root/reference/framework", os.path.join(ROOT, "baseline", "_ref", "framework")):
        if cand and os.path.isfile(os.path.join(cand, "STC_GNN.py")):
            return cand
    return None


CHILD = r"""
import sys, time
sys.path.insert(0, {root!r})
import torch
from stc_gnn_b200.install import run_main
t0 = time.time()
run_main({ref!r}, ['-device', 'cuda:0', '-city', 'SF', '-epoch', {epochs!r}, '-out', {out!r}, '-in', {data!r}],
         hygiene={hygiene}, install_cell={install})
torch.cuda.synchronize()
print('WALL_S', time.time() - t0)
"""


def main():
    epochs = sys.argv[sys.argv.index("--epochs") + 1] if "--epochs" in sys.argv else "2"
    ref = find_reference()
    if ref is None:
        print(json.dumps({"kind": "main_py_epoch", "unavailable": "the unmodified reference is not on this box"}))
        return
    data = os.path.join(os.path.dirname(ref), "data")
    for name, install, hygiene in (("stock cell", False, False), ("b200 cell installed", True, False),
                                   ("b200 cell + loop hygiene", True, True)):
        with tempfile.TemporaryDirectory() as out:
            code = CHILD.format(root=ROOT, ref=ref, epochs=epochs, out=out, data=data, hygiene=hygiene, install=install)
            res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500)
        txt = res.stdout + res.stderr
        times = [float(x) for x in re.findall(r"Epoch \d+: training time: ([0-9.eE+-]+) s/epoch", txt)]   # not the "Average ..." line
        infer = [float(x) for x in re.findall(r"; inference time: ([0-9.eE+-]+) s", txt)]
        loss = re.findall(r"training loss: ([0-9.eE+-]+);", txt)
        wall = re.findall(r"WALL_S ([0-9.]+)", txt)
        rec = {"kind": "main_py_epoch", "variant": name, "epochs": int(epochs), "rc": res.returncode,
               "train_s_per_epoch": times, "validate_s": infer, "train_loss": loss,
               "train_windows": 3834, "samples_per_s_last_epoch": (3834 / times[-1]) if times else None,
               "whole_run_wall_s": float(wall[0]) if wall else None}
        if res.returncode != 0:
            rec["stderr_tail"] = txt[-600:]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
