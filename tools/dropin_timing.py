"""Wall-clock time per HMC sample of the reference's OWN program, CPU build vs GPU drop-in (run on the GPU box):

    python tools/dropin_timing.py --grid 64 --samples 6

oracle/_ref/barcode_cpu (every barlib source unmodified) and oracle/_ref/barcode_gpu (HMC.cc + HMC_momenta.cc
replaced by barcode_b200/csrc/barlib_gpu_glue.cc) run the same input.par (tests/test_dropin.py's) with N_Gibbs = 1
and N_Gibbs = samples; the difference of the two wall times is the cost of samples - 1 samples without the set-up
(mock data, initial guess).  Also with BARCODE_GPU_DEVICE_RNG=1.
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin import CPU, GPU, make_par  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--samples", type=int, default=6)
ap.add_argument("--skip", nargs="*", default=[], help="variants to leave out: cpu, host_rng, separate")
ap.add_argument("--verbose", action="store_true", help="print the glue's per-call wall times of every run")
a = ap.parse_args()
N = a.grid
L = N * 200.0 / 64
tmp = tempfile.mkdtemp(prefix="dropin_timing_")
with np.load(os.path.join(ROOT, "barcode_b200", "data", "pk_table.npz")) as f:
    k, P = f["k"], f["P"]
pk = os.path.join(tmp, "pk.dat")
with open(pk, "w") as o:
    for x, y in zip(k, P):
        o.write(f"{x:.9g} {y:.9g}\n")


def run(exe, tag, n_gibbs, env=None):
    d = os.path.join(tmp, f"{tag}_{n_gibbs}")
    os.makedirs(os.path.join(d, "data"))
    par = make_par(calc_h=0, rsd="true", likelihood=1, eps_fac=0.004, mass_type=1, pk=pk, N=N, L=L,
                           n_gibbs=n_gibbs, masskernel=1)
    open(os.path.join(d, "input.par"), "w").write(par)
    t0 = time.time()
    r = subprocess.run([exe], cwd=d, capture_output=True, text=True,
                       env=dict(os.environ, BARCODE_GPU_TIMING="1", **(env or {})))
    dt = time.time() - t0
    if r.returncode != 0:
        raise SystemExit(r.stdout[-2000:] + r.stderr[-2000:])
    # the glue's own wall times per C-ABI call site (BARCODE_GPU_TIMING): bgpu_create is the CUDA context + plan
    # set-up, 2.3 ... 2.9 s from run to run -- it is taken out of the run's wall time before the two runs are differenced
    for ln in r.stderr.splitlines():
        if ln.startswith("[barcode_gpu]"):
            if a.verbose:
                print(f"      {tag} N_Gibbs={n_gibbs}: {ln}")
            f = ln.split()
            if f[1] == "bgpu_create":
                dt -= float(f[-2]) * 1e-3
    rows = open(os.path.join(d, "performance_log.txt")).read().strip().splitlines()[1:]
    neps = [float(x.split("\t")[2]) for x in rows]
    inside = None   # wall time inside HamiltonianMC without bgpu_create, from the glue's timers
    tm = {ln.split()[1]: float(ln.split()[-2]) for ln in r.stderr.splitlines() if ln.startswith("[barcode_gpu]")}
    if "HamiltonianMC" in tm:
        inside = (tm["HamiltonianMC"] - tm.get("bgpu_create", 0.0)) * 1e-3
    return dt, len(rows), sum(neps), inside


print(f"grid {N}^3, ZA + CIC, Gaussian likelihood, RSD; host cores: {os.cpu_count()}")
variants = [("cpu", CPU, "reference CPU build", None),
            ("host_rng", GPU, "GPU drop-in (GSL stream on the host)", None),
            ("separate", GPU, "GPU drop-in, device RNG, host-array calls",
             {"BARCODE_GPU_DEVICE_RNG": "1", "BARCODE_GPU_FUSED": "0"}),
            ("fused", GPU, "GPU drop-in, device RNG, device-resident candidates", {"BARCODE_GPU_DEVICE_RNG": "1"})]
for key, exe, tag, env in variants:
    if key in a.skip:
        continue
    t1, c1, e1, _ = run(exe, key, 1, env)
    tn, cn, en, inside = run(exe, key, a.samples, env)
    per = (tn - t1) / max(1, (a.samples - 1))
    hmc = "" if inside is None else f"; inside HamiltonianMC {inside / a.samples * 1e3:.1f} ms / sample"
    print(f"  {tag:52s} {per * 1e3:10.1f} ms / sample   ({cn - c1} candidates, {en - e1:.0f} leapfrog steps in "
          f"{tn - t1:.2f} s; set-up without bgpu_create + first sample {t1:.2f} s{hmc})")
