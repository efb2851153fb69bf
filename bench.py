#!/usr/bin/env python
"""bench.py -- HMC gradient evaluations per second on Barcode's hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--grid 256] [--calc-h 0|4] [--impl ours|reference]

A "step" is one `gradient_psi` evaluation (HMC.cc:146-206) on one batch of
synthetic Gaussian-random-field data.  The N=1 workload is BASELINE.json
configs[1]: 256^3, CIC, Gaussian likelihood, plane-parallel RSD, `sfmodel = 2`
requested -- for which the reference runs Zel'dovich (HMC_models.cc:395-400) --
with the reference's own gradient for CIC (`calc_h = 0`, the only one it has).
With N > 1 ranks (torchrun, one per GPU) every rank runs an independent chain
with its own seed (BASELINE.json configs[2]'s sharding: no data-path
collective), so scaling is weak and `value` is the sum over chains.

Prints ONE JSON line on rank 0 (contract in the task statement; DESIGN.md
section "Measurement" says what each key is).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hmc_gradient_evals_per_s"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--calc-h", type=int, default=0, choices=[0, 1, 4])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="chains", choices=["chains", "slab", "e2e_chains"],
                    help="N > 1 GPUs: independent chains (weak scaling, default) or ONE chain, x-slab decomposed "
                         "(strong scaling; BASELINE.json configs[3], [4]).  e2e_chains (internal, one GPU): the "
                         "host-buffer call of several independent chains interleaved on ONE GPU, see run_e2e_chains")
    ap.add_argument("--chains", type=int, default=3,
                    help="e2e_chains: chains (handles, host threads) on the GPU; 3 = one uploading, one computing, one "
                         "downloading (each phase takes about as long at PCIe 5 x16 rates)")
    ap.add_argument("--no-e2e-chains", action="store_true", help="skip the interleaved-chains e2e leg")
    ap.add_argument("--no-sph", action="store_true", help="skip the SPH (masskernel 3, calc_h 2) line in `also`")
    ap.add_argument("--no-f32", action="store_true", help="skip the single-precision (bgpu_f32_*) line in `also`")
    ap.add_argument("--no-512", action="store_true", help="N = 1, default grid: skip the 512^3 line in `also`")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the slab-decomposed leg (512^3 / 1024^3)")
    return ap.parse_args()


def workload(grid: int, calc_h: int):
    from barcode_b200 import inputs
    L = inputs.box_length(grid)
    cfg = dict(N1=grid, L1=L, masskernel=1, likelihood=1, rsd_model=True, sfmodel=2, calc_h=calc_h, mass_type=1)
    which = ("configs[4], slab-decomposed" if grid == 1024 else
             "configs[1]; sfmodel=2 requested, the reference runs Zel'dovich under rsd_model")
    name = (f"{grid}^3 ZA+CIC Gaussian likelihood, plane-parallel RSD, L={L:g} Mpc/h, calc_h={calc_h} "
            f"(BASELINE.json {which})")
    return cfg, name


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            hot = sorted(sm)[len(sm) // 2:]  # the upper half are the samples under load
            out = {"sm_mhz": float(np.median(hot)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------- reference arm
def reference_problem(R, grid, L, seed):
    """Synthetic inputs for the reference arm, made by the reference itself
    (create_GARFIELD / Lag2Eul, barcoderunner.cc:42-205) from the same P(k) table."""
    from oracle import barcode_oracle as bo
    from barcode_b200 import inputs
    k_tab, p_tab = inputs.load_pk_table()
    P = inputs.power_on_grid(k_tab, p_tab, grid, L).ravel()
    one = np.ones(R.N)
    R.set_inputs(Power=P, window=one, noise=one, nobs=one)
    truth = R.create_garfield(seed, P)
    dX = R.forward(truth, want_pos=False)
    nobs = np.maximum(0.0, 1.0 + dX + np.random.default_rng(seed + 1000).standard_normal(R.N))
    R.set_inputs(nobs=nobs)
    s = 0.5 * R.create_garfield(seed + 1, P)
    R.hamiltonian_mass()
    return s


def use_all_host_threads(ref):
    """torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host core."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    ref.lib().ref_set_num_threads(max(1, n))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    cfg, name = workload(args.grid, 0)
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbarcode_ref.so was not built"}))
        return 0
    rc = ref.Config(N1=cfg["N1"], L1=cfg["L1"], masskernel=1, likelihood=1, rsd_model=True, sfmodel=2, calc_h=0,
                    mass_type=1)
    use_all_host_threads(ref)
    R = ref.Reference(rc)
    s = reference_problem(R, args.grid, cfg["L1"], 1)
    for _ in range(max(1, args.warmup)):
        R.gradient_psi(s)
    ref.fft_stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        R.gradient_psi(s)
    dt = time.perf_counter() - t0
    fft_s, fft_calls = ref.fft_stats()
    value = args.steps / dt
    cores = ref.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "fft_backend": ref.fft_backend(),
                   "fft_share_of_step": fft_s / dt, "fft_calls_per_step": fft_calls / args.steps,
                   "note": "the reference's own gradient_psi (calc_h=0) on the host cores, OpenMP threads = cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{args.steps} full gradient_psi calls at {args.grid}^3 after {max(1, args.warmup)} warm-up"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- our arm
def cpu_baseline(args, cfg):
    """The compiled reference (oracle/_ref) timed on this box's host cores: a bounded sample of
    the same workload (full-size gradient_psi calls, about 10-30 s of CPU work)."""
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    rc = ref.Config(N1=cfg["N1"], L1=cfg["L1"], masskernel=1, likelihood=1, rsd_model=True, sfmodel=2, calc_h=0,
                    mass_type=1)
    use_all_host_threads(ref)
    R = ref.Reference(rc)
    s = reference_problem(R, cfg["N1"], cfg["L1"], 1)
    t0 = time.perf_counter()
    R.gradient_psi(s)
    t1 = time.perf_counter() - t0
    reps = int(min(5, max(1, 15.0 / max(t1, 1e-3))))
    sec = R.time_gradient_psi(s, reps)   # one more warm-up call inside, then `reps` timed (omp_get_wtime)
    R.close()
    return {"value": 1.0 / sec, "unit": UNIT, "cores": ref.num_threads(), "kind": "reference",
            "sample": f"{reps} full gradient_psi calls at {cfg['N1']}^3 (calc_h=0) after 2 warm-up calls; "
                      f"FFT backend: {ref.fft_backend()}"}


def run_ours(args):
    import torch
    from barcode_b200 import chain as bc
    from barcode_b200 import inputs, multi

    info = multi.rank_info()
    world, rank, local_rank = info.world, info.rank, info.local_rank
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # each rank on the CPUs / memory of its GPU's NUMA node, before any pinned buffer is allocated
    numa = {"bound": False, "note": "--no-numa-bind"} if args.no_numa_bind else multi.bind_to_gpu_numa(local_rank)
    multi.init("nccl", info, torch.device("cuda", local_rank))

    cfg, name = workload(args.grid, args.calc_h)
    n = args.grid ** 3
    launches0 = bc.kernel_launches()
    ch = bc.Chain(bc.Params(device=local_rank, **cfg))
    prob = inputs.synthetic_problem(ch, seed=multi.chain_seed(1, rank))
    stream = torch.cuda.current_stream()
    ch.set_stream(stream.cuda_stream)
    d_s = torch.from_numpy(np.ascontiguousarray(prob["signal"]).reshape(-1)).cuda()
    d_g = torch.empty_like(d_s)
    d_p = torch.from_numpy(np.ascontiguousarray(prob["momenta"]).reshape(-1)).cuda()

    def barrier():
        multi.barrier(info)
        torch.cuda.synchronize()

    counted = {}

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l_before = bc.kernel_launches()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        counted["launches"] = bc.kernel_launches() - l_before
        barrier()
        return multi.max_over_ranks(e0.elapsed_time(e1), info, "cuda")

    def grad_step():
        ch.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr())

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms = timed(grad_step, args.steps, args.warmup)
    launches_timed = counted["launches"]
    clock_info = clocks.stop() if rank == 0 else {}
    value = world * args.steps / (ms * 1e-3)

    # the other gradient (exact adjoint when the headline is calc_h=0 and vice versa), same steps
    other_h = 4 if args.calc_h != 4 else 0
    ch2 = bc.Chain(bc.Params(device=local_rank, **{**cfg, "calc_h": other_h}))
    ch2.set_static(Power=prob["Power"], nobs=prob["nobs"], noise=prob["noise"], window=prob["window"])
    ch2.set_stream(stream.cuda_stream)
    ms_other = timed(lambda: ch2.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr()), args.steps, args.warmup)
    ch2.close()

    # the reference's SHIPPED default (data/input.par:11-13,134): SPH spline mass assignment with its exact adjoint,
    # calc_h = 2, real space -- same grid, same data, its own per-kernel split
    sph = None
    if not args.no_sph:
        ch3 = bc.Chain(bc.Params(device=local_rank, **{**cfg, "masskernel": 3, "calc_h": 2, "rsd_model": False, "sfmodel": 1}))
        ch3.set_static(Power=prob["Power"], nobs=prob["nobs"], noise=prob["noise"], window=prob["window"])
        ch3.set_stream(stream.cuda_stream)
        sph_steps = max(2, args.steps // 2)
        ms_sph = timed(lambda: ch3.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr()), sph_steps, 1)
        bc.profile_begin()
        ch3.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr())
        prof3 = bc.profile_end()
        ch3.close()
        sph = {"gradient_evals_per_s": world * sph_steps / (ms_sph * 1e-3), "ms_per_eval": ms_sph / sph_steps,
               "config": "masskernel 3 (SPH spline, h = 1 cell), calc_h 2 (likelihood_calc_h_SPH), real space, same grid",
               "per_kernel_ms": {k: v[0] for k, v in prof3.items() if v[1]}}

    # the single-precision mode (reference build option SINGLE_PREC, bgpu_f32_*): the same workload and the same
    # (float-rounded) data, reported BESIDE the FP64 headline, with its distance from the FP64 gradient
    f32 = None
    if not args.no_f32 and args.grid <= 512:
        from barcode_b200.chain_f32 import ChainF32
        c32 = ChainF32(bc.Params(device=local_rank, **cfg))
        c32.set_static(Power=prob["Power"], nobs=prob["nobs"], noise=prob["noise"], window=prob["window"])
        c32.set_stream(stream.cuda_stream)
        d_s32 = d_s.float()
        d_g32 = torch.empty_like(d_s32)
        ms_f32 = timed(lambda: c32.gradient_psi_dev(d_s32.data_ptr(), d_g32.data_ptr()), args.steps, args.warmup)
        bc.profile_begin()
        c32.gradient_psi_dev(d_s32.data_ptr(), d_g32.data_ptr())
        prof32 = bc.profile_end()
        ch.gradient_psi_dev(d_s32.double().data_ptr(), d_g.data_ptr())   # FP64 on the same rounded signal
        torch.cuda.synchronize()
        rel32 = float((d_g32.double() - d_g).norm() / d_g.norm())
        c32.hamiltonian_mass()
        d_s32b, d_p32, d_p32b = d_s32.clone(), d_p.float(), d_p.float()
        def traj32():
            d_s32b.copy_(d_s32)
            d_p32b.copy_(d_p32)
            c32.leapfrog_dev(d_s32b.data_ptr(), d_p32b.data_ptr(), 8, 1e-3)
        calls32 = max(2, args.steps // 4)
        ms_leap32 = timed(traj32, calls32, 1)
        # end to end through bgpu_f32_gradient_psi with pinned host buffers: half the PCIe bytes of the FP64 call
        import ctypes as C32
        fp = C32.POINTER(C32.c_float)
        h_s32 = d_s32.cpu().pin_memory()
        h_g32 = torch.empty(n, dtype=torch.float32).pin_memory()
        def e2e32():
            if c32.L.bgpu_f32_gradient_psi(c32._h, C32.cast(h_s32.data_ptr(), fp), C32.cast(h_g32.data_ptr(), fp)) != 0:
                raise RuntimeError(c32.L.bgpu_last_error().decode())
        for _ in range(2):
            e2e32()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e32()
        torch.cuda.synchronize()
        dt32 = multi.max_over_ranks(time.perf_counter() - t0, info, "cuda")
        c32.close()
        f32 = {"gradient_evals_per_s": world * args.steps / (ms_f32 * 1e-3), "ms_per_eval": ms_f32 / args.steps,
               "leapfrog_steps_per_s": world * calls32 * 8 / (ms_leap32 * 1e-3),
               "rel_l2_vs_fp64_gradient": rel32, "dtype": "f32 arrays and transforms, f64 reductions",
               "e2e": {"value": world * args.steps / dt32, "unit": UNIT, "h2d_bytes_per_step": n * 4,
                       "d2h_bytes_per_step": n * 4, "ms_per_step": 1e3 * dt32 / args.steps,
                       "api": "bgpu_f32_gradient_psi(host signal -> host gradpsi), pinned host buffers"},
               "whole_path_frac": 130 * n * (world * args.steps / (ms_f32 * 1e-3)) / world / 1e9 / 6550.1,
               "whole_path_note": "SURVEY 8(d)'s 260 N bytes halve with 4-byte reals: 130 N bytes per evaluation",
               "per_kernel_ms": {k: v[0] for k, v in prof32.items() if v[1]}}

    # one leapfrog step = kick, M^-1 p, drift, gradient, kick (HMC.cc:289-352)
    ch.hamiltonian_mass()
    d_s2, d_p2 = d_s.clone(), d_p.clone()
    # a trajectory of Neps = 8 steps (9 gradient evaluations, 8 M^-1 p): steps / s = 8 / time per call
    NEPS = 8
    leap_calls = max(2, args.steps // 4)
    def trajectory():  # every call starts from the same state (two device copies, < 1 % of the call)
        d_s2.copy_(d_s)
        d_p2.copy_(d_p)
        ch.leapfrog_dev(d_s2.data_ptr(), d_p2.data_ptr(), NEPS, 1e-3)
    ms_leap = timed(trajectory, leap_calls, 1)
    leap_steps = leap_calls * NEPS
    # momentum draw on the device (bgpu_draw_momenta_device: Philox normals -> r2c -> colour -> c2r); the reference
    # draws 2 N^3 GSL Gaussians serially on the host for every candidate (random.cpp:99-101)
    draw_i = [0]

    def draw():
        ch.draw_momenta_device_dev(1, draw_i[0], d_p2.data_ptr())
        draw_i[0] += 1
    draw_calls = max(2, args.steps // 2)
    ms_draw = timed(draw, draw_calls, 1)

    # one whole HMC candidate on the device (bgpu_candidate: device momentum draw, Neps = 8 trajectory, the four
    # energy evaluations), wall clock including the scalar read-backs -- what the glue's sampler loop costs
    ch.set_signal(prob["signal"])
    ch.candidate(1, 0, NEPS, 1e-5)
    t0 = time.perf_counter()
    n_cand = max(2, args.steps // 4)
    for i in range(n_cand):
        ch.candidate(1, 1 + i, NEPS, 1e-5)
    cand_ms = (time.perf_counter() - t0) * 1e3 / n_cand
    cand_ms = multi.max_over_ranks(cand_ms, info, "cuda")

    # roofline leg: the same K steps again with CUDA events around every kernel launch
    bc.profile_begin()
    for _ in range(args.steps):
        grad_step()
    prof = bc.profile_end()
    total_prof = sum(v[0] for v in prof.values())
    dom = max(prof.items(), key=lambda kv: kv[1][0])
    nh = args.grid * args.grid * (args.grid // 2 + 1)
    # x pass: every launch reads and writes the half-complex array once (32 B / element); the last one of an
    # evaluation also reads its operands -- (V/N)/P (8 B) for calc_h = 1, plus the accumulated h^ (16 B) otherwise
    x_launches = {0: 9, 1: 5, 4: 6}[args.calc_h]   # shared x pass: two x passes per component triple
    if os.environ.get("BGPU_SHARE_X") == "0":
        x_launches = {0: 12, 1: 5, 4: 8}[args.calc_h]
    x_bytes = ((x_launches - 1) * 32 + (40 if args.calc_h == 1 else 56)) * nh / x_launches
    alg_bytes = {  # algorithmic bytes per launch of each kernel class ("Kernels")
        "fft_strided_pass_y": 2 * nh * 16, "fft_strided_pass_x": x_bytes,
        "fft_r2c_zpass": n * 8 + nh * 16,
        "fft_c2r_zpass": n * 8 + nh * 16,
        # fused z+y passes: the real array and the half-complex array cross HBM once each; the
        # intermediate between the two passes stays in L2
        "fft_zy_fused_r2c": n * 8 + nh * 16, "fft_zy_fused_c2r": n * 8 + nh * 16,
        "fft_z_roundtrip": n * 8 + 2 * nh * 16,   # half-complex row in and out, the real multiplier in
        "scatter": 4 * n * 8,                       # Psi_x,y,z in, rho out
        "gather_adjoint": 7 * n * 8,
        "overdens_residual": 5 * n * 8,
        "reduce": n * 8,
        "stream": 3 * n * 8,
        "colour_momenta": n * 16 + nh * 16,
    }
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom_name, (dom_ms, dom_cnt) = dom
    achieved = alg_bytes[dom_name] / (dom_ms * 1e-3 / dom_cnt) / 1e9
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu capture, if one exists
    for tf in ("traffic_r02.json", "traffic_r01.json"):   # the newest capture that has this grid and kernel class
        try:
            with open(os.path.join(ROOT, "profiles", tf)) as f:
                traffic = json.load(f).get(str(args.grid), {}).get(dom_name)
        except (OSError, ValueError):
            traffic = None
        if traffic is not None:
            break
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes[dom_name],
        "peak_source": peak_src,
        "launches_per_step": dom_cnt / args.steps, "avg_launch_ms": dom_ms / dom_cnt,
        "share_of_step": dom_ms / total_prof,
        "per_kernel": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                           "GBps": (alg_bytes[k] / (v[0] * 1e-3 / v[1]) / 1e9) if v[1] else None}
                       for k, v in prof.items() if v[1]},
        "whole_path": {"algorithmic_bytes_per_eval": 260 * n, "achieved_GBps": 260 * n * value / world / 1e9,
                       "frac": 260 * n * value / world / 1e9 / peak_gbs,
                       "note": "SURVEY 8(d): 260 N bytes per exact-adjoint evaluation; calc_h=0 does 12 FFTs instead of 8"},
    }
    exact_rate = value if args.calc_h == 4 else (world * args.steps / (ms_other * 1e-3) if other_h == 4 else None)
    if exact_rate is not None:
        # the evaluation SURVEY 8(d)'s 260 N bytes describe: the exact adjoint (8 FFTs, scatter, gather)
        roofline["whole_path_exact_adjoint"] = {"gradient_evals_per_s": exact_rate,
                                                "achieved_GBps": 260 * n * exact_rate / world / 1e9,
                                                "frac": 260 * n * exact_rate / world / 1e9 / peak_gbs}

    # end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    h_s = torch.from_numpy(np.ascontiguousarray(prob["signal"]).reshape(-1)).pin_memory()
    h_g = torch.empty(n, dtype=torch.float64).pin_memory()
    import ctypes as C
    dp = C.POINTER(C.c_double)

    def e2e_step():
        rc = ch.L.bgpu_gradient_psi(ch._h, C.cast(h_s.data_ptr(), dp), C.cast(h_g.data_ptr(), dp))
        if rc != 0:
            raise RuntimeError(ch.L.bgpu_last_error().decode())

    for _ in range(max(1, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = multi.max_over_ranks(time.perf_counter() - t0, info, "cuda")
    e2e = {"value": world * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * 8,
           "d2h_bytes_per_step": n * 8, "ms_per_step": 1e3 * dt / args.steps,
           "api": "bgpu_gradient_psi(host signal -> host gradpsi), pinned host buffers"}

    # the call the reference's driver makes once per HMC candidate: Hamiltonian_EoM -> bgpu_leapfrog with host
    # s_i, p_i in and s_f, p_f out (Neps = 8); Neps + 1 gradient evaluations per call
    neps = 8
    h_p = torch.from_numpy(np.ascontiguousarray(prob["momenta"]).reshape(-1)).pin_memory()
    h_sf = torch.empty(n, dtype=torch.float64).pin_memory()
    h_pf = torch.empty(n, dtype=torch.float64).pin_memory()

    def traj():
        rc = ch.L.bgpu_leapfrog(ch._h, C.cast(h_s.data_ptr(), dp), C.cast(h_p.data_ptr(), dp), neps, 1e-3,
                                C.cast(h_sf.data_ptr(), dp), C.cast(h_pf.data_ptr(), dp))
        if rc != 0:
            raise RuntimeError(ch.L.bgpu_last_error().decode())

    traj()
    barrier()
    t0 = time.perf_counter()
    ntraj = max(1, args.steps // 5)
    for _ in range(ntraj):
        traj()
    torch.cuda.synchronize()
    dt_traj = multi.max_over_ranks(time.perf_counter() - t0, info, "cuda")
    e2e["trajectory"] = {"api": "bgpu_leapfrog(host s_i, p_i -> host s_f, p_f), Neps = 8", "gradient_evals_per_s":
                         world * ntraj * (neps + 1) / dt_traj, "leapfrog_steps_per_s": world * ntraj * neps / dt_traj,
                         "h2d_bytes_per_call": 2 * n * 8, "d2h_bytes_per_call": 2 * n * 8}
    # the glue's device-resident sampler loop (BARCODE_GPU_DEVICE_RNG=1): one call per HMC candidate, scalars only
    e2e["candidate"] = {"api": "bgpu_candidate(seed, draw, Neps = 8, eps) -> 6 energies; signal and momenta stay on the "
                               "device (momentum draw, trajectory, kinetic + psi at both ends)",
                        "ms_per_candidate": cand_ms, "gradient_evals_per_s": world * (NEPS + 1) / (cand_ms * 1e-3),
                        "leapfrog_steps_per_s": world * NEPS / (cand_ms * 1e-3),
                        "h2d_bytes_per_call": 0, "d2h_bytes_per_call": 6 * 8 + 8}
    e2e["numa"] = numa

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(args, cfg)
    ch.close()
    slab = slab_leg(args, info) if (world > 1 and not args.no_slab) else None
    if rank == 0 and world == 1 and not args.no_e2e_chains:
        # independent chains interleaved on the GPU hide the PCIe transfers of one under the kernels of another;
        # measured in a child process with a timeout, reported beside (not instead of) the single-chain e2e.value
        e2e["interleaved_chains"] = e2e_chains_leg(args)

    grid512 = None
    if rank == 0 and world == 1 and args.grid == 256 and not args.no_512:
        grid512 = grid512_leg(args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "grid": args.grid, "calc_h": args.calc_h,
                       "parallelism": f"{world} independent chain(s), one per GPU, no data-path collective",
                       "l2": f"inputs larger than L2: {12 * n * 8 / 2**20:.0f} MiB of FP64 arrays are streamed per "
                             "evaluation against 126 MB of L2; no explicit flush"},
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "gpu_launches": int(launches_timed),
            "clocks": clock_info,
            "slab": slab,
            "also": {
                f"gradient_evals_per_s_calc_h_{other_h}": world * args.steps / (ms_other * 1e-3),
                "leapfrog_steps_per_s": world * leap_steps / (ms_leap * 1e-3),
                "device_momentum_draw_ms": ms_draw / draw_calls,
                "hmc_candidate_ms_neps8_device_resident": cand_ms,
                "sph_default_config": sph,
                "grid_512": grid512,
                "single_precision_mode": f32,
                "kernel_launches_total": int(bc.kernel_launches() - launches0),
            },
        }
        print(json.dumps(line))
    multi.finalize()
    return 0


def run_e2e_chains(args):
    """Host-buffer gradient evaluations of `--chains` independent chains interleaved on ONE GPU.

    A single synchronous bgpu_gradient_psi(host -> host) call is PCIe-bound: the signal goes up, the evaluation
    runs, the gradient comes down (e2e.value).  Independent chains -- the reference's own multi-chain mode, SURVEY
    8e -- do not depend on each other, so one host thread per chain, each with its own handle, stream and pinned
    buffers, lets chain B's transfers ride under chain A's kernels (two copy engines + the SMs).  Every call is
    still the reference-facing C-ABI call with host buffers, H2D and D2H inside the timed region; value = all
    chains' evaluations / wall time.  Runs in its own process (bench.py spawns it with a timeout) so that the
    headline numbers cannot depend on it."""
    import ctypes as C
    import threading
    import torch
    from barcode_b200 import chain as bc
    from barcode_b200 import inputs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    cfg, name = workload(args.grid, args.calc_h)
    n = args.grid ** 3
    nch = max(1, args.chains)
    chains = [bc.Chain(bc.Params(device=dev, **cfg)) for _ in range(nch)]
    prob = inputs.synthetic_problem(chains[0], seed=1)
    for ch in chains[1:]:
        ch.set_static(Power=prob["Power"], nobs=prob["nobs"], noise=prob["noise"], window=prob["window"])
    dp = C.POINTER(C.c_double)
    sig = np.ascontiguousarray(prob["signal"]).reshape(-1)
    h_s = [torch.from_numpy(sig.copy()).pin_memory() for _ in range(nch)]
    h_g = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(nch)]

    def step(i):
        ch = chains[i]
        rc = ch.L.bgpu_gradient_psi(ch._h, C.cast(h_s[i].data_ptr(), dp), C.cast(h_g[i].data_ptr(), dp))
        if rc != 0:
            raise RuntimeError(ch.L.bgpu_last_error().decode())

    # warm-up one chain at a time (this also initialises the library's per-kernel launch configuration serially)
    for i in range(nch):
        for _ in range(max(1, args.warmup)):
            step(i)
    torch.cuda.synchronize()
    ref_g = h_g[0].clone()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(0)
    torch.cuda.synchronize()
    dt_single = time.perf_counter() - t0

    errors = []
    gate = threading.Barrier(nch + 1)

    def worker(i):
        try:
            gate.wait()
            for _ in range(args.steps):
                step(i)
        except Exception as e:  # noqa: BLE001
            errors.append(f"chain {i}: {e!r}")

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(nch)]
    for t in threads:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if errors:
        raise RuntimeError("; ".join(errors))
    # every chain evaluated the same signal on the same data: the results must agree with the lone call
    # (to rounding: the scatter's atomics are not ordered)
    worst = max(float(torch.linalg.vector_norm(g - ref_g) / torch.linalg.vector_norm(ref_g)) for g in h_g)
    for ch in chains:
        ch.close()
    print(json.dumps({
        "value": nch * args.steps / dt, "unit": UNIT, "chains_per_gpu": nch, "steps_per_chain": args.steps,
        "ms_per_eval": 1e3 * dt / (nch * args.steps), "single_chain_same_process": args.steps / dt_single,
        "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
        "max_rel_l2_vs_lone_call": worst, "workload": name,
        "api": "bgpu_gradient_psi(host signal -> host gradpsi), pinned host buffers, one host thread + handle + "
               "stream per chain"}))
    return 0 if worst < 1e-10 else 1


def e2e_chains_leg(args, timeout_s=240):
    """Run run_e2e_chains in a child process; returns its JSON or {"error": ...} -- never raises, never hangs."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--mode", "e2e_chains", "--grid", str(args.grid),
           "--calc-h", str(args.calc_h), "--steps", str(args.steps), "--warmup", str(args.warmup),
           "--chains", str(args.chains)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if not lines:
            return {"error": f"exit {r.returncode}: {r.stderr.strip()[-300:]}"}
        out = json.loads(lines[-1])
        if r.returncode != 0:
            out["error"] = f"exit {r.returncode}"
        return out
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout_s} s"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def grid512_leg(args, timeout_s=300):
    """The same bench at 512^3 (BASELINE.json's target size: >= 1 evaluation / s on one B200) in a child process, condensed
    to the numbers `also` carries; returns {"error": ...} instead of raising or hanging."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--grid", "512", "--calc-h", str(args.calc_h), "--steps", "5",
           "--warmup", "3", "--no-cpu-baseline", "--no-e2e-chains", "--no-sph", "--no-f32", "--no-512"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if not lines:
            return {"error": f"exit {r.returncode}: {r.stderr.strip()[-300:]}"}
        d = json.loads(lines[-1])
        rf = d["roofline"]
        return {"gradient_evals_per_s": d["value"], "ms_per_eval": d["ms_per_step"], "workload": d["config"]["workload"],
                "whole_path_frac": rf["whole_path"]["frac"],
                "exact_adjoint": rf.get("whole_path_exact_adjoint"),
                "e2e_evals_per_s": d["e2e"]["value"], "leapfrog_steps_per_s": d["also"]["leapfrog_steps_per_s"],
                "dominant_kernel": {"kernel": rf["kernel"], "frac": rf["frac"], "share_of_step": rf["share_of_step"]},
                "per_kernel": {k: {"ms_per_step": round(v["ms_per_step"], 4), "GBps": round(v["GBps"], 1)}
                               for k, v in rf["per_kernel"].items()},
                "clocks": d.get("clocks")}
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout_s} s"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def slab_measure(grid, calc_h, steps, warmup, info, with_clocks=True):
    """ONE chain across all ranks (x-slab decomposition, barcode_b200/slab.py), gradient evaluations of that one
    chain per second, device-timed, max over ranks.  The process group must be up.  Returns the result dict (same
    on every rank up to the per-rank profile)."""
    import torch
    from barcode_b200 import chain as bc
    from barcode_b200 import inputs, multi, slab

    world, rank, local_rank = info.world, info.rank, info.local_rank
    cfg, name = workload(grid, calc_h)
    sc = slab.SlabChain.create(bc.Params(device=local_rank, **cfg), rank, world)
    prob = inputs.slab_problem(sc, seed=1)
    stream = torch.cuda.current_stream()
    sc.set_stream(stream.cuda_stream)
    d_s = torch.from_numpy(np.ascontiguousarray(prob["signal"]).reshape(-1)).cuda()
    d_g = torch.empty_like(d_s)

    def barrier():
        multi.barrier(info)
        torch.cuda.synchronize()

    def step():
        sc.gradient_psi_dev(d_s.data_ptr(), d_g.data_ptr())

    for _ in range(warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0 and with_clocks:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = bc.kernel_launches()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    launches = bc.kernel_launches() - l0
    barrier()
    ms = multi.max_over_ranks(e0.elapsed_time(e1), info, "cuda")
    clock_info = clocks.stop() if (rank == 0 and with_clocks) else {}
    bc.profile_begin()
    for _ in range(steps):
        step()
    prof = bc.profile_end()
    n_loc = sc.N
    nh_loc = sc.Nhalf
    # bytes each rank sends per all-to-all (its whole k-space slab minus the block it keeps)
    a2a_bytes = nh_loc * 16 * (world - 1) / max(world, 1)
    per_kernel = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps}
                  for k, v in prof.items() if v[1]}
    fused_transpose = os.environ.get("BGPU_SLAB_P2P", "1") != "0" and sc.fused_transpose()
    nvlink = {}
    if "all_to_all" in per_kernel and world > 1:
        t = prof["all_to_all"][0] * 1e-3 / prof["all_to_all"][1]
        transposes = prof["all_to_all"][1] / steps
        if fused_transpose:
            # the NVLink traffic rides inside the strided pass (TMA stores into the peers' receive buffers);
            # what is timed under this name is only the cross-rank barrier that follows
            per_kernel["all_to_all"]["note"] = "fused transpose: this entry is the cross-rank barrier only"
            tp = (prof["fft_strided_pass_y"][0] + prof["fft_strided_pass_x"][0]) * 1e-3 / (2 * prof["all_to_all"][1])
            gbs = a2a_bytes / tp / 1e9
            nvlink = {"GBps_sent_per_gpu": gbs, "frac_of_nvlink_900GBps": gbs / 900.0,
                      "where": "inside the transposing strided pass (lower bound: the pass also transforms)",
                      "transposes_per_eval": transposes}
        else:
            gbs = a2a_bytes / t / 1e9
            nvlink = {"GBps_sent_per_gpu": gbs, "frac_of_nvlink_900GBps": gbs / 900.0, "where": "NCCL grouped send/recv",
                      "transposes_per_eval": transposes,
                      "share_of_step": prof["all_to_all"][0] / max(1e-9, sum(v[0] for v in prof.values()))}
    # leapfrog steps / s of the slab chain (k-space form: the trajectory's s^ and p^ stay in the transposed k-space
    # slabs; a step costs 10 transposes instead of 14): 2 trajectories of 4 steps from a device-drawn momentum
    leap = None
    try:
        d_p = torch.empty_like(d_s)
        sc.draw_momenta_device_dev(11, 1, d_p.data_ptr())
        d_s2, d_p2 = d_s.clone(), d_p.clone()

        def traj():
            d_s2.copy_(d_s)
            d_p2.copy_(d_p)
            sc.leapfrog_dev(d_s2.data_ptr(), d_p2.data_ptr(), 4, 1e-6)   # a tiny step: the rate does not depend on it, the halo does

        traj()
        barrier()
        e0.record(stream)
        for _ in range(2):
            traj()
        e1.record(stream)
        barrier()
        ms_leap = multi.max_over_ranks(e0.elapsed_time(e1), info, "cuda")
        leap = {"leapfrog_steps_per_s": 8 / (ms_leap * 1e-3), "ms_per_step": ms_leap / 8,
                "trajectory": "Neps = 4: 5 kicks, 4 drifts, 4 end transforms", "finite": bool(torch.isfinite(d_p2).all().item())}
    except Exception as e:  # noqa: BLE001
        leap = {"error": repr(e)[:300]}
    sc.close()
    return {
        "value": steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "grid": grid, "calc_h": calc_h,
        "leapfrog": leap,
        "workload": name, "n_gpus": world, "steps": steps, "warmup": warmup, "scaling": "strong",
        "parallelism": f"one chain, x-slab decomposed over {world} GPU(s): distributed FFT with "
                       + ("the transpose fused into the strided pass (TMA stores over NVLink peer memory)"
                          if fused_transpose else "NCCL all-to-all transposes") + ", halo-exchanged mass assignment",
        "per_kernel": per_kernel, "nvlink": nvlink, "a2a_bytes_sent_per_gpu_per_transpose": a2a_bytes,
        "local_cells": n_loc, "gpu_launches": int(launches), "clocks": clock_info,
    }


def slab_parity(grid, info):
    """The slab-decomposed chain against the single-GPU chain (which the parity tests pin to the reference) on the
    same inputs: forward density, both energies and the calc_h = 0 / 4 gradients.  Returns the worst relative L2."""
    import torch
    import torch.distributed as dist
    from barcode_b200 import chain as bc
    from barcode_b200 import inputs, slab
    N = grid
    L = inputs.box_length(N)
    rng = np.random.default_rng(3)
    P = inputs.power_on_grid(*inputs.load_pk_table(), N, L)
    n = N ** 3
    nobs = np.maximum(0.0, 1.0 + 0.3 * rng.standard_normal(n)).reshape(N, N, N)
    one = np.ones((N, N, N))
    w = rng.standard_normal((N, N, N))
    s = 0.5 * np.fft.irfftn(np.fft.rfftn(w) * np.sqrt(np.maximum(P[:, :, :N // 2 + 1], 0) * n / L ** 3), s=(N, N, N),
                            axes=(0, 1, 2))

    def gather(local):
        t = torch.from_numpy(np.ascontiguousarray(local)).cuda()
        out = [torch.empty_like(t) for _ in range(info.world)]
        dist.all_gather(out, t)
        return torch.cat(out, 0).cpu().numpy()

    worst = 0.0
    for calc_h in (0, 4):
        kw = dict(N1=N, L1=L, masskernel=1, likelihood=1, rsd_model=True, calc_h=calc_h, mass_type=1, sfmodel=2)
        sc = slab.SlabChain.create(bc.Params(device=info.local_rank, **kw), info.rank, info.world)
        sc.set_static(Power=sc.local(P), nobs=sc.local(nobs), noise=sc.local(one), window=sc.local(one))
        g = gather(sc.gradient_psi(sc.local(s)))
        pp, pl, dX = sc.psi(sc.local(s))
        dX = gather(dX)
        sc.close()
        if info.rank == 0:
            with bc.Chain(bc.Params(device=info.local_rank, **kw)) as ch:
                ch.set_static(Power=P, nobs=nobs, noise=one, window=one)
                g0 = ch.gradient_psi(s)
                pp0, pl0, dX0 = ch.psi(s)
            rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))  # noqa: E731
            worst = max(worst, rel(g, g0), rel(dX, dX0), abs(pp - pp0) / abs(pp0), abs(pl - pl0) / abs(pl0))
    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def slab_leg(args, info):
    """N > 1 GPUs: besides the independent chains, ONE chain slab-decomposed over all ranks (BASELINE.json configs[3],
    [4]): parity against the single-GPU chain at 128^3 and 256^3 first, then 512^3 -- and 1024^3 where it fits (8 GPUs)."""
    out = {}
    try:
        out["parity_max_rel"] = {str(g): slab_parity(g, info) for g in (128, 256)}
        steps = max(3, args.steps // 2)
        out["512"] = slab_measure(512, args.calc_h, steps, 2, info, with_clocks=False)
        if info.world >= 8:
            out["1024"] = slab_measure(1024, args.calc_h, max(2, steps // 2), 1, info, with_clocks=False)
        elif info.world >= 4:
            out["1024"] = {"skipped": "1024^3 needs ~22 N^3 doubles = 190 GB over the ranks plus buffers: run on 8 GPUs"}
    except Exception as e:  # noqa: BLE001 -- the chains leg's numbers must survive a failure here
        out["error"] = repr(e)
    return out


def run_slab(args):
    """`--mode slab`: only the slab-decomposed chain (strong scaling), one JSON line."""
    import torch
    from barcode_b200 import multi

    info = multi.rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    torch.cuda.set_device(info.local_rank)
    multi.init("nccl", info, torch.device("cuda", info.local_rank))
    r = slab_measure(args.grid, args.calc_h, args.steps, args.warmup, info)
    if info.rank == 0:
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": info.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["workload"], "grid": args.grid, "calc_h": args.calc_h, "parallelism": r["parallelism"],
                       "l2": "inputs larger than L2; no explicit flush"},
            "per_kernel": r["per_kernel"], "nvlink": r["nvlink"], "local_cells": r["local_cells"],
            "leapfrog": r.get("leapfrog"), "gpu_launches": r["gpu_launches"], "clocks": r["clocks"],
        }
        print(json.dumps(line))
    multi.finalize()
    return 0


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON result: everything libraries print on file descriptor 1 while the
    # bench runs (NCCL's "NCCL version ..." banner on some boxes) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        rc = run_reference(args)
    elif args.mode == "e2e_chains":
        rc = run_e2e_chains(args)
    elif args.mode == "slab":
        rc = run_slab(args)
    else:
        rc = run_ours(args)
    sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())
