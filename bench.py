#!/usr/bin/env python
"""bench.py -- SimPLIC alpha-advection throughput (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm on the host cores

A "step" = one solveVofEqu::reconstruct() + advect() over the whole mesh (the two calls the
reference driver times, plicVof.H:37-52).  Workload at N=1: BASELINE.json configs[1], the
3-D LeVeque deformation test, sphere r=0.15 in the unit cube, 256^3 hexes, FP64, controls of
tutorials/test/plicVofAdvectionFoam/system/fvSolution, fixed dt = 0.2/N (Co ~ 0.5).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from geometricvofext_b200 import capi, fields, mesh as meshmod  # noqa: E402
from geometricvofext_b200.solver import SolveVofEqu  # noqa: E402

CONTROLS = {"nAlphaBounds": 3, "snapTol": 0, "clip": False, "mixedCellTol": 1e-8, "orientationMethod": "LS",
            "splitWarpedFace": False}
METRIC = "alpha-advection cell-updates/sec"
UNIT = "cell-updates/s"
PERIOD = 6.0


def b_alg(m):
    """Algorithmic bytes of one step (SURVEY.md 8d / BASELINE.md): 24 nC + 16 nF + 4 (2 nIF + nBF)."""
    nC, nF, nIF = m.n_cells, m.n_faces, m.n_internal_faces
    return 24 * nC + 16 * nF + 4 * (2 * nIF + (nF - nIF))


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_case(n, lo=None, hi=None, proc_nbr=None, cut_as_wall=False):
    m = meshmod.hex_block(n, lo=lo or (0, 0, 0), hi=hi, proc_nbr=proc_nbr, cut_as_wall=cut_as_wall)
    a0 = fields.sphere_alpha_quadrature(m)
    return m, a0


def velocity_fields(s, t, dt):
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    f = fields.u_factor(t, dt, PERIOD)
    U = fields.leveque_velocity(C_) * f
    phi = fields.face_flux(Cf, Sf) * f
    return U, phi


def cpu_oracle_rate(m, a0, U, phi, dt, steps, keep=False, controls=None, Ub=None):
    """The CPU restatement of the reference algorithm, timed on this box's host (1 core).  keep=True also returns
    its state after the 1 + steps steps it ran (alpha, alphaPhi, the interface-cell list of that alpha) so that the
    GPU run of the same steps can be compared AT THE SIZE THE NUMBER IS QUOTED ON (checker use of oracle/, as in tests/)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build as oracle_build
    lib = capi.load(oracle_build.build_oracle())
    so = SolveVofEqu(m, controls or CONTROLS, lib=lib)
    so.setAlpha(a0)
    so.setPhi(phi)
    so.setU(U, Ub)
    so.reconstruct()   # warm-up (page faults, first-touch)
    so.advect(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        so.reconstruct()
        so.advect(dt)
    el = time.perf_counter() - t0
    rate = m.n_cells * steps / el
    state = None
    if keep:
        so.reconstruct()
        state = {"alpha": so.alpha(), "alphaPhi": so.alphaPhi(), "mixed": so.mixedCells(), "status": so.cellStatus(),
                 "volume": so.volume(), "steps": steps + 1}
    so.close()
    return rate, el, state


def parity_block(s, m, a0, dt, ref, label):
    """GPU run of the same steps as the oracle's (same inputs), compared field by field."""
    s.setAlpha(a0)
    for _ in range(ref["steps"]):
        s.step(dt)
    s.reconstruct()
    a, ap = s.alpha(), s.alphaPhi()
    mixed_equal = bool(np.array_equal(s.mixedCells(), ref["mixed"]) and np.array_equal(s.cellStatus(), ref["status"]))
    vol = s.volume()
    out = {"case": label, "steps": ref["steps"], "cells": m.n_cells, "mixed_cells": int(ref["mixed"].size),
           "mixed_set_equal": mixed_equal, "max_abs_alpha": float(np.abs(a - ref["alpha"]).max()),
           "max_abs_alphaPhi": float(np.abs(ap - ref["alphaPhi"]).max()),
           "bitwise_alpha": bool(np.array_equal(a, ref["alpha"])), "bitwise_alphaPhi": bool(np.array_equal(ap, ref["alphaPhi"])),
           "volume_rel": float(abs(vol - ref["volume"]) / abs(ref["volume"])), "checker": "oracle/ (CPU restatement), same inputs"}
    return out


SCHED_SETTLE_STEPS = 40


def settle_schedule(s, a_start, dt, args):
    """svof_step_device picks its two-stream schedule at run time (streaming kernel forked at the near sets and uncapped, or
    forked after plane positioning with 4 resident CTAs per SM): it measures 12 steps of each during the first 36 calls.  Those
    calls are made here, untimed, from the same start field; the timed windows then start from that field again.
    Returns None when the selection is off (an explicit --overlap)."""
    if args.overlap >= 0 or args.no_sched_auto:
        return None
    s.setOption("sched_auto", 1)
    s.setAlpha(a_start)
    for _ in range(SCHED_SETTLE_STEPS):
        s.step(dt)
    s.synchronize()
    v = int(s.info(capi.I_SCHEDULE))
    return {"first_window": v, "meaning": "100*fork + resident streaming CTAs per SM (0 = uncapped); chosen by the library from 12 timed "
                                          "steps of each schedule (svof_set_option sched_auto)"}


def advance_on_device(s, U0, phi0, t_end, dt, torch):
    """plicVof.H loop with phi(t), U(t) refreshed on the device every step (updateU.H:59-69 scales the steady field)."""
    dev = torch.device("cuda", 0)
    tU0, tphi0 = torch.as_tensor(U0, device=dev), torch.as_tensor(phi0, device=dev)
    tU, tphi = torch.empty_like(tU0), torch.empty_like(tphi0)
    t, k = 0.0, 0
    while t < t_end - 1e-12:
        t += dt
        f = fields.u_factor(t, dt, PERIOD)
        torch.mul(tU0, f, out=tU)
        torch.mul(tphi0, f, out=tphi)
        torch.cuda.current_stream().synchronize()
        s._chk(s.lib.svof_set_phi_device(s._h, tphi.data_ptr()))
        s._chk(s.lib.svof_set_U_device(s._h, tU.data_ptr(), None))
        s.step(dt)
        s.synchronize()
        k += 1
    return t, k


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from geometricvofext_b200 import multigpu
        return multigpu.bench(args, CONTROLS, METRIC, UNIT)
    n = args.n
    t_setup = time.perf_counter()
    m, a0 = build_case(n)
    t_mesh = time.perf_counter() - t_setup
    s = SolveVofEqu(m, CONTROLS)
    t_create = time.perf_counter() - t_setup - t_mesh
    dt = 0.2 / n
    U, phi = velocity_fields(s, dt, dt)
    Ub = np.zeros((s.nBF, 3))
    s.setAlpha(a0)
    s.setPhi(phi)
    s.setU(U, Ub)
    s.synchronize()
    setup_s = time.perf_counter() - t_setup

    lib, h = s.lib, s._h
    # ---- kernel-timing leg: plain launches; the library keeps CUDA events around every launch of the streaming
    #      kernel (nothing runs beside it), which is what the roofline line below is computed from ----
    s.setOption("overlap", 0)   # the roofline kernel is timed ALONE: single-stream schedule for this leg only
    for _ in range(args.warmup):
        s.reconstruct()
        s.advect(dt)
    s.synchronize()
    clocks = ClockSampler()
    clocks.start()          # sampled over every timed leg below (kernel timing, device-resident, end-to-end)
    d0, dn0 = s.info(capi.I_DENSE_KERNEL_MS), s.info(capi.I_DENSE_KERNEL_LAUNCHES)
    lib.svof_mark(h, 2)
    for _ in range(args.steps):
        s.reconstruct()
        s.advect(dt)
    lib.svof_mark(h, 3)
    ms_plain = C.c_double()
    lib.svof_elapsed_ms(h, 2, 3, C.byref(ms_plain))
    s.synchronize()
    dense_ms = (s.info(capi.I_DENSE_KERNEL_MS) - d0) / max(1.0, s.info(capi.I_DENSE_KERNEL_LAUNCHES) - dn0)
    recon_s, adv_s = s.reconstructionTime(), s.advectionTime()
    s.setOption("overlap", 1 if args.overlap < 0 else args.overlap)   # the product's default schedule (two streams) from here on
    sched_sel = settle_schedule(s, a0, dt, args)
    # ---- device-resident leg: inputs already in HBM, the call a user makes for that case (svof_step_device:
    #      reconstruct + advect replayed as one CUDA graph) ----
    s.setAlpha(a0)          # every leg runs the same steps of the same problem: from t = 0
    for _ in range(2):      # the first step of each buffer parity runs with plain launches and captures its graph
        s.step(dt)
    for _ in range(args.warmup):
        s.step(dt)
    s.synchronize()
    l0 = s.info(capi.I_GPU_LAUNCHES)
    lib.svof_mark(h, 0)
    for _ in range(args.steps):
        s.step(dt)
    lib.svof_mark(h, 1)
    ms = C.c_double()
    lib.svof_elapsed_ms(h, 0, 1, C.byref(ms))
    s.synchronize()
    total_ms = ms.value
    launches = int(s.info(capi.I_GPU_LAUNCHES) - l0)
    value = m.n_cells * args.steps / (total_ms * 1e-3)
    n_mixed, n_near = int(s.info(capi.I_N_MIXED)), int(s.info(capi.I_N_NEAR))

    # ---- end-to-end leg: host buffers in, results out, every step ------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    phi_h, U_h, Ub_h = capi.pinned_array(lib, (s.nF,)), capi.pinned_array(lib, (s.nC, 3)), capi.pinned_array(lib, (max(s.nBF, 1), 3))
    a_out, ap_out = capi.pinned_array(lib, (s.nC,)), capi.pinned_array(lib, (s.nF,))
    phi_h[:] = phi
    U_h[:] = U
    Ub_h[:] = 0
    s.setAlpha(a0)
    for _ in range(1 + args.warmup):   # warm-up (the first call after set_alpha returns full fields, later ones deltas)
        s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
    # the caller's phi and U change on every face / in every cell between calls (as a flow solver's do); only the calls are timed
    e2e_s = 0.0
    for k in range(e2e_steps):
        fk = 1.0 - 1e-3 * (k + 1)
        np.multiply(phi, fk, out=phi_h)
        np.multiply(U, fk, out=U_h)
        t0 = time.perf_counter()
        s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
        e2e_s += time.perf_counter() - t0
    clk = clocks.stop()
    e2e_val = m.n_cells * e2e_steps / e2e_s
    h2d = int(s.info(capi.I_H2D_BYTES))   # bytes the library actually copied in the last step (sparse_io: only the rows of U
    d2h = int(s.info(capi.I_D2H_BYTES))   # the interpolation reads go in; alpha/alphaPhi come back as bitwise deltas)

    # ---- roofline of the dominant (streaming) kernel ------------------------------------------
    peak, peak_src = measured_peak()
    B = b_alg(m)
    achieved = B / (dense_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_dense_update_dram_bytes_per_launch_%d" % n)
        except Exception:
            traffic = None

    # ---- CPU baseline: the oracle on the same workload, bounded sample; its fields are kept to check the GPU run ----
    cpu, parity = None, []
    if not args.no_cpu:
        cs = max(1, args.cpu_steps)
        rate, el, ref = cpu_oracle_rate(m, a0, U, phi, dt, cs, keep=True)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "same %d^3 workload and fields, %d steps (%.1f s) of the single-threaded CPU restatement "
                         "of the reference algorithm (oracle/), OpenFOAM itself is not installable here" % (n, cs, el)}
        s.setPhi(phi)
        s.setU(U, Ub)
        parity.append(parity_block(s, m, a0, dt, ref, "t = 0 window"))

    # ---- second window: the stretched interface near t = 1.5 (maximum deformation: ~3.5x the interface cells) ----
    late = None
    if args.late_steps > 0:
        import torch
        U1, phi1 = fields.leveque_velocity(s.field(capi.F_C)), None
        phi1 = fields.face_flux(s.field(capi.F_CF), s.field(capi.F_SF))
        s.setAlpha(a0)
        t_adv0 = time.perf_counter()
        t_late, k_adv = advance_on_device(s, U1, phi1, args.late_time, dt, torch)
        adv_s_wall = time.perf_counter() - t_adv0
        del U1, phi1
        s.setPhi(phi)        # the timed window uses the same frozen flux field as the first one
        s.setU(U, Ub)
        a_late = s.alpha()
        if sched_sel is not None:       # 4.5 times the interface cells of the first window: let the library measure again,
            s.setOption("sched_retune", 1)   # then start the window from the same field (as settle_schedule does for the first one)
            for _ in range(SCHED_SETTLE_STEPS):
                s.step(dt)
            s.synchronize()
            sched_sel["late_window"] = int(s.info(capi.I_SCHEDULE))
            s.setAlpha(a_late)
        for _ in range(2 + args.warmup):
            s.step(dt)
        s.synchronize()
        lib.svof_mark(h, 4)
        for _ in range(args.late_steps):
            s.step(dt)
        lib.svof_mark(h, 5)
        ms_late = C.c_double()
        lib.svof_elapsed_ms(h, 4, 5, C.byref(ms_late))
        s.synchronize()
        late = {"window": "interface advanced to t = %.3f (%d steps with phi(t), U(t) refreshed on the device, %.1f s), then %d timed "
                          "steps with the frozen flux field of the first window" % (t_late, k_adv, adv_s_wall, args.late_steps),
                "steps": args.late_steps, "ms_per_step": ms_late.value / args.late_steps,
                "value": m.n_cells * args.late_steps / (ms_late.value * 1e-3), "mixed_cells": int(s.info(capi.I_N_MIXED)),
                "near_cells": int(s.info(capi.I_N_NEAR)),
                "step_frac": (B / (ms_late.value / args.late_steps * 1e-3) / 1e9) / peak,
                "error_flags": int(s.info(capi.I_ERROR_FLAGS))}
        if not args.no_cpu and args.late_cpu_steps > 0:
            _, _, ref_late = cpu_oracle_rate(m, a_late, U, phi, dt, max(0, args.late_cpu_steps - 1), keep=True)
            parity.append(parity_block(s, m, a_late, dt, ref_late, "t = %.2f window" % t_late))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "LeVeque 3-D deformation, sphere r=0.15, %d^3 hex blockMesh (BASELINE.json configs[1])" % n,
                   "cells": m.n_cells, "faces": m.n_faces, "dt": dt, "controls": CONTROLS, "mixed_cells": n_mixed,
                   "near_cells": n_near, "l2": "inputs larger than L2 (%.2f GB of fields+connectivity per step)" % (B / 1e9),
                   "timing": "CUDA events on the handle's stream around %d steps" % args.steps,
                   "schedule": "svof_step_device: one CUDA-graph launch per step (%d kernels inside), streaming kernel on a second "
                               "stream beside the interface kernels" % (launches // max(1, args.steps)),
                   "schedule_selected": sched_sel,
                   "ms_per_step_plain_launches": ms_plain.value / args.steps,
                   "reconstruct_ms": 1e3 * recon_s / (args.steps + args.warmup), "advect_ms": 1e3 * adv_s / (args.steps + args.warmup),
                   "setup_s": setup_s,
                   "setup_breakdown_s": {"mesh_and_alpha0_in_python": t_mesh, "svof_create": t_create,
                                         "fields_and_uploads": setup_s - t_mesh - t_create}},
        "clocks": clk,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                "full_field_bytes_per_step": {"h2d": 8 * (s.nF + 3 * s.nC + 3 * s.nBF), "d2h": 8 * (s.nC + s.nF)},
                "note": "svof_step_host with pinned host buffers for phi, U, Ub in and alpha, alphaPhi out; phi and U are rescaled on "
                        "the host between calls (every entry changes), only the calls are timed.  With pinned caller buffers the "
                        "kernels read the caller's phi / U and write its alpha / alphaPhi directly over PCIe (zero copy, one host "
                        "wait per call): phi on boundary faces and on internal faces with alpha != 0 in one of the two cells (the "
                        "others multiply an exactly zero alpha), the rows of U next to cut cells, the alpha cells whose bits "
                        "changed and alphaPhi on the marked faces; the caller's output buffers hold the complete new fields after "
                        "every call (bitwise equal to full-field calls: "
                        "tests/test_gpu_parity.py::test_step_host_sparse_phi_upload_is_bitwise_the_full_upload).  The byte counts "
                        "are those of the last timed call: with snapTol 0 the support of alpha grows by one cell layer per step."},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_dense_update", "kernel_ms": dense_ms,
                     "algorithmic_bytes_per_launch": B, "peak_source": peak_src,
                     "step_frac": (B / (total_ms / args.steps * 1e-3) / 1e9) / peak,
                     "note": "kernel_ms: CUDA events around every launch of the streaming kernel over the %d plain-launch steps timed "
                             "just before the graph leg (same kernel, nothing runs beside it); step_frac: algorithmic bytes of the "
                             "whole step / graph-leg step time / peak" % args.steps},
        "cpu_baseline": cpu,
        "parity": parity,
        "late_window": late,
    }
    print(json.dumps(line))


# ---- the other BASELINE.json workloads (single GPU): --workload kelvin | dambreak --------------------------------
DAMBREAK_CONTROLS = {"nAlphaBounds": 3, "mixedCellTol": 1e-10, "snapTol": 1e-12, "clip": True, "orientationMethod": "LS",
                     "splitWarpedFace": False}   # damBreakWithObstacle/system/fvSolution:20-35


def make_workload(name, size):
    """mesh, controls, alpha0(solver), velocity, dt(solver), description"""
    if name == "kelvin":
        # BASELINE.json configs[2]: polyDualMesh-like cells (14 faces, 24 points), sphere r = 0.15
        # at (0.5, 0.75, 0.5) in solid-body rotation about the z axis; size = cells per axis of each of the two lattices
        m = meshmod.kelvin_mesh_fast(size)
        ctl = dict(CONTROLS)

        def alpha0(s):
            C_, V = s.field(capi.F_C), s.field(capi.F_V)
            return np.clip(0.5 - (np.linalg.norm(C_ - np.array([0.5, 0.75, 0.5]), axis=1) - 0.15) / np.cbrt(V), 0.0, 1.0)

        def dt_of(s, U):
            return 0.25 * float(np.cbrt(s.field(capi.F_V).min())) / float(np.abs(U).max())

        desc = ("polyhedral mesh: %d truncated-octahedron (polyDualMesh-like, 14 faces / 24 points) cells, sphere r=0.15 in solid-body "
                "rotation (BASELINE.json configs[2])" % m.n_cells)
        return m, ctl, alpha0, fields.rotation_velocity, dt_of, desc
    if name == "dambreak":
        # BASELINE.json configs[3] substitute (advection only): the damBreakWithObstacle box
        # (geometryAndMeshDimensions:1-3), water box of setVofFieldDict:33-39, the case's alpha controls, inletOutlet top
        m = meshmod.hex_block(size)
        for p_ in m.patches:
            if p_.name == "top":
                p_.alpha_bc, p_.alpha_value = capi.BC_INLET_OUTLET, 0.0
        ctl = dict(DAMBREAK_CONTROLS)

        def alpha0(s):
            C_ = meshmod.cell_centres_hex(m)
            return ((C_[:, 0] < 0.6) & (C_[:, 1] < 0.1875) & (C_[:, 2] < 0.75)).astype(np.float64)

        def dt_of(s, U):
            return 0.2 / size

        desc = ("damBreakWithObstacle substitute, advection only: %d^3 hex box, water box (0.6 x 0.1875 x 0.75), clip true / snapTol 1e-12 / "
                "mixedCellTol 1e-10, inletOutlet top patch, prescribed 3-D deformation velocity (BASELINE.json configs[3] scaled)" % size)
        return m, ctl, alpha0, fields.leveque_velocity, dt_of, desc
    raise ValueError(name)


def run_workload(args):
    t_setup = time.perf_counter()
    m, ctl, alpha0, velocity, dt_of, desc = make_workload(args.workload, args.n)
    s = SolveVofEqu(m, ctl)
    lib, h = s.lib, s._h
    a0 = alpha0(s)
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    U, phi = velocity(C_), fields.face_flux(Cf, Sf, velocity)
    Ub = velocity(Cf[m.n_internal_faces:]) if args.workload == "kelvin" else np.zeros((s.nBF, 3))
    del C_, Cf, Sf
    dt = dt_of(s, U)
    s.setAlpha(a0)
    s.setPhi(phi)
    s.setU(U, Ub)
    s.synchronize()
    setup_s = time.perf_counter() - t_setup
    # let the interface develop (a sharp box has no mixed cells at t = 0), then keep that field as the start of every leg
    for _ in range(args.develop_steps):
        s.step(dt)
    a1 = s.alpha()
    clocks = ClockSampler()
    clocks.start()
    s.setAlpha(a1)
    s.setOption("overlap", 0)   # the roofline kernel is timed alone
    for _ in range(args.warmup):
        s.reconstruct()
        s.advect(dt)
    s.synchronize()
    d0, dn0 = s.info(capi.I_DENSE_KERNEL_MS), s.info(capi.I_DENSE_KERNEL_LAUNCHES)
    for _ in range(args.steps):
        s.reconstruct()
        s.advect(dt)
    s.synchronize()
    dense_ms = (s.info(capi.I_DENSE_KERNEL_MS) - d0) / max(1.0, s.info(capi.I_DENSE_KERNEL_LAUNCHES) - dn0)
    s.setOption("overlap", 1 if args.overlap < 0 else args.overlap)
    sched_sel = settle_schedule(s, a1, dt, args)
    s.setAlpha(a1)
    for _ in range(2 + args.warmup):
        s.step(dt)
    s.synchronize()
    l0 = s.info(capi.I_GPU_LAUNCHES)
    lib.svof_mark(h, 0)
    for _ in range(args.steps):
        s.step(dt)
    lib.svof_mark(h, 1)
    ms = C.c_double()
    lib.svof_elapsed_ms(h, 0, 1, C.byref(ms))
    s.synchronize()
    launches = int(s.info(capi.I_GPU_LAUNCHES) - l0)
    n_mixed, n_near, err = int(s.info(capi.I_N_MIXED)), int(s.info(capi.I_N_NEAR)), int(s.info(capi.I_ERROR_FLAGS))
    value = m.n_cells * args.steps / (ms.value * 1e-3)
    # end to end through svof_step_host with pinned host buffers
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    phi_h, U_h, Ub_h = capi.pinned_array(lib, (s.nF,)), capi.pinned_array(lib, (s.nC, 3)), capi.pinned_array(lib, (max(s.nBF, 1), 3))
    a_out, ap_out = capi.pinned_array(lib, (s.nC,)), capi.pinned_array(lib, (s.nF,))
    phi_h[:] = phi
    U_h[:] = U
    Ub_h[:s.nBF] = Ub
    s.setAlpha(a1)
    for _ in range(1 + args.warmup):
        s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
    e2e_s = 0.0
    for k in range(e2e_steps):       # phi and U change everywhere between calls; only the calls are timed
        fk = 1.0 - 1e-3 * (k + 1)
        np.multiply(phi, fk, out=phi_h)
        np.multiply(U, fk, out=U_h)
        t0 = time.perf_counter()
        s.step_host(dt, phi_h, U_h, Ub_h, a_out, ap_out)
        e2e_s += time.perf_counter() - t0
    clk = clocks.stop()
    h2d, d2h = int(s.info(capi.I_H2D_BYTES)), int(s.info(capi.I_D2H_BYTES))
    peak, peak_src = measured_peak()
    B = b_alg(m)
    cpu, parity = None, []
    if not args.no_cpu:
        cs = max(1, args.cpu_steps)
        rate, el, ref = cpu_oracle_rate(m, a1, U, phi, dt, cs, keep=True, controls=ctl, Ub=Ub)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "same mesh and fields, %d steps (%.1f s) of the single-threaded CPU restatement of the reference algorithm" % (cs, el)}
        s.setPhi(phi)
        s.setU(U, Ub)
        parity.append(parity_block(s, m, a1, dt, ref, args.workload))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms.value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": desc, "cells": m.n_cells, "faces": m.n_faces, "dt": dt, "controls": ctl, "mixed_cells": n_mixed,
                   "near_cells": n_near, "develop_steps": args.develop_steps, "error_flags": err, "setup_s": setup_s,
                   "l2": "inputs larger than L2 (%.2f GB per step)" % (B / 1e9),
                   "schedule": "svof_step_device: one CUDA-graph launch per step (%d kernels inside)" % (launches // max(1, args.steps)),
                   "schedule_selected": sched_sel},
        "clocks": clk,
        "e2e": {"value": m.n_cells * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": B / (dense_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": B / (dense_ms * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "k_dense_update", "kernel_ms": dense_ms,
                     "algorithmic_bytes_per_launch": B, "peak_source": peak_src,
                     "step_frac": (B / (ms.value / args.steps * 1e-3) / 1e9) / peak},
        "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line))


def _ref_worker(args_tuple):
    """One z-slab sub-domain, single-threaded: the reference's own solveVofEqu class (oracle/_ref/libref_solver.so) when it was
    built, and the oracle port on the same slab and steps for comparison.  -> cells, seconds (reference class or None), seconds (port)"""
    n, lo, hi, dt, steps, warm, want_class = args_tuple
    m, a0 = build_case(n, lo=lo, hi=hi, cut_as_wall=True)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build as oracle_build
    lib = capi.load(oracle_build.build_oracle())
    so = SolveVofEqu(m, CONTROLS, lib=lib)
    U, phi = velocity_fields(so, dt, dt)
    so.setAlpha(a0)
    so.setPhi(phi)
    so.setU(U)
    t_class = None
    if want_class:
        from refsolver import RefSolver
        if RefSolver.lib() is not None:
            Cf = so.field(capi.F_CF)
            ref = RefSolver(m, so._params)
            ref.setState(a0, phi, U, fields.leveque_velocity(Cf[m.n_internal_faces:]) * fields.u_factor(dt, dt, PERIOD))
            for _ in range(warm):
                ref.reconstruct()
                ref.advect(dt)
            t0 = time.perf_counter()
            for _ in range(steps):
                ref.reconstruct()
                ref.advect(dt)
            t_class = time.perf_counter() - t0
            del ref
    for _ in range(warm):
        so.reconstruct()
        so.advect(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        so.reconstruct()
        so.advect(dt)
    return m.n_cells, t_class, time.perf_counter() - t0


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: its solveVofEqu class -- every file of
    src/SimPLIC outside sampling/, compiled unmodified against an OpenFOAM stand-in into oracle/_ref/libref_solver.so
    (OpenFOAM v2312 itself cannot be built here: no wmake/MPI) -- run the way the reference runs in parallel:
    P sub-domains (z-slabs of the same mesh), one single-threaded process each, no halo exchange (cut faces are treated as
    walls, which only removes communication cost from the CPU side).  The oracle port runs the same slabs and steps right
    after it; its figure is reported beside the class's (cpu_baseline.port_value).  Without the prebuilt library the port is
    the arm (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    n = args.n
    strong = args.gpus > 1 and args.scaling == "strong" and not args.ref_n
    if strong:
        n = args.strong_n      # the N-GPU arm runs ONE strong_n^3 problem (BASELINE.json configs[4]): the CPU arm samples that mesh
    P = max(1, min(os.cpu_count() or 1, args.ref_procs or (os.cpu_count() or 1), n))
    # the whole mesh, one z-slab per process (a smaller sample mesh only on request)
    ns = args.ref_n or n
    dt = 0.2 / ns
    bounds = [(ns * p) // P for p in range(P + 1)]
    steps = max(1, args.steps)
    want_class = args.ref_kind != "port"
    if strong:
        # a bounded, unbiased sample of the 512^3 mesh: P thin z-slabs of ~1 M cells spread evenly over the height (the share of
        # interface cells is that of the whole problem; the whole mesh would need ~300 GB with the reference's class)
        thick = max(1, min(ns // P, int(np.ceil(1.05e6 / (ns * ns)))))
        slabs = [(bounds[p], bounds[p] + thick) for p in range(P)]
    else:
        slabs = [(bounds[p], bounds[p + 1]) for p in range(P)]
    jobs = [(ns, (0, 0, lo), (ns, ns, hi), dt, steps, max(1, min(args.warmup, 1)), want_class) for lo, hi in slabs]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(P) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    cells = sum(r[0] for r in res)
    t_port = max(r[2] for r in res)
    have_class = all(r[1] is not None for r in res)
    tmax = max(r[1] for r in res) if have_class else t_port
    value = cells * steps / tmax
    if have_class:
        cpu = {"value": value, "unit": UNIT, "cores": P, "kind": "reference",
               "sample": "%d^3 LeVeque mesh: %d z-slab sub-domains, one single-threaded process each running the "
                         "reference's own solveVofEqu class (src/SimPLIC compiled unmodified against an OpenFOAM stand-in: "
                         "oracle/_ref/libref_solver.so; OpenFOAM v2312/MPI not installable here), %d steps, slowest rank %.1f s"
                         % (ns, P, steps, tmax),
               "port_value": cells * steps / t_port,
               "port_note": "the CPU restatement (oracle/) on the same slabs and steps, slowest rank %.1f s" % t_port}
    else:
        cpu = {"value": value, "unit": UNIT, "cores": P, "kind": "port",
               "sample": "%d^3 LeVeque mesh split into %d z-slab sub-domains, one single-threaded process each "
                         "(CPU restatement of the reference algorithm; oracle/_ref/libref_solver.so not built), "
                         "%d steps, slowest rank %.1f s" % (ns, P, steps, tmax)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tmax / steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "LeVeque 3-D deformation, sphere r=0.15, %d^3 hex blockMesh (BASELINE.json configs[%d])" % (n, 4 if strong else 1),
                   "sample_mesh": ("%d z-slabs of %d planes spread evenly over the %d^3 mesh (%.1f M of %.1f M cells)"
                                   % (P, slabs[0][1] - slabs[0][0], ns, cells / 1e6, ns ** 3 / 1e6)) if strong else "%d^3 in %d z-slabs" % (ns, P),
                   "dt": dt, "controls": CONTROLS, "wall_s": wall},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=int(os.environ.get("SVOF_BENCH_N", "256")),
                    help="cells per axis (per GPU block when --gpus > 1)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-multi", action="store_true", help="N > 1: also time svof_step_host per rank (host buffers in/out)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="leveque", choices=["leveque", "kelvin", "dambreak"],
                    help="leveque: BASELINE.json configs[1] (the headline); kelvin: configs[2] (--size 95 = 1.7 M polyhedral cells); "
                         "dambreak: configs[3] substitute (--size 368 = 50 M cells)")
    ap.add_argument("--develop-steps", type=int, default=40, help="kelvin/dambreak: untimed steps before the timed window")
    ap.add_argument("--late-steps", type=int, default=200, help="timed steps of the second window (0: skip it)")
    ap.add_argument("--late-time", type=float, default=1.5, help="flow time the interface is advanced to before the second window")
    ap.add_argument("--late-cpu-steps", type=int, default=2, help="oracle steps for the parity block of the second window")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-size", dest="ref_n", type=int, default=0)
    ap.add_argument("--ref-kind", default="reference", choices=["reference", "port"],
                    help="--impl reference: the reference's own class from oracle/_ref/libref_solver.so (default; falls back to the "
                         "port when the library is not there) or only the oracle port")
    # N > 1: strong scaling of ONE problem (BASELINE.json configs[4]) unless --scaling weak (one 256^3 unit cube per GPU)
    ap.add_argument("--scaling", default=os.environ.get("SVOF_BENCH_SCALING", "strong"), choices=["strong", "weak"])
    ap.add_argument("--strong-size", dest="strong_n", type=int, default=int(os.environ.get("SVOF_BENCH_STRONG_N", "512")))
    ap.add_argument("--layers", type=int, default=0, help="ghost layers (0: nAlphaBounds + 2)")
    ap.add_argument("--partition", default=os.environ.get("SVOF_BENCH_PARTITION", "rcb"), choices=["interface", "rcb"],
                    help="strong scaling on 4 / 8 ranks: 'interface' = the sphere on half of the ranks with few bulk cells, the other half "
                         "streaming only (multigpu.interface_boxes); 'rcb' = weighted recursive bisection (--mixed-weight)")
    ap.add_argument("--mixed-weight", type=float, default=float(os.environ.get("SVOF_BENCH_MIXED_WEIGHT", "1000")),
                    help="partition weight of an interface cell relative to a bulk cell (strong scaling)")
    ap.add_argument("--overlap", type=int, default=-1, help="two-stream schedule (library option 'overlap'); -1: library default")
    ap.add_argument("--no-sched-auto", action="store_true", help="keep the library's default schedule instead of its run-time selection")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "leveque" and args.gpus <= 1 and int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        run_workload(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
