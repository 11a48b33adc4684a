#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native tensor-train hot path.

Metric (BASELINE.json): MPO x MPS zip-up + truncate sweeps/sec at chi=512 f64.
One step = one `contract(state, mpo, center=0, Zipup{max_bond_dim=chi, SvdTruncationPolicy(0.0)})`
on the synthetic C3 configuration (L=64, d=4, chi=512, MPO bond w=8), i.e. the reference call
stack of SURVEY.md section 3.1: 2 operand canonicalisations + (L-2) zip steps + final block +
final truncation sweep.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A single sweep is sequential over sites (replicas only): with N > 1 every rank runs an
independent TT of the same shape (the batch-of-independent-TTs sharding of the north star, weak
scaling), and NCCL is used only to take the max time and gather the result norms.

`value` is measured with the operands already resident in HBM; `e2e` is the same sweep through
the C ABI from pinned HOST buffers with the H2D upload of both operands and the D2H download of
the result inside the timed region.  `--impl reference` times the CPU restatement of the
reference algorithm (oracle/, NumPy + LAPACK/BLAS on the host cores) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
sys.path.insert(0, ROOT)

METRIC = "MPOxMPS zip-up+truncate sweeps/sec at chi=512 f64"
UNIT = "sweeps/s"


def bond_dims(L, d, chi):
    return [min(d ** (i + 1), d ** (L - 1 - i), chi) for i in range(L - 1)]


def make_c3(seed, L, d, chi, w):
    """Synthetic C3 operands: N(0,1) MPS scaled to O(1) norm, MPO N(0,1)/sqrt(w d)."""
    rng = np.random.default_rng(seed)
    bd = bond_dims(L, d, chi)
    mps, mps_ids, mpo, mpo_ids = [], [], [], []
    for i in range(L):
        shape, ids = [], []
        if i > 0:
            shape.append(bd[i - 1]); ids.append(1000 + i - 1)
        shape.append(d); ids.append(100 + i)
        if i < L - 1:
            shape.append(bd[i]); ids.append(1000 + i)
        a = rng.standard_normal(shape) / np.sqrt(d * (bd[i] if i < L - 1 else 1))   # E||TT||^2 = 1
        mps.append(np.asfortranarray(a)); mps_ids.append(ids)
        shape, ids = [], []
        if i > 0:
            shape.append(w); ids.append(2000 + i - 1)
        shape += [d, d]; ids += [200 + i, 100 + i]
        if i < L - 1:
            shape.append(w); ids.append(2000 + i)
        mpo.append(np.asfortranarray(rng.standard_normal(shape) / np.sqrt(w * d))); mpo_ids.append(ids)
    return mps, mps_ids, mpo, mpo_ids


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- reference arm / cpu baseline (oracle on host cores) -----------------------------------------
def cpu_sample(L, d, chi, w, seed=0x5EED0003):
    """Bounded sample of the CPU restatement: one bulk zip-up step, one bulk two-site truncation
    step and one bulk canonicalisation step at full chi, scaled by the step counts of one sweep."""
    import scipy.linalg as sla
    rng = np.random.default_rng(seed)
    n = chi
    R = rng.standard_normal((n, chi, w))
    A = rng.standard_normal((chi, d, chi))
    B = rng.standard_normal((w, d, d, w))
    t0 = time.perf_counter()
    RA = np.tensordot(R, A, axes=([1], [0]))                      # [n, w, d, chi']
    M = np.tensordot(RA, B, axes=([1, 2], [0, 2]))                # [n, chi', d_out, w']
    M = np.transpose(M, (0, 2, 1, 3)).reshape(n * d, chi * w, order="F")
    u, s, vh = sla.svd(M, full_matrices=False, lapack_driver="gesdd")
    left, right = u[:, :chi], s[:chi, None] * vh[:chi]
    t_zip = time.perf_counter() - t0
    A2 = rng.standard_normal((chi, d, chi))
    t0 = time.perf_counter()
    AB = np.tensordot(A, A2, axes=([2], [0])).reshape(chi * d, d * chi, order="F")
    u, s, vh = sla.svd(AB, full_matrices=False, lapack_driver="gesdd")
    left, right = u[:, :chi], s[:chi, None] * vh[:chi]
    t_two = time.perf_counter() - t0
    t0 = time.perf_counter()
    q, r = sla.qr(A.reshape(chi * d, chi, order="F"), mode="economic")
    nxt = np.tensordot(r, A2, axes=([1], [0]))
    t_qr = time.perf_counter() - t0
    # one sweep: (L-2) zip steps + final block (~ one zip step), 2(L-1) two-site steps,
    # (L-1) QR+absorb for the MPS operand, (L-1) for the result's centre move (MPO QRs are tiny)
    sweep = (L - 1) * t_zip + 2 * (L - 1) * t_two + 2 * (L - 1) * t_qr
    return {"sweep_s": sweep, "t_zip": t_zip, "t_two_site": t_two, "t_qr": t_qr}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        threadpool_limits(limits=n)
    except Exception:
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    if rank != 0:
        return
    L, d, chi, w = args.L, args.d, args.chi, args.w
    use_all_host_threads()
    for _ in range(min(args.warmup, 1)):
        cpu_sample(L, d, min(chi, 64), w)
    samples = [cpu_sample(L, d, chi, w, seed=0x5EED0003 + i) for i in range(max(1, min(args.steps, 2)))]
    sweep_s = statistics.median([s["sweep_s"] for s in samples])
    val = 1.0 / sweep_s
    sample = (f"1 bulk zip step (contract + gesdd {chi*d}x{chi*w}), 1 bulk two-site step (gesdd {chi*d}x{chi*d}), "
              f"1 QR+absorb step ({chi*d}x{chi}) at full chi, scaled by the per-sweep step counts "
              f"({args.L-1} / {2*(args.L-1)} / {2*(args.L-1)})")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sweep_s, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": host_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"C3: MPOxMPS zip-up+truncate, L={args.L} d={args.d} chi={args.chi} MPO bond w={args.w}, "
                        f"f64, center=0, max_bond_dim={args.chi}, SvdTruncationPolicy(0.0) (cap-only)",
            "L": args.L, "d": args.d, "chi": args.chi, "w": args.w,
            "l2_policy": "operands (MPS 0.5 GB at chi=512) and SVD work buffers exceed the 126 MB L2",
            "parallelism": "independent TT replica per GPU (a single sweep is sequential)"}


# ---- our arm ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--d", type=int, default=4)
    ap.add_argument("--chi", type=int, default=512)
    ap.add_argument("--w", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import t4b
    from t4b import tt as t4tt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback "
                         "(use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = t4b.Context(local_rank, stream.cuda_stream)

    L, d, chi, w = args.L, args.d, args.chi, args.w
    mps, mps_ids, mpo, mpo_ids = make_c3(0x5EED0003 + rank, L, d, chi, w)
    policy = t4tt.SvdPolicy(0.0)

    # pinned host staging for the e2e path
    def pinned_like(arrs):
        out = []
        for a in arrs:
            t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
            v = t.numpy().reshape(a.shape, order="F")
            v[...] = a
            out.append((t, v))
        return out
    mps_pin, mpo_pin = pinned_like(mps), pinned_like(mpo)
    h2d_bytes = sum(a.nbytes for a in mps) + sum(a.nbytes for a in mpo)

    a_dev = t4tt.chain_from_arrays(ctx, mps, mps_ids)
    b_dev = t4tt.chain_from_arrays(ctx, mpo, mpo_ids)

    def sweep_resident():
        out = a_dev.contract(b_dev, 0, 0, policy, chi)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sweep_resident().release()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    last = None
    for _ in range(args.steps):
        if last is not None:
            last.release()
        last = sweep_resident()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    norm2 = last.norm_sqr()
    bonds = last.bond_dims()

    # ---- e2e: host buffers in, host buffers out ----------------------------------------------------
    # pinned result buffers (the shapes of the result are fixed by the cap-only truncation)
    out_pin = []
    for i in range(L):
        shp, _ = last.site_shape(i)
        t = torch.empty(int(np.prod(shp)), dtype=torch.float64, pin_memory=True)
        out_pin.append(t.numpy().reshape(shp, order="F"))
    d2h_bytes = 0

    def e2e_step():
        ta = t4tt.chain_from_arrays(ctx, [v for _, v in mps_pin], mps_ids)
        tb = t4tt.chain_from_arrays(ctx, [v for _, v in mpo_pin], mpo_ids)
        out = ta.contract(tb, 0, 0, policy, chi)
        for i in range(L):
            shp, _ = out.site_shape(i)
            if tuple(shp) != out_pin[i].shape:      # never the case for C3; keep the path general
                out_pin[i] = np.empty(shp, dtype=np.float64, order="F")
            out.site_into(i, out_pin[i])
        ta.release(); tb.release(); out.release()
        return sum(a.nbytes for a in out_pin)

    e2e_step()          # untimed warm-up of the end-to-end path (allocator, pinned pages)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        d2h_bytes = e2e_step()
    e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # ---- per-kernel profile of one extra sweep (roofline) ------------------------------------------
    t4b._check(t4b.lib().t4b_ctx_profile_begin(ctx.h))
    sweep_resident().release()
    import ctypes as C
    need = C.c_size_t()
    t4b._check(t4b.lib().t4b_ctx_profile_end(ctx.h, None, C.c_size_t(0), C.byref(need)))
    buf = C.create_string_buffer(need.value)
    t4b._check(t4b.lib().t4b_ctx_profile_end(ctx.h, buf, need, None))
    prof = {}
    for ln in buf.value.decode().splitlines():
        nm, n, tms, work = ln.split()
        prof[nm] = {"launches": int(n), "ms": float(tms), "work": float(work)}
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        norms = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(norms, torch.tensor([norm2], device="cuda", dtype=torch.float64))
        norm_list = [float(x) for x in norms]
        lt = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])
    else:
        norm_list = [norm2]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        f64_peak, f64_src = 37.04, "measured DMMA.8x8x4 issue peak on this pool (profiles/peak_f64_r01.json)"
        try:
            pk = json.load(open(os.path.join(ROOT, "profiles", "peak_f64_r01.json")))
            f64_peak = float(pk["dmma_tflops_w8"])
        except Exception:
            f64_src = "fallback 37.04 TFLOP/s (DMMA issue peak measured round 1)"
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0] if prof else None
        roof = None
        if dom:
            p = prof[dom]
            if dom == "gemm":
                ach = p["work"] / (p["ms"] * 1e-3) / 1e12
                roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": f64_peak, "unit": "TFLOP/s",
                        "frac": ach / f64_peak, "traffic": None, "peak_source": f64_src}
            else:
                ach = p["work"] / (p["ms"] * 1e-3) / 1e9
                roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src}
            # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch)
            try:
                cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("ncu_summary_"))
                ncu = json.load(open(os.path.join(ROOT, "profiles", cands[-1])))
                if dom == "jacobi":
                    cap = ncu["jacobi_2048"][0]
                    roof["traffic"] = cap["dram_bytes"]
                    roof["traffic_note"] = ("dram__bytes_read+write of one jacobi_persistent_kernel launch on the 2048x2048 "
                                            "zip-up factor (profiles/ncu_jacobi_2048_*.csv): the panel is L2-resident, "
                                            "DRAM sees the matrix once while the algorithmic (L2) traffic of that launch is "
                                            "~100 GB; DMMA pipe %.1f%% busy" % cap["fp64_tensor_pct"])
            except Exception:
                pass
            roof["share_of_step"] = p["ms"] / tot_ms
            roof["launches_per_step"] = p["launches"]
            roof["avg_launch_us"] = 1e3 * p["ms"] / max(p["launches"], 1)
        g = prof.get("gemm")
        roof_gemm = None
        if g and g["ms"] > 0:
            ach = g["work"] / (g["ms"] * 1e-3) / 1e12
            roof_gemm = {"kernel": "gemm (DMMA tensordot, all contractions of the sweep)", "bound": "tensor",
                         "achieved": ach, "peak": f64_peak, "unit": "TFLOP/s", "frac": ach / f64_peak,
                         "share_of_step": g["ms"] / tot_ms, "launches_per_step": g["launches"],
                         "peak_source": f64_src}
        value = world * args.steps / (ms * 1e-3)
        e2e_val = world * args.steps / (ms_e2e * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args), "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "roofline": roof, "roofline_contraction": roof_gemm,
                "kernel_profile_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "kernel_profile_note": "per-kernel CUDA-event times of one extra sweep; the QR look-ahead (factor of panel "
                                       "p+1 on a side stream) is serialised while profiling, so qr_* sum to more than they "
                                       "cost in the timed region",
                "result": {"norm_sqr": norm_list, "max_bond": max(bonds) if bonds else 1}}
        if not args.no_cpu_baseline:
            use_all_host_threads()
            cs = cpu_sample(L, d, chi, w)
            line["cpu_baseline"] = {"value": 1.0 / cs["sweep_s"], "unit": UNIT, "cores": host_threads(),
                                    "kind": "port",
                                    "sample": f"1 bulk zip step ({cs['t_zip']:.2f} s), 1 two-site step ({cs['t_two_site']:.2f} s), "
                                              f"1 QR+absorb step ({cs['t_qr']:.2f} s) at full chi, scaled to one sweep"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
