#!/usr/bin/env python3
"""bench.py — the headline metric of BASELINE.json: credential Ed25519 verifies/s (512 B payload) at N B200s.

    python bench.py --gpus N --steps K --warmup W                (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path: oracle C port, all host threads)

A "step" is one pass of the hot path over one batch: per GPU, configs[1] of BASELINE.json — batched Ed25519 verify of
1 M x 512 B credentials (K = 1024 key pairs, 1 % corrupted so the kernel cannot short-circuit).  Weak
scaling: every rank verifies its own 1 M batch, no collective on the data path (§8e).

  value      verifies/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks).  Every step starts from
             an EMPTY issuer-key cache: de-duplication, the 1024 per-key tables and all 10^6 verifications happen inside the step
  e2e        the same work through the C-ABI host call (afc_ed25519_verify_batch): pinned host buffers, H2D of all inputs and
             D2H of the result bitmap inside the timed region (cache emptied before every step as well)
  roofline   the dominant kernel against the measured HBM peak, algorithmic bytes = 609 B / credential
  cpu_baseline  the oracle's C port of Go's algorithm on this box's host cores, bounded sample (rank 0, N = 1 only)
  warm_keycache / no_keycache / keyed   secondary figures: tables kept between steps, cache disabled (generic kernel), explicit key set

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ITEMS = 1_000_000
MSG_LEN = 512
N_KEYS = 1024
ALGO_BYTES = 609            # 512 msg + 32 pk + 64 sig read, 1 result byte written
METRIC = "credential Ed25519 verifies/sec (512B payload)"
UNIT = "verifies/s"


def host_threads():
    """CPU threads this process may actually use: affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for qf, pf in (("/sys/fs/cgroup/cpu.max", None), ("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us")):
        try:
            if pf is None:
                q, per = open(qf).read().split()
            else:
                q, per = open(qf).read().strip(), open(pf).read().strip()
            if q not in ("max", "-1"):
                n = max(1, min(n, int(float(q) / float(per) + 0.999)))
            break
        except Exception:
            continue
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        # NVML in a thread (no process start-up latency, so short timed regions still get samples); nvidia-smi -lms as fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv, self.stop_flag = pynvml, False
            self.th = threading.Thread(target=self._nvml_loop, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _nvml_loop(self):
        nv = self.nv
        bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append("%d, %d, 0, %s" % (sm, mx, ", ".join("Active" if (r & b) else "Not Active" for _, b in bits)))
            except Exception:
                pass
            time.sleep(0.02)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if getattr(self, "nv", None):
            self.stop_flag = True
            self.th.join(timeout=2)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # median of the upper half = clocks under load (idle samples at the edges are dropped)
        s = sorted(sm)
        load = s[len(s) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(ctx, dev, rank):
    """cfg2 inputs, generated on the device (synthetic): returns device tensors + expected bitmap."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(0xAF02 + rank)
    n = N_ITEMS
    rng = np.random.default_rng(0xAF02)
    kseeds = torch.from_numpy(rng.integers(0, 256, (N_KEYS, 32), dtype=np.uint8)).to(dev)
    d_exp = torch.empty((N_KEYS, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(kseeds, N_KEYS, d_exp)
    d_msgs = torch.randint(0, 256, (n, MSG_LEN), dtype=torch.uint8, device=dev, generator=g)
    idx = torch.arange(n, device=dev)
    d_ki = (idx % N_KEYS).to(torch.int32)
    d_off = torch.arange(n + 1, device=dev, dtype=torch.int64) * MSG_LEN
    d_sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ctx.sign_expanded_dev(d_exp, d_ki, d_msgs.view(-1), d_off, n, d_sigs)
    d_pks = d_exp[:, 64:][d_ki.long()].contiguous()
    flip_msg, flip_s = idx % 100 == 0, idx % 100 == 50
    bitpos = (idx // 100) % (MSG_LEN * 8)
    rows, cols = idx[flip_msg], bitpos[flip_msg] // 8
    d_msgs[rows, cols] = d_msgs[rows, cols] ^ (1 << (bitpos[flip_msg] % 8)).to(torch.uint8)
    srows = idx[flip_s]
    d_sigs[srows, 33] = d_sigs[srows, 33] ^ 0x08
    expect = (~(flip_msg | flip_s)).to(torch.uint8)
    torch.cuda.synchronize()
    return d_pks, d_sigs, d_msgs.view(-1), d_off, expect


def cpu_reference_rate(threads, budget_s=12.0):
    """Oracle C port (Go's algorithm: 51-bit limbs, NAF-5/NAF-8 vartime double-scalar mult) on `threads` host threads over a
    bounded sample of the same workload.  Returns (verifies/s, sample size)."""
    from oracle import c_oracle as CO
    rng = np.random.default_rng(0xAF02)
    kseeds = rng.integers(0, 256, (N_KEYS, 32), dtype=np.uint8)
    kpks = CO.ed25519_pubkey_batch(kseeds, threads)

    def sample(m):
        ki = np.arange(m) % N_KEYS
        msgs = rng.integers(0, 256, (m, MSG_LEN), dtype=np.uint8)
        off = np.arange(m + 1, dtype=np.uint64) * MSG_LEN
        sigs = CO.ed25519_sign_batch(kseeds[ki].copy(), msgs.reshape(-1), off, threads)
        msgs[::100, 7] ^= 1
        return kpks[ki].copy(), sigs, msgs.reshape(-1), off

    pk, sg, ms, off = sample(4096 * max(1, threads // 8))
    t0 = time.perf_counter()
    CO.ed25519_verify_batch(pk[:2048], sg[:2048], ms, off[:2049], 1)
    cpu_reference_rate.single_thread = 2048 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    CO.ed25519_verify_batch(pk, sg, ms, off, threads)
    rate0 = (len(off) - 1) / (time.perf_counter() - t0)
    m = int(min(N_ITEMS, max(8192, rate0 * budget_s)))
    pk, sg, ms, off = sample(m)
    t0 = time.perf_counter()
    ok = CO.ed25519_verify_batch(pk, sg, ms, off, threads)
    dt = time.perf_counter() - t0
    assert int(ok.sum()) == m - len(range(0, m, 100))
    return m / dt, m


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    # each step = one bounded sample; scale the per-step budget so warmup+steps end within a few minutes
    budget = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    rates, sample = [], 0
    for i in range(args.warmup + args.steps):
        r, sample = cpu_reference_rate(threads, budget)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": "batched Ed25519 verify, 1 M x 512 B credentials (BASELINE.json configs[1]); CPU sample per step",
                                        "items_per_step": sample, "msg_len": MSG_LEN, "keys": N_KEYS, "corrupted": "1%"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d credentials per step; oracle/afc_oracle.c (C restatement of Go crypto/ed25519: 51-bit limbs, NAF vartime "
                                   "double-scalar mult), %d pthreads (one thread alone: %.0f/s); Go toolchain absent so the reference itself "
                                   "cannot run" % (sample, threads, cpu_reference_rate.single_thread)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time HMAC / sign / Merkle (secondary configs) and microbenchmarks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import agentfield_b200 as afb

    os.environ["NCCL_DEBUG"] = os.environ.get("AFC_NCCL_DEBUG", "WARN")     # keep NCCL banners off stdout: one JSON line only
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = afb.Context(local_rank)
    info = ctx.device_info()
    bad = ctx.selftest(200)
    if bad:
        raise SystemExit("PTX field self-test failed on %d threads" % bad)

    d_pks, d_sigs, d_msgs, d_off, expect = make_workload(ctx, dev, rank)
    n = N_ITEMS
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)          # cold: the issuer-key cache is empty, tables get built in this call
    c1.record()
    torch.cuda.synchronize()
    cold_ms = c0.elapsed_time(c1)
    def step():
        # Nothing is carried from one step to the next: the issuer-key cache is emptied first, so de-duplication, the 1024
        # per-key tables and all 10^6 verifications are redone inside every timed step.
        ctx.keycache_clear()
        ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)

    for _ in range(args.warmup):
        step()
    barrier()
    assert torch.equal(d_ok, expect), "verify bitmap differs from the corruption pattern"
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ctx.profile_begin(64 + 32 * args.steps)      # per-kernel CUDA events on the launching stream (one pair per launch, claimed lazily)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = ctx.profile_end()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    assert torch.equal(d_ok, expect)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---------------- end to end through the C-ABI host call (pinned host buffers, H2D + D2H inside)
    h_pks = d_pks.cpu().pin_memory(); h_sigs = d_sigs.cpu().pin_memory(); h_msgs = d_msgs.cpu().pin_memory()
    h_off = d_off.cpu().pin_memory(); h_ok = torch.empty(n, dtype=torch.uint8).pin_memory()
    lib, H = afb._abi.load(), ctx.handle

    def e2e_step():
        ctx.keycache_clear()
        rc = lib.afc_ed25519_verify_batch(H, h_pks.data_ptr(), h_sigs.data_ptr(), h_msgs.data_ptr(), h_off.data_ptr(), n, h_ok.data_ptr())
        if rc != 0:
            raise afb.AfcError(rc, lib.afc_last_cuda_error(H).decode())

    e2e_step()
    assert torch.equal(h_ok, expect.cpu())
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t_e.item())
    h2d = n * (32 + 64 + MSG_LEN) + (n + 1) * 8
    d2h = n
    # the transfer floor of that call: the same pinned buffers copied to the device with nothing else going on
    scratch = [torch.empty_like(t, device=dev) for t in (h_pks, h_sigs, h_msgs, h_off)]
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        if rep == 1:
            c0.record()
        for dst, src in zip(scratch, (h_pks, h_sigs, h_msgs, h_off)):
            dst.copy_(src, non_blocking=True)
    c1.record(); torch.cuda.synchronize()
    h2d_only_ms = c0.elapsed_time(c1) / 2
    del scratch
    e2e_extra = {"ms_per_step": 1e3 * float(t_e.item()) / args.steps, "h2d_only_ms": h2d_only_ms, "pcie_h2d_GBps": h2d / h2d_only_ms / 1e6,
                 "note": "h2d_only_ms = the step's inputs copied from the same pinned buffers with no compute: the transfer floor of the host call"}

    # ---------------- roofline of the dominant kernel
    hbm_peak, peak_src = peaks()
    zero = {"avg_ms": float("nan"), "count": 0, "total_ms": 0.0}
    dom = "k_ed_verify_cached" if prof.get("k_ed_verify_cached", zero)["total_ms"] > prof.get("k_ed_verify", zero)["total_ms"] else "k_ed_verify"
    kv = prof.get(dom, zero)
    kh = prof.get("k_ed_hram", zero)
    achieved = ALGO_BYTES * n / (kv["avg_ms"] * 1e-3) / 1e9 if kv["count"] else float("nan")
    traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum of k_ed_verify from the committed ncu --set full capture (same 1 M launch)
        tk = json.load(open(os.path.join(ROOT, "profiles", "r01_e_ncu_traffic.json")))["kernels"]
        traffic = tk.get(dom, {}).get("traffic_bytes")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES * n,
                "kernel_avg_ms": kv["avg_ms"], "kernel_share_of_step": kv["total_ms"] / max(ms_total, 1e-9),
                "other_kernels_ms": {k: v["avg_ms"] for k, v in prof.items() if k != dom},
                "note": "integer-multiplier (IMAD.WIDE) bound: see DESIGN.md section 4; HBM fraction reported because the metric asks for it"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "batched Ed25519 verify, 1 M x 512 B credentials per GPU (BASELINE.json configs[1])", "items_per_gpu": n,
                   "msg_len": MSG_LEN, "keys": N_KEYS, "corrupted": "1%", "l2": "inputs (609 MB per step) larger than L2 (126 MB)",
                   "parallelism": "independent shards, no collective" if world > 1 else "single GPU", "sm_count": info["sm_count"]},
        "clocks": clocks, "e2e": dict({"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}, **e2e_extra),
        "gpu_launches": int(launches), "roofline": roofline, "impl": "b200",
    }

    # ---------------- secondary: steady state of a long-running verifier (tables of known issuers stay cached between calls)
    try:
        for _ in range(max(5, args.warmup)):             # the host-call section before this one leaves the SMs mostly idle: let the clocks settle
            ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(args.steps):
            ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
        w1.record()
        barrier()
        wms = w0.elapsed_time(w1) / args.steps
        assert torch.equal(d_ok, expect)
        line["warm_keycache"] = {"value": world * n / (wms * 1e-3), "unit": UNIT, "ms_per_step": wms,
                                 "note": "same call without clearing the issuer-key cache between steps (tables of the 1024 issuers reused)"}
    except Exception as ex:
        line["warm_keycache"] = {"error": repr(ex)}

    # ---------------- secondary: the generic double-scalar kernel alone (issuer-key cache disabled: what a batch of all-distinct keys costs)
    kc_info = ctx.keycache_info()
    kc_info["cold_first_call_ms"] = cold_ms       # first call on an empty cache (1 M credentials, 1024 tables built inside it)
    line["keycache"] = kc_info
    try:
        ctx.keycache_configure(0)
        for _ in range(2):
            ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
        barrier()
        assert torch.equal(d_ok, expect)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1) / args.steps
        line["no_keycache"] = {"value": world * n / (gms * 1e-3), "unit": UNIT, "ms_per_step": gms,
                               "hbm_frac": ALGO_BYTES * n / (gms * 1e-3) / 1e9 / hbm_peak,
                               "note": "same call with afc_keycache_configure(ctx, 0): generic Straus kernel, no per-key tables"}
    except Exception as ex:
        line["no_keycache"] = {"error": repr(ex)}
    finally:
        ctx.keycache_configure(kc_info["max_keys"])
        kc_info["note"] = ("value and e2e empty the issuer-key cache before every step (tables rebuilt inside the timed region); "
                           "warm_keycache keeps them between steps; no_keycache disables the cache (generic kernel); cold_first_call_ms "
                           "additionally includes the one-time device allocations")

    # ---------------- secondary: the same batch verified against a registered key set (identity cache)
    try:
        t0 = time.perf_counter()
        kpks = d_pks[:N_KEYS].cpu().numpy()                      # key i occupies rows i, i + K, ... (key_i = i mod K)
        ks = afb.KeySet([bytes(p) for p in kpks], ctx)
        build_s = time.perf_counter() - t0
        d_ki = (torch.arange(n, device=dev) % N_KEYS).to(torch.int32)
        for _ in range(2):
            ks.verify_dev(d_ki, d_sigs, d_msgs, d_off, n, d_ok)
        barrier()
        assert torch.equal(d_ok, expect), "keyed verify bitmap differs"
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.steps):
            ks.verify_dev(d_ki, d_sigs, d_msgs, d_off, n, d_ok)
        k1.record()
        barrier()
        kms = k0.elapsed_time(k1) / args.steps
        line["keyed"] = {"value": world * n / (kms * 1e-3), "unit": UNIT, "ms_per_step": kms, "keyset_build_s": build_s,
                         "table_bytes": ks.info()["table_bytes"], "hbm_frac": ALGO_BYTES * n / (kms * 1e-3) / 1e9 / hbm_peak,
                         "note": "afc_ed25519_verify_keyed_batch_dev: issuer keys registered once (per-key radix-256 tables), same inputs and bitmap"}
        ks.close()
    except Exception as ex:     # secondary figure only
        line["keyed"] = {"error": repr(ex)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        rate, sample = cpu_reference_rate(threads, 12.0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d credentials of the same workload; oracle/afc_oracle.c (C restatement of Go crypto/ed25519), %d pthreads "
                                          "(one thread alone: %.0f/s)" % (sample, threads, cpu_reference_rate.single_thread)}
    if args.extras:
        line["extras"] = extras(ctx, dev, world, rank)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extras(ctx, dev, world, rank):
    """Secondary configs (not the headline): HMAC (cfg3 shape), sign + Merkle append (cfg4 shape), register-only microbenchmarks."""
    import torch
    import torch.distributed as dist
    import agentfield_b200 as afb
    out = {}
    g = torch.Generator(device=dev); g.manual_seed(0xAF03 + rank)

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    hbm_peak, _ = peaks()
    n = 4_000_000
    bodies = torch.randint(0, 256, (n, 256), dtype=torch.uint8, device=dev, generator=g)
    keys = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    off = torch.arange(n + 1, device=dev, dtype=torch.int64) * 256
    koff = (torch.arange(n + 1, device=dev, dtype=torch.int64) * 32).to(torch.int32)
    tags = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.hmac_sha256_dev(keys.view(-1), koff, bodies.view(-1), off, n, tags))
    out["hmac_sha256_256B"] = {"msgs_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 320 * n / (ms * 1e-3) / 1e9 / hbm_peak}
    dig = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx.sha256_dev(bodies.view(-1), off, n, dig))
    out["sha256_256B"] = {"msgs_per_s": n / (ms * 1e-3), "ms": ms, "hbm_frac": 288 * n / (ms * 1e-3) / 1e9 / hbm_peak}
    del bodies, keys, tags, dig
    # sign + Merkle append (cfg4 shape: N/8 per GPU when 8 ranks; here 2^19 per rank)
    n = 1 << 19
    rng = np.random.default_rng(0xAF04)
    kseeds = torch.from_numpy(rng.integers(0, 256, (N_KEYS, 32), dtype=np.uint8)).to(dev)
    d_exp = torch.empty((N_KEYS, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(kseeds, N_KEYS, d_exp)
    msgs = torch.randint(0, 256, (n, MSG_LEN), dtype=torch.uint8, device=dev, generator=g)
    off = torch.arange(n + 1, device=dev, dtype=torch.int64) * MSG_LEN
    ki = (torch.arange(n, device=dev) % N_KEYS).to(torch.int32)
    sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ms_sign = timed(lambda: ctx.sign_expanded_dev(d_exp, ki, msgs.view(-1), off, n, sigs))
    out["sign_512B_expanded_keys"] = {"signs_per_s": n / (ms_sign * 1e-3), "ms": ms_sign}
    seeds_full = kseeds[ki.long()].contiguous()
    ms_sign2 = timed(lambda: ctx.sign_dev(seeds_full, msgs.view(-1), off, n, sigs), reps=2)
    out["sign_512B_from_seeds"] = {"signs_per_s": n / (ms_sign2 * 1e-3), "ms": ms_sign2}
    soff = torch.arange(n + 1, device=dev, dtype=torch.int64) * 64
    root_dev = torch.empty(32, dtype=torch.uint8, device=dev)

    aud = afb.Auditor(ctx)
    empty_log = aud.save()

    def append_ms():
        aud.load(empty_log)                                   # same log object every time: its level buffers are allocated once
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        aud.append_dev(sigs.view(-1), soff, n)
        aud.root_dev(root_dev)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    append_ms()
    ms_m = min(append_ms() for _ in range(3))
    aud.close()
    out["merkle_append_64B_leaves"] = {"leaves_per_s": n / (ms_m * 1e-3), "ms": ms_m}
    # canonical form on the device: 2^19 VCDocument-shaped documents (23 values of 32 bytes), values as in real credentials (DIDs,
    # base64url hashes, timestamps: nothing to escape) and with 6 % of the bytes being characters Go escapes
    try:
        from agentfield_b200 import canonical as CA
        tmpl = CA.vc_document_template(False, ctx)
        F, flen = tmpl.n_fields, 32
        lib, P, st = afb._abi.load(), afb._abi.ptr, torch.cuda.current_stream().cuda_stream
        plain = list(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-_:/")
        kinds_t = torch.tensor(tmpl.kinds, device=dev)
        raw_cols = torch.nonzero(kinds_t == CA.RAW).flatten()
        d_voff = torch.arange(n * F + 1, device=dev, dtype=torch.int64) * flen
        res = {}
        for label, alphabet in (("plain_values", plain), ("6pct_escaped", plain + [0x22, 0x3c, 0x0a, 0x5c])):
            table = torch.tensor(alphabet, dtype=torch.uint8, device=dev)
            d_vals = table[torch.randint(0, table.numel(), (n * F * flen,), device=dev, generator=torch.Generator(device=dev).manual_seed(5))]
            d_vals.view(n, F, flen)[:, raw_cols, :] = 0x31          # RAW members must be valid JSON on their own: digits
            d_doc, d_doff = tmpl.fill_dev(d_vals, d_voff, n)
            total = int(d_doff[-1].item())

            def both_passes():                               # output buffers allocated once, as a long-running issuer would
                lib.afc_json_fill_sizes_dev(ctx.handle, P(tmpl.d_segs), P(tmpl.d_seg_off), P(tmpl.d_kinds), F, P(d_vals), P(d_voff), n, P(d_doff), None, st)
                lib.afc_json_fill_dev(ctx.handle, P(tmpl.d_segs), P(tmpl.d_seg_off), P(tmpl.d_kinds), F, P(d_vals), P(d_voff), n, P(d_doff), P(d_doc), st)
            both_passes(); torch.cuda.synchronize()
            ctx.profile_begin()                               # per-kernel CUDA events over the same launches the total is taken from
            ms_c = timed(both_passes, reps=5)
            kms = {k: v["avg_ms"] for k, v in ctx.profile_end().items()}
            res[label] = {"docs_per_s": n / (ms_c * 1e-3), "ms": ms_c, "kernels_ms": kms, "bytes_out": total, "bytes_in": int(d_vals.numel()),
                          "hbm_frac": (total + d_vals.numel() * 2 + 8 * (n * F + n)) / (ms_c * 1e-3) / 1e9 / hbm_peak}
            del d_doc, d_doff, d_vals
        res["note"] = "sizes pass + scan + fill pass into preallocated buffers; values are read twice; one document per thread"
        out["canonical_form_vc_documents"] = res
    except Exception as ex:
        out["canonical_form_vc_documents"] = {"error": repr(ex)}
    if world > 1:
        from agentfield_b200 import shard
        roots = shard.allgather_roots(bytes(root_dev.cpu().tolist()))
        out["merkle_global_root"] = afb.fold_roots(np.frombuffer(b"".join(roots), dtype=np.uint8), ctx).hex()
    # configs[4] shape on this rank: open-loop 100 k actions/s (sign + HMAC + audit append) through the native dispatcher
    try:
        ing = afb.Ingest(ctx.expand(rng.integers(0, 256, (64, 32), dtype=np.uint8)), ctx, batch_max=4096, linger_us=500, max_msg=512, max_key=32, max_body=256)
        out["ingest_soak_100k"] = ing.soak(100_000, 5.0, producers=4)
        out["ingest_soak_1M"] = ing.soak(1_000_000, 3.0, producers=8)
        ing.close()
    except Exception as ex:
        out["ingest_soak_100k"] = {"error": repr(ex)}
    mb = {}
    for name, which, iters in (("fe_mul", 0, 4000), ("fe_sq", 1, 4000), ("fe_addsub", 2, 20000), ("fe_mul_portable", 5, 2000),
                               ("fe_sq_via_mul", 6, 4000), ("fe_mul_schoolbook", 7, 4000), ("sha256_compress", 3, 2000), ("sha512_compress", 4, 1000),
                               ("pipe_imad32_x16", 11, 20000), ("pipe_alu_x16", 12, 20000),
                               ("probe_w8_a0", 20, 20000), ("probe_w8_a8", 21, 20000), ("probe_w8_a16", 22, 20000), ("probe_w8_a24", 23, 20000),
                               ("probe_w8_a32", 24, 20000), ("probe_w8_a48", 29, 20000), ("probe_w8_carrypairs16", 25, 20000),
                               ("probe_w8_xor16", 26, 20000), ("probe_wX8_a0", 27, 20000), ("probe_wX8_a16", 28, 20000)):
        ops, ms = ctx.microbench(which, iters)
        mb[name] = {"ops_per_s": ops, "ms": ms}
        if which >= 20:     # cycles per iteration per warp on one SMSP (32 warps/SM in the probe grid: 8 per SMSP)
            mb[name]["cycles_per_iter_per_smsp_warp"] = ms * 1e-3 * 1.965e9 / iters / 8
    out["microbench"] = mb
    return out


if __name__ == "__main__":
    main()
