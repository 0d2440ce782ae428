#!/usr/bin/env python
"""bench.py -- complex-IQ Msamples/s through multi_sniffer (BASELINE.json metric).

One "step" = one pass of the whole receive path (79-channel DDC, squelch, GFSK demod, clock
recovery + slicer, access-code search) over one batch of synthetic 100 Msps IQ.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU)
  python bench.py --impl reference ...                      the reference's own CPU code
                                                            (oracle/_ref/btref) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FS, FC, SNR_DB = 100e6, 2441e6, 10.0
METRIC = "complex-IQ Msamples/s via multi_sniffer"
UNIT = "Msamples/s"
ALGO_BYTES_PER_SAMPLE = 8.0 + 79 * (1e6 / FS) / 8.0      # 8 B in + 1 bit/symbol/channel out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (one streaming nvidia-smi -lms process)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def _run(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.decode(errors="replace").strip().split(",")])

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
            time.sleep(0.15)          # first sample is out before the timed region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.th.join(timeout=5)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_batch(n_slots, seed):
    """Synthetic capture covering n_slots windows that all lie inside the stream (no zero history)."""
    from gr_bluetooth_b200 import synth
    S = int(625 * FS / 1e6)
    lead = 7                                           # ceil((H-1)/S) slots of history in front
    iq, truth = synth.generate(FS, FC, n_slots + lead, seed=seed)
    return iq, truth, lead, S


def cpu_baseline(threads, slots):
    """The reference's CPU path on the same kind of input, bounded sample: `slots` work() calls of
    the 100 Msps / 79-channel configuration on `threads` host threads."""
    from oracle import ref as R
    from oracle import oracle as O
    iq, _, lead, S = synth_batch(slots, seed=99)
    kind = "reference" if R.available() else "port"
    t0 = time.time()
    if kind == "reference":
        with tempfile.NamedTemporaryFile(suffix=".cfile", delete=False) as f:
            iq.tofile(f)
            path = f.name
        try:
            per = [slots // threads + (1 if i < slots % threads else 0) for i in range(threads)]
            procs, first = [], lead
            t0 = time.time()
            for n in per:
                if n:
                    procs.append(subprocess.Popen([R.BTREF, "sniff", "--fs", str(FS), "--fc", str(FC), "--snr", str(SNR_DB),
                                                   "--in", path, "--stateless", "--first-call", str(first),
                                                   "--num-calls", str(n)], stdout=subprocess.DEVNULL,
                                                  stderr=subprocess.DEVNULL))
                    first += n
            for p in procs:
                assert p.wait() == 0
            dt = time.time() - t0
        finally:
            os.unlink(path)
    else:
        P = O.Plan(FS, FC, SNR_DB)
        t0 = time.time()
        P.run(iq, first_call=lead, num_calls=slots, stateless=True, threads=threads)
        dt = time.time() - t0
    return {"value": slots * S / dt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d work() calls (slots) of synthetic 100 Msps / 79-channel IQ, stateless mode, %.1f s wall" % (slots, dt)}


_CPU_PLAN = {}


def cpu_plan(max_threads):
    """How the reference uses this box best: the direct-form 20001-tap FIRs are cache/memory bound, so more threads
    than the memory system feeds do not help.  Probe all, a half, a quarter and an eighth of the host threads (one slot
    per thread each) and keep the fastest; size the timed sample to ~15 s of wall time.  -> (threads, slots)"""
    if "plan" not in _CPU_PLAN:
        best = None
        for t in sorted({max_threads, max(max_threads // 2, 1), max(max_threads // 4, 1), max(max_threads // 8, 1)}, reverse=True):
            r = cpu_baseline(t, t)
            wall = t * int(625 * FS / 1e6) / (r["value"] * 1e6)
            if best is None or r["value"] > best[0]:
                best = (r["value"], t, wall)
        rounds = int(min(8, max(1, round(15.0 / best[2]))))
        _CPU_PLAN["plan"] = (best[1], max(best[1] * rounds, 8))
    return _CPU_PLAN["plan"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads, slots = cpu_plan(os.cpu_count() or 1)
    vals = []
    for i in range(args.warmup + args.steps):
        # warm-up steps only page the binary and the input in: a small sample (8 slots on 8 threads, a few seconds);
        # timed steps use every host thread, one slot each
        cb = cpu_baseline(threads, slots) if i >= args.warmup else cpu_baseline(min(threads, 8), 8)
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    S = int(625 * FS / 1e6)
    cb = dict(vals[-1], value=v)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": slots * S / v / 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "multi_sniffer 79-ch, synthetic 100 Msps IQ (BASELINE configs[2])",
                       "step": "%d slots on %d host threads" % (slots, threads)},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


FP32_NONFUSED_PEAK = 35.6e12     # thread-level FMUL/FADD per second measured on this pool's B200 by tools/ubench_f32x2.cu


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gr_bluetooth_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    B = args.slots
    iq, truth, lead, S = synth_batch(B, seed=1234 + rank)          # every rank its own time shard
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(snr_mode, steps, warmup):
        mk = lambda: g.multi_sniffer.make(FS, FC, SNR_DB, False, mm_mode=g.MM_STATELESS, device=local, max_slots=B,
                                          snr_mode=g.SNR_FAST_GUARDED if snr_mode == "fast" else g.SNR_EXACT,
                                          tail=g.TAIL_FULL if args.tail == "full" else g.TAIL_LAZY)
        NCTX = args.e2e_contexts                  # the e2e loop keeps NCTX - 1 batches in flight behind the one collected
        blks = [mk() for _ in range(NCTX)]
        blk = blks[0]
        H = blk.history()
        w0 = lead * S - (H - 1)
        n_in = (B - 1) * S + H
        pinned = [g.PinnedBuffer(n_in) for _ in range(NCTX)]
        for p in pinned:
            p.array[:] = iq[w0:w0 + n_in]
        d_iq = torch.from_numpy(pinned[0].array.view(np.float32).copy()).to(dev)     # resident copy for `value`

        # ---- device-resident throughput (`value`) ----
        for _ in range(warmup):
            hits, _, _ = blk.process_device(d_iq.data_ptr(), n_in, lead, B, want_symbols=True)
        stage_ms = {}
        barrier()
        l0 = blk.launch_count()
        with ClockSampler(local) as clk:
            dev_ms = 0.0
            for _ in range(steps):
                flush.zero_()                                   # L2 flush between timed iterations
                hits, _, _ = blk.process_device(d_iq.data_ptr(), n_in, lead, B, want_symbols=True)
                tm = blk.last_timing()
                dev_ms += tm["total"]
                for k, v in tm.items():
                    stage_ms[k] = stage_ms.get(k, 0.0) + v / steps
            barrier()
        launches = blk.launch_count() - l0
        dev_ms = allmax(dev_ms)                   # CUDA events on the ctx stream, max over ranks
        total_samples = world * steps * B * S
        value = total_samples / (dev_ms / 1e3) / 1e6

        # ---- end to end through the public calls with HOST buffers: btb200_submit (H2D copy + kernels)
        #      / btb200_collect (hits + symbols D2H), double-buffered over two contexts ----
        def e2e_loop(n):
            # batch i lives in context i % NCTX.  Per step: enqueue the deferred work of the oldest batch, submit a
            # new batch behind it (its input copy overlaps the queued kernels), then wait for the oldest batch.
            res = None
            for i in range(min(NCTX - 1, n)):
                blks[i % NCTX].submit(pinned[i % NCTX].ptr.value, False, n_in, lead, B)
            for j in range(n):
                blks[j % NCTX].collect_begin()
                i = j + NCTX - 1
                if i < n:
                    blks[i % NCTX].submit(pinned[i % NCTX].ptr.value, False, n_in, lead, B)
                res = blks[j % NCTX].collect(want_symbols=True)
            return res
        e2e_loop(NCTX)
        barrier()
        with ClockSampler(local) as clk_e2e:
            t0 = time.perf_counter()
            ehits, esyms, _ = e2e_loop(steps)
            barrier()
            e2e_s = allmax(time.perf_counter() - t0)
        e2e_val = total_samples / e2e_s / 1e6
        d2h = int(len(ehits) * 56 + len(esyms) + 16 + (8 * B * blk.info.n_channels if snr_mode == "fast" else 0))
        nwin = len({(int(h["slot"]), int(h["channel"])) for h in hits})
        res = dict(value=value, ms_per_step=dev_ms / steps, e2e=e2e_val, e2e_ms_per_step=e2e_s / steps * 1e3,
                   stage_ms={k: round(v, 3) for k, v in stage_ms.items()}, launches=int(launches), clocks=clk.summary(),
                   clocks_e2e=clk_e2e.summary(),
                   h2d=int(n_in * 8), d2h=d2h, hits=hits, n_in=n_in, hit_windows=nwin, info=blk.info)
        for bk in blks:
            bk.close()
        for p in pinned:
            p.close()
        del d_iq
        torch.cuda.empty_cache()
        return res

    main = measure(args.snr_mode, args.steps, args.warmup)
    alt = None
    if not args.no_alt:
        other = "fast" if args.snr_mode == "exact" else "exact"
        alt = measure(other, max(2, args.steps // 2), 3)
    # everything that needs the other ranks is done: release them before rank 0 times the CPU baseline
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    if rank == 0:
        I = main["info"]
        peak, peak_src = measured_peaks()
        st = main["stage_ms"]
        dom = max(("chan_fir", "noise_fir", "energy", "demod_mm", "search"), key=lambda k: st[k])
        dom_s = st[dom] / 1e3
        achieved = ALGO_BYTES_PER_SAMPLE * B * S / dom_s / 1e9
        # fp32 work of the two FIR stages (complex MAC = 4 FMUL + 4 FADD, no FMA on the exact path)
        gtot = (B - 1) * (S // I.decimation) + I.ddc_out_per_window
        cmac = {"chan_fir": gtot * I.n_channels * I.chan_taps,
                "noise_fir": main["hit_windows"] * I.noise_out_per_window * I.noise_taps}
        fp32 = {k: {"cmac_per_launch": int(v), "fp32_ops_per_s": 8 * v / (st[k] / 1e3),
                    "frac_of_nonfused_peak": 8 * v / (st[k] / 1e3) / FP32_NONFUSED_PEAK}
                for k, v in cmac.items() if st.get(k, 0) > 0.05}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("slots") == B and dom in tj.get("dram_bytes_per_launch", {}):
                traffic = tj["dram_bytes_per_launch"][dom]
        hits = main["hits"]
        found = {(int(h["channel"]), int(h["lap"])) for h in hits if h["kind"] == 0}
        expect = {(t_["channel"], t_["lap"]) for t_ in truth if t_["slot"] <= B - 2}
        # the CPU leg is timed on rank 0 at N = 1 only (it is the same number at every N)
        cb = cpu_baseline(*cpu_plan(os.cpu_count() or 1)) if (not args.no_cpu and world == 1) else None
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "multi_sniffer 79-ch, synthetic 100 Msps IQ (BASELINE configs[2]), "
                                       "stateless mode, %d slots (%.1f M samples, %.0f MiB) per step per GPU"
                                       % (B, B * S / 1e6, main["n_in"] * 8 / 2**20),
                           "fs": FS, "fc": FC, "channels": I.n_channels, "slots_per_step": B, "snr_mode": args.snr_mode,
                           "l2": "flushed between timed iterations (256 MiB write); input %.0f MiB" % (main["n_in"] * 8 / 2**20),
                           "timing": "value: CUDA events on the ctx stream; e2e: wall clock between barriers; max over ranks",
                           "sharding": "time shards, no collective"},
                "e2e": {"value": main["e2e"], "unit": UNIT, "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"],
                        "ms_per_step": main["e2e_ms_per_step"], "clocks": main["clocks_e2e"],
                        "api": "btb200_submit/btb200_collect with pinned host buffers, %d contexts in flight on the shared compute stream" % args.e2e_contexts},
                "gpu_launches": main["launches"],
                "clocks": main["clocks"],
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "note": "the path is bound by the non-fused fp32 rate (exact-order FIRs), not by HBM: see `fp32`"},
                "fp32": {"peak_ops_per_s": FP32_NONFUSED_PEAK, "peak_source": "tools/ubench_f32x2.cu on this pool (FMUL+FADD)",
                         "kernels": fp32},
                "stage_ms": st,
                "detect": {"truth_bursts": len(expect), "found": len(expect & found), "hit_windows": main["hit_windows"]},
                "cpu_baseline": cb}
        if alt is not None:
            line["alt_snr_mode"] = {"snr_mode": "fast" if args.snr_mode == "exact" else "exact", "value": alt["value"],
                                    "e2e": alt["e2e"], "stage_ms": alt["stage_ms"], "unit": UNIT}
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slots", type=int, default=512, help="slots (625 us each) per step per GPU")
    ap.add_argument("--no-alt", action="store_true", help="skip the second measurement in the other snr mode")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-contexts", type=int, default=3,
                    help="contexts the end-to-end loop cycles through (batches in flight = contexts)")
    ap.add_argument("--tail", default="lazy", choices=["lazy", "full"],
                    help="lazy (default): clock recovery past the searchable prefix only for windows with hits")
    ap.add_argument("--snr-mode", default="exact", choices=["exact", "fast"],
                    help="exact: reference arithmetic for every printed snr; fast: guarded polyphase estimate")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
