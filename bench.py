#!/usr/bin/env python
"""bench.py -- complex-IQ Msamples/s through multi_sniffer (BASELINE.json metric).

One "step" = one pass of the whole receive path (79-channel DDC, squelch, GFSK demod, clock recovery + slicer,
access-code search, symbols of every hit back on the host) over one batch of synthetic wideband IQ.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU)
  python bench.py --impl reference ...                      the reference's own CPU code (oracle/_ref/btref)
  python bench.py --workload ble|hopper                     BASELINE configs[4] / configs[3] instead of configs[1-2]

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

SNR_DB = 10.0
METRIC = "complex-IQ Msamples/s via multi_sniffer"
UNIT = "Msamples/s"
LEAD = 7                                               # ceil((H-1)/S) slots of history in front of the first window

WORKLOADS = {
    # name: (fs, fc, description)
    "sniffer": (100e6, 2441e6, "multi_sniffer 79-ch, synthetic 100 Msps IQ (BASELINE configs[1-2])"),
    "ble": (30e6, 2414e6, "BLE advertising/data channel sniffer path (sniff_aa), 27-ch, synthetic 30 Msps IQ (BASELINE configs[4])"),
}


def algo_bytes_per_sample(fs, nch):
    return 8.0 + nch * (1e6 / fs) / 8.0                # 8 B in + 1 bit per symbol per channel out


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


def pin_to_gpu_numa_node(local):
    """Bind this rank to the CPUs of its GPU's NUMA node BEFORE the pinned host buffers are allocated, so the H2D
    copies do not cross the socket interconnect (8 ranks x ~130 MB per step)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ---------------------------------------------------------------- CPU reference arm

def synth_batch(fs, fc, n_slots, seed, workload="sniffer"):
    """Synthetic capture covering n_slots windows that all lie inside the stream (no zero history)."""
    from gr_bluetooth_b200 import synth
    S = int(625 * fs / 1e6)
    kw = dict(le_adv_occupancy=0.02) if workload == "ble" else {}
    iq, truth = synth.generate(fs, fc, n_slots + LEAD, seed=seed, **kw)
    return iq, truth, LEAD, S


def cpu_baseline(threads, slots, workload="sniffer"):
    """The reference's CPU path on the same kind of input, bounded sample: `slots` work() calls on `threads` host
    threads (its block is single-threaded; the slots are dealt to processes the way our time shards are)."""
    from oracle import ref as R
    from oracle import oracle as O
    fs, fc, _ = WORKLOADS[workload]
    iq, _, lead, S = synth_batch(fs, fc, slots, seed=99, workload=workload)
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
                    procs.append(subprocess.Popen([R.BTREF, "sniff", "--fs", str(fs), "--fc", str(fc), "--snr", str(SNR_DB),
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
        P = O.Plan(fs, fc, SNR_DB)
        t0 = time.time()
        P.run(iq, first_call=lead, num_calls=slots, stateless=True, threads=threads)
        dt = time.time() - t0
    return {"value": slots * S / dt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d work() calls (slots) of synthetic %g Msps IQ, stateless mode, %.1f s wall" % (slots, fs / 1e6, dt)}


_CPU_PLAN = {}


def cpu_plan(max_threads, workload="sniffer"):
    """How the reference uses this box best: the direct-form 20001-tap FIRs are cache/memory bound, so more threads
    than the memory system feeds do not help.  Probe all, a half, a quarter and an eighth of the host threads (one slot
    per thread each) and keep the fastest; size the timed sample to ~15 s of wall time.  -> (threads, slots)"""
    if workload not in _CPU_PLAN:
        best = None
        fs = WORKLOADS[workload][0]
        for t in sorted({max_threads, max(max_threads // 2, 1), max(max_threads // 4, 1), max(max_threads // 8, 1)}, reverse=True):
            r = cpu_baseline(t, t, workload)
            wall = t * int(625 * fs / 1e6) / (r["value"] * 1e6)
            if best is None or r["value"] > best[0]:
                best = (r["value"], t, wall)
        rounds = int(min(8, max(1, round(15.0 / best[2]))))
        _CPU_PLAN[workload] = (best[1], max(best[1] * rounds, 8))
    return _CPU_PLAN[workload]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload if args.workload in WORKLOADS else "sniffer"
    fs, fc, desc = WORKLOADS[wl]
    threads, slots = cpu_plan(os.cpu_count() or 1, wl)
    vals = []
    for i in range(args.warmup + args.steps):
        # warm-up steps only page the binary and the input in: a small sample (8 slots on 8 threads, a few seconds);
        # timed steps use the best thread count, several slots each
        cb = cpu_baseline(threads, slots, wl) if i >= args.warmup else cpu_baseline(min(threads, 8), 8, wl)
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    S = int(625 * fs / 1e6)
    cb = dict(vals[-1], value=v)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": slots * S / v / 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "step": "%d slots on %d host threads" % (slots, threads)},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import gr_bluetooth_b200 as g
    from gr_bluetooth_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)

    if args.workload == "hopper":
        from gr_bluetooth_b200 import hopbench
        return hopbench.run(args, world, rank, local)
    fs, fc, desc = WORKLOADS[args.workload]
    search = g.SEARCH_LE if args.workload == "ble" else (g.SEARCH_BR | g.SEARCH_LE)
    gen_kw = dict(le_adv_occupancy=0.02) if args.workload == "ble" else {}
    B = args.slots
    S = int(625 * fs / 1e6)

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

    def shard(r, nslots=B, occupancy=args.occupancy):
        """ONE synthetic stream; rank r owns its slots [LEAD + r B, LEAD + (r+1) B) and reads LEAD slots (> H-1 samples)
        of guard in front of them.  int16-valued, like the bundled captures."""
        return synth.generate_range(fs, fc, r * nslots, nslots + LEAD, seed=1234, occupancy=occupancy, as_int16=True, **gen_kw)

    xi16, truth = shard(rank)

    def h2d_probe():
        """Plain pinned-host -> device copies of one batch's int16 input on every rank at once: the feed rate this box
        gives each GPU when all of them pull (the ceiling of any end-to-end number at this N)."""
        n = ((B - 1) * S + int(625 * fs / 1e6 * 6.4)) * 4
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        for _ in range(2):
            d.copy_(h, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(8):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = 8 * n / dt / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        lo = t.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        del h, d
        return {"min_gbs_per_gpu": round(float(lo.item()), 1), "aggregate_gbs": round(float(t.item()), 1)}

    feed = h2d_probe()
    numa_all = [numa]
    if world > 1:
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)

    def make_block(ddc, nslots=B):
        return g.multi_sniffer.make(fs, fc, SNR_DB, False, mm_mode=g.MM_STATELESS, device=local, max_slots=nslots,
                                    search=search, ddc=g.DDC_POLYPHASE if ddc == "poly" else g.DDC_EXACT,
                                    snr_mode=g.SNR_FAST_GUARDED if (ddc == "exact" and args.snr_mode == "fast") else g.SNR_EXACT)

    def pipelined(blks, submit, n):
        """batch i lives in context i % NCTX.  Per step: enqueue the deferred work of the oldest batch, submit a new
        batch behind it (its input copy overlaps the queued kernels), then wait for the oldest batch and take its hits
        and their symbols (borrowed: they stay in the library's pinned arena, no second host copy)."""
        nctx, res = len(blks), None
        for i in range(min(nctx - 1, n)):
            submit(i % nctx)
        for j in range(n):
            blks[j % nctx].collect_begin()
            i = j + nctx - 1
            if i < n:
                submit(i % nctx)
            res = blks[j % nctx].collect(want_symbols="borrow")
        return res[0].copy(), res[1].copy(), res[2]

    def measure(ddc, x16, steps, warmup, nslots=B, stages=True, e2e=True, first_slot=None):
        first_slot = LEAD + rank * nslots if first_slot is None else first_slot
        NCTX = args.e2e_contexts
        blks = [make_block(ddc, nslots) for _ in range(NCTX)]
        blk = blks[0]
        H = blk.history()
        w0 = LEAD * S - (H - 1)
        n_in = (nslots - 1) * S + H
        seg16 = x16[2 * w0:2 * (w0 + n_in)]
        d_iq = torch.from_numpy(seg16.astype(np.float32)).to(dev)                    # resident complex64 copy for `value`
        res = dict(n_in=n_in, info=blk.info)

        # ---- stage times of single batches (kernel-level numbers for the roofline; not the throughput figure)
        stage_ms = {}
        if stages:
            for i in range(3 + 3):
                hits, _, _ = blk.process_device(d_iq.data_ptr(), n_in, first_slot, nslots, want_symbols=True)
                if i >= 3:
                    for k, v in blk.last_timing().items():
                        stage_ms[k] = stage_ms.get(k, 0.0) + v / 3
            if ddc == "poly":
                stage_ms["resume"] = stage_ms.pop("energy")       # slot [3] of btb200_last_timing in the throughput mode
            res["stage_ms"] = {k: round(v, 3) for k, v in stage_ms.items()}

        # ---- `value`: input resident in HBM, batches pipelined over NCTX contexts exactly like the e2e loop, timed
        #      with CUDA events on the device's compute stream (which every context of the device shares)
        sub_dev = lambda c: blks[c].submit(d_iq.data_ptr(), True, n_in, first_slot, nslots)
        pipelined(blks, sub_dev, max(warmup, NCTX))
        barrier()
        l0 = sum(b_.launch_count() for b_ in blks)
        with ClockSampler(local) as clk:
            blk.timer_start()
            hits, syms, _ = pipelined(blks, sub_dev, steps)
            dev_ms = blk.timer_stop()
            barrier()
        res["launches"] = sum(b_.launch_count() for b_ in blks) - l0
        dev_ms = allmax(dev_ms)
        total = world * steps * nslots * S
        res.update(value=total / (dev_ms / 1e3) / 1e6, ms_per_step=dev_ms / steps, clocks=clk.summary(), hits=hits,
                   hit_windows=len({(int(h["slot"]), int(h["channel"])) for h in hits}))

        # ---- end to end through the public calls with HOST buffers: btb200_submit[_i16] (H2D copy + kernels) /
        #      btb200_collect (hits + symbols D2H), pipelined over NCTX contexts
        if e2e:
            for kind in (["i16", "c64"] if args.e2e_both else [args.input]):
                pinned = [g.PinnedBuffer(n_in, i16=(kind == "i16")) for _ in range(NCTX)]
                for p in pinned:
                    if kind == "i16":
                        p.array[:] = seg16
                    else:
                        p.array.view(np.float32)[:] = seg16
                if kind == "i16":
                    sub = lambda c: blks[c].submit_i16(pinned[c].ptr.value, False, n_in, first_slot, nslots)
                else:
                    sub = lambda c: blks[c].submit(pinned[c].ptr.value, False, n_in, first_slot, nslots)
                pipelined(blks, sub, NCTX)
                barrier()
                with ClockSampler(local) as clk_e2e:
                    t0 = time.perf_counter()
                    ehits, esyms, _ = pipelined(blks, sub, steps)
                    barrier()
                    e2e_s = allmax(time.perf_counter() - t0)
                h2d = int(n_in * (4 if kind == "i16" else 8))
                d2h = int(len(ehits) * 56 + len(esyms) + 16 + 16 * nslots * blk.info.n_channels * (1 if ddc == "poly" else 0))
                res["e2e_" + kind] = dict(value=total / e2e_s / 1e6, ms_per_step=e2e_s / steps * 1e3, h2d=h2d, d2h=d2h,
                                          h2d_gbs_per_gpu=h2d / (e2e_s / steps) / 1e9, clocks=clk_e2e.summary(),
                                          same_hits=bool(np.array_equal(ehits, hits)))
                for p in pinned:
                    p.close()
        for bk in blks:
            bk.close()
        del d_iq
        torch.cuda.empty_cache()
        return res

    main = measure(args.ddc, xi16, args.steps, args.warmup)
    alt = None
    if not args.no_alt:
        other = "exact" if args.ddc == "poly" else "poly"
        alt = measure(other, xi16, max(3, args.steps // 3), 3)

    # ---- occupancy sweep (N = 1): the throughput figure must not hinge on 5 % of the windows carrying a hit
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = {}
        nsl = B
        for occ in (0.05, 0.2, 0.5):
            x_o, _ = shard(0, nsl, occ)
            r_o = measure(args.ddc, x_o, max(3, args.steps // 2), 3, nslots=nsl, stages=False, e2e=False, first_slot=LEAD)
            sweep["%.2f" % occ] = {"value": round(r_o["value"], 1), "hit_windows_per_step": r_o["hit_windows"]}
        sweep["slots_per_step"] = nsl

    # ---- one stream, N shards: rank 0 recomputes every shard on ITS GPU (= the 1-GPU run of the same stream, batch by
    #      batch) and compares with what the shard's own rank returned; detection vs the generator's ground truth
    shard_equal, detect = None, None
    my_hits = main["hits"]
    all_truth = [truth]
    if world > 1:
        path = "/dev/shm/btb200_shard_%d.i16" % rank
        xi16.tofile(path)
        gathered, all_truth = [None] * world, [None] * world
        dist.all_gather_object(gathered, my_hits)
        dist.all_gather_object(all_truth, truth)
        dist.barrier()
        if rank == 0:
            blk = make_block(args.ddc)
            H = blk.history()
            w0, n_in = LEAD * S - (H - 1), (B - 1) * S + H
            shard_equal = True
            for r in range(world):
                xr_ = np.fromfile("/dev/shm/btb200_shard_%d.i16" % r, dtype=np.int16)
                h_r, _, _ = blk.process_i16(xr_[2 * w0:2 * (w0 + n_in)], LEAD + r * B, B, want_symbols=True)
                keys = ["slot", "channel", "kind", "offset", "n_symbols", "lap", "flags", "snr", "sym_count"]   # not sym_offset: arena layout
                shard_equal = shard_equal and len(h_r) == len(gathered[r]) and bool(np.array_equal(h_r[keys], gathered[r][keys]))
            blk.close()
            my_hits = np.concatenate(gathered)
        dist.barrier()
        os.unlink(path)
    if rank == 0:
        # a burst that starts in slot s is reported by work() call s + 6 or s + 7 (window lag (H-1)/S = 6.3 slots)
        kind = 1 if args.workload == "ble" else 0
        seen = {(int(h["slot"]), int(h["channel"]), int(h["lap"])) for h in my_hits if h["kind"] == kind}
        last_call = LEAD + world * B - 1
        expect = {(t_["slot"], t_["channel"], t_["lap"]) for tl in all_truth for t_ in tl
                  if t_.get("kind", 0) == kind and t_["slot"] + 6 >= LEAD and t_["slot"] + 7 <= last_call}
        found = sum(1 for (s_, c_, l_) in expect if any((s_ + d_, c_, l_) in seen for d_ in (5, 6, 7, 8)))
        detect = {"truth_bursts": len(expect), "found": found, "hits": int(len(my_hits)), "hit_windows_per_step": main["hit_windows"]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

    if rank == 0:
        I = main["info"]
        peak, peak_src = measured_peaks()
        st = main["stage_ms"]
        stage_keys = [k for k in ("chan_fir", "noise_fir", "energy", "resume", "demod_mm", "search") if k in st]
        dom = max(stage_keys, key=lambda k: st[k])
        algo = algo_bytes_per_sample(fs, I.n_channels) * B * S
        ach_dom = algo / (st[dom] / 1e3) / 1e9
        ach_step = algo / (main["ms_per_step"] / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("slots") == B and tj.get("ddc") == args.ddc and tj.get("workload") == args.workload:
                traffic = tj.get("dram_bytes_per_step")
        # fp32 roof from the SM clock sampled during the timed region (128 FMA lanes per SM)
        clk_mhz = main["clocks"]["sm_mhz"] or 1965
        fma_peak = I.sm_count * 128 * clk_mhz * 1e6
        M = int(round(fs / 1e6))
        gtot = (B - 1) * (S // I.decimation) + I.ddc_out_per_window
        fp32 = {"peak_fma_per_s": fma_peak, "peak_source": "sm_count x 128 lanes x sampled SM clock (%d MHz)" % clk_mhz, "kernels": {}}
        if args.ddc == "poly":
            # the estimator evaluates every 4th noise-DDC output by default (BTB200_NEST_FOLD: 0 all, 1 even, 2 every 4th)
            nest_stride = {0: 1, 1: 2}.get(int(os.environ.get("BTB200_NEST_FOLD", "2")), 4)
            if os.environ.get("BTB200_NEST_DENSE"):
                nest_stride = 1
            n_est = (I.noise_out_per_window - 1) // nest_stride + 1 + (2 if nest_stride == 4 else 1 if nest_stride == 2 else 0)
            work = {"chan_fir": gtot * (M * 7 * 2 + 4.0 * M * I.n_channels / (4 if M % 4 == 0 and (M // 4) % 2 else 2 if (M // 2) % 2 else 1)),
                    "noise_fir": B * n_est * (2.0 * I.noise_taps + 4.0 * M * I.n_channels / (4 if M % 4 == 0 and (M // 4) % 2 else 2 if (M // 2) % 2 else 1))}
        else:
            work = {"chan_fir": 4.0 * gtot * I.n_channels * I.chan_taps}
        for k, w in work.items():
            if st.get(k, 0) > 0.02:
                fp32["kernels"][k] = {"fma_per_launch": int(w), "fma_per_s": w / (st[k] / 1e3), "frac_of_fma_peak": w / (st[k] / 1e3) / fma_peak}
        cb = cpu_baseline(*cpu_plan(os.cpu_count() or 1, args.workload), args.workload) if (not args.no_cpu and world == 1) else None
        e2e = main.get("e2e_" + args.input)
        mode_txt = {"poly": "throughput mode: polyphase channelizer + fused demod, tolerance-level floats (tests/test_gpu_polyphase.py)",
                    "exact": "exact mode: direct-form FIRs in the reference's operation order, bit-exact"}
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s, stateless mode, %d slots (%.1f M samples) per step per GPU; %s"
                                       % (desc, B, B * S / 1e6, mode_txt[args.ddc]),
                           "fs": fs, "fc": fc, "channels": I.n_channels, "slots_per_step": B, "ddc": args.ddc,
                           "occupancy": args.occupancy,
                           "l2": "inputs larger than L2: %.0f MiB (complex64) per batch, %d batches in flight on separate buffers"
                                 % (main["n_in"] * 8 / 2**20, args.e2e_contexts),
                           "timing": "value: CUDA events on the device's compute stream around K pipelined steps; "
                                     "e2e: wall clock between barriers; max over ranks",
                           "sharding": "one synthetic stream, contiguous slot ranges per rank with (H-1)-sample guard, no collective",
                           "numa_node_per_rank": numa_all,
                           "h2d_copy_only": feed},
                "e2e": {"value": e2e["value"], "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                        "ms_per_step": e2e["ms_per_step"], "h2d_gbs_per_gpu": round(e2e["h2d_gbs_per_gpu"], 2),
                        "clocks": e2e["clocks"], "same_hits_as_device_run": e2e["same_hits"],
                        "api": "btb200_submit%s/btb200_collect_begin/btb200_collect with pinned host buffers (%s input), "
                               "%d contexts in flight on the shared compute stream"
                               % ("_i16" if args.input == "i16" else "", "int16 interleaved" if args.input == "i16" else "complex64",
                                  args.e2e_contexts)},
                "gpu_launches": main["launches"],
                "clocks": main["clocks"],
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach_dom, "peak": peak, "unit": "GB/s",
                             "frac": ach_dom / peak, "traffic": traffic, "peak_source": peak_src,
                             "achieved_step": ach_step, "frac_step": ach_step / peak,
                             "algorithmic_bytes_per_launch": int(algo),
                             "note": "frac: dominant kernel alone; frac_step: on the whole pipelined step time.  The path is "
                                     "fp32-FMA bound, see `fp32` (SURVEY 8d)"},
                "fp32": fp32,
                "stage_ms": st,
                "detect": detect,
                "cpu_baseline": cb}
        other_in = "c64" if args.input == "i16" else "i16"
        if "e2e_" + other_in in main:
            o = main["e2e_" + other_in]
            line["e2e_" + ("complex64" if other_in == "c64" else "int16")] = {"value": o["value"], "h2d_bytes_per_step": o["h2d"], "ms_per_step": o["ms_per_step"]}
        if shard_equal is not None:
            line["shard_equal"] = shard_equal
        if sweep is not None:
            line["occupancy_sweep"] = sweep
        if alt is not None:
            ae = alt.get("e2e_" + args.input)
            line["exact_mode" if args.ddc == "poly" else "poly_mode"] = {
                "value": alt["value"], "e2e": ae["value"], "ms_per_step": alt["ms_per_step"], "stage_ms": alt["stage_ms"], "unit": UNIT,
                "hits_in_common_with_headline": int(len(set(map(_key, alt["hits"])) & set(map(_key, main["hits"])))),
                "hits": int(len(alt["hits"]))}
        print(json.dumps(line))


def _key(h):
    return (int(h["slot"]), int(h["channel"]), int(h["kind"]), int(h["lap"]) if h["kind"] == 0 else 0, int(h["offset"]) if h["kind"] else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sniffer", choices=["sniffer", "ble", "hopper"])
    ap.add_argument("--slots", type=int, default=512, help="slots (625 us each) per step per GPU")
    ap.add_argument("--ddc", default="poly", choices=["poly", "exact"],
                    help="poly (default): throughput mode, polyphase channelizer; exact: bit-exact direct-form FIRs")
    ap.add_argument("--input", default="i16", choices=["i16", "c64"], help="host sample format of the end-to-end loop")
    ap.add_argument("--e2e-both", action="store_true", help="time the end-to-end loop with both input formats")
    ap.add_argument("--occupancy", type=float, default=0.05, help="probability of a burst per (slot, channel)")
    ap.add_argument("--no-alt", action="store_true", help="skip the second measurement in the other ddc mode")
    ap.add_argument("--no-sweep", action="store_true", help="skip the occupancy sweep")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-contexts", type=int, default=3,
                    help="contexts the pipelined loops cycle through (batches in flight = contexts)")
    ap.add_argument("--tile", type=int, default=200, help="hopper workload: how many times the capture is played (BASELINE: 1000)")
    ap.add_argument("--snr-mode", default="exact", choices=["exact", "fast"],
                    help="exact-mode runs only: exact = reference arithmetic for every printed snr; fast = guarded estimate")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
