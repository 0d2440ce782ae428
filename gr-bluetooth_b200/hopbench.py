"""bench.py --workload hopper: BASELINE configs[3] -- multi_hopper (follow one piconet: LAP + UAP/clock recovery, then
hop-along) on the bundled capture samples/keyboard1.cfile played N times back to back, through the C++ block
gr::bluetooth::multi_hopper::work() (host/btrx_b200 -l 4831dd -p --tile N).

Acquisition (UAP, CLK1-6, CLK1-27 by hop reversal) runs on the reference's chained semantics, one slot per work() call;
once CLK1-27 is known the block follows the piconet in BATCHES of slots (BTB200_MM_MODE=stateless): the hop channel of
every slot is known in advance, so a batch is one masked channel-window per slot (btb200_set_window_mask).  Tile
boundaries are discontinuities of the capture: the work per slot stays what the reference does.
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FS, FC, LAP = 8e6, 2476.5e6, "4831dd"


def _capture_i16():
    p = os.path.join(ROOT, "tests", "golden", "_samples", "keyboard1.i16.npy")
    if os.path.exists(p):
        return np.load(p)
    p = "/root/reference/samples/keyboard1.cfile"
    if os.path.exists(p):
        return np.fromfile(p, dtype=np.float32).astype(np.int16)
    return None


def _run_btrx(path, tile, env_extra, batch):
    exe = os.path.join(ROOT, "gr-bluetooth_b200", "host", "btrx_b200")
    env = dict(os.environ, BTB200_BATCH_SLOTS=str(batch), **env_extra)
    t0 = time.perf_counter()
    p = subprocess.run([exe, "-f", str(FC), "-r", str(FS), "-i", path, "-2", "-l", LAP, "-p", "--tile", str(tile), "--stats"],
                       capture_output=True, env=env)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("btrx_b200 failed: " + p.stderr.decode()[-1500:])
    st = json.loads(p.stderr.decode().strip().splitlines()[-1])
    out = p.stdout.decode()
    return st, out, wall


def cpu_reference(path_i16, tiles):
    """The reference's own multi_hopper (oracle/_ref/btref hop) on `tiles` copies of the capture, one thread: its block is
    one serial chain (piconet state + clock recovery), there is nothing to spread over cores."""
    sys.path.insert(0, ROOT)
    from oracle import ref as R
    if not R.available():
        return None
    x = np.fromfile(path_i16, dtype=np.int16)
    with tempfile.NamedTemporaryFile(suffix=".i16", delete=False) as f:
        for _ in range(tiles):
            x.tofile(f)
        big = f.name
    try:
        t0 = time.time()
        p = subprocess.run([R.BTREF, "hop", "--fs", str(FS), "--fc", str(FC), "--lap", LAP, "--in", big, "--i16"],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.time() - t0
        assert p.returncode == 0
    finally:
        os.unlink(big)
    n = len(x) // 2 * tiles
    return {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "reference",
            "sample": "multi_hopper on keyboard1 x%d (%.1f M samples), %.1f s wall" % (tiles, n / 1e6, dt)}


def run(args, world, rank, local):
    if rank != 0:
        return
    x = _capture_i16()
    if x is None:
        print(json.dumps({"metric": "complex-IQ Msamples/s via multi_hopper", "unavailable": "keyboard1 capture not staged"}))
        return
    tile = args.tile
    batch = 4096 if args.slots == 512 else args.slots      # 8 Msps: a slot is 5000 samples, so batches are long
    with tempfile.NamedTemporaryFile(suffix=".i16", delete=False) as f:
        x.tofile(f)
        path = f.name
    try:
        env = {"BTB200_MM_MODE": "stateless", "BTB200_DEVICE": str(local)}
        if args.ddc == "poly":
            env["BTB200_DDC"] = "polyphase"
        for _ in range(max(1, min(args.warmup, 2))):
            _run_btrx(path, max(1, tile // 20), env, batch)
        runs = [_run_btrx(path, tile, env, batch) for _ in range(max(1, min(args.steps, 3)))]
        st, out, wall = min(runs, key=lambda r: r[0]["seconds"])
        # the chained (reference-exact) run of ONE tile for comparison of what is decoded
        st1, out1, _ = _run_btrx(path, 1, {"BTB200_DEVICE": str(local)}, 16)
        cb = None if args.no_cpu else cpu_reference(path, 10)
    finally:
        os.unlink(path)
    n = st["samples"]
    lines = out.splitlines()
    line = {"metric": "complex-IQ Msamples/s via multi_hopper", "value": n / (st["device_ms"] / 1e3) / 1e6 if st["device_ms"] > 0 else None,
            "unit": "Msamples/s", "n_gpus": 1, "steps": len(runs), "warmup": min(args.warmup, 2),
            "ms_per_step": st["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "bundled capture samples/keyboard1.cfile (8 Msps, 8 channels) x%d" % tile,
            "config": {"workload": "multi_hopper LAP 4831dd: UAP/clock recovery on the chained state, then hop-along in batches of "
                                   "%d slots (one masked channel-window per slot), keyboard1 x%d = %.0f M samples (BASELINE configs[3])"
                                   % (batch, tile, n / 1e6),
                       "fs": FS, "fc": FC, "ddc": args.ddc, "l2": "input %.0f MiB per run, larger than L2" % (n * 8 / 2**20),
                       "timing": "value: sum of the batches' device times (CUDA events); e2e: wall clock inside btrx_b200 around the work() loop"},
            "e2e": {"value": st["msamples_per_s"], "unit": "Msamples/s", "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": None,
                    "api": "gr::bluetooth::multi_hopper::work() through host/btrx_b200 (complex64 host buffers)", "process_wall_s": round(wall, 2)},
            "gpu_launches": None,
            "decoded": {"lines_tiled_run": len(lines), "clock_lines_tiled_run": sum(l.startswith("clock 0x") for l in lines),
                        "clock_lines_one_tile_chained": sum(l.startswith("clock 0x") for l in out1.splitlines()),
                        "acquired": any("Acquired CLK1-27" in l for l in lines)},
            "cpu_baseline": cb}
    print(json.dumps(line))
