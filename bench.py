#!/usr/bin/env python
"""bench.py -- arcs/sec of log-semiring forward-backward on synthetic lattices.

Contract: `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun) prints ONE
JSON line on rank 0.  A "step" is one pass of the hot path over one batch of synthetic lattices:
logZ + beta (first kernel) and the arc posteriors (second kernel) -- the sliced-column pull / flow
kernels for wide lattices, the CSR forward / fused-backward kernels otherwise.

  value      whole-job arcs/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST (pinned) buffers: per step the
             packed batch + scores are copied H2D, logZ[B] is read back D2H
  roofline   dominant kernel of the step: its algorithmic bytes (DESIGN.md section 4 / SURVEY.md 8d)
             / its mean CUDA-event duration, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  oracle/lattice_oracle.c (a C port of the reference recurrence) on the
             host cores, bounded sample of the same workload

`--impl reference` times that CPU port alone (the reference itself is Python + OpenFst
wrappers and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "arcs/sec forward-backward (log semiring)"
UNIT = "arcs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="nfst_b200", choices=["nfst_b200", "reference"])
    ap.add_argument("--workload", default="dag", choices=["dag", "translit", "snips", "cipher-uni", "cipher-bi"])
    ap.add_argument("--arcs", type=int, default=100_000, help="arcs per lattice (dag workload)")
    ap.add_argument("--batch", type=int, default=1024, help="lattices per GPU (weak scaling)")
    ap.add_argument("--levels", type=int, default=64)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-theta", action="store_true", help="skip the theta-mode training-step leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the other points of the config-4 sweep (N=1 only)")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 1, 2, 3 and 5 at their stated batch (N=1 only; --no-sweep skips them too)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_name(a) -> str:
    if a.workload == "dag":
        return (f"config4 random-DAG sweep point: {a.arcs} arcs/lattice, S=A/4, {a.levels} levels, "
                f"in-degree 1+Poisson(3), batch {a.batch}/GPU, scores U(-1,0), seed 3")
    return {
        "translit": f"config1 transliteration edit lattices, batch {a.batch}/GPU",
        "snips": f"config2 SNIPS char x tag grids, batch {a.batch}/GPU",
        "cipher-uni": f"config3 cipher unigram trellis T=1000, batch {a.batch}/GPU",
        "cipher-bi": f"config3 cipher bigram trellis T=1000, batch {a.batch}/GPU",
    }[a.workload]


def make_arcs(a, device, seed_offset=0, batch=None, arcs=None):
    from nfst_b200 import synth

    B = a.batch if batch is None else batch
    if a.workload == "dag":
        return synth.random_dag_batch(B, a.arcs if arcs is None else arcs, levels=a.levels, seed=3 + seed_offset,
                                      device=device)
    if a.workload == "translit":
        return synth.transliteration_batch(B, seed=seed_offset).to(device)
    if a.workload == "snips":
        return synth.snips_batch(B, seed=1 + seed_offset).to(device)
    return synth.cipher_batch(B, T=1000, bigram=a.workload == "cipher-bi", seed=2 + seed_offset, device=device)


def build_packed(a, device, arcs=None, rank=0, world=1):
    """This rank's share of the GLOBAL batch (a.batch x world lattices), generated + packed on the device in
    chunks of lattices (bounds the packer's temporaries) and collated.  At world > 1 every rank draws the same
    global batch (chunk c has the seed of its first lattice), counts the arcs of every lattice and keeps the
    lattices `nfst_b200.dist.shard_by_arcs` assigns to it (LPT bins by arc count)."""
    import torch

    from nfst_b200 import dist as nd
    from nfst_b200.pack import concat_packed

    per_lat = (a.arcs if arcs is None else arcs) if a.workload == "dag" else 700_000
    chunk = max(1, min(a.batch, 60_000_000 // max(per_lat, 1)))
    total = a.batch * world
    spans = [(d, min(chunk, total - d)) for d in range(0, total, chunk)]
    mine = None
    if world > 1:
        counts = []
        for d, n in spans:
            ab = make_arcs(a, device, seed_offset=d, batch=n, arcs=arcs)
            counts.append(torch.bincount(ab.arc_lattice, minlength=n).cpu())
            del ab
        mine = torch.zeros(total, dtype=torch.bool)
        mine[torch.tensor(nd.shard_by_arcs(torch.cat(counts).tolist(), world)[rank], dtype=torch.int64)] = True
    parts, scores = [], []
    build_packed.pack_s, build_packed.chunks = 0.0, []
    for d, n in spans:
        ab = make_arcs(a, device, seed_offset=d, batch=n, arcs=arcs)
        if mine is not None:
            sel = torch.nonzero(mine[d:d + n]).squeeze(1)
            if sel.numel() == 0:
                continue
            ab = ab.select(sel.to(device))
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        p, sc = ab.pack()
        if device != "cpu":
            torch.cuda.synchronize()
        build_packed.pack_s += time.perf_counter() - t0
        build_packed.chunks.append((p.n_arcs, time.perf_counter() - t0))
        p.arc_origin = torch.empty(0, dtype=torch.int64, device=device)  # not needed here; frees 8 B/arc
        parts.append(p)
        scores.append(sc)
        del ab
    packed = concat_packed(parts) if len(parts) > 1 else parts[0]
    return packed, torch.cat(scores)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a thread (every 2 ms -- the timed
    region is tens of milliseconds, far shorter than nvidia-smi's own sampling loop), nvidia-smi as a fallback."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.sm, self.reasons = [], set()
        self.mx = None
        self.proc = None
        self.nvml = None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "samples": len(sm),
                    "reasons": sorted(self.reasons), "source": "nvml, 2 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(kernel: str, arcs_per_gpu: int):
    """DRAM bytes per launch of `kernel` from the newest committed `ncu --set full` capture
    (profiles/r*_traffic.json), if it was taken on this very workload; else None."""
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                d = json.load(f)
            if int(d.get("workload_arcs_per_gpu", -1)) == int(arcs_per_gpu):
                return float(d["kernels"][kernel]["dram_bytes"]), os.path.basename(path)
        except Exception:
            continue
    return None, None


def cpu_baseline(a, seconds: float, threads: int = 0):
    """Time the C oracle (port of the reference recurrence, all host threads) on a bounded
    sample of the workload: as many lattices as fit in ~`seconds`."""
    from oracle import c_oracle

    # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    cores = (os.cpu_count() or c_oracle.max_threads()) if threads <= 0 else threads
    probe_n = max(1, min(a.batch, cores))
    ab = make_arcs(a, "cpu", seed_offset=777, batch=probe_n)
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(),
                        ab.n_states.numpy())
    t0 = time.perf_counter()
    c_oracle.forward_backward(ob, want_post=True, n_threads=cores, want_states=True)
    t_probe = time.perf_counter() - t0
    reps = max(1, int(seconds / max(t_probe, 1e-4)))
    reps = min(reps, 50)
    t0 = time.perf_counter()
    for _ in range(reps):
        c_oracle.forward_backward(ob, want_post=True, n_threads=cores, want_states=True)
    dt = time.perf_counter() - t0
    arcs = ob.n_arcs * reps
    return {"value": arcs / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{probe_n} lattices of the workload ({ob.n_arcs} arcs) x {reps} passes, float64 log-space "
                      f"fwd+bwd+posteriors, oracle/lattice_oracle.c, {cores} OpenMP threads"}, dt / reps, ob.n_arcs


def run_reference(a):
    """--impl reference: the CPU port of the reference recurrence alone, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    cores = os.cpu_count() or c_oracle.max_threads()  # torchrun exports OMP_NUM_THREADS=1
    n = max(1, min(a.batch, cores))
    ab = make_arcs(a, "cpu", seed_offset=777, batch=n)
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(),
                        ab.n_states.numpy())
    for _ in range(min(a.warmup, 2)):
        c_oracle.forward_backward(ob, n_threads=cores)
    steps = a.steps
    t0 = time.perf_counter()
    c_oracle.forward_backward(ob, n_threads=cores)
    t1 = time.perf_counter() - t0
    steps = max(1, min(a.steps, int(120.0 / max(t1, 1e-4))))
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.forward_backward(ob, n_threads=cores)
    dt = time.perf_counter() - t0
    val = ob.n_arcs * steps / dt
    sample = (f"{n} lattices of the workload ({ob.n_arcs} arcs) per step, float64 log-space fwd+bwd+posteriors, "
              f"oracle/lattice_oracle.c (C port of scorers.py:692-751; the Python reference cannot travel), "
              f"{cores} OpenMP threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": min(a.warmup, 2), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def other_configs(dev, steps, peak):
    """BASELINE.json configs 1, 2, 3 and 5 at their stated batch (config 4 is the sweep): the same two-pass step and
    Viterbi + backtrace, CUDA events over `steps` back-to-back calls after 3 warm-ups.  These are latency-bound shapes
    (tens to a thousand dependent levels of a few states each): arcs/s and ms are the figures, the roofline fraction is
    reported because the metric asks for it."""
    import torch

    import nfst_b200 as nb
    from nfst_b200 import ops, synth
    from nfst_b200.pack import concat_packed

    cases = [
        ("config1 transliteration B=32", lambda n, o: synth.transliteration_batch(n, seed=o), 32, 32),
        ("config2 SNIPS B=256", lambda n, o: synth.snips_batch(n, seed=1 + o), 256, 256),
        ("config3 cipher unigram T=1000 B=64", lambda n, o: synth.cipher_batch(n, T=1000, bigram=False, seed=2 + o, device=dev), 64, 64),
        ("config3 cipher bigram T=1000 B=64", lambda n, o: synth.cipher_batch(n, T=1000, bigram=True, seed=2 + o, device=dev), 64, 16),
        ("config5 Viterbi transliteration B=4096", lambda n, o: synth.transliteration_batch(n, seed=4 + o), 4096, 512),
        ("config5 Viterbi integer scores B=4096", lambda n, o: synth.transliteration_batch(n, seed=4 + o, integer_scores=True), 4096, 512),
    ]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > L2: these inputs are smaller than it

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / steps

    rows = []
    for name, gen, B, per_chunk in cases:
        parts, scs = [], []
        for o in range(0, B, per_chunk):
            pk, sc = gen(min(per_chunk, B - o), o).to(dev).pack()
            parts.append(pk)
            scs.append(sc)
        pk = concat_packed(parts) if len(parts) > 1 else parts[0]
        sc = torch.cat(scs)
        A, S = pk.n_arcs, pk.n_states
        cols = all(g.sell or g.tiles for g in pk.groups)
        bb = torch.empty(S, dtype=ops.resolve_state_dtype(pk), device=dev) if pk.has_columns else None

        def fb():
            lz, al, cd = ops.lattice_pull(pk, arc_scores=sc, beta_out=bb)
            nb.lattice_backward(pk, arc_scores=sc, alpha=al, logz=lz, cond=cd, want_beta=not cols, want_post=True)

        ms_fb = timed(fb)
        ms_v = timed(lambda: nb.lattice_viterbi(pk, arc_scores=sc))
        ms_graph = None
        if not (cols or all(g.small_max_arcs > 0 for g in pk.groups)):
            # CSR / level-major groups are many launches per step (one per level for wide levels): the same
            # forward-backward replayed from a CUDA graph (ops.CapturedForwardBackward), scores copied in per step
            cap = ops.CapturedForwardBackward(pk)
            ms_graph = timed(lambda: cap.run(sc))
            del cap
        rows.append({
            "config": name, "arcs": A, "states": S, "levels": pk.max_levels,
            "state_dtype": str(ops.resolve_state_dtype(pk)).replace("torch.", ""),
            "execution": "small-lattice kernel" if all(g.small_max_arcs > 0 for g in pk.groups) else
                         "tile-stream" if all(g.tiles for g in pk.groups) else "CSR / level-major",
            "fwd_bwd_ms": ms_fb, "fwd_bwd_arcs_per_s": A / (ms_fb * 1e-3), "fwd_bwd_frac": (20 * A + 20 * S) / (ms_fb * 1e-3) / 1e9 / peak,
            "fwd_bwd_graph_ms": ms_graph, "viterbi_ms": ms_v, "viterbi_arcs_per_s": A / (ms_v * 1e-3)})
        del pk, sc, parts, scs, bb
        torch.cuda.empty_cache()
    return rows


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    import nfst_b200 as nb
    from nfst_b200 import _lib, ops
    from nfst_b200 import dist as nd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nfst_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().nfst_device_info(local, None, None, None, None))

    packed, scores = build_packed(a, dev, rank=rank, world=world)
    pack_ms = 1e3 * build_packed.pack_s
    pack_chunks = list(build_packed.chunks)
    A, S, B = packed.n_arcs, packed.n_states, packed.n_lattices
    B_global = a.batch * world
    loss = torch.zeros(1, device=dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * a.steps + 2)]

    all_sell = all(g.sell or g.tiles for g in packed.groups)  # column-major groups only: two passes, no alpha
    all_tiles = all(g.tiles for g in packed.groups)
    beta_buf = torch.empty(S, dtype=ops.resolve_state_dtype(packed), device=dev) if packed.has_columns else None

    def sweep_step(pk, sc_):
        """the same two passes on another batch (the --sweep points)"""
        bb = torch.empty(pk.n_states, dtype=ops.resolve_state_dtype(pk), device=dev) if pk.has_columns else None
        sell_only = all(g.sell or g.tiles for g in pk.groups)

        def run():
            lz, al, cd = ops.lattice_pull(pk, arc_scores=sc_, beta_out=bb)
            nb.lattice_backward(pk, arc_scores=sc_, alpha=al, logz=lz, cond=cd, want_beta=not sell_only, want_post=True)
        return run

    def step(i=None):
        """first pass: logZ (+ beta and the arc conditionals for sliced-column groups, alpha for CSR groups);
        second pass: arc posteriors (+ beta for CSR groups)."""
        if i is not None:
            ev[3 * i].record()
        logz, alpha, cond = ops.lattice_pull(packed, arc_scores=scores, beta_out=beta_buf)
        if i is not None:
            ev[3 * i + 1].record()
        r = nb.lattice_backward(packed, arc_scores=scores, alpha=alpha, logz=logz, cond=cond, want_beta=not all_sell,
                                want_post=True)
        if i is not None:
            ev[3 * i + 2].record()
        if world > 1:
            # the path's only collective: nfst_b200.dist.all_reduce_loss_and_grad (here the loss; the theta-mode leg
            # below adds dtheta[V]).  Asynchronous, a ring of buffers deep, so that a step's kernels never queue
            # behind the slowest rank's previous step; every all-reduce is waited for before the timed region ends.
            k = step.count % len(pending)
            step.count += 1
            if pending[k] is not None:
                pending[k].wait()
            pending[k] = nd.all_reduce_loss_and_grad(logz.sum(), None, async_op=True)
        return r

    step.count = 0
    pending = [None] * (4 if world > 1 else 0)

    def drain():
        for k, w in enumerate(pending):
            if w is not None:
                w.wait()
                pending[k] = None

    for _ in range(max(a.warmup, 3)):
        step()
    drain()
    torch.cuda.synchronize()
    launches0 = ops.launch_count
    clocks = ClockSampler(local)  # every rank samples its own GPU
    clocks.start()
    # all ranks enter the timed region together: the barrier is the LAST thing before it (anything rank-local in
    # between -- NVML start-up took milliseconds on some ranks -- becomes waiting time at the first all-reduce)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # ... and once more on the DEVICE timeline: the hosts leave dist.barrier() up to a millisecond or two apart (one
        # rank of eight entered 1.6 ms late in a 12 ms region: the other seven then spend that skew waiting at every loss
        # all-reduce and read 0.09 ms/step "between kernels"); a 4-byte all-reduce that the current stream waits for makes
        # every rank's start event fire when the LAST rank has arrived, while the early hosts already enqueue their steps
        gate = torch.zeros(1, device=dev)
        dist.all_reduce(gate)
    h0 = time.perf_counter()
    t_start.record()
    for i in range(a.steps):
        step(i)
    drain()  # the timed region ends when the last loss all-reduce has completed
    t_end.record()
    host_ms = 1e3 * (time.perf_counter() - h0) / a.steps  # host time to ENQUEUE a step (the device must not starve)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    ms_total = t_start.elapsed_time(t_end)
    launches = (ops.launch_count - launches0) // a.steps
    fwd_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(a.steps)) / a.steps
    bwd_ms = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(a.steps)) / a.steps
    mine = torch.tensor([ms_total / a.steps, fwd_ms, bwd_ms, host_ms, float(A), float(B), float(clk.get("sm_mhz") or 0.0),
                         float(len(clk.get("reasons") or []))], device=dev, dtype=torch.float64)
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = torch.stack(allr).cpu()
    else:
        per_rank = mine.cpu().unsqueeze(0)
    ms_step = float(per_rank[:, 0].max())  # the slowest rank
    arcs_all = float(per_rank[:, 4].sum())
    value = arcs_all / (ms_step * 1e-3)
    ranks = [{"rank": r, "ms_per_step": float(v[0]), "first_kernel_ms": float(v[1]), "second_kernel_ms": float(v[2]),
              "between_kernels_ms": float(v[0] - v[1] - v[2]), "host_enqueue_ms": float(v[3]), "arcs": int(v[4]),
              "lattices": int(v[5]), "sm_mhz": float(v[6]), "throttle_reasons": int(v[7])} for r, v in enumerate(per_rank)]

    # ---- e2e: host (pinned) buffers -> H2D -> fwd+bwd -> D2H logZ ----
    e2e = None
    if not a.no_e2e:
        if all_tiles:  # the tile-stream kernels read the byte stream (16-bit ring slots + headers) and three small tables
            fields = ("state_off", "start_state", "level_off", "tile_stream", "tile_tab", "tile_lw_off", "tile_lat_info")
        elif all_sell:  # the sliced-column kernels read nothing else (no in-order arrays, no chunk lists)
            fields = ("state_off", "start_state", "level_off", "level_ptr", "out_ptr", "dst_out", "out_deg8",
                      "sell_desc", "sell_lvl_slice")
        else:
            fields = ("state_off", "start_state", "sink_off", "sinks", "in_ptr", "src_in", "in2out", "out_ptr", "dst_out",
                      "lanes_in_log2", "lanes_out_log2", "fwd_chunk_off", "fwd_chunks", "bwd_chunk_off", "bwd_chunks",
                      "fwd_gather", "bwd_order", "out_deg8", "level_off", "level_ptr")
        host = {f: getattr(packed, f).cpu().pin_memory() for f in fields}
        host_scores = scores.cpu().pin_memory()
        host_ids = [g.ids.cpu().pin_memory() for g in packed.groups]
        logz_host = torch.empty(B, dtype=torch.float32).pin_memory()
        h2d = sum(t_.numel() * t_.element_size() for t_ in host.values()) + host_scores.numel() * 4 + sum(x.numel() * 4 for x in host_ids)
        d2h = B * 4
        import dataclasses

        from nfst_b200.pack import PackedLattices

        # device-side landing buffers (allocated once; refilled from pinned host memory every step)
        land = {f: (PackedLattices.alloc_padded(host[f].numel(), host[f].dtype, dev) if f in PackedLattices._ARC_FIELDS
                    else torch.empty_like(host[f], device=dev)) for f in fields}
        land_scores = torch.empty_like(host_scores, device=dev)
        land_ids = [torch.empty_like(h, device=dev) for h in host_ids]

        def e2e_step():
            for f in fields:
                land[f].copy_(host[f], non_blocking=True)
            for d_, h_ in zip(land_ids, host_ids):
                d_.copy_(h_, non_blocking=True)
            land_scores.copy_(host_scores, non_blocking=True)
            # arrays the kernels of this batch never read (labels when scores are per-arc, host-side
            # bookkeeping, and for sliced-column batches the in-order arrays and chunk lists) stay resident
            kw = {f: getattr(packed, f) for f in PackedLattices._INT_FIELDS}
            kw.update(lanes_in_log2=packed.lanes_in_log2, lanes_out_log2=packed.lanes_out_log2, out_deg8=packed.out_deg8,
                      tile_stream=packed.tile_stream, src_out=packed.src_out, orig_state=packed.orig_state, arc_origin=packed.arc_origin,
                      arc_off=packed.arc_off, n_levels=packed.n_levels)
            kw.update(land)
            groups = [dataclasses.replace(g, ids=d_) for g, d_ in zip(packed.groups, land_ids)]
            p = PackedLattices(n_lattices=B, n_states=S, n_arcs=A, vocab=packed.vocab, static_scores=None,
                               dense_shape=None, groups=groups, max_levels=packed.max_levels, stats=packed.stats, **kw)
            # the training-step surface: logZ (first pass), arc posteriors (second pass)
            logz, alpha, cond = ops.lattice_pull(p, arc_scores=land_scores)
            nb.lattice_backward(p, arc_scores=land_scores, alpha=alpha, logz=logz, cond=cond, want_beta=not all_sell,
                                want_post=True)
            logz_host.copy_(logz, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(logz_host[0])

        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        n_e2e = max(3, min(a.steps, 10))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": arcs_all / (float(te[0]) / n_e2e), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e, "ms_per_step": 1e3 * float(te[0]) / n_e2e}
        del host, host_scores

    # ---- theta mode: arc score = theta[label], the parametrisation of nFST's WFSTScorer (scorers.py:1663-1687).
    # A training step = logZ (pull pass), d loss / d theta[V] (flow pass with the label histogram, no per-arc output)
    # and the path's collective, nfst_b200.dist.all_reduce_loss_and_grad({sum logZ, dtheta[V]}), every step.  The
    # structure stays resident; end to end only theta[V] travels up and {loss, dtheta[V]} down.
    theta_leg = None
    if not a.no_theta:
        V = packed.vocab
        theta_host = (-torch.rand(V)).pin_memory()
        theta_dev = theta_host.to(dev)
        res_host = torch.empty(1 + V, dtype=torch.float32).pin_memory()

        def theta_step(end_to_end=False):
            if end_to_end:
                theta_dev.copy_(theta_host, non_blocking=True)
            logz, alpha, cond = ops.lattice_pull(packed, theta=theta_dev, beta_out=beta_buf)
            r = nb.lattice_backward(packed, theta=theta_dev, alpha=alpha, logz=logz, cond=cond, want_beta=not all_sell,
                                    want_post=False, want_dtheta=True)
            loss, dth = nd.all_reduce_loss_and_grad(logz.sum(), r["dtheta"])
            if end_to_end:
                res_host[:1].copy_(loss.reshape(1), non_blocking=True)
                res_host[1:].copy_(dth, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return loss

        for _ in range(3):
            theta_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.all_reduce(torch.zeros(1, device=dev))  # the start events aligned on the device timeline, as above
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(a.steps):
            theta_step()
        q1.record()
        torch.cuda.synchronize()
        for _ in range(2):
            theta_step(True)
        if world > 1:
            dist.barrier()
        n_t = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_t):
            theta_step(True)
        dt = time.perf_counter() - t0
        tt = torch.tensor([q0.elapsed_time(q1) / a.steps, 1e3 * dt / n_t], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        theta_leg = {"ms_per_step": float(tt[0]), "arcs_per_s": arcs_all / (float(tt[0]) * 1e-3),
                     "collective": "nfst_b200.dist.all_reduce_loss_and_grad({sum logZ, dtheta[%d]}) every step" % V,
                     "e2e": {"value": arcs_all / (float(tt[1]) * 1e-3), "unit": UNIT, "ms_per_step": float(tt[1]),
                             "h2d_bytes_per_step": 4 * V, "d2h_bytes_per_step": 4 * (V + 1),
                             "note": "structure resident in HBM; theta[V] up, {loss, dtheta[V]} down (pinned host buffers)"}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    if all_sell:
        # column-major passes: pull = dst + score read, cond write per arc, beta write per
        # state; flow = dst + cond read, post write per arc.  Their sum, 24 A + 4 S, equals SURVEY 8(d)'s
        # 20 A + 20 S at S = A/4.  (The tile-stream kernels read destinations as 16-bit ring slots: they move
        # fewer bytes than this algorithmic count.)
        names = (("tile_pull_kernel (beta + logZ + arc conditionals)", "tile_flow_kernel (arc posteriors)") if all_tiles else
                 ("sell_pull_kernel (beta + logZ + arc conditionals)", "sell_flow_kernel (arc posteriors)"))
        fwd_bytes, bwd_bytes = 12 * A + 4 * S, 12 * A
    else:
        names = ("nfst_fwd_kernel", "nfst_bwd_kernel (fused beta + posteriors)")
        fwd_bytes, bwd_bytes = 8 * A + 8 * S, 12 * A + 12 * S
    first = {"kernel": names[0], "achieved": fwd_bytes / (fwd_ms * 1e-3) / 1e9, "kernel_ms": fwd_ms,
             "algorithmic_bytes_per_launch": fwd_bytes}
    second = {"kernel": names[1], "achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "kernel_ms": bwd_ms,
              "algorithmic_bytes_per_launch": bwd_bytes}
    dom, other = (first, second) if fwd_ms >= bwd_ms else (second, first)
    traffic, traffic_src = profiled_traffic(dom["kernel"].split(" ")[0], A)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "lattices_per_gpu": B, "arcs_per_gpu": A, "states_per_gpu": S,
                   "levels": packed.max_levels, "global_batch": B_global, "parallelism": f"dp{world} (global batch of {B_global} lattices sharded by arc count with nfst_b200.dist.shard_by_arcs; one "
                                  f"all_reduce_loss_and_grad per step)",
                   "l2": "inputs larger than L2 (no flush)" if 20 * A > 2 * 126e6 else "inputs fit in L2 (no flush; latency-bound config)",
                   "scores": "per-arc fp32, canonical order",
                   "execution": ("tile-stream (nfst_tiles.cu): " + ", ".join(f"{g.n} lattices x {g.block_threads // 32} warps, ring {g.tile_ring}"
                                                                              for g in packed.groups)) if all_tiles
                   else "sliced-column (nfst_sell.cu)" if all_sell else "CSR kernels (nfst_kernels.cu)"},
        "gpu_launches": launches,
        "clocks": dict(clk, min_sm_mhz_over_ranks=min(r["sm_mhz"] for r in ranks), ranks_with_throttle_reasons=sum(1 for r in ranks if r["throttle_reasons"])),
        "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"],
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom["achieved"] / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "kernel_ms": dom["kernel_ms"],
                     "other_kernel": dict(other, frac=other["achieved"] / peak),
                     "step": {"achieved": (20 * A + 20 * S) / (ms_step * 1e-3) / 1e9,
                              "frac": (20 * A + 20 * S) / (ms_step * 1e-3) / 1e9 / peak}},
    }
    if e2e:
        out["e2e"] = e2e
    if theta_leg:
        theta_leg["frac"] = (20 * arcs_all / world + 20 * S) / (theta_leg["ms_per_step"] * 1e-3) / 1e9 / peak
        out["theta_step"] = theta_leg
    out["ranks"] = ranks
    out["pack"] = {"ms": pack_ms, "arcs_per_s": A / (pack_ms * 1e-3), "when": "once per batch, outside the timed region",
                   "chunks_ms": [round(1e3 * t, 1) for _, t in pack_chunks],
                   "warm_arcs_per_s": (pack_chunks[-1][0] / pack_chunks[-1][1]) if len(pack_chunks) > 1 else None,
                   "note": "the first chunk includes one-time costs (library and torch kernel loading, allocator growth)",
                   "packer": "tensor-op packer (nfst_b200/pack.py + tiles.py) for these wide lattices; the lattices nFST builds "
                             "(configs 1, 2, 5) pack on the device with nfst_pack_small: 1.2 ms for config 1 from its dense tables, 3.5 ms for "
                             "config 5 (tools/pack_profile.py)"}
    if not a.no_cpu and world == 1:  # CPU baseline: rank 0 at N=1 only
        cb, _, _ = cpu_baseline(a, a.cpu_seconds)
        out["cpu_baseline"] = cb
    # ---- Viterbi + backtrace on the same batch (BASELINE metric's second half; bit-exact paths) ----
    if world == 1:
        for _ in range(3):
            nb.lattice_viterbi(packed, arc_scores=scores)
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(a.steps):
            vit = nb.lattice_viterbi(packed, arc_scores=scores)
        v1.record()
        torch.cuda.synchronize()
        vms = v0.elapsed_time(v1) / a.steps
        vbytes = packed.algorithmic_bytes_viterbi(int(vit[1][-1]))
        out["viterbi"] = {"arcs_per_s": A / (vms * 1e-3), "ms_per_step": vms, "algorithmic_bytes": vbytes,
                          "gbs": vbytes / (vms * 1e-3) / 1e9, "frac": vbytes / (vms * 1e-3) / 1e9 / peak,
                          "note": "tropical pull pass + backtrace + ragged path read-out (one host read of the total length)"}
        del vit
    if not a.no_sweep and world == 1 and a.workload == "dag":
        sweep = []
        for arcs in (10_000, 30_000, 100_000, 300_000, 1_000_000):
            if arcs == a.arcs:  # the bench point itself
                sweep.append({"arcs_per_lattice": arcs, "arcs": A, "ms_per_step": ms_step,
                              "execution": out["config"]["execution"].split(" (")[0], "arcs_per_s": value,
                              "gbs": (20 * A + 20 * S) / (ms_step * 1e-3) / 1e9, "frac": (20 * A + 20 * S) / (ms_step * 1e-3) / 1e9 / peak})
                continue
            del packed, scores
            torch.cuda.empty_cache()
            packed, scores = build_packed(a, dev, arcs=arcs)
            run = sweep_step(packed, scores)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(a.steps):
                run()
            s1.record()
            torch.cuda.synchronize()
            ms = s0.elapsed_time(s1) / a.steps
            sweep.append({"arcs_per_lattice": arcs, "arcs": packed.n_arcs, "ms_per_step": ms,
                          "execution": "tile-stream" if all(g.tiles for g in packed.groups) else
                          "sliced-column" if all(g.sell for g in packed.groups) else "CSR/mixed",
                          "arcs_per_s": packed.n_arcs / (ms * 1e-3),
                          "gbs": (20 * packed.n_arcs + 20 * packed.n_states) / (ms * 1e-3) / 1e9,
                          "frac": (20 * packed.n_arcs + 20 * packed.n_states) / (ms * 1e-3) / 1e9 / peak})
        out["sweep"] = sweep
    if not (a.no_configs or a.no_sweep) and world == 1 and a.workload == "dag":
        del packed, scores
        torch.cuda.empty_cache()
        out["configs"] = other_configs(dev, max(a.steps, 5), peak)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
